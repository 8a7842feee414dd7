// rt_materials.cuh — Material::scatter building blocks (materials.rs, tracing.rs:54-97), the ray sort key and fixed-point accumulation
// Part of the single translation unit rt_kernels.cu (everything here is __forceinline__ device code).
#ifndef RT_MATERIALS_CUH
#define RT_MATERIALS_CUH

#include "rt_device_math.cuh"

namespace rt {

// ------------------------------------------------------------------ materials
__device__ __forceinline__ f3 reflectv(f3 v, f3 n) { return v - 2.0f * dot(v, n) * n; }  // tracing.rs:54-56
__device__ __forceinline__ float fresnelf(f3 v, f3 n, float ir) {                         // tracing.rs:58-62
  float q = (ir - 1.0f) / (ir + 1.0f);
  float r0 = q * q;
  float x = 1.0f - fabsf(dot(v, n));
  float x2 = x * x;
  return r0 + (1.0f - r0) * (x * (x2 * x2));
}
__device__ __forceinline__ f3 refractv(f3 v, f3 n, float eta) {  // tracing.rs:64-69
  float cos_theta = fminf(dot(-v, n), 1.0f);
  f3 perp = eta * (v + cos_theta * n);
  f3 par = -sqrtf(fabsf(1.0f - mag2(perp))) * n;
  return perp + par;
}
__device__ __forceinline__ bool ulps_eq(float a, float b) {  // approx::ulps_eq!, epsilon = EPSILON, 4 ulps
  if (fabsf(a - b) <= 1.1920929e-7f) return true;
  if (signbit(a) != signbit(b)) return false;
  long long d = (long long)__float_as_int(a) - (long long)__float_as_int(b);
  return (d < 0 ? -d : d) <= 4;
}
// sample_hemisphere, materials.rs:171-178: ball with y=|y|, rotated by Quaternion::from_arc(unit_y, n)
__device__ __forceinline__ f3 sample_hemisphere(f3 n, f3 ball) {
  f3 dir = mk(ball.x, fabsf(ball.y), ball.z);
  float mag_avg = sqrtf(1.0f * mag2(n));
  float dt = 0.0f * n.x + 1.0f * n.y + 0.0f * n.z;  // dot(unit_y, n)
  float s;
  f3 v;
  if (ulps_eq(dt, mag_avg)) return dir;
  if (ulps_eq(dt, -mag_avg)) {
    s = -4.371139e-8f;
    v = mk(0.0f, 0.0f, 1.0f);
  } else {
    s = mag_avg + dt;
    v = cross(mk(0.0f, 1.0f, 0.0f), n);
    float inv = 1.0f / sqrtf(s * s + mag2(v));
    s = s * inv;
    v = v * inv;
  }
  f3 tmp = cross(v, dir) + dir * s;
  return cross(v, tmp) * 2.0f + dir;
}

// Sort key of a scattered ray: a 15-bit spatial hash of the cell its origin lies in (cells of 1/16 of the TLAS
// box, NOT clamped to the box - floors and walls extend far beyond it and clamping would pile their rays into a few
// boundary bins) and the octant of its direction.  Rays that start close together and head the same way fetch the
// same nodes, so a k_trace warp built from one bin stays together much longer than 32 rays in material-queue order.
__device__ __forceinline__ uint32_t ray_sort_key(const rt_frame& fr, f3 o, f3 d) {
  int cx = __float2int_rd((o.x - fr.sort_min[0]) * fr.sort_scale[0]);
  int cy = __float2int_rd((o.y - fr.sort_min[1]) * fr.sort_scale[1]);
  int cz = __float2int_rd((o.z - fr.sort_min[2]) * fr.sort_scale[2]);
  uint32_t h = ((uint32_t)cx * 73856093u) ^ ((uint32_t)cy * 19349663u) ^ ((uint32_t)cz * 83492791u);
  uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
  if (fr.sort_use_octant == 2u) {
    // 5 direction bits: octant + dominant axis (24 classes = the 6 cube faces x 4 quadrants), 13-bit cell hash
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    uint32_t dom = (ax >= ay && ax >= az) ? 0u : (ay >= az ? 1u : 2u);
    h = (h ^ (h >> 13)) & ((RT_SORT_BINS >> 5) - 1u);  // 13 bits with the default 2^18 bins
    return (h << 5) | (dom << 3) | oct;
  }
  h = (h ^ (h >> 15)) & ((RT_SORT_BINS >> 3) - 1u);
  return (h << 3) | (fr.sort_use_octant ? oct : 0u);
}

// fixed-point accumulation (2^-30 units): order-independent, hence reproducible and shardable.
// Headroom: a sample is clamped to +-65536 = 2^46 units, so a pixel's int64 sum is exact for up to 2^17 samples that
// sit AT the clamp (radiance 65536), i.e. for any realistic image with aa_sample_count * ranks < 2^24.  A NaN sample sets
// a sticky bit per channel in the fourth word (bit 21*k; bits, not counts, so no number of NaN samples can carry into
// the neighbouring channel or wrap to zero; summing the words of up to 2^21 ranks keeps the fields apart).
#define RT_FIX_SCALE 1073741824.0f
#define RT_FIX_CLAMP 65536.0f
__device__ __forceinline__ void accum_add(long long* accum, uint32_t pixel, f3 c) {
  float v[3] = {c.x, c.y, c.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = v[k];
    if (x != x) {
      atomicOr(reinterpret_cast<unsigned long long*>(accum + (size_t)pixel * 4 + 3), 1ull << (21 * k));
    } else if (x != 0.0f) {
      x = fminf(fmaxf(x, -RT_FIX_CLAMP), RT_FIX_CLAMP);
      long long q = __float2ll_rn(x * RT_FIX_SCALE);
      atomicAdd(reinterpret_cast<unsigned long long*>(accum + (size_t)pixel * 4 + k), (unsigned long long)q);
    }
  }
}

}  // namespace rt
#endif
