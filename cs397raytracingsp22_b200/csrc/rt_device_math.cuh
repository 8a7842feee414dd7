// rt_device_math.cuh — vector helpers, the RNG contract, camera rays (Q8) and the work-index -> pixel map
// Part of the single translation unit rt_kernels.cu (everything here is __forceinline__ device code).
#ifndef RT_DEVICE_MATH_CUH
#define RT_DEVICE_MATH_CUH

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "rt_config.cuh"
#include "rt_kernels.h"

namespace rt {

// ------------------------------------------------------------------ small vector helpers
struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 mulv(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // (x+y)+z
__device__ __forceinline__ float mag2(f3 a) { return dot(a, a); }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ f3 normalize(f3 a) { return a * (1.0f / sqrtf(mag2(a))); }  // v * (1/|v|)
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float4 ldq(const void* base, uint32_t quad) {
  return __ldg(reinterpret_cast<const float4*>(base) + quad);
}
// Cache-hinted variants for the two record streams of the traversal loop (compile-time experiments, both off:
// see profiles/r2_notes.md C7).  RT_NODE_EVICT_LAST: BVH nodes ask L1 to keep them longest; RT_TRI_NO_ALLOC:
// triangle records (rarely reused) do not allocate in L1.
#ifndef RT_NODE_EVICT_LAST
#define RT_NODE_EVICT_LAST 0
#endif
#ifndef RT_TRI_NO_ALLOC
#define RT_TRI_NO_ALLOC 0
#endif
__device__ __forceinline__ float4 ldq_node(const float4* p) {
#if RT_NODE_EVICT_LAST
  float4 v;
  asm("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ float4 ldq_tri(const void* base, uint32_t quad) {
#if RT_TRI_NO_ALLOC
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(reinterpret_cast<const float4*>(base) + quad));
  return v;
#else
  return __ldg(reinterpret_cast<const float4*>(base) + quad);
#endif
}
__device__ __forceinline__ uint32_t fbits(float f) { return __float_as_uint(f); }

// ------------------------------------------------------------------ RNG contract
struct u4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return u4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
#define RT_BOUNCE_CAMERA 0xFFFFFFFFu
#define RT_PI 3.14159265358979323846f

// uniform point of the unit ball (stands in for rand_sphere_vec, tracing.rs:71-79; NOT normalised)
__device__ __forceinline__ f3 ball_from(uint32_t a, uint32_t b, uint32_t c) {
  float rad = cbrtf(u01(a));
  float zc = 1.0f - 2.0f * u01(b);
  float s = sqrtf(fmaxf(0.0f, 1.0f - zc * zc));
  float phi = (2.0f * RT_PI) * u01(c);
  float sn, cs;
  sincosf(phi, &sn, &cs);
  return mk(rad * s * cs, rad * zc, rad * s * sn);
}
// uniform point of the unit disk (rand_disk_vec, tracing.rs:81-89)
__device__ __forceinline__ f3 disk_from(uint32_t a, uint32_t b) {
  float rr = sqrtf(u01(a));
  float phi = (2.0f * RT_PI) * u01(b);
  float sn, cs;
  sincosf(phi, &sn, &cs);
  return mk(rr * cs, rr * sn, 0.0f);
}

// ------------------------------------------------------------------ camera (Q8)
__device__ __forceinline__ void camera_ray(const rt_frame& fr, uint32_t x, uint32_t y, uint32_t pixel, uint32_t i,
                                           f3& origin, f3& direction) {
  u4 r = philox4x32_10(pixel, i, RT_BOUNCE_CAMERA, 0u, fr.k0, fr.k1);
  float rand_x = (float)__umulhi(r.x, fr.spp);
  float rand_y = (float)__umulhi(r.y, fr.spp);
  float subpixel_x = (float)(i / fr.rooti);
  float subpixel_y = (float)(i % fr.rooti);
  float ps = fr.pixel_size, n = fr.n, rootn = fr.rootn;
  float off_x = (subpixel_x - 0.5f * rootn) * ps / rootn + (rand_x - 0.5f * n) * ps / n;
  float off_y = (subpixel_y - 0.5f * rootn) * ps / rootn + (rand_y - 0.5f * n) * ps / n;
  f3 center = mk(ps * ((float)x - 0.5f * (float)fr.width + 0.5f) + off_x,
                 ps * (0.5f + 0.5f * (float)fr.height - (float)y) + off_y, -fr.focal_length);
  f3 focus = normalize(center) * fr.focus_dist;
  // lens_radius == 0: 0 * disk is (+-0, +-0, 0) and changes nothing below, so the sin/cos are skipped
  f3 lens = fr.lens_radius != 0.0f ? fr.lens_radius * disk_from(r.z, r.w) : mk(0.0f, 0.0f, 0.0f);
  f3 dcam = normalize(focus - lens);
  f3 c0 = mk(fr.rot0[0], fr.rot0[1], fr.rot0[2]), c1 = mk(fr.rot1[0], fr.rot1[1], fr.rot1[2]),
     c2 = mk(fr.rot2[0], fr.rot2[1], fr.rot2[2]);
  f3 rl = c0 * lens.x + c1 * lens.y + c2 * lens.z;
  origin = mk(fr.eye[0], fr.eye[1], fr.eye[2]) + rl;
  if (fr.ortho) {
    // CameraProjectionMode::Orthographic, tracing.rs:196,200 as written: the camera-space pixel centre is used as a
    // WORLD position (eyepoint ignored) and view_dir is rotated once more by `rotation`
    origin = mk(center.x, center.y, 0.0f);
    dcam = mk(fr.view_dir[0], fr.view_dir[1], fr.view_dir[2]);
  }
  direction = c0 * dcam.x + c1 * dcam.y + c2 * dcam.z;
}

// work index -> (pixel, sample); false when a tile slot falls outside the image
__device__ __forceinline__ bool work_to_pixel(const rt_frame& fr, unsigned long long g, uint32_t& x, uint32_t& y,
                                              uint32_t& sample) {
  unsigned long long pl;
  if (fr.sample_major == 2u) {
    // groups of 32 sample indices: a warp is 32 consecutive samples of ONE pixel (as coherent as camera rays get), but
    // consecutive warps walk the pixels, so the wavefront spans the whole shard instead of a few image rows
    unsigned long long w = g >> 5;
    unsigned long long q = w / fr.pixel_slots;
    pl = w - q * fr.pixel_slots;
    sample = fr.sample_begin + (uint32_t)q * 32u + (uint32_t)(g & 31ull);
  } else if (fr.sample_major) {
    // a warp is 32 neighbouring pixels at one sample index; the wavefront then spans the whole shard at a few
    // sample indices instead of a few pixels at all of theirs
    if (g < 0x100000000ull && fr.pixel_slots < 0x100000000ull) {
      uint32_t q = (uint32_t)g / (uint32_t)fr.pixel_slots;
      pl = (uint32_t)g - q * (uint32_t)fr.pixel_slots;
      sample = fr.sample_begin + q;
    } else {
      unsigned long long q = g / fr.pixel_slots;
      pl = g - q * fr.pixel_slots;
      sample = fr.sample_begin + (uint32_t)q;
    }
  } else if (g < 0x100000000ull) {  // 32-bit division is several times cheaper and covers shards of up to 4 Gi paths
    uint32_t q = (uint32_t)g / fr.sample_count;
    pl = q;
    sample = fr.sample_begin + ((uint32_t)g - q * fr.sample_count);
  } else {
    pl = g / fr.sample_count;
    sample = fr.sample_begin + (uint32_t)(g - pl * fr.sample_count);
  }
  if (fr.shard_mode == RT_SHARD_TILES) {
    uint32_t ts = fr.tile_size, ts2 = ts * ts;
    uint32_t k = (uint32_t)(pl / ts2), within = (uint32_t)(pl - (unsigned long long)k * ts2);
    uint32_t tile = __ldg(fr.tile_list + k);  // which tiles a shard owns is the host's decision (plan_shard)
    uint32_t tx = tile % fr.tiles_x, ty = tile / fr.tiles_x;
    x = tx * ts + within % ts;
    y = ty * ts + within / ts;
    return x < fr.width && y < fr.height;
  }
  uint32_t p = (uint32_t)pl;
  y = p / fr.width;
  x = p - y * fr.width;
  return true;
}

}  // namespace rt
#endif
