// rt_kernels.cu — hand-written sm_100a kernels of the wavefront path tracer.
//
//   k_advance : 1 thread; wavefront bookkeeping between bounces (device-side, no host sync)
//   k_raygen  : Camera::generate_rays (tracing.rs:159-209) for the paths started this iteration
//   k_trace   : closest hit (Scene::intersect_ray tracing.rs:327-346 and everything under it:
//               geometry.rs:50-123,300-366,394-526).  Persistent warps with dynamic ray fetch.
//   k_sort    : material-sorted shade queues (match/ballot compaction)
//   k_shade   : hit resolution (RayHit::new tracing.rs:121-133, geometry.rs:253-298,350-363),
//               Material::scatter / emission (materials.rs:33-166) and the integrator step of
//               Scene::shade_ray (tracing.rs:300-324); one warp-uniform material class per warp
//   k_surface : parity hooks only: hit resolution for every ray, written out
//   k_resolve : mean + output transform (tracing.rs:241-256)
//
// Arithmetic contract: this file is compiled with -fmad=false.  Everything that decides WHICH
// primitive is hit (object-space transform, Möller–Trumbore, sphere, plane, volume entry/exit)
// is a sequence of single IEEE f32 operations in the reference's order, so primary-hit ids are
// bit-exact against the strict-IEEE CPU oracle.  BVH slab tests only cull, are conservative
// (padded boxes, non-strict compare) and use explicit FMAs and approximate reciprocals.
//
// No tensor cores: there is no dense contraction anywhere in this workload.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "rt_kernels.h"

namespace rt {

#ifndef RT_BLOCK
#define RT_BLOCK 128
#endif
#define RT_WARPS (RT_BLOCK / 32)
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 16   // traversal-stack entries per thread kept in shared memory
#endif
#define RT_LOCAL_STACK 48  // overflow entries (local memory; host checks depth <= 62)
// RT_STREAM_HINTS: ray queue / hit record traffic uses the streaming (evict-first) cache operators so that it does
// not push BVH nodes, triangles and texels out of L1/L2.  (Tried and dropped, profiles/r1_notes.md B2, B6: stack
// entries that carry their entry distance, and prefetching the children of the pushed child.)
#ifndef RT_STREAM_HINTS
#define RT_STREAM_HINTS 1
#endif
#if RT_STREAM_HINTS
#define RT_LDS(p) __ldcs(p)
#define RT_STS(p, v) __stcs(p, v)
#else
#define RT_LDS(p) (*(p))
#define RT_STS(p, v) (*(p) = (v))
#endif
#ifndef RT_OCTANT_SORT
#define RT_OCTANT_SORT 1
#endif
#ifndef RT_EXTEND_MIN_BLOCKS
#define RT_EXTEND_MIN_BLOCKS 8  // 64 registers per thread -> 32 resident warps per SM
#endif

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
    uint32_t tile = k * fr.shard_count + fr.shard_rank;
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

// ------------------------------------------------------------------ closest hit
struct Best {
  float t;
  float u, v;    // mesh barycentrics
  int obj;       // top-level object index, -1 = miss
  uint32_t prim; // original triangle index inside the mesh
};
struct Cnt {
  uint32_t nodes, tris, inst, prims, rounds, tlas_nodes;
};

// reference ordering of candidates: smaller t wins; equal t: earlier object wins (strict '<' in
// tracing.rs:335); same mesh and equal t: higher triangle index wins (geometry.rs:105-115,349)
__device__ __forceinline__ bool better(float t, int obj, uint32_t prim, const Best& b) {
  if (b.obj < 0) return true;
  if (t < b.t) return true;
  if (t > b.t) return false;
  if (obj != b.obj) return obj < b.obj;
  return prim > b.prim;
}

// conservative slab test; returns entry distance in tn
__device__ __forceinline__ bool slab(float4 lo, float4 hi, f3 inv, f3 oi, float tmin, float tmax, float& tn) {
  float x0 = __fmaf_rn(lo.x, inv.x, oi.x), x1 = __fmaf_rn(hi.x, inv.x, oi.x);
  float y0 = __fmaf_rn(lo.y, inv.y, oi.y), y1 = __fmaf_rn(hi.y, inv.y, oi.y);
  float z0 = __fmaf_rn(lo.z, inv.z, oi.z), z1 = __fmaf_rn(hi.z, inv.z, oi.z);
  float a = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
  float b = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
  tn = a;
  return a <= b * 1.0000005f;
}
// reciprocal direction for the slab tests only.  A component that is exactly (or nearly) zero is
// replaced by +-1e-20 so that lo*inv + oi never becomes inf - inf: the slab then yields two huge
// finite values of the right signs and does not constrain the interval, which is what a ray
// parallel to (and inside) the slab needs.  Camera rays through the image centre column have
// d.x == 0 exactly, so this case is routine, not exotic.
__device__ __forceinline__ float safe_rcp(float d) {
  float a = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));  // one MUFU.RCP; exactness is not needed here
  return r;
}
__device__ __forceinline__ f3 approx_inv(f3 d) { return mk(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)); }

// Sphere::intersect_ray core (geometry.rs:397-410): returns t or NaN-free miss flag
__device__ __forceinline__ bool sphere_t(f3 center, float radius, f3 o, f3 d, float t_min, float t_max, float& t) {
  f3 f = o - center;
  float a = mag2(d);
  float b = 2.0f * dot(f, d);
  float c = mag2(f) - radius * radius;
  float disc = b * b - 4.0f * a * c;
  if (disc < 0.0f) return false;
  float sq = sqrtf(disc);
  float t1 = (-b - sq) / (2.0f * a);
  float t2 = (-b + sq) / (2.0f * a);
  t = t1 >= t_min ? t1 : t2;
  return !(t < t_min || t > t_max);
}

// per-ray traversal state.  The short stack lives in shared memory ([depth][thread], conflict
// free); entries beyond RT_SMEM_STACK spill to a local-memory array that is almost never touched.
struct Trav {
  // The world-space ray and the RNG key are NOT kept in registers: they stay in the ray queue
  // (slot `slot` of A/B/C) and are re-read on the rare occasions they are needed (leaving an
  // instance, TLAS object tests, volume draws, hit resolution).  That keeps the persistent state
  // of k_extend small enough for 7-8 resident blocks per SM.
  const float4* qA;
  const float4* qB;
  const float4* qC;
  uint32_t slot;
  f3 o, d, inv, oi; // current-space ray (world or instance), reciprocal direction, -o*inv
  float t_min, t_max;
  float ray_t_max;  // the ray's own upper limit (t_max is temporarily replaced during a volume boundary query)
  uint32_t entry;   // packed node link being visited, RT_ENTRY_NONE when a pop is needed
  int sp;
  int cur_obj;
  bool in_blas;
  uint32_t sbase;    // shared-space byte address of this thread's stack column ([depth][thread] layout)
  uint32_t wbase;    // shared-space byte address of this thread's saved world-space (inv, oi), 6 words [k][thread]
  uint32_t vol_phase;  // mesh-bounded volume: 0 = none, 1 = entry query running, 2 = exit query running
  bool volret;         // pop_next just popped the end-of-query marker
  uint32_t* lstack;  // RT_LOCAL_STACK entries of local memory, declared by the kernel
  Best best;
  Cnt cnt;
  uint32_t k0, k1;  // Philox key (kernel constants)
  __device__ __forceinline__ void world_ray(f3& wo, f3& wd) const {
    float4 a = qA[slot], b = qB[slot];
    wo = mk(a.x, a.y, a.z);
    wd = mk(a.w, b.x, b.y);
  }

  __device__ __forceinline__ void push(uint32_t v) {
    if (sp < RT_SMEM_STACK) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + (uint32_t)sp * (RT_BLOCK * 4u)), "r"(v) : "memory");
    else lstack[sp - RT_SMEM_STACK] = v;
    ++sp;
  }
  __device__ __forceinline__ uint32_t pop() {
    --sp;
    uint32_t v;
    if (sp < RT_SMEM_STACK) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + (uint32_t)sp * (RT_BLOCK * 4u)) : "memory");
    else v = lstack[sp - RT_SMEM_STACK];
    return v;
  }
  __device__ __forceinline__ void set_space(f3 no, f3 nd) {
    o = no; d = nd;
    inv = approx_inv(d);
    oi = mk(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
  }
  // the world-space reciprocal direction survives an instance visit in shared memory (cheaper than 3 reciprocals)
  __device__ __forceinline__ void save_world_inv() const {
    const float v[6] = {inv.x, inv.y, inv.z, oi.x, oi.y, oi.z};
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("st.shared.f32 [%0], %1;" ::"r"(wbase + k * (RT_BLOCK * 4u)), "f"(v[k]) : "memory");
  }
  __device__ __forceinline__ void load_world_inv() {
    float v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(wbase + k * (RT_BLOCK * 4u)) : "memory");
    inv = mk(v[0], v[1], v[2]);
    oi = mk(v[3], v[4], v[5]);
  }
  // mesh-bounded volume: the closest hit found so far and t_entr wait in shared memory while the boundary queries run
  __device__ __forceinline__ void vol_save(const Best& b, float t_entr) const {
    const uint32_t v[6] = {__float_as_uint(b.t), __float_as_uint(b.u), __float_as_uint(b.v), (uint32_t)b.obj, b.prim,
                           __float_as_uint(t_entr)};
#pragma unroll
    for (int k = 0; k < 6; ++k)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(wbase + (6 + k) * (RT_BLOCK * 4u)), "r"(v[k]) : "memory");
  }
  __device__ __forceinline__ void vol_load(Best& b, float& t_entr) const {
    uint32_t v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(wbase + (6 + k) * (RT_BLOCK * 4u)) : "memory");
    b.t = __uint_as_float(v[0]); b.u = __uint_as_float(v[1]); b.v = __uint_as_float(v[2]);
    b.obj = (int)v[3]; b.prim = v[4];
    t_entr = __uint_as_float(v[5]);
  }
  // pop the next entry; false when the stack is empty.  A RESTORE marker switches back to the
  // world-space ray and leaves entry = NONE (the caller pops again).
  template <bool VOLMESH>
  __device__ __forceinline__ bool pop_next() {
    if (sp == 0) return false;
    entry = pop();
    if (VOLMESH && entry == RT_ENTRY_VOLRET) {
      volret = true;
      entry = RT_ENTRY_NONE;
      return true;
    }
    if (entry == RT_ENTRY_RESTORE) {
      world_ray(o, d);
      load_world_inv();
      in_blas = false;
      entry = RT_ENTRY_NONE;
    }
    return true;
  }
};

// The reference's AABB::intersect_ray (geometry.rs:52-68), exactly: strict, IEEE division, its own min/max order.
// Used only for GUARD boxes: thin interior boxes of the reference's index-order tree, which that test rejects for
// rays whose origin is far enough away that (min - o) == (max - o) in f32.  A triangle below such a box is a hit
// for the reference only if every guard above it lets the ray in (t_max un-narrowed).
__device__ __noinline__ bool guards_pass(const rt_dev_scene& sc, uint32_t first, uint32_t count, f3 o, f3 d, float t_min,
                                         float t_max) {
  const uint32_t* list = reinterpret_cast<const uint32_t*>(sc.guard_list);
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  for (uint32_t k = 0; k < count; ++k) {
    uint32_t g = __ldg(list + first + k);
    float4 lo = ldq(sc.guards, g * 2u), hi = ldq(sc.guards, g * 2u + 1u);
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
    float tmin = t_min, tmax = t_max;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) {
      float inv_d = 1.0f / dd[axis];
      float t0 = (mn[axis] - oo[axis]) * inv_d;
      float t1 = (mx[axis] - oo[axis]) * inv_d;
      if (inv_d < 0.0f) {
        float tmp = t0;
        t0 = t1;
        t1 = tmp;
      }
      tmin = fmaxf(t0, tmin);
      tmax = fminf(t1, tmax);
      if (tmax <= tmin) return false;
    }
  }
  return true;
}

// interior node: fetch the 64-byte child pair with four 128-bit read-only loads, test both boxes,
// continue with the nearer child and push the other
template <bool COUNT>
__device__ __forceinline__ void trav_interior(const rt_dev_scene& sc, Trav& T) {
  uint32_t e = T.entry;
  const float4* pair = reinterpret_cast<const float4*>(sc.nodes) + (size_t)e * 2u;
  float4 l0 = __ldg(pair), l1 = __ldg(pair + 1), r0 = __ldg(pair + 2), r1 = __ldg(pair + 3);
  if (COUNT) {
    T.cnt.nodes += 2;
    if (!T.in_blas) T.cnt.tlas_nodes += 2;
  }
  float tl, tr;
  bool hl = slab(l0, l1, T.inv, T.oi, T.t_min, T.best.t, tl);
  bool hr = slab(r0, r1, T.inv, T.oi, T.t_min, T.best.t, tr);
  uint32_t el = fbits(l0.w), er = fbits(r0.w);
  if (hl && hr) {
    bool lfirst = tl <= tr;
    uint32_t far = lfirst ? er : el;
    T.push(far);
    T.entry = lfirst ? el : er;
  } else {
    T.entry = hl ? el : (hr ? er : RT_ENTRY_NONE);
  }
}

// leaf: BLAS leaf = up to RT_MAX_LEAF_TRIS triangle records; TLAS leaf = one top-level object
template <bool COUNT, bool VOLMESH>
__device__ __forceinline__ void trav_leaf(const rt_dev_scene& sc, Trav& T) {
  const uint32_t first = (T.entry & ~RT_LEAF_FLAG) >> 4, n = T.entry & 15u;
  const float t_min = T.t_min, t_max = T.t_max;
  Best& best = T.best;
  T.entry = RT_ENTRY_NONE;
  if (T.in_blas) {
    // IndexedTriangle::intersect_ray, geometry.rs:333-349 (object space, un-normalised d)
    const f3 o = T.o, d = T.d;
    for (uint32_t k = 0; k < n; ++k) {
      uint32_t q = (first + k) * RT_TRI_QUADS;
      float4 a0 = ldq(sc.tris, q), a1 = ldq(sc.tris, q + 1), a2 = ldq(sc.tris, q + 2);
      if (COUNT) T.cnt.tris += 1;
      f3 va = mk(a0.x, a0.y, a0.z), e1 = mk(a0.w, a1.x, a1.y), e2 = mk(a1.z, a1.w, a2.x);
      f3 qv = cross(d, e2);
      float g = dot(e1, qv);
      if (fabsf(g) < 0.0001f) continue;
      float f = 1.0f / g;
      f3 s = o - va;
      float u = f * dot(s, qv);
      if (u < 0.0f) continue;
      f3 r = cross(s, e1);
      float v = f * dot(d, r);
      if (v < 0.0f || u + v > 1.0f) continue;
      float t = f * dot(e2, r);
      if (t < t_min || t > t_max) continue;
      uint32_t id = fbits(a2.y);
      if (fbits(a2.w) && !guards_pass(sc, fbits(a2.z), fbits(a2.w), o, d, t_min, t_max)) continue;
      if (better(t, T.cur_obj, id, best)) {
        best.t = t; best.u = u; best.v = v; best.obj = T.cur_obj; best.prim = id;
      }
    }
    return;
  }
  const f3 wo = T.o, wd = T.d;  // not inside an instance: the current space IS world space
  {
    const int obj = (int)first;  // one top-level object per TLAS leaf
    uint32_t q = (uint32_t)obj * RT_OBJ_QUADS;
    float4 h = ldq(sc.objects, q);
    int kind = (int)fbits(h.x);
    if (kind == RT_OBJ_MESH) {  // always alone in its leaf
      float4 r0 = ldq(sc.objects, q + 1), r1 = ldq(sc.objects, q + 2), r2 = ldq(sc.objects, q + 3);
      float4 m7 = ldq(sc.objects, q + 7);
      if (COUNT) T.cnt.inst += 1;
      // StaticMesh::intersect_ray, geometry.rs:304: transform_point / transform_vector
      f3 no = mk(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w * 1.0f, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w * 1.0f,
                 r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w * 1.0f);
      f3 nd = mk(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z + r0.w * 0.0f, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z + r1.w * 0.0f,
                 r2.x * wd.x + r2.y * wd.y + r2.z * wd.z + r2.w * 0.0f);
      uint32_t root = fbits(m7.x);
      if (root != RT_ENTRY_NONE) {
        T.push(RT_ENTRY_RESTORE);
        T.save_world_inv();
        T.set_space(no, nd);
        T.in_blas = true;
        T.cur_obj = obj;
        T.entry = root;
      }
      return;
    }
    if (VOLMESH && kind == RT_OBJ_VOLUME_MESH) {
      // ConvexVolume with a StaticMesh boundary (geometry.rs:505-510): the entry distance is the boundary's closest
      // hit over ALL t (f32::MIN..f32::MAX), found by traversing its BLAS with that range; k_trace's stack is a
      // stack, so the query simply nests: park the closest hit so far, run the query, continue in volume_continue()
      float4 r0 = ldq(sc.objects, q + 1), r1 = ldq(sc.objects, q + 2), r2 = ldq(sc.objects, q + 3);
      float4 m7 = ldq(sc.objects, q + 7);
      if (COUNT) T.cnt.inst += 1;
      f3 no = mk(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w * 1.0f, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w * 1.0f,
                 r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w * 1.0f);
      f3 nd = mk(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z + r0.w * 0.0f, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z + r1.w * 0.0f,
                 r2.x * wd.x + r2.y * wd.y + r2.z * wd.z + r2.w * 0.0f);
      uint32_t root = fbits(m7.x);
      if (root != RT_ENTRY_NONE) {
        T.vol_save(T.best, 0.0f);
        T.push(RT_ENTRY_VOLRET);
        T.save_world_inv();
        T.set_space(no, nd);
        T.in_blas = true;
        T.cur_obj = obj;
        T.vol_phase = 1;
        T.t_min = -CUDART_MAX_NORMAL_F;
        T.t_max = CUDART_MAX_NORMAL_F;
        T.best.t = CUDART_MAX_NORMAL_F;
        T.best.obj = -1;
        T.entry = root;
      }
      return;
    }
    if (COUNT) T.cnt.prims += 1;
    float4 q1 = ldq(sc.objects, q + 1);
    if (kind == RT_OBJ_SPHERE) {
      float t;
      if (sphere_t(mk(q1.x, q1.y, q1.z), q1.w, wo, wd, t_min, t_max, t) && better(t, obj, 0u, best)) {
        best.t = t; best.obj = obj; best.prim = 0;
      }
    } else if (kind == RT_OBJ_TRIANGLE) {
      // Triangle::intersect_ray, geometry.rs:433-447
      float4 q2 = ldq(sc.objects, q + 2), q3 = ldq(sc.objects, q + 3);
      f3 va = mk(q1.x, q1.y, q1.z), e1 = mk(q1.w, q2.x, q2.y), e2 = mk(q2.z, q2.w, q3.x);
      f3 qv = cross(wd, e2);
      float g = dot(e1, qv);
      if (!(fabsf(g) < 0.0001f)) {
        float f = 1.0f / g;
        f3 s = wo - va;
        float u = f * dot(s, qv);
        if (!(u < 0.0f)) {
          f3 r = cross(s, e1);
          float v = f * dot(wd, r);
          if (!(v < 0.0f || u + v > 1.0f)) {
            float t = f * dot(e2, r);
            if (!(t < t_min || t > t_max) && better(t, obj, 0u, best)) {
              best.t = t; best.obj = obj; best.prim = 0;
            }
          }
        }
      }
    } else if (kind == RT_OBJ_VOLUME) {
      // ConvexVolume::intersect_ray with a Sphere boundary, geometry.rs:505-525
      float4 q2 = ldq(sc.objects, q + 2);
      f3 c = mk(q1.x, q1.y, q1.z);
      float t_entr, t_exit;
      if (sphere_t(c, q1.w, wo, wd, -CUDART_MAX_NORMAL_F, CUDART_MAX_NORMAL_F, t_entr) &&
          sphere_t(c, q1.w, wo, wd, t_entr + 0.0001f, CUDART_MAX_NORMAL_F, t_exit) && !(t_exit < t_min || t_entr > t_max)) {
        float t_start = fmaxf(t_entr, t_min);
        float t_end = fminf(t_exit, t_max);
        float dist_in = t_end - t_start;
        uint32_t vi = fbits(q2.y);
        float4 kc = T.qC[T.slot];
        uint32_t pixel = fbits(kc.y), sb = fbits(kc.z);
        u4 rr = philox4x32_10(pixel, sb & 0xFFFFFFu, sb >> 24, 1u + (vi >> 2), T.k0, T.k1);
        uint32_t w = (vi & 3u) == 0 ? rr.x : ((vi & 3u) == 1 ? rr.y : ((vi & 3u) == 2 ? rr.z : rr.w));
        float dist_before = (-1.0f / q2.x) * logf(u01(w));
        if (dist_before < dist_in) {
          float t = t_start + dist_before;
          if (better(t, obj, 0u, best)) {
            best.t = t; best.obj = obj; best.prim = 0;
          }
        }
      }
    }
  }
}

// unbounded objects (planes) and anything the TLAS could not bound: always tested
template <bool COUNT>
__device__ __forceinline__ void test_unbounded(const rt_dev_scene& sc, Trav& T, f3 wo, f3 wd) {
  const int32_t* planes = reinterpret_cast<const int32_t*>(sc.planes);
  Best& best = T.best;
  for (uint32_t pi = 0; pi < sc.n_planes; ++pi) {
    int obj = __ldg(planes + pi);
    uint32_t q = (uint32_t)obj * RT_OBJ_QUADS;
    float4 h = ldq(sc.objects, q);
    int kind = (int)fbits(h.x);
    float4 q1 = ldq(sc.objects, q + 1), q2 = ldq(sc.objects, q + 2);
    if (COUNT) T.cnt.prims += 1;
    if (kind == RT_OBJ_PLANE) {
      // Plane::intersect_ray, geometry.rs:476-485
      f3 nrm = mk(q2.x, q2.y, q2.z);
      f3 to = wo - mk(q1.x, q1.y, q1.z);
      float od = dot(to, nrm);
      float sg = (od != od) ? od : (signbit(od) ? -1.0f : 1.0f);
      f3 n = sg * nrm;
      float dd = dot(wd, n);
      if (!(dd >= 0.0f)) {
        float t = fabsf(od) / fabsf(dd);
        if (!(t < T.t_min || t > T.t_max) && better(t, obj, 0u, best)) {
          best.t = t; best.obj = obj; best.prim = 0;
        }
      }
    }
  }
}

// start the closest-hit query of the ray in T.wo / T.wd
template <bool COUNT>
__device__ __forceinline__ void test_unbounded(const rt_dev_scene& sc, Trav& T, f3 wo, f3 wd);

template <bool COUNT>
__device__ __forceinline__ void trav_begin(const rt_dev_scene& sc, Trav& T, f3 wo, f3 wd) {
  T.set_space(wo, wd);
  T.best.t = T.t_max;
  T.best.obj = -1;
  T.best.prim = 0;
  T.best.u = T.best.v = 0.0f;
  T.cnt = Cnt{0, 0, 0, 0, 0, 0};
  T.sp = 0;
  T.in_blas = false;
  T.cur_obj = -1;
  T.vol_phase = 0;
  T.volret = false;
  // unbounded objects first: a plane hit (the floor is the most common hit of all) shortens the
  // interval before any node is fetched.  Candidate ordering is order independent (see better()).
  test_unbounded<COUNT>(sc, T, wo, wd);
  T.entry = sc.tlas_root;
  if (T.entry != RT_ENTRY_NONE) {
    float tn;
    float4 lo = make_float4(sc.tlas_min[0], sc.tlas_min[1], sc.tlas_min[2], 0.0f);
    float4 hi = make_float4(sc.tlas_max[0], sc.tlas_max[1], sc.tlas_max[2], 0.0f);
    if (!slab(lo, hi, T.inv, T.oi, T.t_min, T.best.t, tn)) T.entry = RT_ENTRY_NONE;
  }
}
// one round: descend while interior, then the leaf this lane reached (if any), then pop.
// Returns true when the stack is exhausted.  (Measured on B200: postponing leaves until every lane
// of the warp holds one - "while-while" - is 30 % slower here, because leaves are cheap (1.9
// triangle tests per ray) compared with the descent they would make the other lanes wait for.)
// A boundary query of a mesh-bounded volume has finished (its end marker was popped).  Phase 1 found t_entr: start
// the exit query from t_entr + 1e-4 (geometry.rs:508).  Phase 2 found t_exit: the rest of ConvexVolume::intersect_ray
// (geometry.rs:512-525) with the ray's own t-range, then back to the world-space traversal.
template <bool COUNT>
__device__ __forceinline__ void volume_continue(const rt_dev_scene& sc, Trav& T, float ray_t_min, float ray_t_max) {
  T.volret = false;
  const int obj = T.cur_obj;
  const uint32_t q = (uint32_t)obj * RT_OBJ_QUADS;
  Best saved;
  float t_entr;
  T.vol_load(saved, t_entr);
  if (T.vol_phase == 1 && T.best.obj >= 0) {
    t_entr = T.best.t;
    T.vol_save(saved, t_entr);  // the parked closest hit stays parked; t_entr joins it
    uint32_t root = fbits(ldq(sc.objects, q + 7).x);
    T.push(RT_ENTRY_VOLRET);
    T.vol_phase = 2;
    T.t_min = t_entr + 0.0001f;
    T.t_max = CUDART_MAX_NORMAL_F;
    T.best.t = CUDART_MAX_NORMAL_F;
    T.best.obj = -1;
    T.entry = root;
    return;
  }
  // either no entry hit, or the exit query is done
  bool have_exit = T.vol_phase == 2 && T.best.obj >= 0;
  float t_exit = T.best.t;
  T.best = saved;
  T.t_min = ray_t_min;
  T.t_max = ray_t_max;
  T.vol_phase = 0;
  T.world_ray(T.o, T.d);
  T.load_world_inv();
  T.in_blas = false;
  T.entry = RT_ENTRY_NONE;
  if (have_exit && !(t_exit < ray_t_min || t_entr > ray_t_max)) {
    float4 q9 = ldq(sc.objects, q + 9);
    float t_start = fmaxf(t_entr, ray_t_min);
    float t_end = fminf(t_exit, ray_t_max);
    float dist_in = t_end - t_start;
    uint32_t vi = fbits(q9.y);
    float4 kc = T.qC[T.slot];
    uint32_t pixel = fbits(kc.y), sb = fbits(kc.z);
    u4 rr = philox4x32_10(pixel, sb & 0xFFFFFFu, sb >> 24, 1u + (vi >> 2), T.k0, T.k1);
    uint32_t w = (vi & 3u) == 0 ? rr.x : ((vi & 3u) == 1 ? rr.y : ((vi & 3u) == 2 ? rr.z : rr.w));
    float dist_before = (-1.0f / q9.x) * logf(u01(w));
    if (dist_before < dist_in) {
      float t = t_start + dist_before;
      if (better(t, obj, 0u, T.best)) {
        T.best.t = t; T.best.obj = obj; T.best.prim = 0;
      }
    }
  }
}

// one round: descend while interior, then the leaf this lane reached (if any), then pop.
// Returns true when the stack is exhausted.  (Measured on B200: postponing leaves until every lane
// of the warp holds one - "while-while" - is 30 % slower here, because leaves are cheap (1.9
// triangle tests per ray) compared with the descent they would make the other lanes wait for.)
template <bool COUNT, bool VOLMESH>
__device__ __forceinline__ bool trav_round(const rt_dev_scene& sc, Trav& T, float ray_t_min, float ray_t_max) {
  if (COUNT) T.cnt.rounds += 1;
  while (T.entry != RT_ENTRY_NONE && !(T.entry & RT_LEAF_FLAG)) trav_interior<COUNT>(sc, T);
  if (T.entry != RT_ENTRY_NONE) trav_leaf<COUNT, VOLMESH>(sc, T);
  if (T.entry == RT_ENTRY_NONE) {
    if (!T.pop_next<VOLMESH>()) return true;
    if (VOLMESH && T.volret) volume_continue<COUNT>(sc, T, ray_t_min, ray_t_max);
  }
  return false;
}

// nearest RGB8 tap, texture.rs:28-31 (Q7)
__device__ __forceinline__ f3 tex_sample(const rt_dev_scene& sc, int tex, float u, float v) {
  uint4 td = __ldg(reinterpret_cast<const uint4*>(sc.textures) + tex);
  float fx = clampf(u, 0.0f, 0.999f) * (float)td.y;
  float fy = (1.0f - clampf(v, 0.0f, 0.999f)) * (float)td.z;
  // Rust `as u32`: saturating, NaN -> 0
  uint32_t x = (fx != fx) ? 0u : __float2uint_rz(fx);
  uint32_t y = (fy != fy) ? 0u : __float2uint_rz(fy);
  x = min(x, td.y - 1u);
  y = min(y, td.z - 1u);
  uint32_t p = __ldg(reinterpret_cast<const uint32_t*>(sc.texels) + td.x + y * td.y + x);
  return mk((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f, (float)((p >> 16) & 255u) / 255.0f);
}

struct Surface {
  f3 hp, n;
  float u, v;
  uint32_t meta;  // class | frontface << 3 | id << 4   (id: material index, or object index for PARAM_TEX)
};

// what the reference attaches to a RayHit: RayHit::new (tracing.rs:121-133), the per-primitive
// normals (geometry.rs:411,449,478,520), and for meshes geometry.rs:350-363 + 274-298 + 307-309
template <bool COUNT>
__device__ __forceinline__ void resolve_hit(const rt_dev_scene& sc, f3 wo, f3 wd, const Best& b, Surface& s,
                                            unsigned long long* counters) {
  uint32_t q = (uint32_t)b.obj * RT_OBJ_QUADS;
  float4 h = ldq(sc.objects, q);
  int kind = (int)fbits(h.x);
  int mat = (int)fbits(h.y);
  uint32_t cls = fbits(h.z);
  s.u = s.v = 0.0f;
  bool front;
  if (kind == RT_OBJ_MESH) {
    float4 r0 = ldq(sc.objects, q + 1), r1 = ldq(sc.objects, q + 2), r2 = ldq(sc.objects, q + 3);
    float4 m0 = ldq(sc.objects, q + 4), m1 = ldq(sc.objects, q + 5), m2 = ldq(sc.objects, q + 6);
    float4 m7 = ldq(sc.objects, q + 7), m8 = ldq(sc.objects, q + 8);
    f3 o = mk(r0.x * wo.x + r0.y * wo.y + r0.z * wo.z + r0.w * 1.0f, r1.x * wo.x + r1.y * wo.y + r1.z * wo.z + r1.w * 1.0f,
              r2.x * wo.x + r2.y * wo.y + r2.z * wo.z + r2.w * 1.0f);
    f3 d = mk(r0.x * wd.x + r0.y * wd.y + r0.z * wd.z + r0.w * 0.0f, r1.x * wd.x + r1.y * wd.y + r1.z * wd.z + r1.w * 0.0f,
              r2.x * wd.x + r2.y * wd.y + r2.z * wd.z + r2.w * 0.0f);
    uint32_t sq = (fbits(m7.y) + b.prim) * RT_SHADE_QUADS;
    float4 s0 = ldq(sc.shade, sq), s1 = ldq(sc.shade, sq + 1), s2 = ldq(sc.shade, sq + 2), s3 = ldq(sc.shade, sq + 3),
           s4 = ldq(sc.shade, sq + 4);
    if (COUNT) atomicAdd(&counters[4], 1ull);
    f3 na = mk(s0.x, s0.y, s0.z), nb = mk(s0.w, s1.x, s1.y), nc = mk(s1.z, s1.w, s2.x);
    float tau = s2.y, tav = s2.z, tbu = s2.w, tbv = s3.x, tcu = s3.y, tcv = s3.z;
    f3 tan = mk(s3.w, s4.x, s4.y);
    float u = b.u, v = b.v, w = 1.0f - u - v;
    f3 mesh_normal = normalize(u * nb + v * nc + w * na);
    front = dot(mesh_normal, d) < 0.0f;
    f3 n = front ? mesh_normal : -mesh_normal;
    f3 hp_obj = o + d * b.t;
    s.u = u * tbu + v * tcu + w * tau;
    s.v = u * tbv + v * tcv + w * tav;
    int tex_normal = (int)fbits(m8.z);
    if (tex_normal >= 0) {
      f3 bitangent = normalize(cross(n, tan));
      f3 tangent = normalize(cross(bitangent, n));
      f3 smp = tex_sample(sc, tex_normal, s.u, s.v);
      if (COUNT) atomicAdd(&counters[8], 1ull);
      f3 nm = 2.0f * smp - mk(1.0f, 1.0f, 1.0f);
      n = tangent * nm.x + bitangent * nm.y + n * nm.z;
    }
    // inv_transform.transpose().transform_vector(n).normalize(): dot with the columns of inv
    f3 wn = mk(r0.x * n.x + r1.x * n.y + r2.x * n.z + 0.0f * 0.0f, r0.y * n.x + r1.y * n.y + r2.y * n.z + 0.0f * 0.0f,
               r0.z * n.x + r1.z * n.y + r2.z * n.z + 0.0f * 0.0f);
    s.n = normalize(wn);
    s.hp = mk(m0.x * hp_obj.x + m0.y * hp_obj.y + m0.z * hp_obj.z + m0.w * 1.0f,
              m1.x * hp_obj.x + m1.y * hp_obj.y + m1.z * hp_obj.z + m1.w * 1.0f,
              m2.x * hp_obj.x + m2.y * hp_obj.y + m2.z * hp_obj.z + m2.w * 1.0f);
    uint32_t id = mat >= 0 ? (uint32_t)mat : (uint32_t)b.obj;
    s.meta = cls | ((front ? 1u : 0u) << 3) | (id << 4);
    return;
  }
  float4 q1 = ldq(sc.objects, q + 1);
  f3 nrm;
  s.hp = wo + wd * b.t;
  if (kind == RT_OBJ_SPHERE) {
    f3 hitpoint = wo + b.t * wd;
    nrm = normalize(hitpoint - mk(q1.x, q1.y, q1.z));
  } else if (kind == RT_OBJ_TRIANGLE) {
    float4 q3 = ldq(sc.objects, q + 3);
    nrm = mk(q3.y, q3.z, q3.w);
  } else if (kind == RT_OBJ_PLANE) {
    float4 q2 = ldq(sc.objects, q + 2);
    f3 pn = mk(q2.x, q2.y, q2.z);
    float od = dot(wo - mk(q1.x, q1.y, q1.z), pn);
    float sg = (od != od) ? od : (signbit(od) ? -1.0f : 1.0f);
    nrm = sg * pn;
  } else {  // volume: zero normal, frontface false (geometry.rs:520)
    nrm = mk(0.0f, 0.0f, 0.0f);
  }
  front = dot(nrm, wd) < 0.0f;
  s.n = front ? nrm : -nrm;
  s.meta = cls | ((front ? 1u : 0u) << 3) | ((uint32_t)mat << 4);
}

// ------------------------------------------------------------------ k_advance
__global__ void k_advance(rt_ctrl* c, uint32_t capacity) {
  uint32_t n_cont = c->n_next;
  unsigned long long remaining = c->total - c->cursor;
  unsigned long long room = capacity - n_cont;
  uint32_t n_new = (uint32_t)(remaining < room ? remaining : room);
  c->n_cont = n_cont;
  c->n_rays = n_cont + n_new;
  c->work_base = c->cursor;
  c->cursor += n_new;
  c->n_next = 0;
  c->next_ray = 0;
#pragma unroll
  for (int i = 0; i < RT_NUM_CLASSES; ++i) c->class_count[i] = 0;
  if (n_cont + n_new == 0) c->done = 1;
  else c->iterations += 1;
  c->n_rays_total += n_cont + n_new;
  c->n_samples += n_new;
}

// ------------------------------------------------------------------ k_raygen
// camera rays for the paths started this iteration: slots [n_cont, n_rays) of the ray queue
__global__ void __launch_bounds__(RT_BLOCK) k_raygen(rt_frame fr, rt_ctrl* __restrict__ ctrl, rt_paths cur) {
  const uint32_t n_rays = ctrl->n_rays, n_cont = ctrl->n_cont;
  const uint32_t k = blockIdx.x * RT_BLOCK + threadIdx.x;
  if (k >= n_rays - n_cont) return;
  const uint32_t i = n_cont + k;
  uint32_t x, y, sample;
  if (work_to_pixel(fr, ctrl->work_base + k, x, y, sample)) {
    uint32_t pixel = y * fr.width + x;
    f3 o, d;
    camera_ray(fr, x, y, pixel, sample, o, d);
    cur.A[i] = make_float4(o.x, o.y, o.z, d.x);
    cur.B[i] = make_float4(d.y, d.z, 1.0f, 1.0f);
    cur.C[i] = make_float4(1.0f, __uint_as_float(pixel), __uint_as_float(sample), 0.0f);  // bounce 0
  } else {
    // tile slot outside the image: a direction of all zeros marks "no ray" for k_trace
    atomicAdd(&ctrl->counters[7], 1ull);
    cur.A[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    cur.B[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    cur.C[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ------------------------------------------------------------------ k_trace
// Persistent warps that claim ray indices in chunks (one atomic per RT_FETCH_CHUNK rays) and do
// nothing but traverse: fetching a ray is two 128-bit loads, retiring one is a 20-byte store.
// A warp can refill idle lanes before the whole batch is done (RT_REFILL_MIN idle lanes trigger a
// refill).  Measured on C4 / B200 (profiles/r1_notes.md): although only 21 % of the lanes of a
// batch are busy in node-visit terms, refilling early is SLOWER (R=4: 535 us, R=8: 509, R=16: 496,
// R=32: 477 per iteration) - a fresh batch of 32 camera rays of one pixel traverses in lockstep,
// and mixing in rays from elsewhere destroys that; the idle tail of a batch is cheap because the
// few lanes left no longer wait for each other.  Hence the default of 32 = refill only when the
// whole warp is idle.
#ifndef RT_REFILL_MIN
#define RT_REFILL_MIN 32
#endif
#ifndef RT_FETCH_CHUNK
#define RT_FETCH_CHUNK 32  // measured: 32 -> 436 us, 64 -> 477 us, 256 -> 703 us per 2 M-ray iteration (load balance)
#endif

template <bool COUNT, bool VOLMESH>
__global__ void __launch_bounds__(RT_BLOCK, RT_EXTEND_MIN_BLOCKS) k_trace(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                                        rt_paths cur, rt_hits hits, const uint32_t* __restrict__ order) {
  __shared__ __align__(8) uint32_t sstack[(RT_SMEM_STACK + 6 + (VOLMESH ? 6 : 0)) * RT_BLOCK];
  const uint32_t n_rays = ctrl->n_rays;
  const uint32_t n_sorted = fr.sort_enabled ? ctrl->n_cont : 0u;  // continuing rays are visited in sorted order
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t lstack[RT_LOCAL_STACK];
  Trav T;
  T.sbase = (uint32_t)__cvta_generic_to_shared(sstack) + threadIdx.x * 4u;
  T.wbase = (uint32_t)__cvta_generic_to_shared(sstack) + RT_SMEM_STACK * (RT_BLOCK * 4u) + threadIdx.x * 4u;
  T.lstack = lstack;
  T.t_min = fr.t_min; T.t_max = fr.t_max;
  T.k0 = fr.k0; T.k1 = fr.k1;
  T.qA = cur.A; T.qB = cur.B; T.qC = cur.C;
  T.slot = 0;
  bool have = false, fin = false;
  uint32_t c_next = 0, c_end = 0;  // this warp's claimed chunk of ray indices (warp uniform)
  bool exhausted = false;          // the queue has no more chunks (warp uniform)

  for (;;) {
    // ---- retire finished lanes: the compact hit record is all k_shade needs to redo the rest
    if (have && fin) {
      const Best& best = T.best;
      RT_STS(&hits.H[T.slot], make_float4(best.t, best.u, best.v, __uint_as_float(best.prim)));
      RT_STS(&hits.obj[T.slot], best.obj);
      if (COUNT) {
        atomicAdd(&ctrl->counters[0], (unsigned long long)T.cnt.nodes);
        atomicAdd(&ctrl->counters[1], (unsigned long long)T.cnt.tris);
        atomicAdd(&ctrl->counters[2], (unsigned long long)T.cnt.inst);
        atomicAdd(&ctrl->counters[3], (unsigned long long)T.cnt.prims);
        atomicAdd(&ctrl->counters[10], (unsigned long long)T.cnt.tlas_nodes);
        // SIMT diagnostic (meaningful with RT_REFILL_MIN == 32, i.e. whole batches retire together): nodes fetched
        // by the lanes of the batch vs 32 x what its slowest lane fetched
        uint32_t m = __activemask();
        uint32_t wmax = __reduce_max_sync(m, T.cnt.nodes);
        if ((m & ((1u << lane) - 1u)) == 0) atomicAdd(&ctrl->counters[9], (unsigned long long)wmax * 32ull);
      }
      have = false;
      fin = false;
    }
    // ---- refill idle lanes
    uint32_t need = __ballot_sync(FULL, !have);
    if ((need == FULL || __popc(need) >= RT_REFILL_MIN) && !(exhausted && c_next >= c_end)) {
      uint32_t n_need = __popc(need);
      if (c_next >= c_end && !exhausted) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctrl->next_ray, (uint32_t)RT_FETCH_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        c_next = base;
        c_end = min(base + (uint32_t)RT_FETCH_CHUNK, n_rays);
        if (base + RT_FETCH_CHUNK >= n_rays) exhausted = true;
        if (base >= n_rays) c_end = c_next = 0;
      }
      uint32_t my = c_next + __popc(need & ((1u << lane) - 1u));
      if (!have && my < c_end) {
        if (my < n_sorted) my = __ldg(order + my);  // queue position -> slot
        float4 a = RT_LDS(&cur.A[my]), b = RT_LDS(&cur.B[my]);
        f3 wo = mk(a.x, a.y, a.z), wd = mk(a.w, b.x, b.y);
        if (wd.x == 0.0f && wd.y == 0.0f && wd.z == 0.0f && b.z == 0.0f) {
          hits.obj[my] = -1;  // "no ray" marker written by k_raygen
        } else {
          T.slot = my;
          T.t_max = T.ray_t_max = fr.ray_tmax_from_c ? RT_LDS(&cur.C[my]).w : fr.t_max;  // Phong shadow rays end at the light
          trav_begin<COUNT>(sc, T, wo, wd);
          have = true;
          fin = false;
        }
      }
      c_next = min(c_next + n_need, c_end);
    }
    uint32_t busy = __ballot_sync(FULL, have);
    if (busy == 0) {
      if (exhausted && c_next >= c_end) break;
      continue;
    }
    // ---- traverse until enough lanes are idle again (or nothing is left to fetch)
    const bool can_refill = !(exhausted && c_next >= c_end);
    for (;;) {
      if (have && !fin) fin = trav_round<COUNT, VOLMESH>(sc, T, fr.t_min, T.ray_t_max);
      uint32_t run = __ballot_sync(FULL, have && !fin);
      if (run == 0) break;
      if (can_refill && 32 - __popc(run) >= RT_REFILL_MIN) break;
    }
  }
}

// ------------------------------------------------------------------ k_sort
// material-sorted shade queues: warp match groups -> per-warp counts -> one atomic per class per
// block -> each hit ray's slot index lands in the queue of its material class
#ifndef RT_SORT_BLOCK
#define RT_SORT_BLOCK 512
#endif
#define RT_SORT_WARPS (RT_SORT_BLOCK / 32)
__global__ void __launch_bounds__(RT_SORT_BLOCK) k_sort(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                   const int32_t* __restrict__ hit_obj, uint32_t* __restrict__ queues) {
  __shared__ uint32_t s_wcount[RT_SORT_WARPS][RT_NUM_CLASSES];
  __shared__ uint32_t s_base[RT_NUM_CLASSES];
  const uint32_t n_rays = ctrl->n_rays;
  const uint32_t i = blockIdx.x * RT_SORT_BLOCK + threadIdx.x;
  if (blockIdx.x * RT_SORT_BLOCK >= n_rays) return;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  int cls = -1;
  if (i < n_rays) {
    int obj = hit_obj[i];
    if (obj >= 0) cls = (int)fbits(ldq(sc.objects, (uint32_t)obj * RT_OBJ_QUADS).z);
  }
  if (threadIdx.x < RT_SORT_WARPS * RT_NUM_CLASSES) (&s_wcount[0][0])[threadIdx.x] = 0;
  static_assert(RT_SORT_WARPS * RT_NUM_CLASSES <= RT_SORT_BLOCK, "zero-fill covers the counters");
  __syncthreads();
  uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
  uint32_t rank = __popc(peers & ((1u << lane) - 1u));
  if (cls >= 0 && rank == 0) s_wcount[warp][cls] = __popc(peers);
  __syncthreads();
  if (threadIdx.x < RT_NUM_CLASSES) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < RT_SORT_WARPS; ++w) {
      uint32_t c = s_wcount[w][threadIdx.x];
      s_wcount[w][threadIdx.x] = total;  // exclusive prefix over warps
      total += c;
    }
    s_base[threadIdx.x] = total ? atomicAdd(&ctrl->class_count[threadIdx.x], total) : 0u;
  }
  __syncthreads();
  if (cls >= 0) queues[(size_t)cls * fr.capacity + s_base[cls] + s_wcount[warp][cls] + rank] = i;
}

// ------------------------------------------------------------------ k_surface (parity hooks only)
__global__ void __launch_bounds__(RT_BLOCK) k_surface(rt_dev_scene sc, rt_ctrl* __restrict__ ctrl, rt_paths cur, rt_hits hits,
                                                      rt_debug dbg) {
  const uint32_t i = blockIdx.x * RT_BLOCK + threadIdx.x;
  if (i >= ctrl->n_rays) return;
  float4 h = hits.H[i];
  Best b;
  b.t = h.x; b.u = h.y; b.v = h.z; b.prim = fbits(h.w);
  b.obj = hits.obj[i];
  Surface s;
  s.hp = s.n = mk(0.f, 0.f, 0.f);
  s.u = s.v = 0.0f;
  s.meta = 0xFFFFFFFFu;
  if (b.obj >= 0) {
    float4 a = cur.A[i], bq = cur.B[i];
    resolve_hit<false>(sc, mk(a.x, a.y, a.z), mk(a.w, bq.x, bq.y), b, s, ctrl->counters);
  }
  dbg.S0[i] = make_float4(s.hp.x, s.hp.y, s.hp.z, s.n.x);
  dbg.S1[i] = make_float4(s.n.y, s.n.z, s.u, s.v);
  dbg.S2[i] = s.meta;
}

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
    h = (h ^ (h >> 13)) & 0x1FFFu;
    return (h << 5) | (dom << 3) | oct;
  }
  h = (h ^ (h >> 15)) & 0x7FFFu;
  return (h << 3) | (fr.sort_use_octant ? oct : 0u);
}

// fixed-point accumulation (2^-30 units): order-independent, hence reproducible and shardable
#define RT_FIX_SCALE 1073741824.0f
#define RT_FIX_CLAMP 65536.0f
__device__ __forceinline__ void accum_add(long long* accum, uint32_t pixel, f3 c) {
  float v[3] = {c.x, c.y, c.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = v[k];
    if (x != x) {
      atomicAdd(reinterpret_cast<unsigned long long*>(accum + (size_t)pixel * 4 + 3), 1ull << (21 * k));
    } else if (x != 0.0f) {
      x = fminf(fmaxf(x, -RT_FIX_CLAMP), RT_FIX_CLAMP);
      long long q = __float2ll_rn(x * RT_FIX_SCALE);
      atomicAdd(reinterpret_cast<unsigned long long*>(accum + (size_t)pixel * 4 + k), (unsigned long long)q);
    }
  }
}

// ------------------------------------------------------------------ k_shade
// k_shade is latency bound on its gathers: 10 resident blocks (48 registers, ~100 B of spills) beat 7 blocks
// (72 registers, no spills) by 1 % on C4 and 7-10 % on the closed scenes C1 / C3 where shading dominates
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS 10
#endif
// MULTI (Camera::path_samples > 1): the kernel runs once per child index fr.branch over the same hits; a child's
// random numbers are keyed by its position in the sample's path tree (carried in C.w), its throughput is divided by
// path_samples (tracing.rs:319), and the hit's emission is added by the first pass only.
template <bool COUNT, bool MULTI>
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MIN_BLOCKS) k_shade(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                    rt_paths cur, rt_paths nxt, rt_hits hits,
                                                    const uint32_t* __restrict__ queues, long long* __restrict__ accum,
                                                    rt_sortbuf sort) {
#if RT_OCTANT_SORT
  __shared__ uint32_t s_ocount[8][RT_WARPS];
#else
  __shared__ uint32_t s_wcount[RT_WARPS];
#endif
  __shared__ uint32_t s_base;
  // which class does this block serve?  class segments are padded to whole blocks
  uint32_t b = blockIdx.x;
  int cls = -1;
  uint32_t count = 0;
  // Blocks serve the classes in the order below, and survivors get their place in the next ray queue in block
  // order.  Rays leaving a textured mesh start inside a BLAS and are the most expensive to trace, rays leaving
  // analytic surfaces are cheap, new camera rays (appended behind all of these) are cheapest: longest first, so
  // k_trace's tail is filled with short, coherent batches.
  const int order[RT_NUM_CLASSES] = {RT_CLASS_PARAM_TEX, RT_CLASS_DIELECTRIC, RT_CLASS_METAL, RT_CLASS_LAMBERT,
                                     RT_CLASS_ISOTROPIC, RT_CLASS_PARAM};
#pragma unroll
  for (int k = 0; k < RT_NUM_CLASSES; ++k) {
    const int c = order[k];
    uint32_t n = ctrl->class_count[c];
    uint32_t nb = (n + RT_BLOCK - 1) / RT_BLOCK;
    if (cls < 0) {
      if (b < nb) {
        cls = c;
        count = n;
      } else {
        b -= nb;
      }
    }
  }
  if (cls < 0) return;
  const uint32_t j = b * RT_BLOCK + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  bool alive = false;
  f3 no, nd, nT;
  uint32_t pixel = 0, sb = 0, tree = 0;
  if (j < count) {
    uint32_t slot = RT_LDS(&queues[(size_t)cls * fr.capacity + j]);
    float4 a = RT_LDS(&cur.A[slot]), bq = RT_LDS(&cur.B[slot]), c = RT_LDS(&cur.C[slot]);
    f3 o = mk(a.x, a.y, a.z);
    f3 d = mk(a.w, bq.x, bq.y);
    f3 T = mk(bq.z, bq.w, c.x);
    pixel = fbits(c.y);
    sb = fbits(c.z);
    if (MULTI) tree = fbits(c.w) * fr.path_samples + fr.branch;
    uint32_t sample = sb & 0xFFFFFFu, bounce = sb >> 24;
    // hit resolution: what the reference attaches to its RayHit
    float4 hr = RT_LDS(&hits.H[slot]);
    Best best;
    best.t = hr.x; best.u = hr.y; best.v = hr.z; best.prim = fbits(hr.w);
    best.obj = RT_LDS(&hits.obj[slot]);
    Surface sf;
    resolve_hit<COUNT>(sc, o, d, best, sf, ctrl->counters);
    f3 hp = sf.hp, n = sf.n;
    float4 h1 = make_float4(0.f, 0.f, sf.u, sf.v);
    uint32_t meta = sf.meta;
    bool front = (meta >> 3) & 1u;
    uint32_t id = meta >> 4;

    // material parameters
    f3 albedo, emission;
    float roughness, metallic;
    if (cls == RT_CLASS_PARAM_TEX) {
      // StaticMesh::get_material_at_uv, geometry.rs:259-269 (Q7 defaults)
      uint32_t q = id * RT_OBJ_QUADS;
      float4 m7 = ldq(sc.objects, q + 7), m8 = ldq(sc.objects, q + 8);
      int ta = (int)fbits(m7.z), te = (int)fbits(m7.w), tm = (int)fbits(m8.x), tr = (int)fbits(m8.y);
      float u = h1.z, v = h1.w;
      albedo = ta >= 0 ? tex_sample(sc, ta, u, v) : mk(0.f, 0.f, 0.f);
      emission = te >= 0 ? tex_sample(sc, te, u, v) : mk(0.f, 0.f, 0.f);
      metallic = tm >= 0 ? tex_sample(sc, tm, u, v).x : 0.0f;
      roughness = tr >= 0 ? tex_sample(sc, tr, u, v).x : 1.0f;
      if (COUNT) atomicAdd(&ctrl->counters[5], (unsigned long long)((ta >= 0) + (te >= 0) + (tm >= 0) + (tr >= 0)));
    } else {
      float4 m0 = ldq(sc.mats, id * RT_MAT_QUADS), m1 = ldq(sc.mats, id * RT_MAT_QUADS + 1);
      if (COUNT) atomicAdd(&ctrl->counters[6], 1ull);
      albedo = mk(m0.x, m0.y, m0.z);
      emission = mk(m1.x, m1.y, m1.z);
      roughness = m0.w;
      metallic = m1.w;
    }

    // emitted light reaches the pixel attenuated by the path throughput (tracing.rs:321)
    if ((!MULTI || fr.branch == 0u) && (emission.x != 0.0f || emission.y != 0.0f || emission.z != 0.0f))
      accum_add(accum, pixel, mulv(T, emission));

    // scatter (materials.rs:33-166)
    u4 r = philox4x32_10(pixel, sample, bounce, 0u, fr.k0, MULTI ? fr.k1 ^ (tree * 0x9E3779B9u) : fr.k1);
    float u_choice = u01(r.x);
    f3 ball = ball_from(r.y, r.z, r.w);
    f3 dir, brdf;
    float pdf;
    const float PI = RT_PI;
    if (cls == RT_CLASS_LAMBERT) {
      dir = sample_hemisphere(n, ball);
      brdf = albedo / PI;
      pdf = 1.0f / (2.0f * PI);
    } else if (cls == RT_CLASS_METAL) {
      dir = reflectv(d, n) + roughness * ball;
      brdf = albedo;
      pdf = 1.0f;
    } else if (cls == RT_CLASS_DIELECTRIC) {
      float ior = roughness;  // stored in the roughness slot
      float eta = front ? 1.0f / ior : ior;
      float cth = fminf(-dot(d, n), 1.0f);
      bool critical = eta * sqrtf(1.0f - cth * cth) > 1.0f;
      float fres = fresnelf(d, n, ior);
      bool will_refract = !critical && u_choice >= fres;
      dir = will_refract ? refractv(d, n, eta) : reflectv(d, n);
      brdf = mk(1.0f, 1.0f, 1.0f);
      pdf = 1.0f;
    } else if (cls == RT_CLASS_ISOTROPIC) {
      dir = ball;
      brdf = albedo;
      pdf = 1.0f;
    } else {  // PARAM / PARAM_TEX, materials.rs:114-145
      float fres = fresnelf(d, n, 1.5f);
      float k_s = fres * (1.0f - roughness);
      float k_d = (1.0f - k_s) * (1.0f - metallic);
      if (u_choice < k_d) {
        dir = sample_hemisphere(n, ball);
        brdf = albedo / PI;
        pdf = 1.0f / (2.0f * PI);
      } else {
        dir = reflectv(d, n) + roughness * ball;
        brdf = (1.0f - metallic) * mk(1.0f, 1.0f, 1.0f) + metallic * albedo;  // lerpvec(1, albedo, metallic)
        pdf = 1.0f;
      }
    }
    // tracing.rs:313-316
    float dot_term = mag2(n) > 0.0f ? clampf(fabsf(dot(dir, n)), 0.0f, 1.0f) : 1.0f;
    f3 w = (dot_term * brdf) / pdf;
    nT = mulv(T, w);
    if (MULTI) nT = nT / (float)fr.path_samples;
    no = hp;
    nd = dir;
    bounce += 1;
    sb = sample | (bounce << 24);
    // the next segment exists only below path_depth (tracing.rs:301); a path whose throughput is
    // exactly zero can add nothing any more
    alive = bounce < fr.path_depth && !(nT.x == 0.0f && nT.y == 0.0f && nT.z == 0.0f);
  }

  // compact survivors into the next ray queue, one atomic per block.  Inside the block's output range the rays are
  // grouped by the octant of their new direction (RT_OCTANT_SORT): rays that agree on the direction signs visit
  // the children of a node in the same order, which keeps more lanes of a k_trace batch together.
#if RT_OCTANT_SORT
  const uint32_t oct = alive ? ((nd.x < 0.0f ? 1u : 0u) | (nd.y < 0.0f ? 2u : 0u) | (nd.z < 0.0f ? 4u : 0u)) : 8u;
  uint32_t my_bal = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; ++k) {
    uint32_t bk = __ballot_sync(0xFFFFFFFFu, oct == k);
    if (lane == 0) s_ocount[k][warp] = __popc(bk);
    if (oct == k) my_bal = bk;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int w = 0; w < RT_WARPS; ++w) {
        uint32_t c = s_ocount[k][w];
        s_ocount[k][w] = total;
        total += c;
      }
    s_base = total ? atomicAdd(&ctrl->n_next, total) : 0u;
  }
  __syncthreads();
  if (alive) {
    uint32_t pos = s_base + s_ocount[oct][warp] + __popc(my_bal & ((1u << lane) - 1u));
#else
  uint32_t bal = __ballot_sync(0xFFFFFFFFu, alive);
  if (lane == 0) s_wcount[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < RT_WARPS; ++w) {
      uint32_t c = s_wcount[w];
      s_wcount[w] = total;
      total += c;
    }
    s_base = total ? atomicAdd(&ctrl->n_next, total) : 0u;
  }
  __syncthreads();
  if (alive) {
    uint32_t pos = s_base + s_wcount[warp] + __popc(bal & ((1u << lane) - 1u));
#endif
    RT_STS(&nxt.A[pos], make_float4(no.x, no.y, no.z, nd.x));
    RT_STS(&nxt.B[pos], make_float4(nd.y, nd.z, nT.x, nT.y));
    RT_STS(&nxt.C[pos], make_float4(nT.z, __uint_as_float(pixel), __uint_as_float(sb), __uint_as_float(tree)));
    if (fr.sort_enabled) {
      // one atomic per distinct key per warp: the rays of a warp mostly leave the same few cells
      uint32_t key = ray_sort_key(fr, no, nd);
      sort.keys[pos] = key;
      uint32_t peers = __match_any_sync(__activemask(), key);
      if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&sort.hist[key], (uint32_t)__popc(peers));
    }
  }
}

// ------------------------------------------------------------------ k_raysort_scan / k_raysort_scatter
// counting sort of the next ray queue by ray_sort_key: k_shade has filled the histogram.  k_raysort_scan: one block
// per 1024 bins turns its slice into exclusive offsets within the slice (coalesced), clears the histogram for the next
// iteration and records the slice total.  k_raysort_scatter: every block first prefixes the slice totals (at most
// RT_SORT_BINS/1024 of them), then every ray claims a place in its bin - one atomic per distinct key per warp.
__global__ void __launch_bounds__(1024) k_raysort_scan(rt_sortbuf sort) {
  __shared__ uint32_t s_warp[32];
  const uint32_t bin = blockIdx.x * 1024u + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t v = sort.hist[bin];
  sort.hist[bin] = 0;
  uint32_t inc = v;  // inclusive scan inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
    if (lane >= (uint32_t)off) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], winc = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, off);
      if (lane >= (uint32_t)off) winc += t;
    }
    s_warp[lane] = winc - w;  // exclusive prefix of the warps
    if (lane == 31) sort.slice_total[blockIdx.x] = winc;
  }
  __syncthreads();
  sort.cursor[bin] = s_warp[warp] + inc - v;  // exclusive offset inside this slice
}
__global__ void __launch_bounds__(256) k_raysort_scatter(rt_ctrl* __restrict__ ctrl, rt_sortbuf sort) {
  __shared__ uint32_t s_slice[RT_SORT_BINS / 1024u];
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (blockIdx.x * 256u >= ctrl->n_next) return;
  const uint32_t lane = threadIdx.x & 31u;
  if (threadIdx.x < 32) {  // exclusive prefix of the slice totals (RT_SORT_BINS / 1024 <= 32 * per-lane count)
    const uint32_t per = (RT_SORT_BINS / 1024u + 31u) / 32u;
    uint32_t sum = 0, loc[per];
#pragma unroll
    for (uint32_t k = 0; k < per; ++k) {
      uint32_t idx = lane * per + k;
      loc[k] = idx < RT_SORT_BINS / 1024u ? sort.slice_total[idx] : 0u;
      sum += loc[k];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
      if (lane >= (uint32_t)off) inc += t;
    }
    uint32_t base = inc - sum;
#pragma unroll
    for (uint32_t k = 0; k < per; ++k) {
      uint32_t idx = lane * per + k;
      if (idx < RT_SORT_BINS / 1024u) s_slice[idx] = base;
      base += loc[k];
    }
  }
  __syncthreads();
  if (i >= ctrl->n_next) return;
  // neighbouring queue positions come from the same k_shade block and share keys: one atomic per distinct key per warp
  uint32_t key = sort.keys[i];
  uint32_t peers = __match_any_sync(__activemask(), key);
  uint32_t leader = __ffs(peers) - 1u;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&sort.cursor[key], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  sort.order[s_slice[key >> 10] + base + __popc(peers & ((1u << lane) - 1u))] = i;
}

// ------------------------------------------------------------------ k_phong_primary / k_phong_shadow
// ShadingMode::Phong (Scene::phong_shade_ray, tracing.rs:277-297), the reference's debug shading: one camera ray,
// one shadow ray towards point_light_pos, no recursion.  Both queries use t_min = 0.
//   k_phong_primary: hit frame, ambient + diffuse * brdf + specular * 0.4 (parked in the T slot of the shadow ray),
//                    shadow ray from hitpoint + 0.01 n towards the light, limited to the distance to the light (C.w)
//   k_phong_shadow : shadow weight (tracing.rs:290-293: 0.3 unless the occluder lies in the far half of the segment)
//                    and accumulation
__global__ void __launch_bounds__(RT_BLOCK) k_phong_primary(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                            rt_paths cur, rt_paths nxt, rt_hits hits) {
  const uint32_t n_rays = ctrl->n_rays;
  const uint32_t i = blockIdx.x * RT_BLOCK + threadIdx.x;
  if (i == 0) ctrl->n_next = n_rays;  // no compaction: shadow ray i belongs to camera ray i
  if (i >= n_rays) return;
  int obj = hits.obj[i];
  {
    uint32_t miss = __ballot_sync(__activemask(), obj < 0);
    if (miss && (threadIdx.x & 31u) == (uint32_t)(__ffs(__activemask()) - 1)) atomicAdd(&ctrl->counters[11], (unsigned long long)__popc(miss));
  }
  if (obj < 0) {  // background_color: black; also the "no ray" marker for k_trace
    nxt.A[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    nxt.B[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    nxt.C[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 a = cur.A[i], bq = cur.B[i], c = cur.C[i];
  f3 o = mk(a.x, a.y, a.z), d = mk(a.w, bq.x, bq.y);
  uint32_t pixel = fbits(c.y), sb = fbits(c.z);
  float4 hr = hits.H[i];
  Best best;
  best.t = hr.x; best.u = hr.y; best.v = hr.z; best.prim = fbits(hr.w);
  best.obj = obj;
  Surface sf;
  resolve_hit<false>(sc, o, d, best, sf, ctrl->counters);
  const uint32_t cls = sf.meta & 7u, id = sf.meta >> 4;
  f3 albedo;
  float roughness, metallic;
  if (cls == RT_CLASS_PARAM_TEX) {
    uint32_t q = id * RT_OBJ_QUADS;
    float4 m7 = ldq(sc.objects, q + 7), m8 = ldq(sc.objects, q + 8);
    int ta = (int)fbits(m7.z), tm = (int)fbits(m8.x), tr = (int)fbits(m8.y);
    albedo = ta >= 0 ? tex_sample(sc, ta, sf.u, sf.v) : mk(0.f, 0.f, 0.f);
    metallic = tm >= 0 ? tex_sample(sc, tm, sf.u, sf.v).x : 0.0f;
    roughness = tr >= 0 ? tex_sample(sc, tr, sf.u, sf.v).x : 1.0f;
  } else {
    float4 m0 = ldq(sc.mats, id * RT_MAT_QUADS), m1 = ldq(sc.mats, id * RT_MAT_QUADS + 1);
    albedo = mk(m0.x, m0.y, m0.z);
    roughness = m0.w;
    metallic = m1.w;
  }
  // hit.material.scatter(&hit, ray).1 : the attenuation term of each material (materials.rs:33-166)
  f3 brdf;
  if (cls == RT_CLASS_LAMBERT) brdf = albedo / RT_PI;
  else if (cls == RT_CLASS_DIELECTRIC) brdf = mk(1.0f, 1.0f, 1.0f);
  else if (cls == RT_CLASS_METAL || cls == RT_CLASS_ISOTROPIC) brdf = albedo;
  else {
    u4 r = philox4x32_10(pixel, sb & 0xFFFFFFu, 0u, 0u, fr.k0, fr.k1);
    float fres = fresnelf(d, sf.n, 1.5f);
    float k_s = fres * (1.0f - roughness);
    float k_d = (1.0f - k_s) * (1.0f - metallic);
    brdf = u01(r.x) < k_d ? albedo / RT_PI : (1.0f - metallic) * mk(1.0f, 1.0f, 1.0f) + metallic * albedo;
  }
  const f3 light = mk(fr.light[0], fr.light[1], fr.light[2]), eye = mk(fr.eye[0], fr.eye[1], fr.eye[2]);
  const f3 n = sf.n, hp = sf.hp;
  f3 to_light = normalize(light - hp);
  f3 to_camera = normalize(eye - hp);
  f3 reflected = -to_light + 2.0f * dot(to_light, n) * n;
  float diffuse_weight = clampf(dot(n, to_light), 0.0f, 1.0f);
  float specular_weight = powf(clampf(dot(to_camera, reflected), 0.0f, 1.0f), 40.0f);
  f3 partial = mk(fr.ambient[0], fr.ambient[1], fr.ambient[2]) + diffuse_weight * brdf + specular_weight * mk(0.4f, 0.4f, 0.4f);
  f3 so = hp + 0.01f * n;
  float dist = sqrtf(mag2(light - hp));
  nxt.A[i] = make_float4(so.x, so.y, so.z, to_light.x);
  nxt.B[i] = make_float4(to_light.y, to_light.z, partial.x, partial.y);
  nxt.C[i] = make_float4(partial.z, __uint_as_float(pixel), __uint_as_float((sb & 0xFFFFFFu) | (1u << 24)), dist);
}
__global__ void __launch_bounds__(RT_BLOCK) k_phong_shadow(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl, rt_paths cur,
                                                           rt_hits hits, long long* __restrict__ accum) {
  const uint32_t i = blockIdx.x * RT_BLOCK + threadIdx.x;
  if (i >= ctrl->n_rays) return;
  float4 a = cur.A[i], bq = cur.B[i], c = cur.C[i];
  f3 o = mk(a.x, a.y, a.z), d = mk(a.w, bq.x, bq.y);
  if (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f && bq.z == 0.0f) return;  // the camera ray missed: black
  f3 partial = mk(bq.z, bq.w, c.x);
  float weight = 1.0f;
  int obj = hits.obj[i];
  if (obj >= 0) {
    float4 hr = hits.H[i];
    Best best;
    best.t = hr.x; best.u = hr.y; best.v = hr.z; best.prim = fbits(hr.w);
    best.obj = obj;
    Surface sf;
    resolve_hit<false>(sc, o, d, best, sf, ctrl->counters);
    const f3 light = mk(fr.light[0], fr.light[1], fr.light[2]);
    weight = best.t * best.t > mag2(light - sf.hp) ? 1.0f : 0.3f;
  }
  accum_add(accum, fbits(c.y), weight * partial);
}

// ------------------------------------------------------------------ k_resolve (Q12)
__global__ void k_resolve(const long long* __restrict__ accum, uint32_t npix, uint32_t spp, float gamma,
                          float* __restrict__ out_linear, uint8_t* __restrict__ out_rgb8) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const longlong2* a = reinterpret_cast<const longlong2*>(accum + (size_t)p * 4);
  longlong2 rg = a[0], bn = a[1];
  unsigned long long nan = (unsigned long long)bn.y;
  float c[3];
  long long s[3] = {rg.x, rg.y, bn.x};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float sum = (float)((double)s[k] * (1.0 / 1073741824.0));
    if ((nan >> (21 * k)) & 0x1FFFFFull) sum = CUDART_NAN_F;
    c[k] = sum / (float)spp;  // tracing.rs:241
  }
  if (out_linear) {
    out_linear[(size_t)p * 3 + 0] = c[0];
    out_linear[(size_t)p * 3 + 1] = c[1];
    out_linear[(size_t)p * 3 + 2] = c[2];
  }
  if (out_rgb8) {
    float f[3] = {c[0], c[1], c[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // tracing.rs:244-251, reads the pre-update copy
      float d = c[k] - 1.0f;
      if (d > 0.0f) {
        f[(k + 1) % 3] += d;
        f[(k + 2) % 3] += d;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = powf(clampf(f[k], 0.0f, 1.0f), 1.0f / gamma) * 255.9999f;
      out_rgb8[(size_t)p * 3 + k] = (v != v || v <= 0.0f) ? 0 : (v >= 255.0f ? 255 : (uint8_t)v);
    }
  }
}

__global__ void k_init_ctrl(rt_ctrl* c, unsigned long long begin, unsigned long long total) {
  c->cursor = begin;
  c->total = total;
  c->n_samples = 0;
  c->n_rays_total = 0;
  c->n_cont = 0;
  c->n_rays = 0;
  c->work_base = 0;
  c->n_next = 0;
  c->next_ray = 0;
  for (int i = 0; i < RT_NUM_CLASSES; ++i) c->class_count[i] = 0;
  c->done = 0;
  c->iterations = 0;
  for (int i = 0; i < 12; ++i) c->counters[i] = 0;
}
// Camera::path_samples > 1: the host walks the path tree depth first and tells the device what the next window is
__global__ void k_set_window(rt_ctrl* c, uint32_t n_cont, uint32_t n_new) {
  c->n_cont = n_cont;
  c->n_rays = n_cont + n_new;
  c->work_base = c->cursor;
  c->cursor += n_new;
  c->n_next = 0;
  c->next_ray = 0;
  for (int i = 0; i < RT_NUM_CLASSES; ++i) c->class_count[i] = 0;
  c->iterations += 1;
  c->n_rays_total += n_cont + n_new;
  c->n_samples += n_new;
}
// ------------------------------------------------------------------ launchers
void launch_set_window(rt_ctrl* ctrl, uint32_t n_cont, uint32_t n_new, cudaStream_t st) {
  k_set_window<<<1, 1, 0, st>>>(ctrl, n_cont, n_new);
}
void launch_init(rt_ctrl* ctrl, unsigned long long begin, unsigned long long end, cudaStream_t st) {
  k_init_ctrl<<<1, 1, 0, st>>>(ctrl, begin, end);
}
void launch_advance(rt_ctrl* ctrl, uint32_t capacity, cudaStream_t st) {
  k_advance<<<1, 1, 0, st>>>(ctrl, capacity);
}
int trace_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<false, false>, RT_BLOCK, 0);
  return n > 0 ? n : 1;
}
void launch_raygen(const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, cudaStream_t st) {
  k_raygen<<<(fr.capacity + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(fr, ctrl, cur);
}
void launch_trace(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_sortbuf sort,
                  bool count, uint32_t persistent_blocks, cudaStream_t st) {
  uint32_t full = (fr.capacity + RT_BLOCK - 1) / RT_BLOCK;
  uint32_t grid = full < persistent_blocks ? full : persistent_blocks;  // never more blocks than there could be rays
  if (sc.n_volume_meshes) {  // the variant that can nest boundary queries (a few more registers)
    if (count) k_trace<true, true><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, hits, sort.order);
    else k_trace<false, true><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, hits, sort.order);
  } else {
    if (count) k_trace<true, false><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, hits, sort.order);
    else k_trace<false, false><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, hits, sort.order);
  }
}
void launch_raysort(const rt_frame& fr, rt_ctrl* ctrl, rt_sortbuf sort, cudaStream_t st) {
  if (!fr.sort_enabled) return;
  k_raysort_scan<<<RT_SORT_BINS / 1024u, 1024, 0, st>>>(sort);
  k_raysort_scatter<<<(fr.capacity + 255) / 256, 256, 0, st>>>(ctrl, sort);
}
void launch_sort(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_hits hits, uint32_t* queues, cudaStream_t st) {
  k_sort<<<(fr.capacity + RT_SORT_BLOCK - 1) / RT_SORT_BLOCK, RT_SORT_BLOCK, 0, st>>>(sc, fr, ctrl, hits.obj, queues);
}
void launch_surface(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_debug dbg,
                    cudaStream_t st) {
  k_surface<<<(fr.capacity + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(sc, ctrl, cur, hits, dbg);
}
void launch_shade(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_paths nxt, rt_hits hits,
                  const uint32_t* queues, long long* accum, rt_sortbuf sort, bool count, cudaStream_t st) {
  uint32_t grid = (fr.capacity + RT_BLOCK - 1) / RT_BLOCK + RT_NUM_CLASSES;
  if (fr.path_samples > 1) k_shade<false, true><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, nxt, hits, queues, accum, sort);
  else if (count) k_shade<true, false><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, nxt, hits, queues, accum, sort);
  else k_shade<false, false><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, nxt, hits, queues, accum, sort);
}
void launch_phong_primary(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_paths nxt, rt_hits hits,
                          cudaStream_t st) {
  k_phong_primary<<<(fr.capacity + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, nxt, hits);
}
void launch_phong_shadow(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, long long* accum,
                         cudaStream_t st) {
  k_phong_shadow<<<(fr.capacity + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(sc, fr, ctrl, cur, hits, accum);
}
void launch_resolve(const long long* accum, uint32_t npix, uint32_t spp, float gamma, float* out_linear,
                    uint8_t* out_rgb8, cudaStream_t st) {
  k_resolve<<<(npix + 255) / 256, 256, 0, st>>>(accum, npix, spp, gamma, out_linear, out_rgb8);
}

}  // namespace rt
