// rt_traverse.cuh — closest hit: candidate order, slab test, traversal state, TLAS/BLAS walk, analytic objects, volumes, hit resolution
// Part of the single translation unit rt_kernels.cu (everything here is __forceinline__ device code).
#ifndef RT_TRAVERSE_CUH
#define RT_TRAVERSE_CUH

#include "rt_device_math.cuh"

namespace rt {

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
#if RT_NODE_CH
// The same test on a box stored as (centre, half extent): per axis the distance to the centre plane, m = c*inv + oi,
// and the half width of the slab along the ray, e = h*|inv|, give near = m - e and far = m + e without a min / max
// pair - three FFMAs per axis instead of two FFMAs and two FMNMXs, i.e. the work moves from the ALU pipe (the busiest
// pipe of k_trace) to the FMA pipe and six instructions per child pair disappear.  Conservative like slab(): the
// lowering rounds h up so that [c - h, c + h] contains the padded box.
template <class V>
__device__ __forceinline__ bool slab_ch(float4 c, V h, f3 inv, f3 ainv, f3 oi, float tmin, float tmax, float& tn) {
  float mx = __fmaf_rn(c.x, inv.x, oi.x), my = __fmaf_rn(c.y, inv.y, oi.y), mz = __fmaf_rn(c.z, inv.z, oi.z);
  float nx = __fmaf_rn(-h.x, ainv.x, mx), ny = __fmaf_rn(-h.y, ainv.y, my), nz = __fmaf_rn(-h.z, ainv.z, mz);
  float fx = __fmaf_rn(h.x, ainv.x, mx), fy = __fmaf_rn(h.y, ainv.y, my), fz = __fmaf_rn(h.z, ainv.z, mz);
  float a = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
  float b = fminf(fminf(fx, fy), fminf(fz, tmax));
  tn = a;
  return a <= b * 1.0000005f;
}
#endif
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
#if RT_NODE_CH
  f3 ainv;          // |inv|
#endif
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
#if RT_TLAS_SMEM
  uint32_t tbase;    // shared-space byte address of the block's copy of the top of the TLAS
#endif
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
#if RT_NODE_CH
    ainv = mk(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
#endif
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
#if RT_NODE_CH
    ainv = mk(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
#endif
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

#if RT_TLAS_SMEM
// Every ray starts at the top of the TLAS, so a persistent block keeps the first RT_TLAS_SMEM TLAS nodes (breadth-first
// order: the top levels) in shared memory: those fetches never miss and leave L1 to the BLAS nodes.
__device__ __forceinline__ uint32_t tlas_cached(const rt_dev_scene& sc) {
  return sc.tlas_count < (uint32_t)RT_TLAS_SMEM ? sc.tlas_count : (uint32_t)RT_TLAS_SMEM;
}
__device__ __forceinline__ void load_tlas_cache(const rt_dev_scene& sc, float4* s_tlas) {
  const uint32_t nq = tlas_cached(sc) * RT_NODE_QUADS;
  const float4* src = reinterpret_cast<const float4*>(sc.nodes) + (size_t)sc.tlas_base * RT_NODE_QUADS;
  for (uint32_t i = threadIdx.x; i < nq; i += blockDim.x) s_tlas[i] = __ldg(src + i);
  __syncthreads();
}
__device__ __forceinline__ float4 lds_quad(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
#endif

#if RT_BVH4
// interior node, 4-wide tree: fetch the 128-byte group of four child records (eight 128-bit read-only loads), test
// the four boxes, continue with the nearest child hit and push the others, farthest first.  Unused slots carry the
// link RT_ENTRY_NONE.
__device__ __forceinline__ void cswap(float& ka, uint32_t& la, float& kb, uint32_t& lb) {
  const bool sw = kb < ka;
  const float k0 = sw ? kb : ka, k1 = sw ? ka : kb;
  const uint32_t l0 = sw ? lb : la, l1 = sw ? la : lb;
  ka = k0; kb = k1; la = l0; lb = l1;
}
template <bool COUNT>
__device__ __forceinline__ void trav_interior(const rt_dev_scene& sc, Trav& T) {
  const uint32_t e = T.entry;
  float4 q[8];
#if RT_TLAS_SMEM
  const uint32_t rel = e - sc.tlas_base;
  if (rel < tlas_cached(sc)) {
    const uint32_t a = T.tbase + rel * 32u;
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = lds_quad(a + 16u * k);
  } else
#endif
  {
    const float4* grp = reinterpret_cast<const float4*>(sc.nodes) + (size_t)e * 2u;
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = __ldg(grp + k);
  }
  if (COUNT) {
    T.cnt.nodes += 4;
    if (!T.in_blas) T.cnt.tlas_nodes += 4;
  }
  float key[4];
  uint32_t link[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float tn;
    link[k] = fbits(q[2 * k].w);
    bool h = slab(q[2 * k], q[2 * k + 1], T.inv, T.oi, T.t_min, T.best.t, tn) && link[k] != RT_ENTRY_NONE;
    key[k] = h ? tn : CUDART_INF_F;
    if (!h) link[k] = RT_ENTRY_NONE;
  }
  // sorting network, ascending entry distance; misses (key = +inf, link = NONE) end up last
  cswap(key[0], link[0], key[1], link[1]);
  cswap(key[2], link[2], key[3], link[3]);
  cswap(key[0], link[0], key[2], link[2]);
  cswap(key[1], link[1], key[3], link[3]);
  cswap(key[1], link[1], key[2], link[2]);
  if (link[3] != RT_ENTRY_NONE) T.push(link[3]);
  if (link[2] != RT_ENTRY_NONE) T.push(link[2]);
  if (link[1] != RT_ENTRY_NONE) T.push(link[1]);
  T.entry = link[0];
}
#else
// interior node: fetch the 64-byte child pair with four 128-bit read-only loads, test both boxes,
// continue with the nearer child and push the other
template <bool COUNT>
__device__ __forceinline__ void trav_interior(const rt_dev_scene& sc, Trav& T) {
  uint32_t e = T.entry;
  float4 l0, l1, r0, r1;
#if RT_TLAS_SMEM
  const uint32_t rel = e - sc.tlas_base;  // BLAS nodes lie below tlas_base: rel wraps to a huge value
  if (rel < tlas_cached(sc)) {            // groups never straddle the end: the cached count is a multiple of 4
    const uint32_t a = T.tbase + rel * 32u;
    l0 = lds_quad(a); l1 = lds_quad(a + 16u); r0 = lds_quad(a + 32u); r1 = lds_quad(a + 48u);
  } else
#endif
  {
    const float4* pair = reinterpret_cast<const float4*>(sc.nodes) + (size_t)e * 2u;
    l0 = ldq_node(pair); l1 = ldq_node(pair + 1); r0 = ldq_node(pair + 2);
#if RT_NODE_CH != 2
    r1 = ldq_node(pair + 3);
#endif
  }
  if (COUNT) {
    T.cnt.nodes += 2;
    if (!T.in_blas) T.cnt.tlas_nodes += 2;
  }
#if RT_EXTRA_LDG
  // diagnostic only (profiles/r2_notes.md C9): RT_EXTRA_LDG more 128-bit fetches per visit, of the adjacent child pair
  // (the triangle array follows the nodes in the same allocation, so the last pair's neighbour is readable) - no extra arithmetic - to see whether the visit is bound by L1 data-pipe wavefronts
  {
    const float4* pair = reinterpret_cast<const float4*>(sc.nodes) + (size_t)e * 2u;
#pragma unroll
    for (int k = 0; k < RT_EXTRA_LDG; ++k) {
      float4 x;
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(pair + 4 + k));  // the NEXT pair (other addresses would be merged with the real loads)
      if (fbits(x.x) == 0x7fc12345u) T.t_min = x.y + x.z + x.w;  // never true: keeps all four words of the load alive
    }
  }
#endif
  float tl, tr;
#if RT_NODE_CH == 2
  // packed pair: l0 = (left centre, left link), l1 = (right centre, right link), r0 = the six half extents as bf16
  // (rounded up by the lowering); the fourth quad of the 64-byte slot is not fetched
  const uint32_t w0 = fbits(r0.x), w1 = fbits(r0.y), w2 = fbits(r0.z);
  const f3 hL = mk(__uint_as_float(w0 << 16), __uint_as_float(w0 & 0xFFFF0000u), __uint_as_float(w1 << 16));
  const f3 hR = mk(__uint_as_float(w1 & 0xFFFF0000u), __uint_as_float(w2 << 16), __uint_as_float(w2 & 0xFFFF0000u));
  bool hl = slab_ch(l0, hL, T.inv, T.ainv, T.oi, T.t_min, T.best.t, tl);
  bool hr = slab_ch(l1, hR, T.inv, T.ainv, T.oi, T.t_min, T.best.t, tr);
  r0.w = l1.w;  // the right child's link, where the code below expects it
#elif RT_NODE_CH
  bool hl = slab_ch(l0, l1, T.inv, T.ainv, T.oi, T.t_min, T.best.t, tl);
  bool hr = slab_ch(r0, r1, T.inv, T.ainv, T.oi, T.t_min, T.best.t, tr);
#else
  bool hl = slab(l0, l1, T.inv, T.oi, T.t_min, T.best.t, tl);
  bool hr = slab(r0, r1, T.inv, T.oi, T.t_min, T.best.t, tr);
#endif
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
#endif

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
      float4 a0 = ldq_tri(sc.tris, q), a1 = ldq_tri(sc.tris, q + 1), a2 = ldq_tri(sc.tris, q + 2);
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

// start the closest-hit query of the world-space ray (wo, wd)
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

}  // namespace rt
#endif
