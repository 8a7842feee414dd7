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

#include "rt_config.cuh"
#include "rt_device_math.cuh"
#include "rt_traverse.cuh"
#include "rt_materials.cuh"
#include "rt_shade.cuh"

namespace rt {

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
  c->poll[0] = c->done;
  c->poll[1] = c->cursor == c->total ? 1u : 0u;
  c->poll[2] = n_cont + n_new;
  c->poll[3] = c->iterations;
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
// ray indices per claim (one atomic on one address).  Round 1, 2 Mi rays per launch: 32 -> 436 us, 64 -> 477 us,
// 256 -> 703 us (load balance at the end of the launch).  With 32 Mi rays per launch the end of the launch matters
// less and the atomic more: 32 / 64 / 128 -> C4 2723 / 2756 / 2688, C5 1811 / 1830 / 1821 Msamples/s (r2_notes.md C11).
#define RT_FETCH_CHUNK 64
#endif

template <bool COUNT, bool VOLMESH>
__global__ void __launch_bounds__(RT_BLOCK, RT_EXTEND_MIN_BLOCKS) k_trace(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                                        rt_paths cur, rt_hits hits, const uint32_t* __restrict__ order) {
  __shared__ __align__(8) uint32_t sstack[(RT_SMEM_STACK + 6 + (VOLMESH ? 6 : 0)) * RT_BLOCK];
#if RT_TLAS_SMEM
  __shared__ float4 s_tlas[RT_TLAS_SMEM * RT_NODE_QUADS];
  load_tlas_cache(sc, s_tlas);
#endif
  const uint32_t n_rays = ctrl->n_rays;
  const uint32_t n_sorted = fr.sort_enabled ? ctrl->n_cont : 0u;  // continuing rays are visited in sorted order
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t lstack[RT_LOCAL_STACK];
  Trav T;
  T.sbase = (uint32_t)__cvta_generic_to_shared(sstack) + threadIdx.x * 4u;
  T.wbase = (uint32_t)__cvta_generic_to_shared(sstack) + RT_SMEM_STACK * (RT_BLOCK * 4u) + threadIdx.x * 4u;
  T.lstack = lstack;
#if RT_TLAS_SMEM
  T.tbase = (uint32_t)__cvta_generic_to_shared(s_tlas);
#endif
  T.t_min = fr.t_min; T.t_max = fr.t_max;
  T.k0 = fr.k0; T.k1 = fr.k1;
  T.qA = cur.A; T.qB = cur.B; T.qC = cur.C;
  T.slot = 0;
  bool have = false, fin = false;
  uint32_t c_next = 0, c_end = 0;  // this warp's claimed chunk of ray indices (warp uniform)
  bool exhausted = false;          // the queue has no more chunks (warp uniform)
#if RT_CLAIM_PREFETCH
  // the chunk after the current one is claimed ahead of time (lane 0 holds the result): by the time the warp needs
  // it, the atomic has long returned.  Each warp ends up with one claim past the end of the queue, which is harmless
  // (k_advance resets the cursor every iteration).
  uint32_t pf = 0;
  if (lane == 0) pf = atomicAdd(&ctrl->next_ray, (uint32_t)RT_FETCH_CHUNK);
#endif

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
#if RT_CLAIM_PREFETCH
        uint32_t base = __shfl_sync(FULL, pf, 0);
        if (lane == 0 && base + RT_FETCH_CHUNK < n_rays) pf = atomicAdd(&ctrl->next_ray, (uint32_t)RT_FETCH_CHUNK);
#else
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctrl->next_ray, (uint32_t)RT_FETCH_CHUNK);
        base = __shfl_sync(FULL, base, 0);
#endif
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

// ------------------------------------------------------------------ k_shade
// k_shade is latency bound on its gathers: 40 resident warps (48 registers, ~100 B of spills) beat 28 warps
// (72 registers, no spills) by 1 % on C4 and 7-10 % on the closed scenes C1 / C3 where shading dominates
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS (1280 / RT_BLOCK)  // 40 resident warps per SM
#endif
#ifndef RT_SHADE_WARP_SCAN
#define RT_SHADE_WARP_SCAN 1
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
    PathState p;
    p.o = mk(a.x, a.y, a.z);
    p.d = mk(a.w, bq.x, bq.y);
    p.T = mk(bq.z, bq.w, c.x);
    p.pixel = fbits(c.y);
    p.sb = fbits(c.z);
    p.tree = MULTI ? fbits(c.w) : 0u;
    float4 hr = RT_LDS(&hits.H[slot]);
    Best best;
    best.t = hr.x; best.u = hr.y; best.v = hr.z; best.prim = fbits(hr.w);
    best.obj = RT_LDS(&hits.obj[slot]);
    alive = shade_hit<COUNT, MULTI>(sc, fr, p, best, cls, accum, ctrl->counters);
    no = p.o; nd = p.d; nT = p.T;
    pixel = p.pixel; sb = p.sb; tree = p.tree;
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
#if RT_SHADE_WARP_SCAN
  if (warp == 0) {  // exclusive prefix of the 8 x RT_WARPS counters (octant-major) by one warp: each lane sums its run
                    // of consecutive counters, the lane sums are scanned with shuffles, each lane writes its run back
    constexpr int N = 8 * RT_WARPS, PER = (N + 31) / 32;
    uint32_t* cnt = &s_ocount[0][0];
    uint32_t c[PER], run = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = (int)lane * PER + k;
      c[k] = i < N ? cnt[i] : 0u;
      run += c[k];
    }
    uint32_t inc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
      if (lane >= (uint32_t)off) inc += t;
    }
    uint32_t base = inc - run;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = (int)lane * PER + k;
      if (i < N) cnt[i] = base;
      base += c[k];
    }
    if (lane == 31) s_base = inc ? atomicAdd(&ctrl->n_next, inc) : 0u;
  }
#else
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
#endif
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
      uint32_t peers = __match_any_sync(__activemask(), key);
#if RT_SORT_RANKED
      // the histogram atomic hands back the ray's rank inside its bin, so the scatter pass needs no atomics of its own
      const uint32_t leader = __ffs(peers) - 1u;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&sort.hist[key], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      reinterpret_cast<uint2*>(sort.keys)[pos] = make_uint2(key, base + __popc(peers & ((1u << lane) - 1u)));
#else
      sort.keys[pos] = key;
      if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&sort.hist[key], (uint32_t)__popc(peers));
#endif
    }
  }
}

// ------------------------------------------------------------------ k_path (megakernel engine)
// One persistent thread per path: the whole of Scene::shade_ray's recursion (tracing.rs:300-324) runs in place.
// A lane generates its camera ray (Camera::generate_rays, tracing.rs:159-209), finds the closest hit with the same
// traversal code as k_trace, shades it with the same shade_hit() as k_shade, and loops with the scattered ray; when
// its path ends it claims the next work index and starts over ("path regeneration").  Nothing but the lowered scene
// is read from HBM and nothing but the accumulator is written: the 148 bytes of queue state per path and six of the
// wavefront engine's seven kernels per bounce are gone.  The price is that the 32 lanes of a warp hold paths at
// different depths, so a warp's rays are less coherent than a sorted wavefront batch.
//
// The ray and the path's bookkeeping (throughput, pixel, sample | bounce) wait in shared memory while the lane
// traverses, in the layout of the wavefront's ray queue (A, B, C quads), so the traversal code re-reads the world
// ray and the RNG coordinates through the same pointers it uses there.  Images are bit-identical to the wavefront
// engine's: same functions, same Philox keys, integer accumulation.
// Measured (profiles/r2_notes.md C1, C3): 8 resident blocks (64 registers) beat 6 (80 registers) by 15-18 % on the scenes
// this engine is for - occupancy hides the shared-memory and L1 latency of the phase changes; and refilling only when
// at least 8 lanes are idle is within noise of refilling at once on C4 and happens to steer ptxas away from ~100 B of
// spills in the traversal loop, which alone is worth 20 % (same source, RT_PATH_REGEN_MIN 1 vs 8).
#ifndef RT_PATH_MIN_BLOCKS
#define RT_PATH_MIN_BLOCKS (1024 / RT_BLOCK)  // 32 resident warps per SM
#endif
#ifndef RT_PATH_REGEN_MIN
#define RT_PATH_REGEN_MIN 8   // idle lanes that trigger a regeneration round (an all-idle warp always regenerates)
#endif
#ifndef RT_PATH_CHUNK_MAX
#define RT_PATH_CHUNK_MAX 256  // work indices a warp claims per global atomic (guided: shrinks towards the end of the shard)
#endif

template <bool VOLMESH>
__global__ void __launch_bounds__(RT_BLOCK, RT_PATH_MIN_BLOCKS) k_path(rt_dev_scene sc, rt_frame fr, rt_ctrl* __restrict__ ctrl,
                                                                       long long* __restrict__ accum) {
  __shared__ __align__(16) uint32_t sstack[(RT_SMEM_STACK + 6 + (VOLMESH ? 6 : 0)) * RT_BLOCK];
  __shared__ float4 sA[RT_BLOCK], sB[RT_BLOCK], sC[RT_BLOCK];
#if RT_TLAS_SMEM
  __shared__ float4 s_tlas[RT_TLAS_SMEM * RT_NODE_QUADS];
  load_tlas_cache(sc, s_tlas);
#endif
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t lstack[RT_LOCAL_STACK];
  Trav T;
  T.sbase = (uint32_t)__cvta_generic_to_shared(sstack) + tid * 4u;
  T.wbase = (uint32_t)__cvta_generic_to_shared(sstack) + RT_SMEM_STACK * (RT_BLOCK * 4u) + tid * 4u;
  T.lstack = lstack;
#if RT_TLAS_SMEM
  T.tbase = (uint32_t)__cvta_generic_to_shared(s_tlas);
#endif
  T.t_min = fr.t_min; T.t_max = fr.t_max;
  T.k0 = fr.k0; T.k1 = fr.k1;
  T.qA = sA; T.qB = sB; T.qC = sC;
  T.slot = tid;
  bool have = false;
  // this warp's claimed chunk of work indices lives in shared memory (warp-uniform values; registers are what limits
  // the occupancy of this kernel): s_chunk[warp] = {base lo, base hi, next offset, count}
  __shared__ uint32_t s_chunk[RT_WARPS][4];
  uint32_t* const chunk = s_chunk[tid >> 5];
  if (lane < 4) chunk[lane] = 0u;
  __syncwarp();
  bool exhausted = false;              // the shard has no more chunks (warp uniform)
  uint32_t n_rays = 0;

  for (;;) {
    // ---- regeneration: idle lanes start the next camera paths of this warp's chunk
    uint32_t need = __ballot_sync(FULL, !have);
    if (need == FULL || __popc(need) >= RT_PATH_REGEN_MIN) {
#pragma unroll 1
      for (int round = 0; round < 2 && need; ++round) {
        uint32_t c_off = chunk[2], c_cnt = chunk[3];
        if (c_off >= c_cnt) {
          if (exhausted) break;
          const unsigned long long total = ctrl->total;
          unsigned long long base = 0;
          uint32_t cnt = 0;
          if (lane == 0) {
            // guided self-scheduling: large chunks while there is plenty of work, 32 at the end (load balance of the tail)
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&ctrl->cursor);
            unsigned long long want = cur < total ? (total - cur) / (4ull * gridDim.x * RT_WARPS) : 0ull;
            cnt = want >= RT_PATH_CHUNK_MAX ? (uint32_t)RT_PATH_CHUNK_MAX : (want < 32ull ? 32u : ((uint32_t)want & ~31u));
            base = atomicAdd(&ctrl->cursor, (unsigned long long)cnt);
          }
          base = __shfl_sync(FULL, base, 0);
          cnt = __shfl_sync(FULL, cnt, 0);
          if (base >= total) {
            exhausted = true;
            break;
          }
          unsigned long long left = total - base;
          c_off = 0;
          c_cnt = left < (unsigned long long)cnt ? (uint32_t)left : cnt;
          __syncwarp();
          if (lane == 0) {
            chunk[0] = (uint32_t)base;
            chunk[1] = (uint32_t)(base >> 32);
            chunk[3] = c_cnt;
          }
          __syncwarp();
        }
        const uint32_t avail = c_cnt - c_off;
        const uint32_t rank = __popc(need & ((1u << lane) - 1u));
        if (!have && rank < avail) {
          const unsigned long long c_base = (unsigned long long)chunk[0] | ((unsigned long long)chunk[1] << 32);
          uint32_t x, y, sample;
          if (work_to_pixel(fr, c_base + c_off + rank, x, y, sample)) {
            uint32_t pixel = y * fr.width + x;
            f3 o, d;
            camera_ray(fr, x, y, pixel, sample, o, d);
            sA[tid] = make_float4(o.x, o.y, o.z, d.x);
            sB[tid] = make_float4(d.y, d.z, 1.0f, 1.0f);
            sC[tid] = make_float4(1.0f, __uint_as_float(pixel), __uint_as_float(sample), 0.0f);  // bounce 0
            have = true;
          } else {
            atomicAdd(&ctrl->counters[7], 1ull);  // tile slot outside the image: the index is spent, the lane asks again
          }
        }
        __syncwarp();
        if (lane == 0) chunk[2] = c_off + min((uint32_t)__popc(need), avail);
        __syncwarp();
        need = __ballot_sync(FULL, !have);
      }
    }
    if (!__any_sync(FULL, have)) {
      if (exhausted && chunk[2] >= chunk[3]) break;
      continue;
    }

    // ---- closest hit (Scene::intersect_ray, tracing.rs:327-346)
    bool fin = true;
    if (have) {
      float4 a = sA[tid], b = sB[tid];
      T.t_max = T.ray_t_max = fr.t_max;
      trav_begin<false>(sc, T, mk(a.x, a.y, a.z), mk(a.w, b.x, b.y));
      fin = false;
      ++n_rays;
    }
    while (!fin) fin = trav_round<false, VOLMESH>(sc, T, fr.t_min, T.ray_t_max);
    __syncwarp();

    // ---- shade and scatter (tracing.rs:305-322); a miss is the black background (tracing.rs:302-303)
    if (have) {
      if (T.best.obj < 0) {
        have = false;
      } else {
        float4 a = sA[tid], b = sB[tid], c = sC[tid];
        PathState p;
        p.o = mk(a.x, a.y, a.z);
        p.d = mk(a.w, b.x, b.y);
        p.T = mk(b.z, b.w, c.x);
        p.pixel = fbits(c.y);
        p.sb = fbits(c.z);
        p.tree = 0u;
        have = shade_hit<false, false>(sc, fr, p, T.best, -1, accum, ctrl->counters);
        if (have) {
          sA[tid] = make_float4(p.o.x, p.o.y, p.o.z, p.d.x);
          sB[tid] = make_float4(p.d.y, p.d.z, p.T.x, p.T.y);
          sC[tid] = make_float4(p.T.z, __uint_as_float(p.pixel), __uint_as_float(p.sb), 0.0f);
        }
      }
    }
  }
  // ---- bookkeeping: one atomic per warp.  Every claimed index was started or counted as an invalid tile slot, so the
  // number of camera paths is the number of claimed indices.
  n_rays = __reduce_add_sync(FULL, n_rays);
  if (lane == 0) atomicAdd(&ctrl->n_rays_total, (unsigned long long)n_rays);
  if (blockIdx.x == 0 && tid == 0) ctrl->n_samples = ctrl->total;
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
#if RT_SORT_RANKED
  const uint2 kr = reinterpret_cast<const uint2*>(sort.keys)[i];
  sort.order[s_slice[kr.x >> 10] + sort.cursor[kr.x] + kr.y] = i;
  return;
#endif
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
  for (int i = 0; i < 4; ++i) c->poll[i] = 0;
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
int path_blocks_per_sm() {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_path<false>, RT_BLOCK, 0);
  return n > 0 ? n : 1;
}
void launch_path(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, long long* accum, unsigned long long total,
                 uint32_t persistent_blocks, cudaStream_t st) {
  unsigned long long full = (total + RT_BLOCK - 1) / RT_BLOCK;
  uint32_t grid = (uint32_t)(full < persistent_blocks ? full : persistent_blocks);
  if (grid == 0) return;
  if (sc.n_volume_meshes) k_path<true><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, accum);
  else k_path<false><<<grid, RT_BLOCK, 0, st>>>(sc, fr, ctrl, accum);
}
// Launch grids cover fr.grid_rays rays: the wavefront width normally, less once the host knows that the shard is
// exhausted and only a few rays are left to drain (empty blocks of a 131 072-block grid are not free).
static inline uint32_t grid_rays(const rt_frame& fr) { return fr.grid_rays ? fr.grid_rays : fr.capacity; }
void launch_raygen(const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, cudaStream_t st) {
  k_raygen<<<(grid_rays(fr) + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(fr, ctrl, cur);
}
void launch_trace(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_sortbuf sort,
                  bool count, uint32_t persistent_blocks, cudaStream_t st) {
  uint32_t full = (grid_rays(fr) + RT_BLOCK - 1) / RT_BLOCK;
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
  k_raysort_scatter<<<(grid_rays(fr) + 255) / 256, 256, 0, st>>>(ctrl, sort);
}
void launch_sort(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_hits hits, uint32_t* queues, cudaStream_t st) {
  k_sort<<<(grid_rays(fr) + RT_SORT_BLOCK - 1) / RT_SORT_BLOCK, RT_SORT_BLOCK, 0, st>>>(sc, fr, ctrl, hits.obj, queues);
}
void launch_surface(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_debug dbg,
                    cudaStream_t st) {
  k_surface<<<(fr.capacity + RT_BLOCK - 1) / RT_BLOCK, RT_BLOCK, 0, st>>>(sc, ctrl, cur, hits, dbg);
}
void launch_shade(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_paths nxt, rt_hits hits,
                  const uint32_t* queues, long long* accum, rt_sortbuf sort, bool count, cudaStream_t st) {
  uint32_t grid = (grid_rays(fr) + RT_BLOCK - 1) / RT_BLOCK + RT_NUM_CLASSES;
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
