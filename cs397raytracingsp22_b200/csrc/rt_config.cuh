// rt_config.cuh — compile-time knobs of the kernels (each one was measured, see profiles/r1_notes.md and r2_notes.md)
#ifndef RT_CONFIG_CUH
#define RT_CONFIG_CUH

// Threads per block of the ray kernels (k_raygen, k_trace, k_shade, k_path, ...).  Measured at the same number of
// resident warps per SM (profiles/r2_notes.md C11): 64 / 128 / 256 threads -> C4 2557 / 2658 / 2677, C3 984 / 1080 /
// 1101 Msamples/s (k_shade pays one atomic on the queue cursor and two barriers per block; a bigger block halves the
// atomics, and k_trace does not care).
#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif
#define RT_WARPS (RT_BLOCK / 32)
// Traversal-stack entries per thread kept in shared memory; the others live in local memory, i.e. in L1.  Measured
// (profiles/r2_notes.md C7, C11): 16 entries mean a 100 KB carve-out; 4 entries (64 KB) are worth 0.4-0.6 % of that,
// none at all (32 KB carve-out, 224 KB of L1; the stack's hot lines stay in L1 anyway) another 0.2 % on C4 and 0.5 %
// on C5; 2 and 8 entries are slower than either neighbour.
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 0
#endif
#define RT_LOCAL_STACK (64 - RT_SMEM_STACK)  // overflow entries (local memory; 64 in all, the host checks depth <= 62)
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
// Children per interior node: 2 (64-byte pairs) or 4 with -DRT_BVH4=1 (128-byte groups, collapsed from the binary tree
// by rt_lower.cpp, which must be compiled with the same switch).
#ifndef RT_BVH4
#define RT_BVH4 0
#endif
// TLAS node records (32 B each, breadth-first from the root) every persistent block keeps in shared memory; 0 = none.
// Must be a multiple of 4.  Measured (profiles/r2_notes.md): 128 records cost 4 % on C4 and C5 (the shared memory
// comes out of L1, where the BLAS nodes live), 512 records 11 %; so it is off.
#ifndef RT_TLAS_SMEM
#define RT_TLAS_SMEM 0
#endif
// k_trace can claim its next chunk of ray indices while it traverses the current one (the latency of the
// single-address atomic hides behind the traversal).  Measured: -0.4 % on C4, -0.7 % on C5 - the 32 resident warps
// already hide it, and a chunk held back by a slow warp lengthens the tail; so it is off.
#ifndef RT_CLAIM_PREFETCH
#define RT_CLAIM_PREFETCH 0
#endif
// Ray sort: k_shade keeps the value its histogram atomic returns (= the ray's rank inside its bin), which frees the
// scatter pass from a second round of atomics at the price of 4 more bytes per ray and an atomic whose result is waited
// for.  Measured: +2.1 % on C4, +2.3 % on C5 (profiles/r2_notes.md C6).
#ifndef RT_SORT_RANKED
#define RT_SORT_RANKED 1
#endif
// How a child pair of the binary tree is stored (RT_BVH4 above selects the tree).
//   0: (min, max) per child, four 128-bit fetches per visit, slab test = 12 FFMA + 20 FMNMX(3)
//   1: (centre, half extent) per child, four fetches, slab test = 18 FFMA + 8 FMNMX(3) (see slab_ch(), rt_traverse.cuh)
//   2: packed pair - two centres with their links, six half extents as bf16 (rounded up) - THREE fetches per visit;
//      the fourth quad of the 64-byte slot is never read.
// Measured (profiles/r2_notes.md C9): every extra 128-bit fetch per visit costs k_trace 6-9 % (the visit sits at the
// knee of the L1 data pipe: 4 wavefronts per LDG.128 whatever the number of active lanes), six ALU instructions
// fewer per visit are worth 0.6 %; the packed pair is worth 1.1 % of k_trace on C4 and 2.8 % on C5 for 1 % more node
// visits (bf16 inflates a box by < 0.8 % of its half extent).  rt_lower.cpp is compiled with the same switch (one nvcc
// command).  The 4-wide tree keeps (min, max).
#ifndef RT_NODE_CH
#if RT_BVH4
#define RT_NODE_CH 0
#else
#define RT_NODE_CH 2
#endif
#endif
#if RT_NODE_CH && RT_BVH4
#error "RT_NODE_CH is implemented for the binary tree only"
#endif
#ifndef RT_EXTRA_LDG
#define RT_EXTRA_LDG 0  // diagnostic: extra node fetches per visit (rt_traverse.cuh)
#endif
#ifndef RT_EXTEND_MIN_BLOCKS
#define RT_EXTEND_MIN_BLOCKS (1024 / RT_BLOCK)  // 64 registers per thread -> 32 resident warps per SM
#endif

#endif
