// rt_types.h — layouts shared by the host lowering (rt_lower.cpp) and the sm_100a kernels
// (rt_kernels.cu).  Everything the kernels read is a flat array of 16-byte quads so that
// each record is fetched with 128-bit read-only loads.
//
// Lowered scene, all in HBM (and in practice L2-resident):
//   nodes   : 2 quads (32 B) per BVH node  lo=(min.xyz, link) hi=(max.xyz, count)
//             count==0: interior, link = index of the left child; children are nodes link and link+1
//                       (one 64-byte pair, fetched together)
//             count>0 : leaf, link = RT_LEAF_FLAG | first << 4 | count;
//                       BLAS leaf => `count` triangle records from `first`,
//                       TLAS leaf => count==1, first = top-level object index
//   tris    : 3 quads (48 B) per triangle, leaf order: (v0.xyz,e1.x)(e1.yz,e2.xy)(e2.z,id,guard first,guard count)
//   guards  : 2 quads per guard box + a uint32 index list: thin interior boxes of the REFERENCE's tree whose strict
//             slab test must be replayed before a hit below them is accepted (rare; see rt_lower.cpp)
//   shade   : 5 quads (80 B) per triangle, ORIGINAL order: na nb nc uva uvb uvc tangent
//   objects : 10 quads (160 B) per top-level object, insertion order (tie-breaking!)
//   mats    : 2 quads (32 B) per material
//   texels  : RGBA8, one 32-bit word per texel; textures: (offset, w, h, _)
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stdint.h>

#define RT_NODE_QUADS 2
#define RT_TRI_QUADS 3
#define RT_SHADE_QUADS 5
#define RT_OBJ_QUADS 10
#define RT_MAT_QUADS 2

#define RT_LEAF_FLAG 0x80000000u
#define RT_ENTRY_NONE 0xFFFFFFFFu
#define RT_ENTRY_RESTORE 0xFFFFFFFEu
#define RT_ENTRY_VOLRET 0xFFFFFFFDu  /* end of a boundary query of a mesh-bounded volume */
#ifndef RT_SORT_BINS
#define RT_SORT_BINS 262144u  /* ray-sort key space: (spatial hash of the origin cell) << 5 | 5-bit direction class; a power of two >= 2^15 */
#endif
#define RT_MAX_LEAF_TRIS 8  /* fits the 4-bit count of a packed stack entry */

enum rt_obj_kind { RT_OBJ_MESH = 0, RT_OBJ_SPHERE = 1, RT_OBJ_TRIANGLE = 2, RT_OBJ_PLANE = 3, RT_OBJ_VOLUME = 4,
                   RT_OBJ_VOLUME_MESH = 5 };

// shade classes: the material tag, plus one class for texture-driven mesh hits
enum { RT_CLASS_LAMBERT = 0, RT_CLASS_METAL = 1, RT_CLASS_DIELECTRIC = 2, RT_CLASS_PARAM = 3,
       RT_CLASS_ISOTROPIC = 4, RT_CLASS_PARAM_TEX = 5, RT_NUM_CLASSES = 6 };

// object record, quad 0 (as ints): kind, material id (-1: textured), class of that material, aux
//   MESH     q1..q3 rows of inv_transform (3x4), q4..q6 rows of transform (3x4),
//            q7 = (blas root entry, shade base, tex albedo, tex emission) ints
//            q8 = (tex metallic, tex roughness, tex normal, _) ints
//   SPHERE   q1 = (center.xyz, radius)
//   TRIANGLE q1 = (a.xyz, e1.x) q2 = (e1.yz, e2.xy) q3 = (e2.z, n.xyz)   n = normalize(e1 x e2)
//   PLANE    q1 = (point.xyz, _) q2 = (normal.xyz, _)
//   VOLUME   q1 = (center.xyz, radius) q2 = (density, vol_index(int), _, _)
//   VOLUME_MESH  like MESH (q1..q3 inverse transform, q7.x BLAS root) + q9 = (density, vol_index(int), _, _)

struct rt_dev_scene {
  const void* nodes;     // float4*
  const void* tris;      // float4*
  const void* shade;     // float4*
  const void* objects;   // float4*
  const void* mats;      // float4*
  const void* textures;  // uint4*  (texel offset, w, h, 0)
  const void* texels;    // uint32_t* RGBA8
  const void* planes;    // int32_t* indices of unbounded objects (always tested)
  const void* guards;    // float4* guard boxes (min, max)
  const void* guard_list;// uint32_t* per-triangle guard indices
  uint32_t tlas_root;    // packed entry of the TLAS root, RT_ENTRY_NONE when there is none
  uint32_t tlas_base;    // node index of the first TLAS node (the TLAS is stored breadth first behind all BLASes)
  uint32_t tlas_count;   // number of TLAS nodes (incl. the padding node 1)
  uint32_t n_planes;
  uint32_t n_objects;
  uint32_t n_volumes;
  uint32_t n_volume_meshes;  // > 0 selects the k_trace variant that can run boundary queries
  float tlas_min[3];     // root box of the TLAS
  float tlas_max[3];
};

// device control block for the wavefront loop (one per scene, lives in HBM)
struct rt_ctrl {
  unsigned long long cursor;       // next work index to hand out
  unsigned long long total;        // work indices in this shard
  unsigned long long n_samples;    // camera paths started
  unsigned long long n_rays_total; // closest-hit queries issued
  uint32_t n_cont;                 // rays carried over from the previous bounce
  uint32_t n_rays;                 // rays this iteration (n_cont + newly generated)
  unsigned long long work_base;    // work index of the first newly generated ray
  uint32_t n_next;                 // shade's output counter (next iteration's n_cont)
  uint32_t next_ray;               // k_extend's dynamic ray fetch cursor
  uint32_t pad0_;
  uint32_t class_count[RT_NUM_CLASSES];
  uint32_t done;                   // 1 when cursor==total and nothing is in flight
  uint32_t iterations;
  uint32_t pad_;
  unsigned long long counters[12]; // nodes, tris, instances, prims, mesh_hits, shade taps, mats, invalid tile slots,
                                   // normal-map taps, warp node slots, TLAS nodes, -
  uint32_t poll[4];                // what the host peeks at: done, exhausted (cursor == total), n_rays, iterations
};

struct rt_frame {
  // camera, precomputed once per frame on the host exactly as tracing.rs:160-191 does per sample
  float eye[3];
  float rot0[3], rot1[3], rot2[3];  // columns of `rotation` (tracing.rs:187-191)
  float pixel_size, n, rootn, focal_length, focus_dist, lens_radius;
  float t_min, t_max;
  uint32_t rooti, spp, width, height, path_depth;
  uint32_t k0, k1;                  // Philox key
  // shard
  uint32_t shard_mode, shard_rank, shard_count, tile_size, tiles_x, tiles_y;
  const uint32_t* tile_list;        // RT_SHARD_TILES: row-major tile ids of this shard, in the order it walks them (device pointer)
  uint32_t sample_begin, sample_count;
  uint32_t sample_major;            // 1: consecutive work indices walk the pixels (sample index changes slowest)
  unsigned long long pixel_slots;   // pixel slots of this shard (work items = pixel_slots * sample_count)
  uint32_t capacity;                // wavefront width P (also the stride of the shade queues)
  uint32_t grid_rays;               // host-side upper bound of the rays of this iteration: sizes the launch grids (0 => capacity)
  // debug modes of the reference: orthographic projection (tracing.rs:196,200), Phong shading (tracing.rs:277-297)
  uint32_t path_samples, branch;    // Camera::path_samples > 1 (tracing.rs:146,308-319): child index of this k_shade pass
  uint32_t phong;                   // host-side switch: camera ray + shadow ray pairs instead of path iterations
  uint32_t ortho, ray_tmax_from_c;  // ray_tmax_from_c: a ray's own t_max travels in C.w (Phong shadow rays)
  float view_dir[3], light[3], ambient[3];
  // secondary-ray sorting (k_shade -> k_raysort_*): origin cell (16^3 over the TLAS box, Morton order) | direction octant
  uint32_t sort_enabled;
  float sort_min[3], sort_scale[3];
  float sort_cells_m1;              // cells per axis - 1 (<= 15)
  uint32_t sort_use_octant;
};

#endif
