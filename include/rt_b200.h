/* rt_b200.h — C ABI of librt_b200.so, the B200 (sm_100a) path-tracing back end.
 *
 * This is the drop-in boundary for the hot path of mbk6/CS397RayTracingSP22:
 * "whole scene in, whole image out".  The reference has no FFI; its hot path is
 * entered through one Rust call, Scene::render_to_image (src/util/tracing.rs:221),
 * over a scene made of struct literals (tracing.rs:138-155,213-218;
 * geometry.rs:389-393,424-429,468-472,495-500; materials.rs:20-23,51-55,74-76,
 * 107-112,152-157) and StaticMesh::load_from_file (geometry.rs:138).  Each entry
 * point below names the reference item it replaces.  Plain pointers and sizes
 * only; no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every function returns >= 0 on success (0, or the id of the thing added)
 *     and a negative rt_status on failure; nothing aborts or throws across the ABI
 *     (the reference panics instead: geometry.rs:149-151,168; tracing.rs:546);
 *   - rt_last_error() gives a thread-local message for the last failure;
 *   - input buffers are caller-owned and copied during the call;
 *   - outputs are caller-allocated;
 *   - object insertion order is preserved: it defines tie-breaking between
 *     top-level objects exactly as the reference's linear scan does
 *     (tracing.rs:330-341, strict '<' => first object wins);
 *   - a scene handle is not thread-safe; render calls block until done;
 *   - there is NO CPU fallback: without a CUDA device every device call fails
 *     with RT_ERR_CUDA.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 3

typedef enum rt_status {
  RT_OK = 0,
  RT_ERR_INVALID = -1,     /* bad argument / bad id / bad state            */
  RT_ERR_UNSUPPORTED = -2, /* reference feature that is out of scope here   */
  RT_ERR_CUDA = -3,        /* CUDA runtime error, no device, out of memory  */
  RT_ERR_IO = -4,          /* file could not be read / parsed               */
  RT_ERR_NOT_COMMITTED = -5
} rt_status;

/* ---- materials: the `Material` trait objects (materials.rs:12-15) as a tagged union */
typedef enum rt_material_tag {
  RT_MAT_LAMBERTIAN = 0,    /* materials.rs:20-48   albedo, emission                  */
  RT_MAT_METAL = 1,         /* materials.rs:51-71   albedo, emission, roughness       */
  RT_MAT_DIELECTRIC = 2,    /* materials.rs:74-104  ior                               */
  RT_MAT_PARAMETERIZED = 3, /* materials.rs:107-149 albedo, emission, roughness, metallic */
  RT_MAT_ISOTROPIC = 4      /* materials.rs:152-166 albedo, emission (phase function) */
} rt_material_tag;

typedef struct rt_material_desc {
  uint32_t tag; /* rt_material_tag */
  float albedo[3];
  float emission[3];
  float roughness;
  float metallic;
  float ior; /* Dielectric::idx_of_refraction */
} rt_material_desc;

/* ---- camera: field-for-field mirror of `Camera` (tracing.rs:138-155) */
enum { RT_PROJ_ORTHOGRAPHIC = 0, RT_PROJ_PERSPECTIVE = 1 };
enum { RT_SHADE_PHONG = 0, RT_SHADE_PATHTRACE = 1 };

typedef struct rt_camera {
  float eyepoint[3];
  float view_dir[3];
  float up[3];
  uint32_t projection_mode; /* RT_PROJ_ORTHOGRAPHIC is the reference's debug projection (tracing.rs:196,200) */
  uint32_t shading_mode;    /* RT_SHADE_PHONG is the reference's debug shading (tracing.rs:277-297)          */
  uint32_t path_depth;
  uint32_t path_samples; /* scattered rays per hit (tracing.rs:146,308-319); > 1 walks the path tree depth first */
  uint32_t screen_width;
  uint32_t screen_height;
  float focal_length;
  float focus_dist;
  float lens_radius;
  uint32_t aa_sample_count; /* perfect square (tracing.rs:152) */
  float max_trace_dist;
  float gamma;
} rt_camera;

/* ---- how one render call is cut out of the frame (replaces the rayon row loop,
 *      tracing.rs:228).  Every (pixel, sample) is independent and the RNG is keyed
 *      on (pixel, sample, bounce), so any partition gives the same sums. */
enum { RT_SHARD_ALL = 0, RT_SHARD_SAMPLES = 1, RT_SHARD_TILES = 2 };

enum {
  RT_OPT_COUNTERS = 1u << 0, /* count nodes / triangles / taps on the device (slower) */
  RT_OPT_NO_EVENTS = 1u << 1 /* do not bracket kernels with CUDA events               */
};

/* How the device walks the paths of one render call.  Both engines run the same device functions on the same
 * Philox keys and add into the same integer accumulators, so their images are bit-identical.
 *   WAVEFRONT : ray-gen / closest-hit / material-sorted shade kernels over a queue of paths in HBM
 *   MEGAKERNEL: one persistent kernel, one thread per path, the path's state in registers and shared memory,
 *               finished lanes start the next camera path at once
 *   AUTO      : chosen per call (pick_engine() in csrc/rt_api.cu; measurements in DESIGN.md).  The debug modes of
 *               the reference (ShadingMode::Phong, path_samples > 1) and the device counters exist in the
 *               wavefront engine only and always select it. */
enum { RT_ENGINE_AUTO = 0, RT_ENGINE_WAVEFRONT = 1, RT_ENGINE_MEGAKERNEL = 2 };
/* wavefront engine only: sort the scattered rays by (origin cell, direction class) before the next closest-hit pass */
enum { RT_RAYSORT_AUTO = 0, RT_RAYSORT_OFF = 1, RT_RAYSORT_ON = 2 };
/* order in which work indices walk the shard: AUTO picks per shard mode */
enum { RT_ORDER_AUTO = 0, RT_ORDER_PIXEL_MAJOR = 1, RT_ORDER_SAMPLE_MAJOR = 2, RT_ORDER_GROUPED = 3 };

typedef struct rt_render_opts {
  uint64_t seed;          /* Philox key                                              */
  uint32_t shard_mode;    /* RT_SHARD_*                                              */
  uint32_t shard_rank;    /* this rank, 0..shard_count-1                             */
  uint32_t shard_count;   /* number of ranks the frame is cut into (>=1)             */
  uint32_t tile_size;     /* RT_SHARD_TILES: square tile edge in pixels (0 => 64)    */
  uint32_t sample_begin;  /* RT_SHARD_ALL/TILES: [begin,end) sample indices;         */
  uint32_t sample_end;    /*   both 0 => [0, aa_sample_count)                        */
  uint32_t wavefront;     /* wavefront engine: paths in flight, 148 bytes of device state each (0 => library default:
                             32 Mi paths = 5 GB, halved until the allocation succeeds) */
  uint32_t flags;         /* RT_OPT_*                                                */
  float point_light_pos[3]; /* Scene::point_light_pos, used by ShadingMode::Phong only (tracing.rs:216) */
  float ambient[3];         /* Scene::ambient, Phong only (tracing.rs:217)                              */
  uint32_t engine;        /* RT_ENGINE_*  (0 => AUTO)                                */
  uint32_t ray_sort;      /* RT_RAYSORT_* (0 => AUTO: on when the instanced triangle count is >= 16 Ki) */
  uint32_t work_order;    /* RT_ORDER_*   (0 => AUTO)                                */
  uint32_t blocks_per_sm; /* resident blocks per SM of the persistent kernels (0 => what the occupancy query says) */
  uint32_t reserved[4];   /* must be 0                                               */
} rt_render_opts;

typedef struct rt_stats {
  uint64_t samples;           /* camera paths started                               */
  uint64_t rays;              /* closest-hit queries (= Scene::intersect_ray calls) */
  uint64_t iterations;        /* wavefront iterations                               */
  uint64_t kernel_launches;   /* kernels launched by this call                      */
  uint64_t extend_launches;   /* k_trace launches that had rays                     */
  uint64_t shade_launches;    /* k_shade launches that had rays                     */
  /* device counters, filled only with RT_OPT_COUNTERS */
  uint64_t nodes_visited;     /* 32-byte BVH nodes fetched                          */
  uint64_t tlas_nodes_visited;/* of those, nodes of the top-level BVH               */
  uint64_t tris_tested;       /* 48-byte triangle records fetched                   */
  uint64_t instances_entered; /* 96 bytes of transforms fetched                     */
  uint64_t prims_tested;      /* analytic sphere/triangle/plane/volume records      */
  uint64_t mesh_hits;         /* 80-byte shading records fetched                    */
  uint64_t texel_taps;        /* RGB8 texel fetches (all kernels)                   */
  uint64_t extend_texel_taps; /* of those, normal-map taps made inside k_extend     */
  uint64_t material_fetches;  /* 32-byte material records fetched                   */
  uint64_t warp_node_slots;   /* 32 x (most nodes fetched by any lane) summed over warp batches:
                                 nodes_visited / warp_node_slots = SIMT efficiency of traversal */
  /* CUDA-event times on the render stream, milliseconds */
  double ms_total;
  double ms_extend;           /* k_trace only (the dominant kernel)                 */
  double ms_shade;            /* k_sort + k_shade                                   */
  double ms_resolve;
  uint64_t h2d_bytes;         /* bytes copied host->device by this call             */
  uint64_t d2h_bytes;         /* bytes copied device->host by this call             */
  uint64_t engine;            /* RT_ENGINE_* that rendered this call (never AUTO)    */
} rt_stats;

typedef struct rt_scene rt_scene; /* opaque, library-owned: replaces `Scene.objects` (tracing.rs:215) */

int rt_abi_version(void);
const char* rt_last_error(void);
int rt_device_count(void); /* number of CUDA devices, or RT_ERR_CUDA */

int rt_scene_create(rt_scene** out);
void rt_scene_destroy(rt_scene* s);

/* Texture::load_from_file result (texture.rs:16-25) after decode: RGB8, row 0 = top. */
int rt_add_texture(rt_scene* s, const uint8_t* rgb8, uint32_t width, uint32_t height);
/* Arc<dyn Material> (materials.rs:12-166) */
int rt_add_material(rt_scene* s, const rt_material_desc* desc);
/* tobj::Mesh as StaticMesh uses it (geometry.rs:157,220-243): single-index arrays.
 * pos/nrm are 3*nverts floats, uv 2*nverts, idx 3*ntris.  Builds the BLAS once. */
int rt_add_mesh(rt_scene* s, const float* pos, const float* nrm, const float* uv, uint32_t nverts,
                const uint32_t* idx, uint32_t ntris);
/* StaticMesh (geometry.rs:127-134,300-314).  xform / inv_xform are column-major 4x4
 * (cgmath Matrix4 layout); inv_xform may be NULL (computed by cofactors, like
 * cgmath's inverse_transform, geometry.rs:168).  material = -1 => textures drive a
 * ParameterizedMaterial (geometry.rs:253-271).  tex = {albedo, emission, metallic,
 * roughness, normal} texture ids or -1 (geometry.rs:130). */
int rt_add_instance(rt_scene* s, int mesh, const float xform[16], const float* inv_xform,
                    int material, const int tex[5]);
int rt_add_sphere(rt_scene* s, const float center[3], float radius, int material);   /* geometry.rs:389-413 */
int rt_add_triangle(rt_scene* s, const float a[3], const float b[3], const float c[3],
                    int material);                                                   /* geometry.rs:424-450 */
int rt_add_plane(rt_scene* s, const float point[3], const float normal[3], int material); /* geometry.rs:468-489 */
/* ConvexVolume with a Sphere boundary (geometry.rs:495-526) */
int rt_add_volume_sphere(rt_scene* s, const float center[3], float radius, float density,
                         int phase_material);

/* ConvexVolume with a StaticMesh boundary (geometry.rs:495-526 with boundary = a StaticMesh, geometry.rs:300-314):
 * entry = the boundary's closest hit over all t, exit = its closest hit beyond entry + 1e-4.  The boundary's own
 * material and textures are ignored, as in the reference.  xform / inv_xform as for rt_add_instance. */
int rt_add_volume_mesh(rt_scene* s, int mesh, const float xform[16], const float* inv_xform, float density,
                       int phase_material);

/* Lower the scene (reachability mask, binned-SAH BLAS/TLAS, tables) and upload it to
 * CUDA device `device`.  May be called again after more rt_add_* calls. */
int rt_commit(rt_scene* s, int device);
/* Host-only half of rt_commit: lower the scene (no CUDA needed) and report what was built. */
typedef struct rt_lower_info {
  uint64_t bytes;          /* size of the lowered scene                              */
  uint32_t nodes;          /* 32-byte BVH nodes, all BLASes + TLAS                   */
  uint32_t tris;           /* packed (reachable) triangle records                    */
  uint32_t objects;        /* top-level objects                                      */
  uint32_t unbounded;      /* objects tested for every ray (planes)                  */
  uint32_t tlas_depth;     /* longest root-to-leaf path of the TLAS, in nodes        */
  uint32_t max_blas_depth; /* same for the deepest BLAS                              */
  uint32_t guard_boxes;    /* thin interior boxes of the reference tree kept as guards */
  uint32_t guarded_tris;   /* triangles that have at least one guard                  */
} rt_lower_info;
int rt_scene_lower(rt_scene* s, rt_lower_info* info);
/* Bytes of lowered scene data resident on the device (what rt_commit uploads). */
uint64_t rt_scene_device_bytes(const rt_scene* s);
/* Re-upload the already lowered scene from host memory (used to time host->device). */
int rt_scene_upload(rt_scene* s);

/* Scene::render_to_image (tracing.rs:221-263), host buffers.
 * out_linear_rgb: W*H*3 floats, per-pixel mean radiance BEFORE the output transform
 * (tracing.rs:241); out_rgb8: W*H*3 bytes AFTER saturate/clamp/gamma/quantise
 * (tracing.rs:243-256).  Either may be NULL.  Row 0 = top, like RgbImage. */
int rt_render(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts,
              float* out_linear_rgb, uint8_t* out_rgb8, rt_stats* stats);

/* Device-resident pieces of the same call, for one-process-per-GPU sharding.
 * d_accum is a DEVICE pointer to W*H*4 int64: fixed-point (2^-30) radiance sums for
 * R,G,B and a word of per-channel NaN flags; rt_render_accum ADDS this shard's samples into it
 * (zero it first).  Because the sums are integers they are exact and independent of
 * order, so summing shards (e.g. with an NCCL int64 all-reduce) is bit-identical to
 * one GPU rendering everything.  `stream` is a cudaStream_t (0 => default stream). */
int rt_render_accum(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts,
                    void* d_accum, void* stream, rt_stats* stats);
/* mean + output transform (tracing.rs:241-256) from an accumulator holding
 * `total_spp` samples per pixel; d_out_* are DEVICE pointers (either may be NULL). */
int rt_resolve(rt_scene* s, const rt_camera* cam, const void* d_accum, uint32_t total_spp,
               float* d_out_linear_rgb, uint8_t* d_out_rgb8, void* stream);
size_t rt_accum_bytes(uint32_t width, uint32_t height);

/* Progressive / checkpointed rendering, host buffers.  Adds the sample indices
 * [opts->sample_begin, opts->sample_end) of every pixel to the HOST accumulator `accum` (W*H*4 int64, rt_accum_bytes();
 * zero it before the first call) and, if an output pointer is given, resolves the image of the `spp_in_accum` samples
 * per pixel the accumulator holds after this call.  The accumulator IS the checkpoint: write it to disk, read it back
 * in another process, continue with the next sample range.  Because the sums are integers and the RNG is keyed on
 * (pixel, sample, bounce), a render that was interrupted and resumed any number of times is bit-identical to one that
 * was not.  The reference has no counterpart: its render_to_image (tracing.rs:221-263) is all or nothing. */
int rt_render_progressive(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts, int64_t* accum,
                          uint32_t spp_in_accum, float* out_linear_rgb, uint8_t* out_rgb8, rt_stats* stats);

/* Parity hooks (no counterpart in the reference; they expose what
 * Scene::intersect_ray, tracing.rs:327-346, returns).
 * rt_trace_primary: for every pixel, the camera ray of sample index `sample`
 * (Camera::generate_rays, tracing.rs:159-209) and its closest hit in
 * [0.001, max_trace_dist].  All outputs optional (NULL), host buffers, W*H entries:
 * obj_id = index of the top-level object (insertion order) or -1, prim_id = triangle
 * index inside a mesh (0 otherwise), t, world normal (3 floats), ray origin+direction
 * (6 floats). */
int rt_trace_primary(rt_scene* s, const rt_camera* cam, uint64_t seed, uint32_t sample,
                     int32_t* obj_id, int32_t* prim_id, float* t, float* normal_xyz,
                     float* ray_od);
/* Closest hit for n caller-supplied rays (origin xyz, direction xyz; directions are
 * NOT normalised, like the reference's scattered rays).  Ray i uses RNG key
 * (pixel=i, sample=0, bounce=0) for volume free-flight draws. */
int rt_intersect_rays(rt_scene* s, uint64_t seed, uint32_t n, const float* ray_od, float t_min,
                      float t_max, int32_t* obj_id, int32_t* prim_id, float* t,
                      float* normal_xyz, float* hitpoint_xyz, float* uv, int32_t* frontface);

/* ---- asset readers: what tobj::load_obj (geometry.rs:140-148) and image::open
 *      (texture.rs:17) do for the reference, for callers without those crates. */
typedef struct rt_obj_mesh {
  uint32_t nverts;
  uint32_t ntris;
  float* pos;    /* 3*nverts */
  float* nrm;    /* 3*nverts (zeros when the file has no vn) */
  float* uv;     /* 2*nverts (zeros when the file has no vt) */
  uint32_t* idx; /* 3*ntris  */
  uint32_t has_normals;
  uint32_t has_texcoords;
} rt_obj_mesh;
/* single_index + triangulate (fan), first model only (geometry.rs:143-144,157) */
int rt_obj_parse(const char* text, size_t len, rt_obj_mesh* out);
int rt_obj_load(const char* path, rt_obj_mesh* out);
void rt_obj_free(rt_obj_mesh* m);
/* TGA (types 2,3,10,11; 8/24/32 bpp) -> RGB8 top-down; *rgb is malloc'ed, free with rt_free */
int rt_tga_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h);
int rt_tga_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len);
/* PNG (non-interlaced; grey / RGB / palette / alpha, 1-16 bit) -> RGB8 top-down, and RGB8 -> PNG (stored deflate).
 * Self-contained inflate, no zlib.  What image::open (texture.rs:17) and save_with_format (tracing.rs:546) do. */
int rt_png_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h);
int rt_png_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len);
/* JPEG (baseline, extended sequential and progressive Huffman; 8 bit; grey or YCbCr/RGB; any sampling) -> RGB8
 * top-down.  What image::open (texture.rs:17) does for the reference's .jpg textures through jpeg-decoder 0.1.22. */
int rt_jpeg_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h);
void rt_free(void* p);

/* Reachability mask of the reference's index-order BVH (geometry.rs:190-217 with the
 * strict slab test geometry.rs:63-67): mask[i] = 1 when triangle i can be hit.
 * Exposed so tests can compare it with the oracle's replay of the reference tree. */
int rt_mesh_reachability(const float* pos, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                         uint8_t* mask);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
