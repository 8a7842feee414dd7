// rt_api.cu — implementation of the C ABI in include/rt_b200.h: scene IR, commit (lower +
// upload), the host side of the wavefront loop, resolve, parity hooks, asset readers.
// There is deliberately no CPU rendering path in this library: every device entry point fails
// with RT_ERR_CUDA when no CUDA device is usable.

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rt_kernels.h"
#include "rt_lower.h"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) return fail(RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};
}  // namespace

struct rt_scene {
  std::vector<rt::HostTexture> textures;
  std::vector<rt_material_desc> materials;
  std::vector<rt::HostMesh> meshes;
  std::vector<rt::HostObject> objects;
  int n_volumes = 0;

  // lowered (host) + resident (device)
  rt::Lowered low;
  std::vector<void*> pinned;  // lowered arrays registered as page-locked memory, so rt_scene_upload copies at bus speed
  bool pin_tried = false;
  bool lowered = false;
  bool committed = false;
  int device = -1;
  DevBuf d_geom;  // BVH nodes, then (256-byte aligned) triangle records: one range, so one L2 access-policy window covers both
  DevBuf d_shade, d_objects, d_mats, d_textures, d_texels, d_planes, d_guards, d_guard_list;
  rt_dev_scene dev{};

  // wavefront engine state: ray queues, hit records, shade queues, sort buffers, control block
  struct Wavefront {
    uint32_t capacity = 0;
    rt::rt_paths paths[2] = {};
    rt::rt_hits hits = {};
    uint32_t* queues = nullptr;
    rt::rt_sortbuf sort = {};
    rt_ctrl* ctrl = nullptr;     // also the control block of the megakernel engine
    rt_ctrl* h_ctrl = nullptr;   // pinned
    uint32_t* h_done = nullptr;  // pinned poll slots: 2 x rt_ctrl::poll
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;
  };
  Wavefront wf;
  cudaStream_t own_stream = nullptr;
  size_t geom_bytes = 0;
  uint32_t sm_count = 148;
  uint32_t trace_blocks_per_sm = 8;  // resident blocks per SM of the persistent kernels (occupancy query)
  uint32_t path_blocks_per_sm = 6;
  // scratch for host-buffer entry points
  DevBuf d_accum, d_linear, d_rgb8, d_dbg;
  DevBuf d_tree[3];  // ray pool of the depth-first walk (path_samples > 1)
  DevBuf d_tiles;    // RT_SHARD_TILES: this shard's tile ids
  std::vector<uint32_t> h_tiles;
};

namespace {

void free_buf(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}
// two timing events that do not outlive an early error return
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  int create() {
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    return RT_OK;
  }
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

int ensure_buf(DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return RT_OK;
  free_buf(b);
  CUDA_TRY(cudaMalloc(&b.p, std::max<size_t>(bytes, 16)));
  b.bytes = std::max<size_t>(bytes, 16);
  return RT_OK;
}
// The big lowered arrays (texels, nodes, triangle and shading records) are page-locked in place on the first upload:
// a copy from pageable memory is staged through the driver's bounce buffers at a fraction of the bus speed.
void unpin_lowered(rt_scene* s) {
  for (void* p : s->pinned) cudaHostUnregister(p);
  s->pinned.clear();
  s->pin_tried = false;
}
void pin_lowered(rt_scene* s) {
  if (s->pin_tried) return;
  s->pin_tried = true;
  const rt::Lowered& L = s->low;
  const std::pair<const void*, size_t> arrays[] = {
      {L.texels.data(), L.texels.size() * 4}, {L.nodes.data(), L.nodes.size() * 16}, {L.tris.data(), L.tris.size() * 16},
      {L.shade.data(), L.shade.size() * 16}};
  for (const auto& a : arrays) {
    if (a.second < (1u << 16)) continue;
    if (cudaHostRegister(const_cast<void*>(a.first), a.second, cudaHostRegisterDefault) == cudaSuccess) s->pinned.push_back(const_cast<void*>(a.first));
    else cudaGetLastError();  // not fatal: that array is copied from pageable memory
  }
}

int upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
  int rc = ensure_buf(b, bytes);
  if (rc != RT_OK) return rc;
  if (bytes) CUDA_TRY(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
  return RT_OK;
}

void free_wavefront(rt_scene* s) {
  rt_scene::Wavefront& L = s->wf;
  for (int k = 0; k < 2; ++k) {
    if (L.paths[k].A) cudaFree(L.paths[k].A);
    if (L.paths[k].B) cudaFree(L.paths[k].B);
    if (L.paths[k].C) cudaFree(L.paths[k].C);
    L.paths[k] = rt::rt_paths{};
  }
  if (L.hits.H) cudaFree(L.hits.H);
  if (L.hits.obj) cudaFree(L.hits.obj);
  L.hits = rt::rt_hits{};
  if (L.queues) cudaFree(L.queues);
  L.queues = nullptr;
  if (L.sort.keys) cudaFree(L.sort.keys);
  if (L.sort.order) cudaFree(L.sort.order);
  L.sort.keys = L.sort.order = nullptr;
  L.capacity = 0;
}

// stream, control block, poll slots: what both engines need
int ensure_runtime(rt_scene* s) {
  if (s->own_stream) return RT_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  s->sm_count = (uint32_t)std::max(1, sms);
  s->trace_blocks_per_sm = (uint32_t)rt::trace_blocks_per_sm();
  s->path_blocks_per_sm = (uint32_t)rt::path_blocks_per_sm();
  rt_scene::Wavefront& L = s->wf;
  CUDA_TRY(cudaMalloc((void**)&L.ctrl, sizeof(rt_ctrl)));
  CUDA_TRY(cudaMallocHost((void**)&L.h_ctrl, sizeof(rt_ctrl)));
  CUDA_TRY(cudaMallocHost((void**)&L.h_done, 2 * 4 * sizeof(uint32_t)));
  CUDA_TRY(cudaEventCreateWithFlags(&L.poll_ev[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&L.poll_ev[1], cudaEventDisableTiming));
  return RT_OK;
}
uint32_t persistent_grid(const rt_scene* s, uint32_t per_sm_default, uint32_t requested) {
  uint32_t per_sm = requested ? std::min(requested, per_sm_default) : per_sm_default;
  return s->sm_count * std::max(1u, per_sm);
}

// RT_L2_PERSIST: while an engine runs, the BVH nodes and triangle records (one range, a few MB) are marked
// "persisting" in L2 through the stream's access-policy window, so the ray queues streaming through L2 cannot evict
// them.  The window is taken off the stream again when the engine returns (the stream may be the caller's).
#ifndef RT_L2_PERSIST
#define RT_L2_PERSIST 0
#endif
struct L2Window {
  cudaStream_t st = nullptr;
  bool on = false;
  void open(const rt_scene* s, cudaStream_t stream) {
#if RT_L2_PERSIST
    if (!s->d_geom.p || !s->geom_bytes) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, s->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, s->device);
    size_t bytes = std::min<size_t>(s->geom_bytes, (size_t)std::max(0, max_window));
    if (!bytes || !max_persist) return;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)max_persist)) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.base_ptr = s->d_geom.p;
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = bytes <= (size_t)max_persist ? 1.0f : (float)max_persist / (float)bytes;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    st = stream;
    on = true;
#else
    (void)s; (void)stream;
#endif
  }
  ~L2Window() {
    if (!on) return;
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaCtxResetPersistingL2Cache();
  }
};

int ensure_wavefront(rt_scene* s, uint32_t capacity) {
  int rc = ensure_runtime(s);
  if (rc != RT_OK) return rc;
  rt_scene::Wavefront& L = s->wf;
  if (L.capacity >= capacity && L.queues) return RT_OK;
  free_wavefront(s);
  size_t n = capacity;
  for (int k = 0; k < 2; ++k) {
    CUDA_TRY(cudaMalloc((void**)&L.paths[k].A, n * 16));
    CUDA_TRY(cudaMalloc((void**)&L.paths[k].B, n * 16));
    CUDA_TRY(cudaMalloc((void**)&L.paths[k].C, n * 16));
  }
  CUDA_TRY(cudaMalloc((void**)&L.hits.H, n * 16));
  CUDA_TRY(cudaMalloc((void**)&L.hits.obj, n * 4));
  CUDA_TRY(cudaMalloc((void**)&L.queues, n * 4 * RT_NUM_CLASSES));
  CUDA_TRY(cudaMalloc((void**)&L.sort.keys, n * 8));  // key (+ rank inside the bin with RT_SORT_RANKED)
  CUDA_TRY(cudaMalloc((void**)&L.sort.order, n * 4));
  if (!L.sort.hist) {
    CUDA_TRY(cudaMalloc((void**)&L.sort.hist, RT_SORT_BINS * 4));
    CUDA_TRY(cudaMalloc((void**)&L.sort.cursor, RT_SORT_BINS * 4));
    CUDA_TRY(cudaMalloc((void**)&L.sort.slice_total, (RT_SORT_BINS / 1024u) * 4));
    CUDA_TRY(cudaMemset(L.sort.hist, 0, RT_SORT_BINS * 4));
  }
  L.capacity = capacity;
  return RT_OK;
}

int check_camera(const rt_camera* cam) {
  if (!cam) return fail(RT_ERR_INVALID, "camera is NULL");
  if (cam->projection_mode != RT_PROJ_PERSPECTIVE && cam->projection_mode != RT_PROJ_ORTHOGRAPHIC)
    return fail(RT_ERR_INVALID, "unknown CameraProjectionMode (tracing.rs:25-28)");
  if (cam->shading_mode != RT_SHADE_PATHTRACE && cam->shading_mode != RT_SHADE_PHONG)
    return fail(RT_ERR_INVALID, "unknown ShadingMode (tracing.rs:29-32)");
  if (cam->path_samples == 0) return fail(RT_ERR_INVALID, "path_samples must be >= 1 (tracing.rs:146)");
  if (cam->path_samples > 1 && cam->shading_mode == RT_SHADE_PATHTRACE) {
    // every hit spawns path_samples children (tracing.rs:308-319); the position of a path in that tree keys its
    // random numbers and must fit 32 bits
    double nodes = std::pow((double)cam->path_samples, (double)std::max(1u, cam->path_depth) - 1.0);
    if (nodes >= 4294967296.0) return fail(RT_ERR_UNSUPPORTED, "path_samples^(path_depth-1) must be below 2^32");
  }
  if (cam->screen_width == 0 || cam->screen_height == 0) return fail(RT_ERR_INVALID, "empty image");
  if ((uint64_t)cam->screen_width * cam->screen_height > (1ull << 31)) return fail(RT_ERR_INVALID, "image too large");
  if (cam->aa_sample_count == 0 || cam->aa_sample_count >= (1u << 24))
    return fail(RT_ERR_INVALID, "aa_sample_count must be in [1, 2^24)");
  if (cam->path_depth >= 256) return fail(RT_ERR_INVALID, "path_depth must be < 256");
  if ((uint32_t)std::sqrt((float)cam->aa_sample_count) == 0) return fail(RT_ERR_INVALID, "bad aa_sample_count");
  return RT_OK;
}

// camera constants exactly as Camera::generate_rays computes them per sample (tracing.rs:160-191)
void fill_frame_camera(const rt_camera& cam, uint64_t seed, rt_frame& fr) {
  std::memset(&fr, 0, sizeof fr);
  for (int k = 0; k < 3; ++k) fr.eye[k] = cam.eyepoint[k];
  const float* v = cam.view_dir;
  const float* u = cam.up;
  float cx = v[1] * u[2] - v[2] * u[1], cy = v[2] * u[0] - v[0] * u[2], cz = v[0] * u[1] - v[1] * u[0];
  float inv = 1.0f / std::sqrt(cx * cx + cy * cy + cz * cz);
  fr.rot0[0] = cx * inv; fr.rot0[1] = cy * inv; fr.rot0[2] = cz * inv;
  for (int k = 0; k < 3; ++k) {
    fr.rot1[k] = u[k];
    fr.rot2[k] = -v[k];
  }
  fr.pixel_size = 1.0f / (float)cam.screen_height;
  fr.n = (float)cam.aa_sample_count;
  fr.rootn = std::sqrt(fr.n);
  fr.rooti = (uint32_t)fr.rootn;
  fr.focal_length = cam.focal_length;
  fr.focus_dist = cam.focus_dist;
  fr.lens_radius = cam.lens_radius;
  fr.t_min = 0.001f;  // tracing.rs:305
  if (cam.shading_mode == RT_SHADE_PHONG) {
    fr.phong = 1;
    fr.t_min = 0.0f;  // tracing.rs:279,290
  }
  if (cam.projection_mode == RT_PROJ_ORTHOGRAPHIC) {
    fr.ortho = 1;
    for (int k = 0; k < 3; ++k) fr.view_dir[k] = cam.view_dir[k];
  }
  fr.t_max = cam.max_trace_dist;
  fr.spp = cam.aa_sample_count;
  fr.width = cam.screen_width;
  fr.height = cam.screen_height;
  fr.path_depth = cam.path_depth;
  fr.path_samples = cam.shading_mode == RT_SHADE_PHONG ? 1u : cam.path_samples;  // phong_shade_ray scatters once
  fr.k0 = (uint32_t)seed;
  fr.k1 = (uint32_t)(seed >> 32);
  fr.shard_mode = RT_SHARD_ALL;
  fr.shard_count = 1;
  fr.sample_begin = 0;
  fr.sample_count = cam.aa_sample_count;
}

// resolves the shard description into (pixel-slot count, sample range); returns work item count
// `tiles` (optional) receives the tile ids of a tile shard.
int plan_shard(const rt_camera& cam, const rt_render_opts& o, rt_frame& fr, unsigned long long& total,
               std::vector<uint32_t>* tiles = nullptr) {
  uint32_t spp = cam.aa_sample_count;
  uint32_t count = o.shard_count ? o.shard_count : 1;
  if (o.shard_rank >= count) return fail(RT_ERR_INVALID, "shard_rank >= shard_count");
  uint32_t sb = o.sample_begin, se = o.sample_end;
  if (sb == 0 && se == 0) se = spp;
  if (se > spp || sb > se) return fail(RT_ERR_INVALID, "bad sample range");
  unsigned long long npix = (unsigned long long)cam.screen_width * cam.screen_height;
  fr.shard_mode = o.shard_mode;
  fr.shard_rank = o.shard_rank;
  fr.shard_count = count;
  switch (o.shard_mode) {
    case RT_SHARD_ALL:
      break;
    case RT_SHARD_SAMPLES: {
      // contiguous sample ranges of [sb, se), as even as possible
      uint32_t n = se - sb, base = n / count, rem = n % count;
      uint32_t b = sb + o.shard_rank * base + std::min(o.shard_rank, rem);
      uint32_t e = b + base + (o.shard_rank < rem ? 1u : 0u);
      sb = b;
      se = e;
      fr.shard_mode = RT_SHARD_ALL;  // on the device this is just a sample range
      break;
    }
    case RT_SHARD_TILES: {
      uint32_t ts = o.tile_size ? o.tile_size : 64;
      if (ts > 1024) return fail(RT_ERR_INVALID, "tile_size too large");
      fr.tile_size = ts;
      fr.tiles_x = (cam.screen_width + ts - 1) / ts;
      fr.tiles_y = (cam.screen_height + ts - 1) / ts;
      // Which tiles: tile (tx, ty) belongs to rank (tx + stride * ty) mod count - diagonals, not columns.  (Round-robin
      // over the row-major index degenerates into columns whenever tiles_x is a multiple of count: at 1080p, 16-pixel
      // tiles and 8 ranks every rank owned the same eight-th of every row, vertical structures of the scene - the
      // columns of spheres, the drone's flanks - aliased with the 128-pixel period and the slowest rank took 11 % longer
      // than the fastest.)  stride is the smallest odd number >= 3 that is coprime to count.
      uint32_t stride = 3;
      auto gcd = [](uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; };
      while (gcd(stride, count) != 1) stride += 2;
      uint32_t mine = 0;
      for (uint32_t ty = 0; ty < fr.tiles_y; ++ty)
        for (uint32_t tx = 0; tx < fr.tiles_x; ++tx)
          if ((tx + stride * ty) % count == o.shard_rank) {
            ++mine;
            if (tiles) tiles->push_back(ty * fr.tiles_x + tx);
          }
      npix = (unsigned long long)mine * ts * ts;  // slots; those outside the image are skipped
      break;
    }
    default:
      return fail(RT_ERR_INVALID, "unknown shard_mode");
  }
  fr.sample_begin = sb;
  fr.sample_count = se - sb;
  fr.pixel_slots = npix ? npix : 1;
  // Work order.  0: pixel-major (32 consecutive samples of a pixel per warp, pixels in order) - best on whole frames
  // and sample-range shards.  2: the same warps, but consecutive warps walk the shard's pixels - a tile shard at high
  // spp otherwise keeps the whole wavefront inside half a tile, which piles the ray-sort keys into a few bins (C5 on
  // 1/8 of the tiles: 1399 -> 1645 Msamples/s).  1: sample-major (a warp = 32 neighbouring pixels), slower everywhere.
  switch (o.work_order) {
    // pixel-major keeps only capacity / spp pixels in flight; at very high spp that is a few thousand pixels, the
    // ray-sort keys pile into a few bins and the shade queues see one corner of the scene (C5, 4096 spp, full frame:
    // 1465 pixel-major vs 1688 Msamples/s grouped; C4, 1024 spp: 2586 vs 2567)
    case RT_ORDER_AUTO: fr.sample_major = (o.shard_mode == RT_SHARD_TILES || se - sb >= 2048u) ? 2u : 0u; break;
    case RT_ORDER_PIXEL_MAJOR: fr.sample_major = 0u; break;
    case RT_ORDER_SAMPLE_MAJOR: fr.sample_major = 1u; break;
    case RT_ORDER_GROUPED: fr.sample_major = 2u; break;
    default: return fail(RT_ERR_INVALID, "unknown work_order");
  }
  if (fr.sample_major == 2u && (fr.sample_count % 32u) != 0u) fr.sample_major = 0;  // needs whole groups of 32 samples
  total = fr.sample_count ? npix * fr.sample_count : 0;
  if (fr.sample_count == 0) fr.sample_count = 1;  // never divide by zero on the device
  return RT_OK;
}

// paths in flight.  Every k_trace launch ends with a tail in which the last, longest ray batches finish on a nearly
// empty machine (~125 us on C4, independent of the width), so wider is better: 1 M -> 1250, 2 M -> 1574, 4 M -> 1795,
// 8 M -> 1903 Msamples/s in round 1; on the final kernels 16 Mi -> 2684, 32 Mi -> 2732, 64 Mi -> 2735 Msamples/s on a
// whole C4 frame, 104.3 -> 100.3 ms on one eighth of it, C5 1763 -> 1796 (profiles/r2_notes.md C11).  148 bytes of
// state per path: 32 Mi paths are 5 GB of a B200's 180 GB.  If that cannot be allocated the default is halved until it
// can (a width the caller asked for is taken as given).
const uint32_t kDefaultWavefront = 1u << 25;

uint32_t pick_capacity(unsigned long long total, uint32_t requested) {
  unsigned long long cap = requested ? requested : kDefaultWavefront;
  cap = std::min<unsigned long long>(cap, std::max<unsigned long long>(total, 1));
  cap = (cap + 127) / 128 * 128;
  return (uint32_t)std::min<unsigned long long>(cap, 1ull << 27);
}

int set_device(rt_scene* s) {
  if (!s) return fail(RT_ERR_INVALID, "scene is NULL");
  if (!s->committed) return fail(RT_ERR_NOT_COMMITTED, "rt_commit has not been called on this scene");
  CUDA_TRY(cudaSetDevice(s->device));
  return RT_OK;
}

// what one engine run leaves in rt_stats
void stats_from_ctrl(const rt_ctrl& c, rt_stats* stats) {
  stats->samples += c.n_samples - c.counters[7];
  stats->rays += c.n_rays_total - c.counters[7] - c.counters[11];  // [11]: Phong shadow slots of missed camera rays
  stats->nodes_visited += c.counters[0];
  stats->tris_tested += c.counters[1];
  stats->instances_entered += c.counters[2];
  stats->prims_tested += c.counters[3];
  stats->mesh_hits += c.counters[4];
  stats->texel_taps += c.counters[5] + c.counters[8];
  stats->extend_texel_taps += c.counters[8];
  stats->material_fetches += c.counters[6];
  stats->warp_node_slots += c.counters[9];
  stats->tlas_nodes_visited += c.counters[10];
}

// The wavefront loop.  Launches are asynchronous; the device decides how many rays each iteration has.  The host
// only peeks at a `done` flag every few iterations, two polls deep, so the stream never drains.
int run_wavefront(rt_scene* s, const rt_frame& fr_in, unsigned long long total, long long* d_accum, bool count,
                  bool use_events, cudaStream_t st, uint32_t blocks_per_sm, bool capacity_is_default, rt_stats* stats) {
  int rc;
  rt_frame fr = fr_in;
  if (fr.phong) fr.sort_enabled = 0;  // camera ray + shadow ray pairs: slot i of the shadow pass belongs to slot i of the camera pass
  fr.capacity = std::max<uint32_t>(128u, (fr.capacity + 127u) / 128u * 128u);
  while ((rc = ensure_wavefront(s, fr.capacity)) != RT_OK) {
    if (!capacity_is_default || fr.capacity <= (1u << 20)) return rc;
    cudaGetLastError();  // out of memory for the default width: try half of it
    free_wavefront(s);
    fr.capacity = std::max<uint32_t>(128u, (fr.capacity / 2u + 127u) / 128u * 128u);
  }
  rt_scene::Wavefront& L = s->wf;
  L2Window l2;
  l2.open(s, st);
  const uint32_t grid = persistent_grid(s, s->trace_blocks_per_sm, blocks_per_sm);
  // Iterations are enqueued two at a time; after each pair the host asks for the control block's poll words and reads the
  // answer to the request before last, so two to four iterations are always queued ahead of the device and at most
  // that many empty ones run after the last ray has died (an "empty" iteration is not free: four of its kernels have
  // grids sized for the whole wavefront).  Once a poll shows the shard exhausted, the ray count it reports bounds
  // every later iteration (no new paths, survivors only), and the launch grids shrink to it.
  const int kChunk = 2;
  const size_t kMaxTimedIters = 1u << 14;
  size_t ev_used = 0, timed_iters = 0;
  uint64_t launches = 0, ext = 0, shd = 0;
  unsigned long long it = 0;
  auto next_event = [&](cudaEvent_t& ev) -> int {
    if (ev_used == L.events.size()) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreate(&e));
      L.events.push_back(e);
    }
    ev = L.events[ev_used++];
    return RT_OK;
  };
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  if ((rc = next_event(ev_begin)) != RT_OK || (rc = next_event(ev_end)) != RT_OK) return rc;
  const size_t ev_iter_base = ev_used;
  CUDA_TRY(cudaEventRecord(ev_begin, st));
  rt::launch_init(L.ctrl, 0ull, total, st);
  launches = 1;
  std::memset(L.h_done, 0, 2 * 4 * sizeof(uint32_t));
  bool done = false, exhausted = false;
  fr.grid_rays = fr.capacity;
  for (int chunk = 0; !done; ++chunk) {
    for (int k = 0; k < kChunk; ++k) {
      int cur = (int)(it & 1), nxt = cur ^ 1;
      rt::launch_advance(L.ctrl, fr.capacity, st);
      bool timed = use_events && timed_iters < kMaxTimedIters;
      cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
      if (!exhausted) rt::launch_raygen(fr, L.ctrl, L.paths[cur], st);  // an exhausted shard starts no more camera paths
      if (timed) {
        if ((rc = next_event(e0)) != RT_OK || (rc = next_event(e1)) != RT_OK || (rc = next_event(e2)) != RT_OK) return rc;
        CUDA_TRY(cudaEventRecord(e0, st));
      }
      rt::launch_trace(s->dev, fr, L.ctrl, L.paths[cur], L.hits, L.sort, count, grid, st);  // dominant kernel
      if (timed) CUDA_TRY(cudaEventRecord(e1, st));
      if (fr.phong) {
        // ShadingMode::Phong (tracing.rs:277-297).  k_phong_primary leaves n_next = n_rays, so the second
        // k_advance admits no new work (a batch is either full or the last one) and the shadow rays keep their slots.
        rt::launch_phong_primary(s->dev, fr, L.ctrl, L.paths[cur], L.paths[nxt], L.hits, st);
        rt::launch_advance(L.ctrl, fr.capacity, st);
        rt_frame fs = fr;
        fs.ray_tmax_from_c = 1;
        rt::launch_trace(s->dev, fs, L.ctrl, L.paths[nxt], L.hits, L.sort, count, grid, st);
        rt::launch_phong_shadow(s->dev, fr, L.ctrl, L.paths[nxt], L.hits, d_accum, st);
        if (timed) {
          CUDA_TRY(cudaEventRecord(e2, st));
          ++timed_iters;
        }
        launches += 7; ext += 2; shd += 2; it += 2;
        continue;
      }
      rt::launch_sort(s->dev, fr, L.ctrl, L.hits, L.queues, st);
      rt::launch_shade(s->dev, fr, L.ctrl, L.paths[cur], L.paths[nxt], L.hits, L.queues, d_accum, L.sort, count, st);
      rt::launch_raysort(fr, L.ctrl, L.sort, st);
      if (timed) {
        CUDA_TRY(cudaEventRecord(e2, st));
        ++timed_iters;
      }
      launches += (fr.sort_enabled ? 6 : 4) + (exhausted ? 0 : 1); ++ext; ++shd; ++it;
    }
    // poll: the words k_advance wrote, two chunks deep
    int slot = chunk & 1;
    uint32_t* h = L.h_done + 4 * slot;
    if (chunk >= 2) {
      CUDA_TRY(cudaEventSynchronize(L.poll_ev[slot]));
      if (h[0]) done = true;
      if (h[1] && !fr.phong) {  // (Phong's second k_advance per iteration reports the shadow pass, whose slots do not shrink)
        exhausted = true;
        fr.grid_rays = std::min(fr.grid_rays, std::max(h[2], 1u));
      }
    }
    if (!done) {
      CUDA_TRY(cudaMemcpyAsync(h, L.ctrl->poll, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(L.poll_ev[slot], st));
    }
  }
  CUDA_TRY(cudaMemcpyAsync(L.h_ctrl, L.ctrl, sizeof(rt_ctrl), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(ev_end, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  const rt_ctrl& c = *L.h_ctrl;
  if (!c.done && c.cursor != c.total) return fail(RT_ERR_CUDA, "wavefront loop ended before all work was issued");
  if (stats) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    stats->ms_total += ms;
    stats_from_ctrl(c, stats);
    stats->engine = RT_ENGINE_WAVEFRONT;
    stats->iterations += c.iterations;
    stats->kernel_launches += launches;
    // report the launches that did work, not the no-op tail queued behind the `done` poll
    stats->extend_launches += std::min<uint64_t>(ext, c.iterations);
    stats->shade_launches += std::min<uint64_t>(shd, c.iterations);
    // only iterations that actually had rays count as launches of the dominant kernel
    double me = 0.0, msd = 0.0;
    size_t live_it = std::min<size_t>(timed_iters, c.iterations);
    for (size_t i = 0; i < live_it; ++i) {
      float a = 0.0f, b = 0.0f;
      cudaEventElapsedTime(&a, L.events[ev_iter_base + 3 * i], L.events[ev_iter_base + 3 * i + 1]);
      cudaEventElapsedTime(&b, L.events[ev_iter_base + 3 * i + 1], L.events[ev_iter_base + 3 * i + 2]);
      me += a;
      msd += b;
    }
    if (live_it && live_it < c.iterations) {  // more iterations than event slots: scale up
      double f = (double)c.iterations / (double)live_it;
      me *= f;
      msd *= f;
    }
    stats->ms_extend += me;
    stats->ms_shade += msd;
  }
  return RT_OK;
}

// The megakernel engine: one launch of k_path renders the whole shard.  The kernel hands out work indices itself
// (warps claim chunks of ctrl->cursor), so there is nothing for the host to do between the launch and the end.
int run_megakernel(rt_scene* s, const rt_frame& fr_in, unsigned long long total, long long* d_accum, cudaStream_t st,
                   uint32_t blocks_per_sm, rt_stats* stats) {
  int rc;
  if ((rc = ensure_runtime(s)) != RT_OK) return rc;
  rt_frame fr = fr_in;
  fr.sort_enabled = 0;
  fr.sample_major = 0;  // 32 consecutive work indices = 32 samples of one pixel: a freshly filled warp runs in lockstep
  rt_scene::Wavefront& L = s->wf;
  EventPair ev;
  if ((rc = ev.create()) != RT_OK) return rc;
  L2Window l2;
  l2.open(s, st);
  CUDA_TRY(cudaEventRecord(ev.a, st));
  rt::launch_init(L.ctrl, 0ull, total, st);
  rt::launch_path(s->dev, fr, L.ctrl, d_accum, total, persistent_grid(s, s->path_blocks_per_sm, blocks_per_sm), st);
  CUDA_TRY(cudaMemcpyAsync(L.h_ctrl, L.ctrl, sizeof(rt_ctrl), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(ev.b, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  const rt_ctrl& c = *L.h_ctrl;
  if (c.cursor < c.total) return fail(RT_ERR_CUDA, "megakernel ended before all work was claimed");
  if (c.n_samples != total) return fail(RT_ERR_CUDA, "megakernel: work items started != work items in the shard");
  if (stats) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev.a, ev.b);
    stats->ms_total += ms;
    stats->ms_extend += ms;  // closest hit and shading are one kernel here
    stats->engine = RT_ENGINE_MEGAKERNEL;
    stats_from_ctrl(c, stats);
    stats->iterations += 1;
    stats->kernel_launches += 2;
    stats->extend_launches += 1;
  }
  return RT_OK;
}

// Camera::path_samples > 1 (tracing.rs:308-319): every hit spawns path_samples scattered rays, so a camera sample is
// a tree of up to path_samples^(path_depth-1) paths.  The tree is walked depth first so that memory stays bounded:
// level d holds the pending rays of bounce d; an iteration pops up to P rays from the deepest non-empty level, traces
// them and runs k_shade once per child index, appending the children to level d+1 (which is empty at that moment and
// can hold path_samples * P rays).  The host reads the child count back every iteration - this mode is exponential
// in path_depth anyway and is not a bench path.
int run_branching(rt_scene* s, const rt_frame& fr_in, unsigned long long total, long long* d_accum, cudaStream_t st,
                  rt_stats* stats) {
  int rc;
  rt_frame fr = fr_in;
  fr.sort_enabled = 0;
  const uint64_t S = fr.path_samples, D = std::max(1u, fr.path_depth);
  uint64_t P = std::min<uint64_t>(fr.capacity, 1u << 20);
  while ((P + (D - 1) * S * P) * 48ull > (4ull << 30) && P > 4096) P /= 2;
  P = std::max<uint64_t>(128, P / 128 * 128);
  fr.capacity = (uint32_t)P;
  if ((rc = ensure_wavefront(s, fr.capacity)) != RT_OK) return rc;
  const uint64_t slots = P + (D - 1) * S * P;
  for (auto& b : s->d_tree)
    if ((rc = ensure_buf(b, slots * 16)) != RT_OK) return rc;
  auto level_base = [&](uint64_t d) { return d == 0 ? 0ull : P + (d - 1) * S * P; };
  auto at = [&](uint64_t off) {
    return rt::rt_paths{(float4*)s->d_tree[0].p + off, (float4*)s->d_tree[1].p + off, (float4*)s->d_tree[2].p + off};
  };
  rt_scene::Wavefront& L = s->wf;
  const uint32_t grid = persistent_grid(s, s->trace_blocks_per_sm, 0);
  EventPair ev;
  if ((rc = ev.create()) != RT_OK) return rc;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  CUDA_TRY(cudaEventRecord(e0, st));
  rt::launch_init(L.ctrl, 0ull, total, st);
  std::vector<uint64_t> pending(D, 0);
  unsigned long long cursor = 0;
  uint64_t launches = 1, iters = 0;
  for (;;) {
    int d = (int)D - 1;
    while (d >= 0 && pending[d] == 0) --d;
    uint32_t n_cont = 0, n_new = 0;
    rt::rt_paths cur;
    if (d < 0) {
      if (cursor >= total) break;
      d = 0;
      n_new = (uint32_t)std::min<unsigned long long>(P, total - cursor);
      cursor += n_new;
      cur = at(level_base(0));
    } else {
      n_cont = (uint32_t)std::min<uint64_t>(pending[d], P);
      pending[d] -= n_cont;
      cur = at(level_base(d) + pending[d]);
    }
    rt::rt_paths nxt = at((uint64_t)d + 1 < D ? level_base(d + 1) : 0);  // the last level has no children
    rt::launch_set_window(L.ctrl, n_cont, n_new, st);
    if (n_new) rt::launch_raygen(fr, L.ctrl, cur, st);
    rt::launch_trace(s->dev, fr, L.ctrl, cur, L.hits, L.sort, false, grid, st);
    rt::launch_sort(s->dev, fr, L.ctrl, L.hits, L.queues, st);
    for (uint32_t b = 0; b < (uint32_t)S; ++b) {
      fr.branch = b;
      rt::launch_shade(s->dev, fr, L.ctrl, cur, nxt, L.hits, L.queues, d_accum, L.sort, false, st);
    }
    fr.branch = 0;
    uint32_t n_next = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_next, &L.ctrl->n_next, sizeof n_next, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint64_t)d + 1 < D) pending[d + 1] += n_next;
    else if (n_next) return fail(RT_ERR_CUDA, "branching walk: rays survived the last level");
    launches += 4 + S + (n_new ? 1 : 0);
    ++iters;
  }
  CUDA_TRY(cudaMemcpyAsync(L.h_ctrl, L.ctrl, sizeof(rt_ctrl), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(e1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (stats) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const rt_ctrl& c = *L.h_ctrl;
    stats->ms_total += ms;
    stats->samples += c.n_samples - c.counters[7];
    stats->rays += c.n_rays_total - c.counters[7];
    stats->engine = RT_ENGINE_WAVEFRONT;
    stats->iterations += iters;
    stats->kernel_launches += launches;
    stats->extend_launches += iters;
    stats->shade_launches += iters * S;
  }
  return RT_OK;
}

// Last row of a column-major 4x4.  The forward transform must be affine exactly; a caller-supplied inverse (cgmath's
// general Matrix4::invert, geometry.rs:168) comes out as (0,0,0,1) only up to rounding - w.w is minor3x3 * (1/det4) -
// so a few ulps are accepted there and the row is then set to (0,0,0,1): the reference's divide by w (= 1 +- ulp)
// is dropped, see DESIGN.md.
bool affine_row(const float* m, bool exact) {
  const float tol = exact ? 0.0f : 8.0f * 1.1920929e-7f;
  return std::fabs(m[3]) <= tol && std::fabs(m[7]) <= tol && std::fabs(m[11]) <= tol && std::fabs(m[15] - 1.0f) <= tol;
}
void set_affine_row(float* m) {
  m[3] = m[7] = m[11] = 0.0f;
  m[15] = 1.0f;
}

// RT_ENGINE_AUTO.  Measured on B200 (profiles/r2_notes.md C1/C3), megakernel vs wavefront: C1 +64 %,
// C2 +8 %, C3 -7 %, C4 -47 %, C5 -38 %.  The megakernel wins where there is little to traverse and a wavefront
// iteration is mostly path state moving through HBM; as soon as the rays of a warp diverge in the BVH (big instanced
// meshes) or in the material code (many material classes in one scene), warps of sorted wavefront batches beat warps
// of unrelated paths.
uint32_t pick_engine(bool phong, bool counters, unsigned long long inst_tris, uint32_t material_classes) {
  if (phong || counters) return RT_ENGINE_WAVEFRONT;
  return (inst_tris < 16384ull && material_classes <= 2) ? RT_ENGINE_MEGAKERNEL : RT_ENGINE_WAVEFRONT;
}

}  // namespace

// Nothing may throw across the C ABI: std::bad_alloc / length_error from a huge or hostile asset becomes an error code.
#define RT_CATCH(who)                                                                                   \
  catch (const std::bad_alloc&) { return fail(RT_ERR_IO, who ": out of memory"); }                      \
  catch (const std::exception& e) { return fail(RT_ERR_INVALID, std::string(who ": ") + e.what()); }    \
  catch (...) { return fail(RT_ERR_INVALID, who ": unknown exception"); }

// ====================================================================== C ABI
extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }
const char* rt_last_error(void) { return g_err.c_str(); }
int rt_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  return n;
}

int rt_scene_create(rt_scene** out) {
  if (!out) return fail(RT_ERR_INVALID, "out is NULL");
  *out = new (std::nothrow) rt_scene();
  return *out ? RT_OK : fail(RT_ERR_INVALID, "out of memory");
}

void rt_scene_destroy(rt_scene* s) {
  if (!s) return;
  if (s->device >= 0 && cudaSetDevice(s->device) == cudaSuccess) {
    unpin_lowered(s);
    free_wavefront(s);
    free_buf(s->d_geom); free_buf(s->d_shade); free_buf(s->d_objects);
    free_buf(s->d_mats); free_buf(s->d_textures); free_buf(s->d_texels); free_buf(s->d_planes); free_buf(s->d_guards); free_buf(s->d_guard_list);
    free_buf(s->d_accum); free_buf(s->d_linear); free_buf(s->d_rgb8); free_buf(s->d_dbg);
    for (auto& b : s->d_tree) free_buf(b);
    free_buf(s->d_tiles);
    {
      rt_scene::Wavefront& L = s->wf;
      if (L.ctrl) cudaFree(L.ctrl);
      if (L.sort.hist) cudaFree(L.sort.hist);
      if (L.sort.cursor) cudaFree(L.sort.cursor);
      if (L.sort.slice_total) cudaFree(L.sort.slice_total);
      if (L.h_ctrl) cudaFreeHost(L.h_ctrl);
      if (L.h_done) cudaFreeHost(L.h_done);
      for (auto e : L.events) cudaEventDestroy(e);
      for (auto e : L.poll_ev) if (e) cudaEventDestroy(e);
    }
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
  }
  delete s;
}

int rt_add_texture(rt_scene* s, const uint8_t* rgb8, uint32_t w, uint32_t h) try {
  if (!s || !rgb8 || !w || !h) return fail(RT_ERR_INVALID, "rt_add_texture: bad argument");
  rt::HostTexture t;
  t.w = w;
  t.h = h;
  t.rgba.resize((size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; ++i)
    t.rgba[i] = (uint32_t)rgb8[3 * i] | ((uint32_t)rgb8[3 * i + 1] << 8) | ((uint32_t)rgb8[3 * i + 2] << 16) | 0xFF000000u;
  s->textures.push_back(std::move(t));
  s->lowered = false;
  return (int)s->textures.size() - 1;
}
RT_CATCH("rt_add_texture")

int rt_add_material(rt_scene* s, const rt_material_desc* d) try {
  if (!s || !d) return fail(RT_ERR_INVALID, "rt_add_material: bad argument");
  if (d->tag > RT_MAT_ISOTROPIC) return fail(RT_ERR_INVALID, "rt_add_material: unknown material tag");
  s->materials.push_back(*d);
  s->lowered = false;
  return (int)s->materials.size() - 1;
}
RT_CATCH("rt_add_material")

int rt_add_mesh(rt_scene* s, const float* pos, const float* nrm, const float* uv, uint32_t nverts, const uint32_t* idx,
                uint32_t ntris) try {
  if (!s || !pos || !nrm || !uv || !idx || !nverts || !ntris) return fail(RT_ERR_INVALID, "rt_add_mesh: bad argument");
  for (size_t i = 0; i < 3 * (size_t)ntris; ++i)
    if (idx[i] >= nverts) return fail(RT_ERR_INVALID, "rt_add_mesh: index out of range");
  rt::HostMesh m;
  m.pos.assign(pos, pos + 3 * (size_t)nverts);
  m.nrm.assign(nrm, nrm + 3 * (size_t)nverts);
  m.uv.assign(uv, uv + 2 * (size_t)nverts);
  m.idx.assign(idx, idx + 3 * (size_t)ntris);
  rt::build_mesh(m);
  s->meshes.push_back(std::move(m));
  s->lowered = false;
  return (int)s->meshes.size() - 1;
}
RT_CATCH("rt_add_mesh")

int rt_add_instance(rt_scene* s, int mesh, const float xform[16], const float* inv_xform, int material, const int tex[5]) try {
  if (!s || !xform) return fail(RT_ERR_INVALID, "rt_add_instance: bad argument");
  if (mesh < 0 || mesh >= (int)s->meshes.size()) return fail(RT_ERR_INVALID, "rt_add_instance: bad mesh id");
  if (material >= (int)s->materials.size()) return fail(RT_ERR_INVALID, "rt_add_instance: bad material id");
  if (!affine_row(xform, true) || (inv_xform && !affine_row(inv_xform, false)))
    return fail(RT_ERR_UNSUPPORTED, "rt_add_instance: only affine transforms (last row 0 0 0 1) are supported");
  rt::HostObject o;
  o.kind = RT_OBJ_MESH;
  o.mesh = mesh;
  std::memcpy(o.xform, xform, 64);
  if (inv_xform) {
    std::memcpy(o.inv_xform, inv_xform, 64);
    set_affine_row(o.inv_xform);
  } else if (!rt::invert_affine_cofactor(xform, o.inv_xform)) {
    return fail(RT_ERR_INVALID, "rt_add_instance: transform is singular (the reference panics here, geometry.rs:168)");
  }
  o.material = material < 0 ? -1 : material;
  for (int k = 0; k < 5; ++k) {
    o.tex[k] = tex ? tex[k] : -1;
    if (o.tex[k] >= (int)s->textures.size()) return fail(RT_ERR_INVALID, "rt_add_instance: bad texture id");
    if (o.tex[k] < 0) o.tex[k] = -1;
  }
  s->objects.push_back(o);
  s->lowered = false;
  return (int)s->objects.size() - 1;
}
RT_CATCH("rt_add_instance")

static int add_simple(rt_scene* s, rt::HostObject& o, int material, const char* who) {
  if (material < 0 || material >= (int)s->materials.size()) return fail(RT_ERR_INVALID, std::string(who) + ": bad material id");
  o.material = material;
  s->objects.push_back(o);
  s->lowered = false;
  return (int)s->objects.size() - 1;
}
int rt_add_sphere(rt_scene* s, const float c[3], float radius, int material) try {
  if (!s || !c) return fail(RT_ERR_INVALID, "rt_add_sphere: bad argument");
  rt::HostObject o;
  o.kind = RT_OBJ_SPHERE;
  std::memcpy(o.a, c, 12);
  o.radius = radius;
  return add_simple(s, o, material, "rt_add_sphere");
}
RT_CATCH("rt_add_sphere")
int rt_add_triangle(rt_scene* s, const float a[3], const float b[3], const float c[3], int material) try {
  if (!s || !a || !b || !c) return fail(RT_ERR_INVALID, "rt_add_triangle: bad argument");
  rt::HostObject o;
  o.kind = RT_OBJ_TRIANGLE;
  std::memcpy(o.a, a, 12);
  std::memcpy(o.b, b, 12);
  std::memcpy(o.c, c, 12);
  return add_simple(s, o, material, "rt_add_triangle");
}
RT_CATCH("rt_add_triangle")
int rt_add_plane(rt_scene* s, const float p[3], const float n[3], int material) try {
  if (!s || !p || !n) return fail(RT_ERR_INVALID, "rt_add_plane: bad argument");
  rt::HostObject o;
  o.kind = RT_OBJ_PLANE;
  std::memcpy(o.a, p, 12);
  std::memcpy(o.b, n, 12);
  return add_simple(s, o, material, "rt_add_plane");
}
RT_CATCH("rt_add_plane")
int rt_add_volume_sphere(rt_scene* s, const float c[3], float radius, float density, int phase_material) try {
  if (!s || !c) return fail(RT_ERR_INVALID, "rt_add_volume_sphere: bad argument");
  rt::HostObject o;
  o.kind = RT_OBJ_VOLUME;
  std::memcpy(o.a, c, 12);
  o.radius = radius;
  o.density = density;
  o.vol_index = s->n_volumes;
  int rc = add_simple(s, o, phase_material, "rt_add_volume_sphere");
  if (rc >= 0) s->n_volumes++;
  return rc;
}
RT_CATCH("rt_add_volume_sphere")

int rt_add_volume_mesh(rt_scene* s, int mesh, const float xform[16], const float* inv_xform, float density,
                       int phase_material) try {
  if (!s || !xform) return fail(RT_ERR_INVALID, "rt_add_volume_mesh: bad argument");
  if (mesh < 0 || mesh >= (int)s->meshes.size()) return fail(RT_ERR_INVALID, "rt_add_volume_mesh: bad mesh id");
  if (!affine_row(xform, true) || (inv_xform && !affine_row(inv_xform, false)))
    return fail(RT_ERR_UNSUPPORTED, "rt_add_volume_mesh: only affine transforms (last row 0 0 0 1) are supported");
  rt::HostObject o;
  o.kind = RT_OBJ_VOLUME_MESH;
  o.mesh = mesh;
  std::memcpy(o.xform, xform, 64);
  if (inv_xform) {
    std::memcpy(o.inv_xform, inv_xform, 64);
    set_affine_row(o.inv_xform);
  } else if (!rt::invert_affine_cofactor(xform, o.inv_xform)) return fail(RT_ERR_INVALID, "rt_add_volume_mesh: transform is singular");
  o.density = density;
  o.vol_index = s->n_volumes;
  int rc = add_simple(s, o, phase_material, "rt_add_volume_mesh");
  if (rc >= 0) s->n_volumes++;
  return rc;
}
RT_CATCH("rt_add_volume_mesh")

int rt_scene_upload(rt_scene* s) try {
  if (!s || !s->lowered) return fail(RT_ERR_NOT_COMMITTED, "rt_scene_upload: scene has not been lowered (call rt_commit)");
  CUDA_TRY(cudaSetDevice(s->device));
  const rt::Lowered& L = s->low;
  int rc;
  cudaStream_t st = 0;
  pin_lowered(s);
  const size_t node_bytes = (L.nodes.size() * 16 + 255) / 256 * 256, tri_bytes = L.tris.size() * 16;
  if ((rc = ensure_buf(s->d_geom, node_bytes + tri_bytes)) != RT_OK) return rc;
  s->geom_bytes = node_bytes + tri_bytes;
  if (!L.nodes.empty()) CUDA_TRY(cudaMemcpyAsync(s->d_geom.p, L.nodes.data(), L.nodes.size() * 16, cudaMemcpyHostToDevice, st));
  if (tri_bytes) CUDA_TRY(cudaMemcpyAsync((char*)s->d_geom.p + node_bytes, L.tris.data(), tri_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = upload(s->d_shade, L.shade.data(), L.shade.size() * 16, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_objects, L.objects.data(), L.objects.size() * 16, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_mats, L.mats.data(), L.mats.size() * 16, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_textures, L.textures.data(), L.textures.size() * 16, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_texels, L.texels.data(), L.texels.size() * 4, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_planes, L.planes.data(), L.planes.size() * 4, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_guards, L.guards.data(), L.guards.size() * 16, st)) != RT_OK) return rc;
  if ((rc = upload(s->d_guard_list, L.guard_list.data(), L.guard_list.size() * 4, st)) != RT_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  rt_dev_scene& d = s->dev;
  d.nodes = s->d_geom.p; d.tris = (char*)s->d_geom.p + node_bytes; d.shade = s->d_shade.p; d.objects = s->d_objects.p;
  d.mats = s->d_mats.p; d.textures = s->d_textures.p; d.texels = s->d_texels.p; d.planes = s->d_planes.p; d.guards = s->d_guards.p; d.guard_list = s->d_guard_list.p;
  d.tlas_root = L.tlas_root;
  d.tlas_base = L.tlas_base;
  d.tlas_count = L.tlas_count & ~3u;  // whole child groups (pairs, or fours with RT_BVH4)
  d.n_planes = (L.planes.size() == 1 && L.planes[0] < 0) ? 0u : (uint32_t)L.planes.size();
  d.n_objects = (uint32_t)s->objects.size();
  d.n_volumes = L.n_volumes;
  d.n_volume_meshes = 0;
  for (const auto& o : s->objects) d.n_volume_meshes += o.kind == RT_OBJ_VOLUME_MESH ? 1u : 0u;
  for (int k = 0; k < 3; ++k) {
    d.tlas_min[k] = L.tlas_min[k];
    d.tlas_max[k] = L.tlas_max[k];
  }
  s->committed = true;
  return RT_OK;
}
RT_CATCH("rt_scene_upload")

int rt_scene_lower(rt_scene* s, rt_lower_info* info) try {
  if (!s) return fail(RT_ERR_INVALID, "scene is NULL");
  std::string err;
  if (!s->pinned.empty()) unpin_lowered(s);  // the arrays are about to be rebuilt
  s->pin_tried = false;
  int rc = rt::lower_scene(s->textures, s->materials, s->meshes, s->objects, s->low, err);
  if (rc != RT_OK) return fail(rc, err);
  s->lowered = true;
  if (info) {
    const rt::Lowered& L = s->low;
    info->bytes = L.bytes();
    info->nodes = (uint32_t)(L.nodes.size() / RT_NODE_QUADS);
    info->tris = 0;
    info->max_blas_depth = 0;
    for (const auto& m : s->meshes) {
      info->tris += m.n_reachable;
      info->max_blas_depth = std::max(info->max_blas_depth, m.depth);
    }
    info->objects = (uint32_t)s->objects.size();
    info->unbounded = (L.planes.size() == 1 && L.planes[0] < 0) ? 0u : (uint32_t)L.planes.size();
    info->tlas_depth = L.tlas_depth;
    info->guard_boxes = 0;
    info->guarded_tris = 0;
    for (const auto& m : s->meshes) {
      info->guard_boxes += (uint32_t)(m.guards.size() / 2);
      for (size_t q = 0; q < m.tris.size(); q += RT_TRI_QUADS) info->guarded_tris += m.tris[q + 2].u[3] ? 1u : 0u;
    }
  }
  return RT_OK;
}
RT_CATCH("rt_scene_lower")

int rt_commit(rt_scene* s, int device) try {
  if (!s) return fail(RT_ERR_INVALID, "scene is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(RT_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= n) return fail(RT_ERR_INVALID, "rt_commit: bad device index");
  if (s->device >= 0 && s->device != device) return fail(RT_ERR_INVALID, "rt_commit: scene is already bound to another device");
  int rc = rt_scene_lower(s, nullptr);
  if (rc != RT_OK) return rc;
  s->device = device;
  return rt_scene_upload(s);
}
RT_CATCH("rt_commit")

uint64_t rt_scene_device_bytes(const rt_scene* s) { return s && s->lowered ? s->low.bytes() : 0; }

size_t rt_accum_bytes(uint32_t width, uint32_t height) { return (size_t)width * height * 4 * sizeof(long long); }

int rt_render_accum(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts, void* d_accum, void* stream,
                    rt_stats* stats) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if ((rc = check_camera(cam)) != RT_OK) return rc;
  if (!d_accum) return fail(RT_ERR_INVALID, "d_accum is NULL");
  rt_render_opts o{};
  if (opts) o = *opts;
  rt_frame fr;
  fill_frame_camera(*cam, o.seed, fr);
  for (int k = 0; k < 3; ++k) {
    fr.light[k] = o.point_light_pos[k];
    fr.ambient[k] = o.ambient[k];
  }
  unsigned long long total = 0;
  s->h_tiles.clear();
  if ((rc = plan_shard(*cam, o, fr, total, &s->h_tiles)) != RT_OK) return rc;
  fr.capacity = pick_capacity(total, o.wavefront);
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (total == 0) return RT_OK;
  // path_depth == 0: shade_ray returns the (black) background before it looks for a hit (tracing.rs:301-303)
  if (cam->shading_mode == RT_SHADE_PATHTRACE && cam->path_depth == 0) return RT_OK;
  if ((rc = ensure_runtime(s)) != RT_OK) return rc;
  // NULL means the (legacy) default stream, as documented: work must be ordered after whatever the caller has
  // already enqueued there (e.g. the memset of d_accum), which a private non-blocking stream would not be
  cudaStream_t st = (cudaStream_t)stream;
  if (fr.shard_mode == RT_SHARD_TILES) {  // the shard's tile list; h_tiles lives until the engines have synchronised `st`
    if ((rc = ensure_buf(s->d_tiles, s->h_tiles.size() * 4)) != RT_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->d_tiles.p, s->h_tiles.data(), s->h_tiles.size() * 4, cudaMemcpyHostToDevice, st));
    fr.tile_list = (const uint32_t*)s->d_tiles.p;
  }
  for (uint32_t r : o.reserved)
    if (r) return fail(RT_ERR_INVALID, "rt_render_opts.reserved must be zero");
  if (o.engine > RT_ENGINE_MEGAKERNEL) return fail(RT_ERR_INVALID, "unknown engine");
  if (o.ray_sort > RT_RAYSORT_ON) return fail(RT_ERR_INVALID, "unknown ray_sort");
  unsigned long long inst_tris = 0;
  for (const auto& ob : s->objects)
    if (ob.kind == RT_OBJ_MESH) inst_tris += s->meshes[ob.mesh].n_reachable;
  // Secondary-ray sorting (k_shade keys -> k_raysort_*), wavefront engine.  It pays where traversal dominates: on C4
  // k_trace drops from 1230 to 990 us per 8 Mi rays for ~150 us of sorting (+9 %); on C2 (two 240-triangle teapots in a
  // closed box, 5 nodes per ray) the indirection costs more than it saves (-14 %).  Hence the switch on the amount of
  // instanced geometry.
  {
    bool worth = o.ray_sort == RT_RAYSORT_ON || (o.ray_sort == RT_RAYSORT_AUTO && inst_tris >= 16384ull);
    fr.sort_enabled = (worth && s->low.tlas_root != RT_ENTRY_NONE) ? 1u : 0u;
#ifndef RT_SORT_CELLS
#define RT_SORT_CELLS 32
#endif
    const int cells = RT_SORT_CELLS;    // cells per axis of the TLAS box (measured: 8 / 16 / 32 / 64 -> 2253 / 2314 / 2348 / 2284 Msamples/s)
    fr.sort_cells_m1 = (float)(cells - 1);
    fr.sort_use_octant = 2;  // direction class = octant + dominant axis (measured best of: none, octant, octant + axis)
    for (int k = 0; k < 3; ++k) {
      float ext = s->low.tlas_max[k] - s->low.tlas_min[k];
      fr.sort_min[k] = s->low.tlas_min[k];
      fr.sort_scale[k] = ext > 0.0f ? (float)cells / ext : 0.0f;
    }
  }
  const bool counters = (o.flags & RT_OPT_COUNTERS) != 0;
  if (fr.path_samples > 1) {
    if (o.engine == RT_ENGINE_MEGAKERNEL) return fail(RT_ERR_UNSUPPORTED, "path_samples > 1 runs on the wavefront engine only");
    return run_branching(s, fr, total, (long long*)d_accum, st, stats);
  }
  // Engine.  The debug modes and the device counters exist in the wavefront engine only.
  uint32_t engine = o.engine;
  if (engine == RT_ENGINE_MEGAKERNEL && (fr.phong || counters))
    return fail(RT_ERR_UNSUPPORTED, "ShadingMode::Phong and RT_OPT_COUNTERS run on the wavefront engine only");
  if (engine == RT_ENGINE_AUTO) {
    uint32_t class_mask = 0;  // material classes that can be hit in this scene
    for (const auto& ob : s->objects)
      class_mask |= 1u << (ob.material >= 0 ? s->materials[ob.material].tag : (uint32_t)RT_CLASS_PARAM_TEX);
    uint32_t classes = 0;
    for (uint32_t m = class_mask; m; m &= m - 1) ++classes;
    engine = pick_engine(fr.phong != 0, counters, inst_tris, classes);
  }
  if (engine == RT_ENGINE_MEGAKERNEL) return run_megakernel(s, fr, total, (long long*)d_accum, st, o.blocks_per_sm, stats);
  return run_wavefront(s, fr, total, (long long*)d_accum, counters, (o.flags & RT_OPT_NO_EVENTS) == 0, st, o.blocks_per_sm,
                       o.wavefront == 0, stats);
}
RT_CATCH("rt_render_accum")

int rt_resolve(rt_scene* s, const rt_camera* cam, const void* d_accum, uint32_t total_spp, float* d_out_linear,
               uint8_t* d_out_rgb8, void* stream) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if (!cam || !d_accum || !total_spp) return fail(RT_ERR_INVALID, "rt_resolve: bad argument");
  cudaStream_t st = stream ? (cudaStream_t)stream : 0;
  rt::launch_resolve((const long long*)d_accum, cam->screen_width * cam->screen_height, total_spp, cam->gamma,
                     d_out_linear, d_out_rgb8, st);
  CUDA_TRY(cudaGetLastError());
  return RT_OK;
}
RT_CATCH("rt_resolve")

int rt_render(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts, float* out_linear, uint8_t* out_rgb8,
              rt_stats* stats) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if ((rc = check_camera(cam)) != RT_OK) return rc;
  size_t npix = (size_t)cam->screen_width * cam->screen_height;
  if ((rc = ensure_buf(s->d_accum, rt_accum_bytes(cam->screen_width, cam->screen_height))) != RT_OK) return rc;
  if ((rc = ensure_runtime(s)) != RT_OK) return rc;
  cudaStream_t st = s->own_stream;
  CUDA_TRY(cudaMemsetAsync(s->d_accum.p, 0, rt_accum_bytes(cam->screen_width, cam->screen_height), st));
  rt_stats local{};
  if ((rc = rt_render_accum(s, cam, opts, s->d_accum.p, st, &local)) != RT_OK) return rc;
  // how many samples per pixel ended up in the accumulator
  rt_render_opts o{};
  if (opts) o = *opts;
  rt_frame fr;
  fill_frame_camera(*cam, o.seed, fr);
  unsigned long long total = 0;
  if ((rc = plan_shard(*cam, o, fr, total)) != RT_OK) return rc;
  uint32_t spp = total ? fr.sample_count : 1;
  EventPair ev;
  if ((rc = ev.create()) != RT_OK) return rc;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  if (out_linear && (rc = ensure_buf(s->d_linear, npix * 12)) != RT_OK) return rc;
  if (out_rgb8 && (rc = ensure_buf(s->d_rgb8, npix * 3)) != RT_OK) return rc;
  CUDA_TRY(cudaEventRecord(e0, st));
  rt::launch_resolve((const long long*)s->d_accum.p, (uint32_t)npix, spp, cam->gamma,
                     out_linear ? (float*)s->d_linear.p : nullptr, out_rgb8 ? (uint8_t*)s->d_rgb8.p : nullptr, st);
  CUDA_TRY(cudaEventRecord(e1, st));
  local.kernel_launches += 1;
  if (out_linear) {
    CUDA_TRY(cudaMemcpyAsync(out_linear, s->d_linear.p, npix * 12, cudaMemcpyDeviceToHost, st));
    local.d2h_bytes += npix * 12;
  }
  if (out_rgb8) {
    CUDA_TRY(cudaMemcpyAsync(out_rgb8, s->d_rgb8.p, npix * 3, cudaMemcpyDeviceToHost, st));
    local.d2h_bytes += npix * 3;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  local.ms_resolve = ms;
  local.h2d_bytes += sizeof(rt_frame) + sizeof(rt_dev_scene);  // kernel parameters
  local.d2h_bytes += sizeof(rt_ctrl);
  if (stats) *stats = local;
  return RT_OK;
}
RT_CATCH("rt_render")

int rt_render_progressive(rt_scene* s, const rt_camera* cam, const rt_render_opts* opts, int64_t* accum,
                          uint32_t spp_in_accum, float* out_linear, uint8_t* out_rgb8, rt_stats* stats) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if ((rc = check_camera(cam)) != RT_OK) return rc;
  if (!accum) return fail(RT_ERR_INVALID, "rt_render_progressive: accum is NULL");
  if ((out_linear || out_rgb8) && !spp_in_accum) return fail(RT_ERR_INVALID, "rt_render_progressive: spp_in_accum is 0");
  const size_t npix = (size_t)cam->screen_width * cam->screen_height, bytes = rt_accum_bytes(cam->screen_width, cam->screen_height);
  if ((rc = ensure_buf(s->d_accum, bytes)) != RT_OK) return rc;
  if ((rc = ensure_runtime(s)) != RT_OK) return rc;
  cudaStream_t st = s->own_stream;
  CUDA_TRY(cudaMemcpyAsync(s->d_accum.p, accum, bytes, cudaMemcpyHostToDevice, st));   // resume from the checkpoint
  rt_stats local{};
  if ((rc = rt_render_accum(s, cam, opts, s->d_accum.p, st, &local)) != RT_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(accum, s->d_accum.p, bytes, cudaMemcpyDeviceToHost, st));   // the new checkpoint
  local.h2d_bytes += bytes;
  local.d2h_bytes += bytes;
  if (out_linear && (rc = ensure_buf(s->d_linear, npix * 12)) != RT_OK) return rc;
  if (out_rgb8 && (rc = ensure_buf(s->d_rgb8, npix * 3)) != RT_OK) return rc;
  if (out_linear || out_rgb8) {
    rt::launch_resolve((const long long*)s->d_accum.p, (uint32_t)npix, spp_in_accum, cam->gamma,
                       out_linear ? (float*)s->d_linear.p : nullptr, out_rgb8 ? (uint8_t*)s->d_rgb8.p : nullptr, st);
    local.kernel_launches += 1;
    if (out_linear) CUDA_TRY(cudaMemcpyAsync(out_linear, s->d_linear.p, npix * 12, cudaMemcpyDeviceToHost, st));
    if (out_rgb8) CUDA_TRY(cudaMemcpyAsync(out_rgb8, s->d_rgb8.p, npix * 3, cudaMemcpyDeviceToHost, st));
    local.d2h_bytes += (out_linear ? npix * 12 : 0) + (out_rgb8 ? npix * 3 : 0);
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (stats) *stats = local;
  return RT_OK;
}
RT_CATCH("rt_render_progressive")

// ---- parity hooks: one k_extend launch in debug mode, results copied back
static int trace_common(rt_scene* s, const rt_frame& fr, unsigned long long total, uint32_t n, const float* ray_od,
                        int32_t* obj_id, int32_t* prim_id, float* t, float* normal_xyz, float* hitpoint_xyz, float* uv,
                        int32_t* frontface, float* ray_out) {
  int rc = ensure_wavefront(s, fr.capacity);
  if (rc != RT_OK) return rc;
  rt_scene::Wavefront& L0 = s->wf;
  cudaStream_t st = s->own_stream;
  const size_t cap = fr.capacity;
  if ((rc = ensure_buf(s->d_dbg, cap * 36)) != RT_OK) return rc;
  rt::rt_debug dbg;
  dbg.S0 = (float4*)s->d_dbg.p;
  dbg.S1 = dbg.S0 + cap;
  dbg.S2 = (uint32_t*)(dbg.S1 + cap);
  std::vector<float> A, B, C;
  if (ray_od) {
    // caller-supplied rays become "continuing" rays: pixel = i, sample 0, bounce 0
    A.resize((size_t)n * 4); B.resize((size_t)n * 4); C.resize((size_t)n * 4);
    for (uint32_t i = 0; i < n; ++i) {
      const float* p = ray_od + (size_t)i * 6;
      A[4 * i] = p[0]; A[4 * i + 1] = p[1]; A[4 * i + 2] = p[2]; A[4 * i + 3] = p[3];
      B[4 * i] = p[4]; B[4 * i + 1] = p[5]; B[4 * i + 2] = 1.0f; B[4 * i + 3] = 1.0f;
      C[4 * i] = 1.0f;
      uint32_t px = i, sb = 0;
      std::memcpy(&C[4 * i + 1], &px, 4);
      std::memcpy(&C[4 * i + 2], &sb, 4);
      C[4 * i + 3] = 0.0f;
    }
    CUDA_TRY(cudaMemcpyAsync(L0.paths[0].A, A.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(L0.paths[0].B, B.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(L0.paths[0].C, C.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
  }
  rt_frame f2 = fr;
  // single iteration; with supplied rays the control block starts with n_next = n and no new work
  rt::launch_init(L0.ctrl, 0ull, ray_od ? 0ull : total, st);
  if (ray_od) CUDA_TRY(cudaMemcpyAsync(&L0.ctrl->n_next, &n, 4, cudaMemcpyHostToDevice, st));
  rt::launch_advance(L0.ctrl, f2.capacity, st);
  if (!ray_od) rt::launch_raygen(f2, L0.ctrl, L0.paths[0], st);
  rt::launch_trace(s->dev, f2, L0.ctrl, L0.paths[0], L0.hits, L0.sort, false, persistent_grid(s, s->trace_blocks_per_sm, 0), st);
  rt::launch_surface(s->dev, f2, L0.ctrl, L0.paths[0], L0.hits, dbg, st);
  CUDA_TRY(cudaGetLastError());
  std::vector<float> H0((size_t)n * 4), H1((size_t)n * 4), HH((size_t)n * 4);
  std::vector<uint32_t> H2(n);
  std::vector<int32_t> o(n), p(n);
  std::vector<float> tt(n);
  CUDA_TRY(cudaMemcpyAsync(H0.data(), dbg.S0, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(H1.data(), dbg.S1, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(H2.data(), dbg.S2, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(HH.data(), L0.hits.H, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(o.data(), L0.hits.obj, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (ray_out) {
    A.resize((size_t)n * 4); B.resize((size_t)n * 4);
    CUDA_TRY(cudaMemcpyAsync(A.data(), L0.paths[0].A, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(B.data(), L0.paths[0].B, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < n; ++i) {
    bool hit = o[i] >= 0;
    tt[i] = hit ? HH[4 * i] : 0.0f;
    uint32_t pr;
    std::memcpy(&pr, &HH[4 * i + 3], 4);
    p[i] = hit ? (int32_t)pr : 0;
    if (obj_id) obj_id[i] = o[i];
    if (prim_id) prim_id[i] = p[i];
    if (t) t[i] = tt[i];
    if (normal_xyz) {
      normal_xyz[3 * i] = hit ? H0[4 * i + 3] : 0.0f;
      normal_xyz[3 * i + 1] = hit ? H1[4 * i] : 0.0f;
      normal_xyz[3 * i + 2] = hit ? H1[4 * i + 1] : 0.0f;
    }
    if (hitpoint_xyz)
      for (int k = 0; k < 3; ++k) hitpoint_xyz[3 * i + k] = hit ? H0[4 * i + k] : 0.0f;
    if (uv) {
      uv[2 * i] = hit ? H1[4 * i + 2] : 0.0f;
      uv[2 * i + 1] = hit ? H1[4 * i + 3] : 0.0f;
    }
    if (frontface) frontface[i] = hit ? (int32_t)((H2[i] >> 3) & 1u) : 0;
    if (ray_out) {
      float* r = ray_out + (size_t)i * 6;
      r[0] = A[4 * i]; r[1] = A[4 * i + 1]; r[2] = A[4 * i + 2]; r[3] = A[4 * i + 3];
      r[4] = B[4 * i]; r[5] = B[4 * i + 1];
    }
  }
  return RT_OK;
}

int rt_trace_primary(rt_scene* s, const rt_camera* cam, uint64_t seed, uint32_t sample, int32_t* obj_id,
                     int32_t* prim_id, float* t, float* normal_xyz, float* ray_od) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if ((rc = check_camera(cam)) != RT_OK) return rc;
  if (sample >= cam->aa_sample_count) return fail(RT_ERR_INVALID, "rt_trace_primary: sample >= aa_sample_count");
  rt_frame fr;
  fill_frame_camera(*cam, seed, fr);
  fr.sample_begin = sample;
  fr.sample_count = 1;
  unsigned long long total = (unsigned long long)cam->screen_width * cam->screen_height;
  fr.pixel_slots = total;
  if (total > (1ull << 27)) return fail(RT_ERR_INVALID, "rt_trace_primary: image too large for one wavefront");
  fr.capacity = (uint32_t)((total + 127) / 128 * 128);
  return trace_common(s, fr, total, (uint32_t)total, nullptr, obj_id, prim_id, t, normal_xyz, nullptr, nullptr, nullptr,
                      ray_od);
}
RT_CATCH("rt_trace_primary")

int rt_intersect_rays(rt_scene* s, uint64_t seed, uint32_t n, const float* ray_od, float t_min, float t_max,
                      int32_t* obj_id, int32_t* prim_id, float* t, float* normal_xyz, float* hitpoint_xyz, float* uv,
                      int32_t* frontface) try {
  int rc = set_device(s);
  if (rc != RT_OK) return rc;
  if (!ray_od) return fail(RT_ERR_INVALID, "rt_intersect_rays: ray_od is NULL");
  if (n == 0) return RT_OK;
  if (n > (1u << 27)) return fail(RT_ERR_INVALID, "rt_intersect_rays: too many rays for one wavefront");
  rt_frame fr;
  std::memset(&fr, 0, sizeof fr);
  fr.k0 = (uint32_t)seed;
  fr.k1 = (uint32_t)(seed >> 32);
  fr.t_min = t_min;
  fr.t_max = t_max;
  fr.width = n; fr.height = 1; fr.spp = 1; fr.rooti = 1; fr.sample_count = 1; fr.path_depth = 1;
  fr.shard_count = 1;
  fr.capacity = (n + 127) / 128 * 128;
  return trace_common(s, fr, 0, n, ray_od, obj_id, prim_id, t, normal_xyz, hitpoint_xyz, uv, frontface, nullptr);
}
RT_CATCH("rt_intersect_rays")

// ---- assets
int rt_obj_parse(const char* text, size_t len, rt_obj_mesh* out) try {
  if (!text || !out) return fail(RT_ERR_INVALID, "rt_obj_parse: bad argument");
  std::string err;
  int rc = rt::obj_parse(text, len, out, err);
  return rc == RT_OK ? RT_OK : fail(rc, err);
}
RT_CATCH("rt_obj_parse")
int rt_obj_load(const char* path, rt_obj_mesh* out) try {
  if (!path || !out) return fail(RT_ERR_INVALID, "rt_obj_load: bad argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(RT_ERR_IO, std::string("cannot open ") + path);
  std::string data;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
  std::fclose(f);
  return rt_obj_parse(data.data(), data.size(), out);
}
RT_CATCH("rt_obj_load")
void rt_obj_free(rt_obj_mesh* m) {
  if (!m) return;
  std::free(m->pos); std::free(m->nrm); std::free(m->uv); std::free(m->idx);
  std::memset(m, 0, sizeof *m);
}
int rt_tga_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h) try {
  if (!bytes || !rgb || !w || !h) return fail(RT_ERR_INVALID, "rt_tga_decode: bad argument");
  std::string err;
  int rc = rt::tga_decode(bytes, len, rgb, w, h, err);
  return rc == RT_OK ? RT_OK : fail(rc, err);
}
RT_CATCH("rt_tga_decode")
int rt_tga_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len) try {
  int rc = rt::tga_encode_rgb8(rgb, w, h, bytes, len);
  return rc == RT_OK ? RT_OK : fail(rc, "rt_tga_encode_rgb8: bad argument");
}
RT_CATCH("rt_tga_encode_rgb8")
int rt_png_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h) try {
  if (!bytes || !rgb || !w || !h) return fail(RT_ERR_INVALID, "rt_png_decode: bad argument");
  std::string err;
  int rc = rt::png_decode(bytes, len, rgb, w, h, err);
  return rc == RT_OK ? RT_OK : fail(rc, err);
}
RT_CATCH("rt_png_decode")
int rt_jpeg_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h) try {
  if (!bytes || !rgb || !w || !h) return fail(RT_ERR_INVALID, "rt_jpeg_decode: bad argument");
  std::string err;
  int rc = rt::jpeg_decode(bytes, len, rgb, w, h, err);
  return rc == RT_OK ? RT_OK : fail(rc, err);
}
RT_CATCH("rt_jpeg_decode")
int rt_png_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len) try {
  if (!bytes || !len) return fail(RT_ERR_INVALID, "rt_png_encode_rgb8: bad argument");
  int rc = rt::png_encode_rgb8(rgb, w, h, bytes, len);
  return rc == RT_OK ? RT_OK : fail(rc, "rt_png_encode_rgb8: bad argument");
}
RT_CATCH("rt_png_encode_rgb8")
void rt_free(void* p) { std::free(p); }

int rt_mesh_reachability(const float* pos, uint32_t nverts, const uint32_t* idx, uint32_t ntris, uint8_t* mask) try {
  if (!pos || !idx || !mask) return fail(RT_ERR_INVALID, "rt_mesh_reachability: bad argument");
  for (size_t i = 0; i < 3 * (size_t)ntris; ++i)
    if (idx[i] >= nverts) return fail(RT_ERR_INVALID, "rt_mesh_reachability: index out of range");
  rt::mesh_reachability(pos, idx, ntris, mask);
  return RT_OK;
}
RT_CATCH("rt_mesh_reachability")

}  // extern "C"
