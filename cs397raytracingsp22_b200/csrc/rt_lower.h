// rt_lower.h — host-side scene IR and lowering to the flat arrays of rt_types.h.
// Replaces the reference's one-time asset build (tobj::load_obj geometry.rs:140-148, the
// pointer BVH geometry.rs:175-217, image::open texture.rs:17) — never in the per-ray loop.
#ifndef RT_LOWER_H
#define RT_LOWER_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_types.h"

namespace rt {

struct Quad {
  union {
    float f[4];
    uint32_t u[4];
    int32_t i[4];
  };
};
static_assert(sizeof(Quad) == 16, "quad");

struct HostTexture {
  uint32_t w, h;
  std::vector<uint32_t> rgba;  // one word per texel, R in the low byte
};

struct HostMesh {
  std::vector<float> pos, nrm, uv;
  std::vector<uint32_t> idx;
  std::vector<uint8_t> reach;  // 1 = reachable in the reference's index-order tree
  // BLAS over the reachable triangles, local node indices (rebased when concatenated)
  std::vector<Quad> nodes;     // 2 per node
  std::vector<Quad> tris;      // 3 per packed triangle, leaf order
  std::vector<Quad> shade;     // 5 per ORIGINAL triangle
  std::vector<Quad> guards;    // 2 per guard box (thin interior boxes of the reference tree, see rt_lower.cpp)
  std::vector<uint32_t> guard_list;  // per-triangle lists of guard indices
  float root_min[3], root_max[3];
  uint32_t root_entry_local;   // packed entry of the root (local indices)
  uint32_t n_reachable;
  uint32_t depth = 0;          // longest root-to-leaf path of the BLAS, in nodes
  float pad_base = 0, amax = 0, pad = 0;  // box padding: base (4e-6 of the largest coordinate), applied value
  uint32_t ntris() const { return (uint32_t)(idx.size() / 3); }
};

struct HostObject {
  int kind;
  int mesh = -1;
  float xform[16], inv_xform[16];
  int tex[5] = {-1, -1, -1, -1, -1};
  int material = -1;
  float a[3] = {0, 0, 0}, b[3] = {0, 0, 0}, c[3] = {0, 0, 0};
  float radius = 0, density = 0;
  int vol_index = -1;
};

struct Lowered {
  std::vector<Quad> nodes, tris, shade, objects, mats;
  std::vector<Quad> textures;
  std::vector<uint32_t> texels;
  std::vector<int32_t> planes;
  std::vector<Quad> guards;
  std::vector<uint32_t> guard_list;
  uint32_t tlas_root = RT_ENTRY_NONE;
  uint32_t tlas_base = 0, tlas_count = 0;  // where the TLAS nodes sit in `nodes` (breadth-first order)
  float tlas_min[3] = {0, 0, 0}, tlas_max[3] = {0, 0, 0};
  uint32_t n_volumes = 0;
  uint32_t tlas_depth = 0;
  uint64_t bytes() const;
};

// reference tree replay: which triangles can the strict slab test ever reach
void mesh_reachability(const float* pos, const uint32_t* idx, uint32_t ntris, uint8_t* mask);
// reachability mask + binned-SAH BLAS + packed records
void build_mesh(HostMesh& m);
// TLAS + tables
int lower_scene(const std::vector<HostTexture>& textures, const std::vector<rt_material_desc>& materials,
                std::vector<HostMesh>& meshes, const std::vector<HostObject>& objects, Lowered& out,
                std::string& err);

bool invert_affine_cofactor(const float m[16], float out[16]);

// assets
int obj_parse(const char* text, size_t len, rt_obj_mesh* out, std::string& err);
int tga_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err);
int tga_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len);
int png_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err);  // rt_png.cpp
int jpeg_decode(const uint8_t* bytes, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err);  // rt_jpeg.cpp
int png_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len);

}  // namespace rt
#endif
