// rt_scene_api.hpp — header-only C++ mirror of the reference's scene-construction API on top of the C ABI
// (rt_b200.h).  The reference is compiled code (Rust); Rust is not available in this image, so this is the compiled
// host side: same type names, field names and call shapes as src/util/{tracing,geometry,materials,texture}.rs, with
// one added method per type, lower(), which is exactly what INTEGRATION.md asks the Rust traits to gain.
//
//   Rust                                              here
//   Arc<dyn Material + Send + Sync>                   std::shared_ptr<const rt::Material>
//   Arc<dyn Intersectable + Send + Sync>              std::shared_ptr<const rt::Intersectable>
//   Scene{camera, objects, ..}.render_to_image()      rt::Scene{camera, objects}.render_to_image()
//   panics (geometry.rs:149-151,168)                  rt::Error exceptions carrying the ABI's error code
#ifndef RT_SCENE_API_HPP
#define RT_SCENE_API_HPP

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "rt_b200.h"

namespace rt {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline int check(int rc) {
  if (rc < 0) throw Error(rc, rt_last_error());
  return rc;
}

struct Vec3 {
  float x = 0, y = 0, z = 0;
};
inline Vec3 vec3(float x, float y, float z) { return Vec3{x, y, z}; }
using Color = Vec3;

// ---- cgmath::Matrix4<f32>, column-major, the constructors the reference's scene uses (tracing.rs:383,393,403)
struct Matrix4 {
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  static Matrix4 from_translation(Vec3 v) {
    Matrix4 r;
    r.m[12] = v.x; r.m[13] = v.y; r.m[14] = v.z;
    return r;
  }
  static Matrix4 from_scale(float s) {
    Matrix4 r;
    r.m[0] = r.m[5] = r.m[10] = s;
    return r;
  }
  static Matrix4 from_angle_x(float deg) {
    float a = deg * 0.017453292f, s = std::sin(a), c = std::cos(a);
    Matrix4 r;
    r.m[5] = c; r.m[6] = s; r.m[9] = -s; r.m[10] = c;
    return r;
  }
  static Matrix4 from_angle_y(float deg) {
    float a = deg * 0.017453292f, s = std::sin(a), c = std::cos(a);
    Matrix4 r;
    r.m[0] = c; r.m[2] = -s; r.m[8] = s; r.m[10] = c;
    return r;
  }
  Matrix4 operator*(const Matrix4& b) const {  // column j = a0*b[j][0] + a1*b[j][1] + a2*b[j][2] + a3*b[j][3]
    Matrix4 r;
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i)
        r.m[j * 4 + i] = m[i] * b.m[j * 4] + m[4 + i] * b.m[j * 4 + 1] + m[8 + i] * b.m[j * 4 + 2] + m[12 + i] * b.m[j * 4 + 3];
    return r;
  }
};

struct LowerCtx;

// ---- materials (materials.rs)
struct Material {
  virtual ~Material() = default;
  virtual rt_material_desc desc() const = 0;
};
struct Lambertian : Material {  // materials.rs:20-32
  Vec3 albedo{1, 1, 1}, emission{0, 0, 0};
  Lambertian() = default;
  Lambertian(Vec3 a, Vec3 e = Vec3{}) : albedo(a), emission(e) {}
  rt_material_desc desc() const override {
    return rt_material_desc{RT_MAT_LAMBERTIAN, {albedo.x, albedo.y, albedo.z}, {emission.x, emission.y, emission.z}, 0, 0, 1};
  }
};
struct Metal : Material {  // materials.rs:51-55
  Color albedo{1, 1, 1}, emission{0, 0, 0};
  float roughness = 0;
  Metal(Color a, Color e, float r) : albedo(a), emission(e), roughness(r) {}
  rt_material_desc desc() const override {
    return rt_material_desc{RT_MAT_METAL, {albedo.x, albedo.y, albedo.z}, {emission.x, emission.y, emission.z}, roughness, 0, 1};
  }
};
struct Dielectric : Material {  // materials.rs:74-76
  float idx_of_refraction = 1.5f;
  explicit Dielectric(float ior) : idx_of_refraction(ior) {}
  rt_material_desc desc() const override { return rt_material_desc{RT_MAT_DIELECTRIC, {0, 0, 0}, {0, 0, 0}, 0, 0, idx_of_refraction}; }
};
struct ParameterizedMaterial : Material {  // materials.rs:107-112
  Color albedo{1, 1, 1}, emission{0, 0, 0};
  float roughness = 1, metallic = 0;
  ParameterizedMaterial(Color a, Color e, float r, float m) : albedo(a), emission(e), roughness(r), metallic(m) {}
  rt_material_desc desc() const override {
    return rt_material_desc{RT_MAT_PARAMETERIZED, {albedo.x, albedo.y, albedo.z}, {emission.x, emission.y, emission.z}, roughness, metallic, 1};
  }
};
struct Isotropic : Material {  // materials.rs:152-157
  Color albedo{1, 1, 1}, emission{0, 0, 0};
  Isotropic(Color a, Color e = Color{}) : albedo(a), emission(e) {}
  rt_material_desc desc() const override {
    return rt_material_desc{RT_MAT_ISOTROPIC, {albedo.x, albedo.y, albedo.z}, {emission.x, emission.y, emission.z}, 0, 0, 1};
  }
};
using MaterialRef = std::shared_ptr<const Material>;

// ---- texture (texture.rs).  PNG, JPEG and TGA are decoded by the library's own readers (format by magic number).
struct Texture {
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> rgb8;
  static std::optional<Texture> load_from_file(const std::string& file_name) {  // texture.rs:16-25: None on failure
    std::ifstream f(file_name, std::ios::binary);
    if (!f) return std::nullopt;
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    uint8_t* px = nullptr;
    uint32_t w = 0, h = 0;
    bool png = bytes.size() > 8 && bytes[0] == 137 && bytes[1] == 'P' && bytes[2] == 'N' && bytes[3] == 'G';
    bool jpeg = bytes.size() > 3 && bytes[0] == 0xFF && bytes[1] == 0xD8;
    int rc = png    ? rt_png_decode(bytes.data(), bytes.size(), &px, &w, &h)
             : jpeg ? rt_jpeg_decode(bytes.data(), bytes.size(), &px, &w, &h)
                    : rt_tga_decode(bytes.data(), bytes.size(), &px, &w, &h);
    if (rc != RT_OK) return std::nullopt;
    Texture t;
    t.width = w; t.height = h;
    t.rgb8.assign(px, px + (size_t)w * h * 3);
    rt_free(px);
    return t;
  }
};
using TextureRef = std::shared_ptr<const Texture>;

// ---- de-duplication of shared Arcs while lowering
struct LowerCtx {
  rt_scene* s;
  std::map<const Material*, int> mats;
  std::map<const Texture*, int> texs;
  std::map<const void*, int> meshes;
  int material(const MaterialRef& m) {
    auto it = mats.find(m.get());
    if (it != mats.end()) return it->second;
    rt_material_desc d = m->desc();
    return mats[m.get()] = check(rt_add_material(s, &d));
  }
  int texture(const TextureRef& t) {
    if (!t) return -1;
    auto it = texs.find(t.get());
    if (it != texs.end()) return it->second;
    return texs[t.get()] = check(rt_add_texture(s, t->rgb8.data(), t->width, t->height));
  }
};

// ---- geometry (geometry.rs)
struct Intersectable {  // tracing.rs:42-47; intersect_ray itself runs in k_trace
  virtual ~Intersectable() = default;
  virtual int lower(LowerCtx& c) const = 0;
};
struct Sphere : Intersectable {  // geometry.rs:389-393
  Vec3 center;
  float radius;
  MaterialRef material;
  Sphere(Vec3 c, float r, MaterialRef m) : center(c), radius(r), material(std::move(m)) {}
  int lower(LowerCtx& c) const override { return check(rt_add_sphere(c.s, &center.x, radius, c.material(material))); }
};
struct Triangle : Intersectable {  // geometry.rs:424-429
  Vec3 a, b, c;
  MaterialRef material;
  Triangle(Vec3 a_, Vec3 b_, Vec3 c_, MaterialRef m) : a(a_), b(b_), c(c_), material(std::move(m)) {}
  int lower(LowerCtx& x) const override { return check(rt_add_triangle(x.s, &a.x, &b.x, &c.x, x.material(material))); }
};
struct Plane : Intersectable {  // geometry.rs:468-472
  Vec3 point, normal;
  MaterialRef material;
  Plane(Vec3 p, Vec3 n, MaterialRef m) : point(p), normal(n), material(std::move(m)) {}
  int lower(LowerCtx& c) const override { return check(rt_add_plane(c.s, &point.x, &normal.x, c.material(material))); }
};
struct StaticMesh;
struct ConvexVolume : Intersectable {  // geometry.rs:495-500; the boundary is a Sphere or a StaticMesh
  std::shared_ptr<const Sphere> boundary;
  std::shared_ptr<const StaticMesh> mesh_boundary;
  MaterialRef phase_function;
  float density;
  ConvexVolume(std::shared_ptr<const Sphere> b, MaterialRef p, float d) : boundary(std::move(b)), phase_function(std::move(p)), density(d) {}
  ConvexVolume(std::shared_ptr<const StaticMesh> b, MaterialRef p, float d) : mesh_boundary(std::move(b)), phase_function(std::move(p)), density(d) {}
  int lower(LowerCtx& c) const override;
};
struct MeshData {  // tobj::Mesh
  std::vector<float> positions, normals, texcoords;
  std::vector<uint32_t> indices;
};
struct StaticMesh : Intersectable {  // geometry.rs:127-134
  std::shared_ptr<const MeshData> mesh;
  MaterialRef material;       // may be null: textures drive a ParameterizedMaterial
  TextureRef textures[5];     // albedo, emission, metallic, roughness, normal
  Matrix4 transform;
  // StaticMesh::load_from_file, geometry.rs:138-172
  static StaticMesh load_from_file(const std::string& file_name, const char* albedo_path, const char* emission_path,
                                   const char* metallic_path, const char* roughness_path, const char* normal_path,
                                   MaterialRef material, const Matrix4& transform) {
    rt_obj_mesh m;
    check(rt_obj_load(file_name.c_str(), &m));  // the reference asserts/panics here (geometry.rs:149-151)
    if (!m.has_normals || !m.has_texcoords) {
      rt_obj_free(&m);
      throw Error(RT_ERR_IO, file_name + ": vn and vt are required on every face corner (geometry.rs:230-243)");
    }
    auto md = std::make_shared<MeshData>();
    md->positions.assign(m.pos, m.pos + 3 * (size_t)m.nverts);
    md->normals.assign(m.nrm, m.nrm + 3 * (size_t)m.nverts);
    md->texcoords.assign(m.uv, m.uv + 2 * (size_t)m.nverts);
    md->indices.assign(m.idx, m.idx + 3 * (size_t)m.ntris);
    rt_obj_free(&m);
    StaticMesh sm;
    sm.mesh = md;
    sm.material = std::move(material);
    const char* paths[5] = {albedo_path, emission_path, metallic_path, roughness_path, normal_path};
    for (int i = 0; i < 5; ++i)
      if (paths[i])
        if (auto t = Texture::load_from_file(paths[i])) sm.textures[i] = std::make_shared<Texture>(std::move(*t));
    sm.transform = transform;
    return sm;
  }
  int lower(LowerCtx& c) const override {
    auto it = c.meshes.find(mesh.get());
    int mid = it != c.meshes.end() ? it->second
                                   : (c.meshes[mesh.get()] = check(rt_add_mesh(c.s, mesh->positions.data(), mesh->normals.data(),
                                                                              mesh->texcoords.data(), (uint32_t)(mesh->positions.size() / 3),
                                                                              mesh->indices.data(), (uint32_t)(mesh->indices.size() / 3))));
    int tex[5];
    for (int i = 0; i < 5; ++i) tex[i] = c.texture(textures[i]);
    // inv_xform = NULL: the library inverts by cofactors, like transform.inverse_transform().unwrap() (geometry.rs:168)
    return check(rt_add_instance(c.s, mid, transform.m, nullptr, material ? c.material(material) : -1, tex));
  }
};

inline int ConvexVolume::lower(LowerCtx& c) const {
  if (boundary) return check(rt_add_volume_sphere(c.s, &boundary->center.x, boundary->radius, density, c.material(phase_function)));
  const StaticMesh& m = *mesh_boundary;
  auto it = c.meshes.find(m.mesh.get());
  int mid = it != c.meshes.end() ? it->second
                                 : (c.meshes[m.mesh.get()] = check(rt_add_mesh(c.s, m.mesh->positions.data(), m.mesh->normals.data(),
                                                                                m.mesh->texcoords.data(), (uint32_t)(m.mesh->positions.size() / 3),
                                                                                m.mesh->indices.data(), (uint32_t)(m.mesh->indices.size() / 3))));
  return check(rt_add_volume_mesh(c.s, mid, m.transform.m, nullptr, density, c.material(phase_function)));
}

// ---- camera and scene (tracing.rs)
enum class CameraProjectionMode { Orthographic = RT_PROJ_ORTHOGRAPHIC, Perspective = RT_PROJ_PERSPECTIVE };
enum class ShadingMode { Phong = RT_SHADE_PHONG, PathTrace = RT_SHADE_PATHTRACE };
struct Camera {  // tracing.rs:138-155, same field names, HEAD's defaults (tracing.rs:357-373)
  Vec3 eyepoint{0, 2, 5.5f}, view_dir{0, 0, -1}, up{0, 1, 0};
  CameraProjectionMode projection_mode = CameraProjectionMode::Perspective;
  ShadingMode shading_mode = ShadingMode::PathTrace;
  uint32_t path_depth = 10, path_samples = 1, screen_width = 100, screen_height = 100;
  float focal_length = 0.6f, focus_dist = 5.0f, lens_radius = 0.0f;
  uint32_t aa_sample_count = 100;
  float max_trace_dist = 100.0f, gamma = 2.0f;
  rt_camera to_c() const {
    rt_camera c;
    std::memset(&c, 0, sizeof c);
    std::memcpy(c.eyepoint, &eyepoint.x, 12);
    std::memcpy(c.view_dir, &view_dir.x, 12);
    std::memcpy(c.up, &up.x, 12);
    c.projection_mode = (uint32_t)projection_mode;
    c.shading_mode = (uint32_t)shading_mode;
    c.path_depth = path_depth; c.path_samples = path_samples;
    c.screen_width = screen_width; c.screen_height = screen_height;
    c.focal_length = focal_length; c.focus_dist = focus_dist; c.lens_radius = lens_radius;
    c.aa_sample_count = aa_sample_count; c.max_trace_dist = max_trace_dist; c.gamma = gamma;
    return c;
  }
};
struct RgbImage {  // image::RgbImage: row 0 = top
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> data;
  void save_png(const std::string& path) const {  // save_with_format("render.png", ImageFormat::Png), tracing.rs:546
    uint8_t* bytes = nullptr;
    size_t len = 0;
    check(rt_png_encode_rgb8(data.data(), width, height, &bytes, &len));
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)bytes, (std::streamsize)len);
    rt_free(bytes);
  }
  void save_tga(const std::string& path) const {
    uint8_t* bytes = nullptr;
    size_t len = 0;
    check(rt_tga_encode_rgb8(data.data(), width, height, &bytes, &len));
    std::ofstream f(path, std::ios::binary);
    f.write((const char*)bytes, (std::streamsize)len);
    rt_free(bytes);
  }
};
struct Scene {  // tracing.rs:213-218
  Camera camera;
  std::vector<std::shared_ptr<const Intersectable>> objects;
  Vec3 point_light_pos{0, 1, 5}, ambient{0.1f, 0.1f, 0.1f};  // ShadingMode::Phong only
  uint64_t seed = 0x5EED;
  // Scene::render_to_image, tracing.rs:221-263: lower -> commit -> render on CUDA device `device`
  RgbImage render_to_image(int device = 0, rt_stats* stats = nullptr) const {
    rt_scene* s = nullptr;
    check(rt_scene_create(&s));
    struct Guard {
      rt_scene* s;
      ~Guard() { rt_scene_destroy(s); }
    } guard{s};
    LowerCtx ctx{s, {}, {}, {}};
    for (const auto& o : objects) o->lower(ctx);  // insertion order = the reference's tie-break order
    check(rt_commit(s, device));
    rt_camera cam = camera.to_c();
    rt_render_opts opts;
    std::memset(&opts, 0, sizeof opts);
    opts.seed = seed;
    opts.point_light_pos[0] = point_light_pos.x; opts.point_light_pos[1] = point_light_pos.y; opts.point_light_pos[2] = point_light_pos.z;
    opts.ambient[0] = ambient.x; opts.ambient[1] = ambient.y; opts.ambient[2] = ambient.z;
    RgbImage img;
    img.width = camera.screen_width;
    img.height = camera.screen_height;
    img.data.resize((size_t)img.width * img.height * 3);
    check(rt_render(s, &cam, &opts, nullptr, img.data.data(), stats));
    return img;
  }
};

}  // namespace rt
#endif
