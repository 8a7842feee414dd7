// node_dump.cpp — development / test helper: lower a small scene (two instances of an OBJ mesh with different
// transforms, three spheres, a loose triangle) with the library's own lowering code and write the node array
// (16-byte quads, four per child pair) to stdout.  tests/test_node_layout.py builds it twice, with -DRT_NODE_CH=0
// ((min, max) per child) and with the default packed child pairs, and checks that every packed box contains the
// (min, max) box it came from, tightly, and that the links are the same.
//   g++ -std=c++17 -O2 [-DRT_NODE_CH=0] -Ics397raytracingsp22_b200/csrc -I/usr/local/cuda/include tools/node_dump.cpp \
//       cs397raytracingsp22_b200/csrc/rt_lower.cpp cs397raytracingsp22_b200/csrc/rt_png.cpp cs397raytracingsp22_b200/csrc/rt_jpeg.cpp -o node_dump
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rt_lower.h"

static void affine(float m[16], float s, float tx, float ty, float tz) {
  std::memset(m, 0, 16 * sizeof(float));
  m[0] = m[5] = m[10] = s;
  m[15] = 1.0f;
  m[12] = tx; m[13] = ty; m[14] = tz;  // column-major 4x4 like cgmath (include/rt_b200.h)
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 1;
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
  std::fclose(f);
  rt_obj_mesh om;
  std::string err;
  if (rt::obj_parse(text.data(), text.size(), &om, err) != 0) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  std::vector<rt::HostMesh> meshes(1);
  rt::HostMesh& m = meshes[0];
  m.pos.assign(om.pos, om.pos + 3 * (size_t)om.nverts);
  m.nrm.assign(om.nrm, om.nrm + 3 * (size_t)om.nverts);
  m.uv.assign(om.uv, om.uv + 2 * (size_t)om.nverts);
  m.idx.assign(om.idx, om.idx + 3 * (size_t)om.ntris);
  std::free(om.pos); std::free(om.nrm); std::free(om.uv); std::free(om.idx);
  rt::build_mesh(m);

  std::vector<rt_material_desc> mats(1);
  std::memset(&mats[0], 0, sizeof mats[0]);
  mats[0].tag = RT_MAT_LAMBERTIAN;
  mats[0].albedo[0] = mats[0].albedo[1] = mats[0].albedo[2] = 0.5f;
  std::vector<rt::HostObject> objects;
  const float place[2][4] = {{0.25f, -1.0f, 0.0f, -3.0f}, {0.003f, 40.0f, 2.0f, 7.0f}};
  for (int k = 0; k < 2; ++k) {
    rt::HostObject o;
    o.kind = RT_OBJ_MESH;
    o.mesh = 0;
    o.material = 0;
    affine(o.xform, place[k][0], place[k][1], place[k][2], place[k][3]);
    if (!rt::invert_affine_cofactor(o.xform, o.inv_xform)) return 1;
    objects.push_back(o);
  }
  for (int k = 0; k < 3; ++k) {
    rt::HostObject o;
    o.kind = RT_OBJ_SPHERE;
    o.material = 0;
    o.a[0] = 2.0f * k - 1.0f; o.a[1] = 0.5f; o.a[2] = -1.0f - k;
    o.radius = 0.3f + 0.2f * k;
    objects.push_back(o);
  }
  {
    rt::HostObject o;
    o.kind = RT_OBJ_TRIANGLE;
    o.material = 0;
    o.a[0] = -2; o.a[1] = 5; o.a[2] = -2;
    o.b[0] = 2; o.b[1] = 5; o.b[2] = -2;
    o.c[0] = 0; o.c[1] = 5; o.c[2] = 2;
    objects.push_back(o);
  }
  rt::Lowered L;
  std::vector<rt::HostTexture> textures;
  if (rt::lower_scene(textures, mats, meshes, objects, L, err) != 0) {
    std::fprintf(stderr, "lower_scene: %s\n", err.c_str());
    return 1;
  }
  std::fprintf(stderr, "%zu quads, tlas_base %u, tlas_count %u, tlas_root %u\n", L.nodes.size(), L.tlas_base, L.tlas_count, L.tlas_root);
  std::fwrite(L.nodes.data(), sizeof(rt::Quad), L.nodes.size(), stdout);
  return 0;
}
