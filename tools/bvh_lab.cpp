// bvh_lab.cpp — development tool, CPU only: how much traversal work would a wider tree save?
//
// Builds the BLAS of a mesh with the product's own builder (rt_lower.cpp), replays the binary traversal of k_trace on the
// CPU (child pair per step, nearer child first, the other pushed, interval shrinks with every hit), then collapses the
// same tree into 4- and 8-wide nodes and traverses those.  Reports, per ray: dependent node fetches (the latency
// rounds of a GPU lane), box tests, stack pushes, triangle tests, and an instruction estimate built from the SASS
// counts of the real kernel (binary pair visit: ~50 instructions).  Every tree must find the same closest hit.
//
//   g++ -std=c++17 -O2 -Ics397raytracingsp22_b200/csrc -I/usr/local/cuda/include tools/bvh_lab.cpp \
//       cs397raytracingsp22_b200/csrc/rt_lower.cpp cs397raytracingsp22_b200/csrc/rt_png.cpp \
//       cs397raytracingsp22_b200/csrc/rt_jpeg.cpp -o build/bvh_lab
//   gunzip -c assets/obj/drone.obj.gz > /tmp/drone.obj && build/bvh_lab /tmp/drone.obj
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "rt_lower.h"

using rt::Quad;

struct V3 {
  float x, y, z;
};
static V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static V3 norm(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct Ray {
  V3 o, d;
};
// HostMesh keeps (leftFirst, count) per node; the packed link is made when the BLASes are concatenated
static uint32_t link_of(const Quad* node) {
  uint32_t lf = node[0].u[3], cnt = node[1].u[3];
  return cnt ? (RT_LEAF_FLAG | (lf << 4) | cnt) : lf;
}
struct Stats {
  double fetches = 0, boxes = 0, pushes = 0, tris = 0, leaves = 0, rays = 0, hits = 0, maxstack = 0;
};

static bool slab(const float lo[3], const float hi[3], V3 o, V3 inv, float tmin, float tmax, float& tn) {
  float x0 = (lo[0] - o.x) * inv.x, x1 = (hi[0] - o.x) * inv.x;
  float y0 = (lo[1] - o.y) * inv.y, y1 = (hi[1] - o.y) * inv.y;
  float z0 = (lo[2] - o.z) * inv.z, z1 = (hi[2] - o.z) * inv.z;
  float a = std::max(std::max(std::min(x0, x1), std::min(y0, y1)), std::max(std::min(z0, z1), tmin));
  float b = std::min(std::min(std::max(x0, x1), std::max(y0, y1)), std::min(std::max(z0, z1), tmax));
  tn = a;
  return a <= b * 1.0000005f;
}

struct Mesh {
  rt::HostMesh m;
  bool tri(uint32_t i, const Ray& r, float tmin, float tmax, float& t) const {
    const Quad* q = &m.tris[(size_t)i * RT_TRI_QUADS];
    V3 a = {q[0].f[0], q[0].f[1], q[0].f[2]}, e1 = {q[0].f[3], q[1].f[0], q[1].f[1]}, e2 = {q[1].f[2], q[1].f[3], q[2].f[0]};
    V3 p = cross(r.d, e2);
    float det = dot(e1, p);
    if (std::fabs(det) < 1e-12f) return false;
    float f = 1.0f / det;
    V3 s = r.o - a;
    float u = f * dot(s, p);
    if (u < 0.0f) return false;
    V3 qv = cross(s, e1);
    float v = f * dot(r.d, qv);
    if (v < 0.0f || u + v > 1.0f) return false;
    t = f * dot(e2, qv);
    return !(t < tmin || t > tmax);
  }
};

// ---------------------------------------------------------------- binary tree, as k_trace walks it
static int trace2(const Mesh& M, const Ray& r, float tmin, float tmax, float& best_t, Stats& st) {
  const std::vector<Quad>& N = M.m.nodes;
  V3 inv = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  uint32_t stack[128];
  int sp = 0, best = -1;
  best_t = tmax;
  uint32_t entry = M.m.root_entry_local;
  for (;;) {
    while (entry != RT_ENTRY_NONE && !(entry & RT_LEAF_FLAG)) {
      const Quad* p = &N[(size_t)entry * 2];
      st.fetches += 1;
      st.boxes += 2;
      float tl, tr;
      bool hl = slab(p[0].f, p[1].f, r.o, inv, tmin, best_t, tl), hr = slab(p[2].f, p[3].f, r.o, inv, tmin, best_t, tr);
      uint32_t el = link_of(p), er = link_of(p + 2);
      if (hl && hr) {
        bool lf = tl <= tr;
        stack[sp++] = lf ? er : el;
        st.pushes += 1;
        st.maxstack = std::max<double>(st.maxstack, sp);
        entry = lf ? el : er;
      } else {
        entry = hl ? el : (hr ? er : RT_ENTRY_NONE);
      }
    }
    if (entry != RT_ENTRY_NONE) {
      uint32_t first = (entry & ~RT_LEAF_FLAG) >> 4, n = entry & 15u;
      st.leaves += 1;
      for (uint32_t k = 0; k < n; ++k) {
        float t;
        st.tris += 1;
        if (M.tri(first + k, r, tmin, best_t, t)) {
          best_t = t;
          best = (int)(first + k);
        }
      }
    }
    if (sp == 0) break;
    entry = stack[--sp];
  }
  return best;
}

// ---------------------------------------------------------------- wide tree: collapse of the binary one
struct WNode {
  int n = 0;
  float lo[8][3], hi[8][3];
  uint32_t child[8];  // index of a WNode, or RT_LEAF_FLAG | first << 4 | count
};
struct Wide {
  int width;
  std::vector<WNode> nodes;
  uint32_t root;
};
static float area(const float lo[3], const float hi[3]) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}
// entry = link to a child pair of the binary tree; returns the index of the wide node made from it
static uint32_t collapse(const Mesh& M, uint32_t entry, int width, Wide& W) {
  const std::vector<Quad>& N = M.m.nodes;
  struct Slot {
    float lo[3], hi[3];
    uint32_t link;
  };
  std::vector<Slot> slots;
  auto add_pair = [&](uint32_t e) {
    for (int c = 0; c < 2; ++c) {
      const Quad* p = &N[((size_t)e + c) * 2];
      Slot s;
      for (int k = 0; k < 3; ++k) {
        s.lo[k] = p[0].f[k];
        s.hi[k] = p[1].f[k];
      }
      s.link = link_of(p);
      slots.push_back(s);
    }
  };
  add_pair(entry);
  while ((int)slots.size() < width) {  // open the interior child with the largest surface area
    int pick = -1;
    float best = -1.0f;
    for (size_t i = 0; i < slots.size(); ++i)
      if (!(slots[i].link & RT_LEAF_FLAG) && area(slots[i].lo, slots[i].hi) > best) {
        best = area(slots[i].lo, slots[i].hi);
        pick = (int)i;
      }
    if (pick < 0) break;
    uint32_t e = slots[pick].link;
    slots.erase(slots.begin() + pick);
    add_pair(e);
  }
  uint32_t id = (uint32_t)W.nodes.size();
  W.nodes.emplace_back();
  WNode node;
  node.n = (int)slots.size();
  for (int i = 0; i < node.n; ++i) {
    std::memcpy(node.lo[i], slots[i].lo, 12);
    std::memcpy(node.hi[i], slots[i].hi, 12);
    node.child[i] = slots[i].link;
  }
  for (int i = 0; i < node.n; ++i)
    if (!(node.child[i] & RT_LEAF_FLAG)) node.child[i] = collapse(M, node.child[i], width, W);
  W.nodes[id] = node;
  return id;
}
static int tracew(const Mesh& M, const Wide& W, const Ray& r, float tmin, float tmax, float& best_t, Stats& st) {
  V3 inv = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
  uint32_t stack[256];
  int sp = 0, best = -1;
  best_t = tmax;
  uint32_t entry = W.root;
  for (;;) {
    while (entry != RT_ENTRY_NONE && !(entry & RT_LEAF_FLAG)) {
      const WNode& n = W.nodes[entry];
      st.fetches += 1;
      st.boxes += n.n;
      float tn[8];
      int idx[8], k = 0;
      for (int i = 0; i < n.n; ++i) {
        float t;
        if (slab(n.lo[i], n.hi[i], r.o, inv, tmin, best_t, t)) {
          tn[k] = t;
          idx[k++] = i;
        }
      }
      for (int a = 1; a < k; ++a)  // insertion sort by entry distance
        for (int b = a; b > 0 && tn[b] < tn[b - 1]; --b) {
          std::swap(tn[b], tn[b - 1]);
          std::swap(idx[b], idx[b - 1]);
        }
      for (int a = k - 1; a >= 1; --a) {
        stack[sp++] = n.child[idx[a]];
        st.pushes += 1;
      }
      st.maxstack = std::max<double>(st.maxstack, sp);
      entry = k ? n.child[idx[0]] : RT_ENTRY_NONE;
    }
    if (entry != RT_ENTRY_NONE) {
      uint32_t first = (entry & ~RT_LEAF_FLAG) >> 4, cnt = entry & 15u;
      st.leaves += 1;
      for (uint32_t q = 0; q < cnt; ++q) {
        float t;
        st.tris += 1;
        if (M.tri(first + q, r, tmin, best_t, t)) {
          best_t = t;
          best = (int)(first + q);
        }
      }
    }
    if (sp == 0) break;
    entry = stack[--sp];
  }
  return best;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: bvh_lab mesh.obj [rays]\n");
    return 2;
  }
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
  Mesh M;
  M.m.pos.assign(om.pos, om.pos + 3 * (size_t)om.nverts);
  M.m.nrm.assign(om.nrm, om.nrm + 3 * (size_t)om.nverts);
  M.m.uv.assign(om.uv, om.uv + 2 * (size_t)om.nverts);
  M.m.idx.assign(om.idx, om.idx + 3 * (size_t)om.ntris);
  std::free(om.pos); std::free(om.nrm); std::free(om.uv); std::free(om.idx);
  rt::build_mesh(M.m);
  const size_t ntri = M.m.tris.size() / RT_TRI_QUADS;
  std::printf("%s: %u triangles, %zu reachable, %zu binary nodes, depth %u\n", argv[1], om.ntris, ntri, M.m.nodes.size() / 2, M.m.depth);
  if (M.m.root_entry_local == RT_ENTRY_NONE || (M.m.root_entry_local & RT_LEAF_FLAG)) return 0;

  Wide W4{4, {}, 0}, W8{8, {}, 0};
  W4.root = collapse(M, M.m.root_entry_local, 4, W4);
  W8.root = collapse(M, M.m.root_entry_local, 8, W8);
  auto fill = [](const Wide& W) {
    double c = 0;
    for (const WNode& n : W.nodes) c += n.n;
    return c / std::max<size_t>(W.nodes.size(), 1);
  };
  std::printf("4-wide: %zu nodes, %.2f children per node;  8-wide: %zu nodes, %.2f children per node\n", W4.nodes.size(), fill(W4),
              W8.nodes.size(), fill(W8));

  // rays: (a) from outside towards points inside the box, like camera rays that reach the instance; (b) from points on
  // the surface into the hemisphere above it, like scattered rays leaving the mesh; (c) long rays that cross the box
  const size_t n_rays = argc > 2 ? (size_t)std::atol(argv[2]) : 200000;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(0.0f, 1.0f);
  V3 lo = {M.m.root_min[0], M.m.root_min[1], M.m.root_min[2]}, hi = {M.m.root_max[0], M.m.root_max[1], M.m.root_max[2]};
  V3 c = (lo + hi) * 0.5f, ext = hi - lo;
  float rad = 0.5f * std::sqrt(dot(ext, ext));
  auto ball = [&]() {
    for (;;) {
      V3 v = {2 * U(rng) - 1, 2 * U(rng) - 1, 2 * U(rng) - 1};
      float l = dot(v, v);
      if (l <= 1.0f && l > 1e-6f) return v;
    }
  };
  const char* names[3] = {"towards the mesh from outside", "leaving the surface", "crossing the box"};
  for (int kind = 0; kind < 3; ++kind) {
    Stats s2, s4, s8;
    size_t mismatch = 0;
    double t_sum = 0.0;           // checksum of the closest hits: must not depend on how the tree was built
    unsigned long long id_sum = 0;
    for (size_t i = 0; i < n_rays; ++i) {
      Ray r;
      if (kind == 0) {
        r.o = c + norm(ball()) * (3.0f * rad);
        V3 target = {lo.x + U(rng) * ext.x, lo.y + U(rng) * ext.y, lo.z + U(rng) * ext.z};
        r.d = norm(target - r.o);
      } else if (kind == 1) {
        size_t t = (size_t)(U(rng) * ntri) % ntri;
        const Quad* q = &M.m.tris[t * RT_TRI_QUADS];
        V3 a = {q[0].f[0], q[0].f[1], q[0].f[2]}, e1 = {q[0].f[3], q[1].f[0], q[1].f[1]}, e2 = {q[1].f[2], q[1].f[3], q[2].f[0]};
        float u = U(rng), v = U(rng);
        if (u + v > 1) {
          u = 1 - u;
          v = 1 - v;
        }
        V3 n = cross(e1, e2);
        if (dot(n, n) < 1e-20f) {
          --i;
          continue;
        }
        n = norm(n);
        if (U(rng) < 0.5f) n = n * -1.0f;
        V3 d = ball();
        if (dot(d, n) < 0) d = d * -1.0f;
        r.o = a + e1 * u + e2 * v + n * (1e-3f * rad);
        r.d = d;  // not normalised, like the reference's scattered rays (Q1)
      } else {
        r.o = c + norm(ball()) * (3.0f * rad);
        r.d = norm((c + ball() * rad) - r.o);
      }
      float t2, t4, t8;
      int h2 = trace2(M, r, 0.001f, 3.0e38f, t2, s2), h4 = tracew(M, W4, r, 0.001f, 3.0e38f, t4, s4), h8 = tracew(M, W8, r, 0.001f, 3.0e38f, t8, s8);
      if (h2 >= 0) {
        s2.hits += 1;
        t_sum += t2;
        id_sum += M.m.tris[(size_t)h2 * RT_TRI_QUADS + 2].u[1];  // original triangle id
      }
      if ((h2 >= 0) != (h4 >= 0) || (h2 >= 0) != (h8 >= 0) || (h2 >= 0 && (t2 != t4 || t2 != t8))) ++mismatch;
    }
    std::printf("\nrays %s (%zu, %.1f %% hit, %zu closest-hit mismatches between the trees; checksum t %.9g ids %llu)\n", names[kind], n_rays,
                100.0 * s2.hits / n_rays, mismatch, t_sum, id_sum);
    std::printf("  tree     node fetches  box tests  pushes  leaves  tri tests  max stack   est. instructions\n");
    auto row = [&](const char* nm, const Stats& s, double per_fetch, double per_box) {
      double n = (double)n_rays;
      // SASS of k_trace: a binary pair visit is ~50 instructions = ~14 of loop / address / push overhead + 2 x ~18 per box
      // (6 FFMA, 8 FMNMX, 2 FMNMX3, 1 FMUL, 1 FSETP); a wide visit keeps the overhead, pays the same per box, and
      // orders its hits with ~3 instructions per box
      double instr = s.fetches * per_fetch + s.boxes * per_box + s.tris * 45.0 + s.pushes * 4.0;
      std::printf("  %-8s %9.2f  %9.2f  %6.2f  %6.2f  %9.2f  %9.0f   %9.0f\n", nm, s.fetches / n, s.boxes / n, s.pushes / n, s.leaves / n, s.tris / n,
                  s.maxstack, instr / n);
    };
    row("binary", s2, 14.0, 18.0);
    row("4-wide", s4, 14.0, 21.0);
    row("8-wide", s8, 14.0, 21.0);
  }
  return 0;
}
