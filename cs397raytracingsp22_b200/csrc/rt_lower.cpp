// rt_lower.cpp — host lowering: reference-reachability mask, binned-SAH BVH2 (32-byte nodes,
// children in adjacent pairs), leaf-packed triangle records, TLAS over bounded objects,
// tagged-union material / texture tables; plus the asset readers (OBJ, TGA).
//
// Must be compiled WITHOUT floating-point contraction (-ffp-contract=off): the records
// precomputed here (edges, tangents, normals) have to be bit-identical to what the
// reference computes per ray (geometry.rs:336-337,245-250,449).

#include "rt_lower.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>

namespace rt {

// ------------------------------------------------------------------ small helpers
static inline float fmin3(float a, float b, float c) { return std::fmin(a, std::fmin(b, c)); }
static inline float fmax3(float a, float b, float c) { return std::fmax(a, std::fmax(b, c)); }

uint64_t Lowered::bytes() const {
  return (nodes.size() + tris.size() + shade.size() + objects.size() + mats.size() + textures.size() + guards.size()) * 16ull +
         texels.size() * 4ull + planes.size() * 4ull + guard_list.size() * 4ull;
}

// ------------------------------------------------------------------ reachability (Q3)
// The reference tree (geometry.rs:190-217): node over [start,end), mid = start+(end-start)/2,
// leaf = triangle `start`.  Its slab test (geometry.rs:63-67) rejects when tmax <= tmin, so an
// INTERIOR node whose box has zero thickness on an axis is never entered, and every triangle
// below it can never be hit.  Leaves are not box-tested (geometry.rs:95-98).
namespace {
struct Box3 {
  float mn[3], mx[3];
};
struct GuardRange {
  Box3 box;
  uint32_t start, end;  // triangles below the thin interior node
};
// An interior box that is thin but not flat is rejected by the same strict test only for rays whose origin is so far
// away that (min - o) == (max - o) in f32 - ray dependent, so it cannot go into the static mask.  Such boxes are
// recorded as GUARDS: the kernel re-runs the reference's slab test on them (exactly) before it accepts a hit on a
// triangle below.  Threshold: thickness below 2^-10 of the largest coordinate; a guard that passes costs nothing
// but time, so over-flagging is harmless.
Box3 reach_rec(const float* pos, const uint32_t* idx, uint32_t start, uint32_t end, uint8_t* mask,
               std::vector<GuardRange>* guards) {
  Box3 b;
  if (end - start == 1) {
    const float* a = pos + 3 * (size_t)idx[3 * start];
    const float* p = pos + 3 * (size_t)idx[3 * start + 1];
    const float* c = pos + 3 * (size_t)idx[3 * start + 2];
    for (int k = 0; k < 3; ++k) {
      b.mn[k] = fmin3(a[k], p[k], c[k]);
      b.mx[k] = fmax3(a[k], p[k], c[k]);
    }
    mask[start] = 1;
    return b;
  }
  uint32_t mid = start + (end - start) / 2;
  Box3 l = reach_rec(pos, idx, start, mid, mask, guards);
  Box3 r = reach_rec(pos, idx, mid, end, mask, guards);
  for (int k = 0; k < 3; ++k) {
    b.mn[k] = std::fmin(l.mn[k], r.mn[k]);
    b.mx[k] = std::fmax(l.mx[k], r.mx[k]);
  }
  bool flat = b.mn[0] == b.mx[0] || b.mn[1] == b.mx[1] || b.mn[2] == b.mx[2];
  if (flat) {
    for (uint32_t t = start; t < end; ++t) mask[t] = 0;
  } else if (guards) {
    float amax = 0.0f, thin = FLT_MAX;
    for (int k = 0; k < 3; ++k) {
      amax = std::max(amax, std::max(std::fabs(b.mn[k]), std::fabs(b.mx[k])));
      thin = std::min(thin, b.mx[k] - b.mn[k]);
    }
    if (thin < amax * (1.0f / 1024.0f)) guards->push_back(GuardRange{b, start, end});
  }
  return b;
}
}  // namespace

void mesh_reachability(const float* pos, const uint32_t* idx, uint32_t ntris, uint8_t* mask) {
  if (ntris == 0) return;
  reach_rec(pos, idx, 0, ntris, mask, nullptr);
}

// ------------------------------------------------------------------ binned SAH BVH2
namespace {
struct BPrim {
  float mn[3], mx[3], c[3];
  uint32_t id;
  uint32_t solo = 0;  // 1: must be alone in its leaf (TLAS: mesh instances)
};
struct BNode {
  float mn[3], mx[3];
  uint32_t leftFirst, count;
};
inline float half_area(const float* mn, const float* mx) {
  float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
  return ex * ey + ey * ez + ez * ex;
}

// builder settings, compile-time (each was measured on the device, profiles/r1_notes.md: leaf size 2 / 4 / 8 and
// traversal cost 0.5 / 1 / 2 move the frame time by < 1 %; the full sweep beats 16 bins by 0.6 %)
#ifndef RT_LEAF_MAX
#define RT_LEAF_MAX 4      // triangles per BLAS leaf, 1..RT_MAX_LEAF_TRIS
#endif
#ifndef RT_SAH_CT
#define RT_SAH_CT 1.0f     // cost of a node-pair visit in triangle tests
#endif
#ifndef RT_SAH_SWEEP
#define RT_SAH_SWEEP 1     // 1: full-sweep SAH (every split position on every axis), 0: 16 bins
#endif
static_assert(RT_LEAF_MAX >= 1 && RT_LEAF_MAX <= RT_MAX_LEAF_TRIS, "leaf size must fit the packed entry");
const float g_sah_ct = RT_SAH_CT;
const uint32_t g_leaf_max = RT_LEAF_MAX;
const int g_sweep = RT_SAH_SWEEP;
float g_prim_cost = 1.0f;       // set per build

// full-sweep SAH: best (axis, position) over all count-1 splits of the centroid-sorted order.
// On success prims[first, first+count) is sorted along the best axis and `mid` is the split index.
bool sweep_split(std::vector<BPrim>& prims, uint32_t first, uint32_t count, float& best_cost, uint32_t& mid) {
  std::vector<float> right_area(count);
  int best_axis = -1;
  uint32_t best_pos = 0;
  best_cost = FLT_MAX;
  std::vector<BPrim> tmp(prims.begin() + first, prims.begin() + first + count);
  for (int axis = 0; axis < 3; ++axis) {
    std::sort(tmp.begin(), tmp.end(), [axis](const BPrim& a, const BPrim& b) {
      return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.id < b.id);
    });
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = count; i-- > 1;) {
      for (int k = 0; k < 3; ++k) {
        mn[k] = std::min(mn[k], tmp[i].mn[k]);
        mx[k] = std::max(mx[k], tmp[i].mx[k]);
      }
      right_area[i] = half_area(mn, mx);
    }
    for (int k = 0; k < 3; ++k) {
      mn[k] = FLT_MAX;
      mx[k] = -FLT_MAX;
    }
    for (uint32_t i = 1; i < count; ++i) {
      for (int k = 0; k < 3; ++k) {
        mn[k] = std::min(mn[k], tmp[i - 1].mn[k]);
        mx[k] = std::max(mx[k], tmp[i - 1].mx[k]);
      }
      float cost = half_area(mn, mx) * (float)i + right_area[i] * (float)(count - i);
      // ties go to the more balanced split, so that coincident primitives give a log-depth tree, not a chain
      auto off = [count](uint32_t p) { return p > count / 2 ? p - count / 2 : count / 2 - p; };
      if (cost < best_cost || (cost == best_cost && best_axis >= 0 && off(i) < off(best_pos))) {
        best_cost = cost;
        best_axis = axis;
        best_pos = i;
      }
    }
  }
  if (best_axis < 0) return false;
  std::sort(prims.begin() + first, prims.begin() + first + count, [best_axis](const BPrim& a, const BPrim& b) {
    return a.c[best_axis] < b.c[best_axis] || (a.c[best_axis] == b.c[best_axis] && a.id < b.id);
  });
  mid = first + best_pos;
  return true;
}

// `recurse` false: split this node once (its two children are appended, their boxes are still unset) and stop
void subdivide(std::vector<BPrim>& prims, std::vector<BNode>& nodes, uint32_t ni, uint32_t max_leaf, bool recurse = true) {
  const int NB = 16;
  uint32_t first = nodes[ni].leftFirst, count = nodes[ni].count;
  // bounds
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any_solo = false;
  for (uint32_t i = first; i < first + count; ++i) {
    any_solo = any_solo || prims[i].solo;
    for (int k = 0; k < 3; ++k) {
      mn[k] = std::min(mn[k], prims[i].mn[k]);
      mx[k] = std::max(mx[k], prims[i].mx[k]);
      cmn[k] = std::min(cmn[k], prims[i].c[k]);
      cmx[k] = std::max(cmx[k], prims[i].c[k]);
    }
  }
  for (int k = 0; k < 3; ++k) {
    nodes[ni].mn[k] = mn[k];
    nodes[ni].mx[k] = mx[k];
  }
  if (count <= 1) return;

  if (g_sweep) {
    float cost;
    uint32_t mid;
    float parent_area = half_area(mn, mx);
    bool ok = sweep_split(prims, first, count, cost, mid);
    bool must_split = count > max_leaf || any_solo;
    bool sah_split = ok && (cost * g_prim_cost + g_sah_ct * parent_area) < parent_area * (float)count * g_prim_cost;
    if (!must_split && !sah_split) return;
    if (!ok) mid = first + count / 2;
    uint32_t left = (uint32_t)nodes.size();
    nodes.push_back(BNode{});
    nodes.push_back(BNode{});
    nodes[left].leftFirst = first;
    nodes[left].count = mid - first;
    nodes[left + 1].leftFirst = mid;
    nodes[left + 1].count = first + count - mid;
    nodes[ni].leftFirst = left;
    nodes[ni].count = 0;
    if (recurse) {
      subdivide(prims, nodes, left, max_leaf);
      subdivide(prims, nodes, left + 1, max_leaf);
    }
    return;
  }

  int best_axis = -1, best_bin = -1;
  float best_cost = FLT_MAX;
  for (int axis = 0; axis < 3; ++axis) {
    float ext = cmx[axis] - cmn[axis];
    if (!(ext > 0.0f)) continue;
    struct Bin {
      float mn[3], mx[3];
      uint32_t n;
    } bins[NB];
    for (auto& b : bins) {
      b.n = 0;
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = FLT_MAX;
        b.mx[k] = -FLT_MAX;
      }
    }
    float scale = (float)NB / ext;
    for (uint32_t i = first; i < first + count; ++i) {
      int bi = std::min(NB - 1, (int)((prims[i].c[axis] - cmn[axis]) * scale));
      Bin& b = bins[bi];
      b.n++;
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = std::min(b.mn[k], prims[i].mn[k]);
        b.mx[k] = std::max(b.mx[k], prims[i].mx[k]);
      }
    }
    float la[NB - 1], ra[NB - 1];
    uint32_t ln[NB - 1], rn[NB - 1];
    float amn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, amx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t acc = 0;
    for (int i = 0; i < NB - 1; ++i) {
      acc += bins[i].n;
      for (int k = 0; k < 3; ++k) {
        amn[k] = std::min(amn[k], bins[i].mn[k]);
        amx[k] = std::max(amx[k], bins[i].mx[k]);
      }
      ln[i] = acc;
      la[i] = acc ? half_area(amn, amx) : 0.0f;
    }
    for (int k = 0; k < 3; ++k) {
      amn[k] = FLT_MAX;
      amx[k] = -FLT_MAX;
    }
    acc = 0;
    for (int i = NB - 1; i > 0; --i) {
      acc += bins[i].n;
      for (int k = 0; k < 3; ++k) {
        amn[k] = std::min(amn[k], bins[i].mn[k]);
        amx[k] = std::max(amx[k], bins[i].mx[k]);
      }
      rn[i - 1] = acc;
      ra[i - 1] = acc ? half_area(amn, amx) : 0.0f;
    }
    for (int i = 0; i < NB - 1; ++i) {
      if (ln[i] == 0 || rn[i] == 0) continue;
      float cost = la[i] * (float)ln[i] + ra[i] * (float)rn[i];
      if (cost < best_cost) {
        best_cost = cost;
        best_axis = axis;
        best_bin = i;
      }
    }
  }

  float parent_area = half_area(mn, mx);
  bool must_split = count > max_leaf || any_solo;
  // cost model: one node-pair visit costs g_sah_ct triangle tests
  bool sah_split = best_axis >= 0 && (best_cost * g_prim_cost + g_sah_ct * parent_area) < parent_area * (float)count * g_prim_cost;
  if (!must_split && !sah_split) return;

  uint32_t mid;
  if (best_axis >= 0) {
    float ext = cmx[best_axis] - cmn[best_axis];
    float scale = (float)NB / ext;
    auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const BPrim& p) {
      int bi = std::min(NB - 1, (int)((p.c[best_axis] - cmn[best_axis]) * scale));
      return bi <= best_bin;
    });
    mid = (uint32_t)(it - prims.begin());
  } else {
    mid = first;  // all centroids coincide
  }
  if (mid == first || mid == first + count) {
    // fall back to an object median split along the widest axis (keeps ids ascending otherwise)
    int axis = 0;
    float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
    if (e1 > e0 && e1 >= e2) axis = 1;
    if (e2 > e0 && e2 > e1) axis = 2;
    mid = first + count / 2;
    std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
                     [axis](const BPrim& a, const BPrim& b) {
                       return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.id < b.id);
                     });
  }
  uint32_t left = (uint32_t)nodes.size();
  nodes.push_back(BNode{});
  nodes.push_back(BNode{});
  nodes[left].leftFirst = first;
  nodes[left].count = mid - first;
  nodes[left + 1].leftFirst = mid;
  nodes[left + 1].count = first + count - mid;
  nodes[ni].leftFirst = left;
  nodes[ni].count = 0;
  if (recurse) {
    subdivide(prims, nodes, left, max_leaf);
    subdivide(prims, nodes, left + 1, max_leaf);
  }
}

// nodes[0] = root, nodes[1] = padding so that child pairs start on even indices (64-byte pairs)
// Big inputs are built in parallel (the reference's one-time BVH build, geometry.rs:175-217, is serial): the top of
// the tree is split serially, largest node first, until there are kBuildTasks open subtrees; each subtree is then built
// by its own thread into its own node array (the primitive ranges are disjoint) and spliced in.  The tree is the one
// the serial build produces - same splits, same boxes - only the numbering of the nodes differs.
const uint32_t kParallelMin = 8192, kBuildTasks = 16;
void build_bvh(std::vector<BPrim>& prims, uint32_t max_leaf, std::vector<BNode>& nodes) {
  nodes.clear();
  nodes.reserve(2 * prims.size() + 2);
  BNode root{};
  root.leftFirst = 0;
  root.count = (uint32_t)prims.size();
  nodes.push_back(root);
  nodes.push_back(BNode{});
  for (int k = 0; k < 3; ++k) nodes[1].mn[k] = nodes[1].mx[k] = 0.0f;
  if (prims.empty()) return;
  const unsigned hw = std::thread::hardware_concurrency();
  if (prims.size() < kParallelMin || hw < 2) {
    subdivide(prims, nodes, 0, max_leaf);
    return;
  }
  // phase 1: open subtrees, largest first
  std::vector<uint32_t> open{0}, closed;
  while (!open.empty() && open.size() + closed.size() < kBuildTasks) {
    size_t bi = 0;
    for (size_t i = 1; i < open.size(); ++i)
      if (nodes[open[i]].count > nodes[open[bi]].count) bi = i;
    const uint32_t ni = open[bi];
    open.erase(open.begin() + bi);
    if (nodes[ni].count < kParallelMin / 8) {  // small enough: finish it in phase 2 as it is
      closed.push_back(ni);
      continue;
    }
    subdivide(prims, nodes, ni, max_leaf, false);
    if (nodes[ni].count == 0) {  // it was split: its children are open now
      open.push_back(nodes[ni].leftFirst);
      open.push_back(nodes[ni].leftFirst + 1);
    }                            // else: it stays a leaf, done
  }
  closed.insert(closed.end(), open.begin(), open.end());
  // phase 2: one thread per subtree, each into its own array (local node 0 = the subtree's root, 1 = padding)
  std::vector<std::vector<BNode>> sub(closed.size());
  std::vector<std::thread> workers;
  for (size_t t = 0; t < closed.size(); ++t)
    workers.emplace_back([&, t] {
      sub[t].reserve(2 * (size_t)nodes[closed[t]].count + 2);
      sub[t].push_back(nodes[closed[t]]);
      sub[t].push_back(BNode{});
      subdivide(prims, sub[t], 0, max_leaf);
    });
  for (auto& w : workers) w.join();
  // phase 3: splice (local child indices start at 2)
  for (size_t t = 0; t < closed.size(); ++t) {
    const uint32_t base = (uint32_t)nodes.size();
    for (size_t i = 2; i < sub[t].size(); ++i) {
      BNode n = sub[t][i];
      if (!n.count) n.leftFirst = n.leftFirst - 2 + base;
      nodes.push_back(n);
    }
    BNode r = sub[t][0];
    if (!r.count) r.leftFirst = r.leftFirst - 2 + base;
    nodes[closed[t]] = r;
  }
}

// Children per interior node: 2 (adjacent pairs, 64 B per fetch) or, with -DRT_BVH4=1, 4 (adjacent groups of four
// 32-byte records, 128 B per fetch; built by collapsing the binary tree).  Measured on the device: DESIGN.md §2.
#ifndef RT_BVH4
#define RT_BVH4 0
#endif
#ifndef RT_NODE_CH  // how a child pair is stored: see rt_config.cuh (same default here)
#if RT_BVH4
#define RT_NODE_CH 0
#else
#define RT_NODE_CH 2
#endif
#endif
const uint32_t kWidth = RT_BVH4 ? 4u : 2u;
const uint32_t kNoChild = 0xFFFFFFFFu;  // link of an unused slot of a 4-wide group (count == 0)
inline bool empty_slot(const BNode& n) { return n.count == 0 && n.leftFirst == kNoChild; }

// 1 + the most entries the traversal stack can hold below node `ni`: visiting a group pushes all its children but
// the one it descends into.  For the binary tree this is the longest root-to-leaf path, in nodes.
uint32_t bvh_depth(const std::vector<BNode>& nodes, uint32_t ni = 0) {
  if (nodes.empty() || nodes[ni].count || empty_slot(nodes[ni])) return 1;
  uint32_t g = nodes[ni].leftFirst, live = 0, below = 0;
  for (uint32_t k = 0; k < kWidth; ++k) {
    if (empty_slot(nodes[g + k])) continue;
    ++live;
    below = std::max(below, bvh_depth(nodes, g + k));
  }
  return (live ? live - 1 : 0) + below;
}

// Binary tree (root at 0, child pairs) -> 4-wide tree (root at 0, records 1..3 padding, then one group of four adjacent
// child records per interior node, breadth first).  A group starts as the two children of its node; while it has a free
// slot, the interior child with the largest surface area is replaced by its own two children.
void collapse4(std::vector<BNode>& nodes) {
  BNode none{};
  none.leftFirst = kNoChild;
  none.count = 0;
  BNode pad{};
  std::vector<BNode> out;
  out.reserve(nodes.size() * 2 + 4);
  out.push_back(nodes.empty() ? pad : nodes[0]);
  for (int k = 0; k < 3; ++k) out.push_back(pad);
  if (nodes.size() <= 2 || nodes[0].count) {
    nodes.swap(out);
    return;
  }
  std::vector<uint32_t> queue{0};  // records of `out` whose link still points into the binary tree
  for (size_t qi = 0; qi < queue.size(); ++qi) {
    const uint32_t oi = queue[qi];
    if (out[oi].count || empty_slot(out[oi])) continue;
    std::vector<uint32_t> slots{out[oi].leftFirst, out[oi].leftFirst + 1};
    while (slots.size() < 4) {
      int best = -1;
      float best_area = -1.0f;
      for (size_t k = 0; k < slots.size(); ++k) {
        const BNode& c = nodes[slots[k]];
        if (c.count) continue;
        float a = half_area(c.mn, c.mx);
        if (a > best_area) {
          best_area = a;
          best = (int)k;
        }
      }
      if (best < 0) break;
      uint32_t l = nodes[slots[best]].leftFirst;
      slots[best] = l;
      slots.push_back(l + 1);
    }
    const uint32_t g = (uint32_t)out.size();
    for (uint32_t k = 0; k < 4; ++k) out.push_back(k < slots.size() ? nodes[slots[k]] : none);
    out[oi].leftFirst = g;
    for (uint32_t k = 0; k < 4; ++k) queue.push_back(g + k);
  }
  nodes.swap(out);
}

// breadth-first order (root, padding, then child pairs level by level): the first N nodes are the top of the tree,
// which is what a kernel wants to keep in shared memory
void reorder_bfs(std::vector<BNode>& nodes) {
  if (nodes.size() <= 2) return;
  std::vector<BNode> out;
  out.reserve(nodes.size());
  out.push_back(nodes[0]);
  out.push_back(nodes[1]);
  std::vector<uint32_t> queue{0};  // indices into `out` of nodes whose children still have to be placed
  for (size_t qi = 0; qi < queue.size(); ++qi) {
    uint32_t ni = queue[qi];
    if (out[ni].count) continue;
    uint32_t old_left = out[ni].leftFirst, new_left = (uint32_t)out.size();
    out.push_back(nodes[old_left]);
    out.push_back(nodes[old_left + 1]);
    out[ni].leftFirst = new_left;
    queue.push_back(new_left);
    queue.push_back(new_left + 1);
  }
  nodes.swap(out);
}

inline uint32_t pack_entry(uint32_t leftFirst, uint32_t count) {
  return count ? (RT_LEAF_FLAG | (leftFirst << 4) | count) : leftFirst;
}

void nodes_to_quads(const std::vector<BNode>& nodes, float pad, std::vector<Quad>& out) {
  out.resize(nodes.size() * 2);
  for (size_t i = 0; i < nodes.size(); ++i) {
    Quad lo, hi;
    for (int k = 0; k < 3; ++k) {
      lo.f[k] = nodes[i].mn[k] - pad;
      hi.f[k] = nodes[i].mx[k] + pad;
    }
    lo.u[3] = nodes[i].leftFirst;
    hi.u[3] = nodes[i].count;
    out[2 * i] = lo;
    out[2 * i + 1] = hi;
  }
}
}  // namespace

// ------------------------------------------------------------------ mesh -> BLAS + records
void build_mesh(HostMesh& m) {
  const uint32_t nt = m.ntris();
  m.reach.assign(nt, 1);
  std::vector<GuardRange> granges;
  if (nt) reach_rec(m.pos.data(), m.idx.data(), 0, nt, m.reach.data(), &granges);
  // per-triangle guard lists: every thin interior box above a reachable triangle, outermost first
  // (reach_rec emits children before parents, so walk the ranges backwards)
  std::vector<std::vector<uint32_t>> tri_guards(nt);
  m.guards.clear();
  for (size_t gi = granges.size(); gi-- > 0;) {
    const GuardRange& g = granges[gi];
    bool any = false;
    for (uint32_t t = g.start; t < g.end; ++t) any = any || m.reach[t];
    if (!any) continue;
    uint32_t id = (uint32_t)(m.guards.size() / 2);
    Quad lo{}, hi{};
    for (int k = 0; k < 3; ++k) {
      lo.f[k] = g.box.mn[k];
      hi.f[k] = g.box.mx[k];
    }
    m.guards.push_back(lo);
    m.guards.push_back(hi);
    for (uint32_t t = g.start; t < g.end; ++t)
      if (m.reach[t]) tri_guards[t].push_back(id);
  }
  // flatten: guard_list holds, per triangle, the indices of its guard boxes
  m.guard_list.clear();
  std::vector<uint32_t> g_first(nt, 0), g_count(nt, 0);
  for (uint32_t t = 0; t < nt; ++t) {
    g_first[t] = (uint32_t)m.guard_list.size();
    g_count[t] = (uint32_t)tri_guards[t].size();
    m.guard_list.insert(m.guard_list.end(), tri_guards[t].begin(), tri_guards[t].end());
  }

  // shading records, original order (geometry.rs:230-250,350-363)
  m.shade.assign((size_t)nt * RT_SHADE_QUADS, Quad{});
  std::vector<BPrim> prims;
  prims.reserve(nt);
  float amax = 0.0f;
  for (uint32_t t = 0; t < nt; ++t) {
    uint32_t i0 = m.idx[3 * t], i1 = m.idx[3 * t + 1], i2 = m.idx[3 * t + 2];
    const float *pa = &m.pos[3 * (size_t)i0], *pb = &m.pos[3 * (size_t)i1], *pc = &m.pos[3 * (size_t)i2];
    const float *na = &m.nrm[3 * (size_t)i0], *nb = &m.nrm[3 * (size_t)i1], *nc = &m.nrm[3 * (size_t)i2];
    const float *ta = &m.uv[2 * (size_t)i0], *tb = &m.uv[2 * (size_t)i1], *tc = &m.uv[2 * (size_t)i2];
    // StaticMesh::get_tangent, geometry.rs:245-250
    float u1 = ta[0], u2 = tb[0], u3 = tc[0], v1 = ta[1], v2 = tb[1], v3 = tc[1];
    float den = (u2 - u1) * (v3 - v1) - (v2 - v1) * (u3 - u1);
    float tan[3];
    for (int k = 0; k < 3; ++k) tan[k] = ((v3 - v1) * (pb[k] - pa[k]) - (v2 - v1) * (pc[k] - pa[k])) / den;
    float rec[20] = {na[0], na[1], na[2], nb[0], nb[1], nb[2], nc[0], nc[1], nc[2], ta[0],
                     ta[1], tb[0], tb[1], tc[0], tc[1], tan[0], tan[1], tan[2], 0.0f, 0.0f};
    std::memcpy(&m.shade[(size_t)t * RT_SHADE_QUADS], rec, sizeof rec);
    if (!m.reach[t]) continue;
    BPrim p;
    for (int k = 0; k < 3; ++k) {
      p.mn[k] = fmin3(pa[k], pb[k], pc[k]);
      p.mx[k] = fmax3(pa[k], pb[k], pc[k]);
      p.c[k] = 0.5f * (p.mn[k] + p.mx[k]);
      amax = std::max(amax, std::max(std::fabs(p.mn[k]), std::fabs(p.mx[k])));
    }
    p.id = t;
    prims.push_back(p);
  }
  m.n_reachable = (uint32_t)prims.size();

  std::vector<BNode> nodes;
  g_prim_cost = 1.0f;
  build_bvh(prims, g_leaf_max, nodes);
  if (RT_BVH4) collapse4(nodes);
  // Conservative traversal: boxes are padded by a few ulps of the largest coordinate so that a
  // triangle the reference's Möller–Trumbore test accepts is never culled by rounding in the
  // slab test (flat boxes of coplanar triangles get thickness this way, cf. Q3).
  m.depth = prims.empty() ? 0 : bvh_depth(nodes);
  // the padding itself is applied when the scene is lowered, because it depends on how the mesh is instanced
  m.pad_base = 4e-6f * amax + 1e-30f;
  m.amax = amax;
  nodes_to_quads(nodes, 0.0f, m.nodes);
  for (int k = 0; k < 3; ++k) {
    m.root_min[k] = prims.empty() ? 0.0f : nodes[0].mn[k];
    m.root_max[k] = prims.empty() ? 0.0f : nodes[0].mx[k];
  }
  m.root_entry_local = prims.empty() ? RT_ENTRY_NONE : pack_entry(nodes[0].leftFirst, nodes[0].count);

  // intersection records in leaf order: v0, e1 = b - a, e2 = c - a (geometry.rs:336-337), id
  m.tris.assign(prims.size() * RT_TRI_QUADS, Quad{});
  for (size_t i = 0; i < prims.size(); ++i) {
    uint32_t t = prims[i].id;
    uint32_t i0 = m.idx[3 * t], i1 = m.idx[3 * t + 1], i2 = m.idx[3 * t + 2];
    const float *pa = &m.pos[3 * (size_t)i0], *pb = &m.pos[3 * (size_t)i1], *pc = &m.pos[3 * (size_t)i2];
    Quad* q = &m.tris[i * RT_TRI_QUADS];
    q[0].f[0] = pa[0]; q[0].f[1] = pa[1]; q[0].f[2] = pa[2];
    q[0].f[3] = pb[0] - pa[0];
    q[1].f[0] = pb[1] - pa[1]; q[1].f[1] = pb[2] - pa[2];
    q[1].f[2] = pc[0] - pa[0]; q[1].f[3] = pc[1] - pa[1];
    q[2].f[0] = pc[2] - pa[2];
    q[2].u[1] = t;
    q[2].u[2] = g_first[t];   // guard list (rebased when the BLASes are concatenated)
    q[2].u[3] = g_count[t];
  }
}

// ------------------------------------------------------------------ matrices
// general 4x4 inverse by cofactors (what cgmath's Matrix4::invert does), f32
bool invert_affine_cofactor(const float m[16], float out[16]) {
  float inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0f || det != det) return false;
  float id = 1.0f / det;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * id;
  return true;
}

// ------------------------------------------------------------------ scene -> flat arrays
int lower_scene(const std::vector<HostTexture>& textures, const std::vector<rt_material_desc>& materials,
                std::vector<HostMesh>& meshes, const std::vector<HostObject>& objects, Lowered& L,
                std::string& err) {
  L = Lowered();
  // BLASes, concatenated; child and triangle indices rebased to the global arrays
  std::vector<uint32_t> node_base(meshes.size()), tri_base(meshes.size()), shade_base(meshes.size());
  // Box padding per mesh.  The slab test's rounding error is a few ulps of the OBJECT-SPACE ray origin, which for a
  // tiny instance in a big scene is far larger than the mesh itself (C5: drones scaled by 6e-4, origins 3e4 units
  // away from a mesh of extent 500).  So the base padding (4e-6 of the largest coordinate) grows with the ratio of
  // the scene's extent, measured in each instance's object units, to the mesh's own extent.
  float world_mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, world_mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (const HostObject& o : objects) {
    float c[3], r = 0.0f;
    bool has = true;
    if (o.kind == RT_OBJ_MESH || o.kind == RT_OBJ_VOLUME_MESH) {
      if (o.mesh < 0 || o.mesh >= (int)meshes.size()) continue;
      const HostMesh& m = meshes[o.mesh];
      for (int k = 0; k < 3; ++k) c[k] = o.xform[12 + k];
      float sc = 0.0f;
      for (int k = 0; k < 9; ++k) sc = std::max(sc, std::fabs(o.xform[(k / 3) * 4 + k % 3]));
      r = 1.8f * sc * m.amax;
    } else if (o.kind == RT_OBJ_SPHERE || o.kind == RT_OBJ_VOLUME) {
      for (int k = 0; k < 3; ++k) c[k] = o.a[k];
      r = std::fabs(o.radius);
    } else if (o.kind == RT_OBJ_TRIANGLE) {
      for (int k = 0; k < 3; ++k) {
        c[k] = o.a[k];
        r = std::max(r, std::max(std::fabs(o.b[k] - o.a[k]), std::fabs(o.c[k] - o.a[k])));
      }
    } else {
      has = false;
    }
    if (!has) continue;
    for (int k = 0; k < 3; ++k) {
      if (std::isfinite(c[k]) && std::isfinite(r)) {
        world_mn[k] = std::min(world_mn[k], c[k] - r);
        world_mx[k] = std::max(world_mx[k], c[k] + r);
      }
    }
  }
  float world_ext = 0.0f;
  for (int k = 0; k < 3; ++k)
    if (world_mx[k] >= world_mn[k]) world_ext = std::max(world_ext, world_mx[k] - world_mn[k]);
  std::vector<float> pad_factor(meshes.size(), 1.0f);
  for (const HostObject& o : objects) {
    if ((o.kind != RT_OBJ_MESH && o.kind != RT_OBJ_VOLUME_MESH) || o.mesh < 0 || o.mesh >= (int)meshes.size()) continue;
    float isc = 0.0f;  // largest entry of the inverse's linear part ~ object units per world unit
    for (int k = 0; k < 9; ++k) isc = std::max(isc, std::fabs(o.inv_xform[(k / 3) * 4 + k % 3]));
    float ratio = 2.0f * world_ext * isc / std::max(meshes[o.mesh].amax, 1e-30f);  // rays may start ~2 extents away
    if (std::isfinite(ratio)) pad_factor[o.mesh] = std::max(pad_factor[o.mesh], std::min(ratio / 8.0f, 256.0f));
  }
  for (size_t mi = 0; mi < meshes.size(); ++mi) {
    HostMesh& m = meshes[mi];
    m.pad = m.pad_base * pad_factor[mi];
    node_base[mi] = (uint32_t)(L.nodes.size() / RT_NODE_QUADS);
    tri_base[mi] = (uint32_t)(L.tris.size() / RT_TRI_QUADS);
    shade_base[mi] = (uint32_t)(L.shade.size() / RT_SHADE_QUADS);
    size_t n0 = L.nodes.size();
    L.nodes.insert(L.nodes.end(), m.nodes.begin(), m.nodes.end());
    for (size_t q = n0; q < L.nodes.size(); q += 2) {
      uint32_t count = L.nodes[q + 1].u[3];
      if (count || L.nodes[q].u[3] != kNoChild) L.nodes[q].u[3] += count ? tri_base[mi] : node_base[mi];
      for (int k = 0; k < 3; ++k) {
        L.nodes[q].f[k] -= m.pad;
        L.nodes[q + 1].f[k] += m.pad;
      }
    }
    // guards: boxes and per-triangle index lists, rebased to the global arrays
    uint32_t guard_base = (uint32_t)(L.guards.size() / 2), glist_base = (uint32_t)L.guard_list.size();
    L.guards.insert(L.guards.end(), m.guards.begin(), m.guards.end());
    for (uint32_t g : m.guard_list) L.guard_list.push_back(g + guard_base);
    size_t t0 = L.tris.size();
    L.tris.insert(L.tris.end(), m.tris.begin(), m.tris.end());
    for (size_t q = t0; q < L.tris.size(); q += RT_TRI_QUADS)
      if (L.tris[q + 2].u[3]) L.tris[q + 2].u[2] += glist_base;
    L.shade.insert(L.shade.end(), m.shade.begin(), m.shade.end());
    if ((L.tris.size() / RT_TRI_QUADS) >= (1u << 27)) {
      err = "too many triangles";
      return RT_ERR_UNSUPPORTED;
    }
  }

  // materials: (albedo.xyz, roughness | ior) (emission.xyz, metallic)
  L.mats.assign(materials.size() * RT_MAT_QUADS, Quad{});
  for (size_t i = 0; i < materials.size(); ++i) {
    const rt_material_desc& d = materials[i];
    Quad* q = &L.mats[i * RT_MAT_QUADS];
    for (int k = 0; k < 3; ++k) {
      q[0].f[k] = d.albedo[k];
      q[1].f[k] = d.tag == RT_MAT_DIELECTRIC ? 0.0f : d.emission[k];
    }
    q[0].f[3] = d.tag == RT_MAT_DIELECTRIC ? d.ior : d.roughness;
    q[1].f[3] = d.metallic;
  }

  // textures
  L.textures.assign(textures.size(), Quad{});
  for (size_t i = 0; i < textures.size(); ++i) {
    L.textures[i].u[0] = (uint32_t)L.texels.size();
    L.textures[i].u[1] = textures[i].w;
    L.textures[i].u[2] = textures[i].h;
    L.texels.insert(L.texels.end(), textures[i].rgba.begin(), textures[i].rgba.end());
  }
  if (L.texels.empty()) L.texels.push_back(0);

  // objects + TLAS primitives
  L.objects.assign(objects.size() * RT_OBJ_QUADS, Quad{});
  std::vector<BPrim> tprims;
  int nvol = 0;
  for (size_t oi = 0; oi < objects.size(); ++oi) {
    const HostObject& o = objects[oi];
    Quad* q = &L.objects[oi * RT_OBJ_QUADS];
    q[0].i[0] = o.kind;
    q[0].i[1] = o.material;
    if (o.material >= (int)materials.size()) {
      err = "object references a material id that does not exist";
      return RT_ERR_INVALID;
    }
    q[0].i[2] = o.material >= 0 ? (int)materials[o.material].tag : RT_CLASS_PARAM_TEX;
    q[0].i[3] = 0;
    BPrim p;
    bool bounded = true;
    switch (o.kind) {
      case RT_OBJ_VOLUME_MESH:
      case RT_OBJ_MESH: {
        if (o.mesh < 0 || o.mesh >= (int)meshes.size()) {
          err = "instance references a mesh id that does not exist";
          return RT_ERR_INVALID;
        }
        for (int k = 0; k < 5; ++k)
          if (o.tex[k] >= (int)textures.size()) {
            err = "instance references a texture id that does not exist";
            return RT_ERR_INVALID;
          }
        const HostMesh& m = meshes[o.mesh];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 4; ++c) {
            q[1 + r].f[c] = o.inv_xform[c * 4 + r];
            q[4 + r].f[c] = o.xform[c * 4 + r];
          }
        uint32_t re = m.root_entry_local;
        if (re != RT_ENTRY_NONE) {
          if (re & RT_LEAF_FLAG) {
            uint32_t first = (re & ~RT_LEAF_FLAG) >> 4, cnt = re & 15u;
            re = pack_entry(first + tri_base[o.mesh], cnt);
          } else {
            re += node_base[o.mesh];
          }
        }
        q[7].u[0] = re;
        q[7].u[1] = shade_base[o.mesh];
        q[7].i[2] = o.tex[0];
        q[7].i[3] = o.tex[1];
        q[8].i[0] = o.tex[2];
        q[8].i[1] = o.tex[3];
        q[8].i[2] = o.tex[4];
        q[8].i[3] = 0;
        if (o.kind == RT_OBJ_VOLUME_MESH) {
          q[9].f[0] = o.density;
          q[9].i[1] = nvol++;
        }
        if (m.n_reachable == 0) {
          bounded = false;  // nothing to hit: not in the TLAS, not in the plane list either
          break;
        }
        // world box = box of the transformed vertices of the reachable triangles (a rotated instance gets a much
        // tighter box than the box of its object-space box's corners would), grown by the object-space padding
        // pushed through the linear part of the transform
        for (int k = 0; k < 3; ++k) {
          p.mn[k] = FLT_MAX;
          p.mx[k] = -FLT_MAX;
        }
        for (uint32_t t = 0; t < m.ntris(); ++t) {
          if (!m.reach[t]) continue;
          for (int c = 0; c < 3; ++c) {
            const float* v = &m.pos[3 * (size_t)m.idx[3 * t + c]];
            for (int k = 0; k < 3; ++k) {
              float w = o.xform[k] * v[0] + o.xform[4 + k] * v[1] + o.xform[8 + k] * v[2] + o.xform[12 + k];
              p.mn[k] = std::min(p.mn[k], w);
              p.mx[k] = std::max(p.mx[k], w);
            }
          }
        }
        {
          float amax_obj = 0.0f;
          for (int k = 0; k < 3; ++k) amax_obj = std::max(amax_obj, std::max(std::fabs(m.root_min[k]), std::fabs(m.root_max[k])));
          float pad_obj = 2.0f * m.pad + 8e-6f * amax_obj;
          for (int k = 0; k < 3; ++k) {
            float g = (std::fabs(o.xform[k]) + std::fabs(o.xform[4 + k]) + std::fabs(o.xform[8 + k])) * pad_obj;
            p.mn[k] -= g;
            p.mx[k] += g;
          }
        }
        break;
      }
      case RT_OBJ_SPHERE:
      case RT_OBJ_VOLUME: {
        q[1].f[0] = o.a[0]; q[1].f[1] = o.a[1]; q[1].f[2] = o.a[2]; q[1].f[3] = o.radius;
        if (o.kind == RT_OBJ_VOLUME) {
          q[2].f[0] = o.density;
          q[2].i[1] = nvol++;
        }
        float r = std::fabs(o.radius);
        for (int k = 0; k < 3; ++k) {
          p.mn[k] = o.a[k] - r;
          p.mx[k] = o.a[k] + r;
        }
        break;
      }
      case RT_OBJ_TRIANGLE: {
        float e1[3], e2[3];
        for (int k = 0; k < 3; ++k) {
          e1[k] = o.b[k] - o.a[k];
          e2[k] = o.c[k] - o.a[k];
        }
        // normalize(e1 x e2), geometry.rs:449
        float cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        float inv = 1.0f / std::sqrt(cx * cx + cy * cy + cz * cz);
        q[1].f[0] = o.a[0]; q[1].f[1] = o.a[1]; q[1].f[2] = o.a[2]; q[1].f[3] = e1[0];
        q[2].f[0] = e1[1]; q[2].f[1] = e1[2]; q[2].f[2] = e2[0]; q[2].f[3] = e2[1];
        q[3].f[0] = e2[2]; q[3].f[1] = cx * inv; q[3].f[2] = cy * inv; q[3].f[3] = cz * inv;
        for (int k = 0; k < 3; ++k) {
          p.mn[k] = fmin3(o.a[k], o.b[k], o.c[k]);
          p.mx[k] = fmax3(o.a[k], o.b[k], o.c[k]);
        }
        break;
      }
      case RT_OBJ_PLANE: {
        q[1].f[0] = o.a[0]; q[1].f[1] = o.a[1]; q[1].f[2] = o.a[2];
        q[2].f[0] = o.b[0]; q[2].f[1] = o.b[1]; q[2].f[2] = o.b[2];
        bounded = false;
        L.planes.push_back((int32_t)oi);
        break;
      }
      default:
        err = "unknown object kind";
        return RT_ERR_INVALID;
    }
    if (bounded) {
      bool finite = true;
      for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(p.mn[k]) && std::isfinite(p.mx[k]);
      if (!finite) {
        // NaN / infinite centre, radius, vertex or transform: such an object cannot be bounded, and the always-tested
        // list only knows planes.  Refused loudly rather than silently never intersected.
        err = "object " + std::to_string(oi) + " has non-finite coordinates (it cannot be placed in the TLAS)";
        return RT_ERR_UNSUPPORTED;
      }
      float amax = 0.0f;
      for (int k = 0; k < 3; ++k) amax = std::max(amax, std::max(std::fabs(p.mn[k]), std::fabs(p.mx[k])));
      float pad = 1e-5f * amax + 1e-30f;
      for (int k = 0; k < 3; ++k) {
        p.mn[k] -= pad;
        p.mx[k] += pad;
        p.c[k] = 0.5f * (p.mn[k] + p.mx[k]);
      }
      p.id = (uint32_t)oi;
      p.solo = (o.kind == RT_OBJ_MESH || o.kind == RT_OBJ_VOLUME_MESH) ? 1u : 0u;
      tprims.push_back(p);
    }
  }
  L.n_volumes = (uint32_t)nvol;
  if (L.planes.empty()) L.planes.push_back(-1);  // keep the buffer non-empty; n_planes stays 0
  if (objects.empty()) L.objects.push_back(Quad{});
  if (L.mats.empty()) L.mats.assign(RT_MAT_QUADS, Quad{});
  if (L.textures.empty()) L.textures.push_back(Quad{});
  if (L.tris.empty()) L.tris.assign(RT_TRI_QUADS, Quad{});
  if (L.guards.empty()) L.guards.assign(2, Quad{});
  if (L.guard_list.empty()) L.guard_list.push_back(0);
  if (L.shade.empty()) L.shade.assign(RT_SHADE_QUADS, Quad{});

  // TLAS: one object per leaf, the leaf's link holds the object index directly.  (Measured: leaves of 2-8 analytic
  // objects behind an index list cut TLAS node visits from 9.8 to 9.0 per ray but were 1-2 % slower.)
  if (!tprims.empty()) {
    g_prim_cost = 1.0f;
    std::vector<BNode> tn;
    build_bvh(tprims, 1, tn);
    if (RT_BVH4) collapse4(tn);  // emits breadth first as well
    else reorder_bfs(tn);
    L.tlas_depth = bvh_depth(tn);
    for (auto& n : tn)
      if (n.count) n.leftFirst = tprims[n.leftFirst].id;  // leaf -> object index
    uint32_t base = (uint32_t)(L.nodes.size() / RT_NODE_QUADS);
    std::vector<Quad> tq;
    nodes_to_quads(tn, 0.0f, tq);
    for (size_t q = 0; q < tq.size(); q += 2)
      if (tq[q + 1].u[3] == 0 && tq[q].u[3] != kNoChild) tq[q].u[3] += base;
    L.nodes.insert(L.nodes.end(), tq.begin(), tq.end());
    L.tlas_base = base;
    L.tlas_count = (uint32_t)tn.size();
    L.tlas_root = tn[0].count ? pack_entry(tn[0].leftFirst, tn[0].count) : (tn[0].leftFirst + base);
    for (int k = 0; k < 3; ++k) {
      L.tlas_min[k] = tn[0].mn[k];
      L.tlas_max[k] = tn[0].mx[k];
    }
  }
  {
    // the traversal stack (64 entries: RT_SMEM_STACK in shared memory, the rest in local memory) must hold the deepest path:
    // TLAS depth + one RESTORE marker + the deepest BLAS
    uint32_t deepest = 0;
    for (const HostObject& o : objects)
      if (o.kind == RT_OBJ_MESH || o.kind == RT_OBJ_VOLUME_MESH) deepest = std::max(deepest, meshes[o.mesh].depth);
    if (L.tlas_depth + 1 + deepest > 62) {
      err = "BVH too deep for the traversal stack (" + std::to_string(L.tlas_depth + 1 + deepest) + " > 62)";
      return RT_ERR_UNSUPPORTED;
    }
  }
  if (L.nodes.empty()) L.nodes.assign(RT_NODE_QUADS * 4, Quad{});
  // final form of the link word: the packed traversal entry (leaf flag | first << 4 | count, or the
  // index of the child pair), so the kernel uses lo.w as is; hi.w keeps the plain count
  for (size_t q = 0; q < L.nodes.size(); q += 2) L.nodes[q].u[3] = pack_entry(L.nodes[q].u[3], L.nodes[q + 1].u[3]);
#if RT_NODE_CH
  // (min, max) -> (centre, half extent), the half extent rounded up so that [c - h, c + h] contains [min, max]
  auto to_ch = [](float lo, float hi, float& c, float& h) {
    if (!(lo <= hi)) {  // empty or NaN slot: can never be hit
      c = 0.0f;
      h = -1.0f;
      return;
    }
    if (!std::isfinite(lo) || !std::isfinite(hi)) {  // unbounded along this axis: never culls
      c = 0.0f;
      h = INFINITY;
      return;
    }
    c = 0.5f * lo + 0.5f * hi;
    const double need = std::max((double)hi - (double)c, (double)c - (double)lo);
    h = (float)need;
    if ((double)h < need) h = std::nextafter(h, INFINITY);
  };
#if RT_NODE_CH == 2
  // packed child pairs: quad 0 = (left centre, left link), quad 1 = (right centre, right link), quad 2 = the six half
  // extents as bf16, rounded up; quad 3 of the 64-byte slot is never fetched (the plain counts stay there)
  auto bf16_up = [](float h) -> uint32_t {
    uint32_t b;
    std::memcpy(&b, &h, 4);
    if (h > 0.0f && (b & 0xFFFFu)) b = (b | 0xFFFFu) + 1u;  // next bf16 above (carries into the exponent as it should)
    return b >> 16;
  };
  for (size_t q = 0; q + 3 < L.nodes.size(); q += 4) {
    float c[2][3], h[2][3];
    for (int s = 0; s < 2; ++s)
      for (int k = 0; k < 3; ++k) to_ch(L.nodes[q + 2 * s].f[k], L.nodes[q + 2 * s + 1].f[k], c[s][k], h[s][k]);
    const uint32_t link_l = L.nodes[q].u[3], link_r = L.nodes[q + 2].u[3];
    const uint32_t count_l = L.nodes[q + 1].u[3], count_r = L.nodes[q + 3].u[3];
    Quad q0, q1, q2, q3;
    for (int k = 0; k < 3; ++k) {
      q0.f[k] = c[0][k];
      q1.f[k] = c[1][k];
    }
    q0.u[3] = link_l;
    q1.u[3] = link_r;
    const uint32_t hb[6] = {bf16_up(h[0][0]), bf16_up(h[0][1]), bf16_up(h[0][2]), bf16_up(h[1][0]), bf16_up(h[1][1]), bf16_up(h[1][2])};
    q2.u[0] = hb[0] | (hb[1] << 16);
    q2.u[1] = hb[2] | (hb[3] << 16);
    q2.u[2] = hb[4] | (hb[5] << 16);
    q2.u[3] = 0;
    q3.u[0] = count_l; q3.u[1] = count_r; q3.u[2] = q3.u[3] = 0;
    L.nodes[q] = q0; L.nodes[q + 1] = q1; L.nodes[q + 2] = q2; L.nodes[q + 3] = q3;
  }
#else
  for (size_t q = 0; q < L.nodes.size(); q += 2)
    for (int k = 0; k < 3; ++k) {
      float c, h;
      to_ch(L.nodes[q].f[k], L.nodes[q + 1].f[k], c, h);
      L.nodes[q].f[k] = c;
      L.nodes[q + 1].f[k] = h;
    }
#endif
#endif
  return RT_OK;
}

// ------------------------------------------------------------------ OBJ (tobj 3.2.0 semantics)
// load_obj with single_index + triangulate (geometry.rs:140-148): faces are fan-triangulated
// (0,i,i+1); vertices are de-duplicated on the (v,vt,vn) triple in first-seen order; a new
// `o`/`g` statement after faces have been seen starts a new model and only the first model is
// used (geometry.rs:157).
namespace {
struct Cursor {
  const char* p;
  const char* end;
};
inline void skip_ws(Cursor& c) {
  while (c.p < c.end && (*c.p == ' ' || *c.p == '\t' || *c.p == '\r')) ++c.p;
}
inline bool parse_float(Cursor& c, float& out) {
  skip_ws(c);
  if (c.p >= c.end || *c.p == '\n') return false;
  char buf[64];
  size_t n = 0;
  while (c.p < c.end && n < 63 && *c.p != ' ' && *c.p != '\t' && *c.p != '\r' && *c.p != '\n') buf[n++] = *c.p++;
  buf[n] = 0;
  char* e = nullptr;
  out = std::strtof(buf, &e);
  return e != buf;
}
inline bool parse_int(Cursor& c, long& out) {
  if (c.p >= c.end) return false;
  char* e = nullptr;
  char buf[32];
  size_t n = 0;
  const char* q = c.p;
  while (q < c.end && n < 31 && (*q == '-' || *q == '+' || (*q >= '0' && *q <= '9'))) buf[n++] = *q++;
  buf[n] = 0;
  if (n == 0) return false;
  out = std::strtol(buf, &e, 10);
  if (e == buf) return false;
  c.p = q;
  return true;
}
struct VKey {
  long v, vt, vn;
  bool operator==(const VKey& o) const { return v == o.v && vt == o.vt && vn == o.vn; }
};
struct VKeyHash {
  size_t operator()(const VKey& k) const {
    uint64_t h = (uint64_t)k.v * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.vt * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (uint64_t)k.vn * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};
}  // namespace

int obj_parse(const char* text, size_t len, rt_obj_mesh* out, std::string& err) {
  std::vector<float> P, T, N;           // file-order v / vt / vn
  std::vector<float> pos, uv, nrm;      // de-duplicated outputs
  std::vector<uint32_t> idx;
  std::unordered_map<VKey, uint32_t, VKeyHash> map;
  bool any_vt = false, any_vn = false, missing_vt = false, missing_vn = false;
  bool have_faces = false;
  Cursor c{text, text + len};
  std::vector<VKey> face;
  auto add_vertex = [&](const VKey& k) {
    auto it = map.find(k);
    if (it != map.end()) {
      idx.push_back(it->second);
      return;
    }
    uint32_t id = (uint32_t)map.size();
    map.emplace(k, id);
    pos.push_back(P[3 * k.v]); pos.push_back(P[3 * k.v + 1]); pos.push_back(P[3 * k.v + 2]);
    if (k.vt >= 0) { uv.push_back(T[2 * k.vt]); uv.push_back(T[2 * k.vt + 1]); any_vt = true; }
    else { uv.push_back(0.0f); uv.push_back(0.0f); missing_vt = true; }
    if (k.vn >= 0) { nrm.push_back(N[3 * k.vn]); nrm.push_back(N[3 * k.vn + 1]); nrm.push_back(N[3 * k.vn + 2]); any_vn = true; }
    else { nrm.push_back(0.0f); nrm.push_back(0.0f); nrm.push_back(0.0f); missing_vn = true; }
    idx.push_back(id);
  };
  bool stop = false;
  while (c.p < c.end && !stop) {
    skip_ws(c);
    const char* line = c.p;
    const char* eol = line;
    while (eol < c.end && *eol != '\n') ++eol;
    Cursor l{line, eol};
    size_t n = (size_t)(eol - line);
    if (n >= 2 && line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
      l.p += 2;
      float x, y, z;
      if (!parse_float(l, x) || !parse_float(l, y) || !parse_float(l, z)) { err = "bad v line"; return RT_ERR_IO; }
      P.push_back(x); P.push_back(y); P.push_back(z);
    } else if (n >= 3 && line[0] == 'v' && line[1] == 't' && (line[2] == ' ' || line[2] == '\t')) {
      l.p += 3;
      float u, v;
      if (!parse_float(l, u) || !parse_float(l, v)) { err = "bad vt line"; return RT_ERR_IO; }
      T.push_back(u); T.push_back(v);
    } else if (n >= 3 && line[0] == 'v' && line[1] == 'n' && (line[2] == ' ' || line[2] == '\t')) {
      l.p += 3;
      float x, y, z;
      if (!parse_float(l, x) || !parse_float(l, y) || !parse_float(l, z)) { err = "bad vn line"; return RT_ERR_IO; }
      N.push_back(x); N.push_back(y); N.push_back(z);
    } else if (n >= 2 && line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
      l.p += 2;
      face.clear();
      for (;;) {
        skip_ws(l);
        if (l.p >= l.end) break;
        VKey k{-1, -1, -1};
        long v;
        if (!parse_int(l, v)) { err = "bad f line"; return RT_ERR_IO; }
        k.v = v < 0 ? (long)(P.size() / 3) + v : v - 1;
        if (l.p < l.end && *l.p == '/') {
          ++l.p;
          long t;
          if (parse_int(l, t)) k.vt = t < 0 ? (long)(T.size() / 2) + t : t - 1;
          if (l.p < l.end && *l.p == '/') {
            ++l.p;
            long nn;
            if (parse_int(l, nn)) k.vn = nn < 0 ? (long)(N.size() / 3) + nn : nn - 1;
          }
        }
        if (k.v < 0 || k.v >= (long)(P.size() / 3) || k.vt >= (long)(T.size() / 2) || k.vn >= (long)(N.size() / 3)) {
          err = "face index out of range";
          return RT_ERR_IO;
        }
        face.push_back(k);
      }
      if (face.empty()) { err = "empty face"; return RT_ERR_IO; }
      have_faces = true;
      if (face.size() == 1) {  // point -> degenerate triangle (ignore_points: false)
        add_vertex(face[0]); add_vertex(face[0]); add_vertex(face[0]);
      } else if (face.size() == 2) {  // line -> degenerate triangle (ignore_lines: false)
        add_vertex(face[0]); add_vertex(face[1]); add_vertex(face[1]);
      } else {
        for (size_t i = 1; i + 1 < face.size(); ++i) {
          add_vertex(face[0]); add_vertex(face[i]); add_vertex(face[i + 1]);
        }
      }
    } else if (n >= 2 && (line[0] == 'o' || line[0] == 'g') && (line[1] == ' ' || line[1] == '\t')) {
      if (have_faces) stop = true;  // second model starts: the reference only uses models[0]
    } else if (n == 1 && (line[0] == 'o' || line[0] == 'g')) {
      if (have_faces) stop = true;
    }
    c.p = eol < c.end ? eol + 1 : eol;
  }
  if (!have_faces) { err = "OBJ has no faces"; return RT_ERR_IO; }
  std::memset(out, 0, sizeof *out);
  out->nverts = (uint32_t)(pos.size() / 3);
  out->ntris = (uint32_t)(idx.size() / 3);
  out->has_normals = any_vn && !missing_vn;
  out->has_texcoords = any_vt && !missing_vt;
  out->pos = (float*)std::malloc(pos.size() * 4 + 4);
  out->nrm = (float*)std::malloc(nrm.size() * 4 + 4);
  out->uv = (float*)std::malloc(uv.size() * 4 + 4);
  out->idx = (uint32_t*)std::malloc(idx.size() * 4 + 4);
  if (!out->pos || !out->nrm || !out->uv || !out->idx) { err = "out of memory"; return RT_ERR_IO; }
  std::memcpy(out->pos, pos.data(), pos.size() * 4);
  std::memcpy(out->nrm, nrm.data(), nrm.size() * 4);
  std::memcpy(out->uv, uv.data(), uv.size() * 4);
  std::memcpy(out->idx, idx.data(), idx.size() * 4);
  return RT_OK;
}

// ------------------------------------------------------------------ TGA
int tga_decode(const uint8_t* b, size_t len, uint8_t** rgb, uint32_t* w, uint32_t* h, std::string& err) {
  if (len < 18) { err = "TGA too short"; return RT_ERR_IO; }
  uint32_t idlen = b[0], cmaptype = b[1], type = b[2];
  uint32_t cm_first = b[3] | (b[4] << 8), cm_len = b[5] | (b[6] << 8), cm_bpp = b[7];
  uint32_t W = b[12] | (b[13] << 8), H = b[14] | (b[15] << 8), bpp = b[16], desc = b[17];
  bool rle = type == 9 || type == 10 || type == 11;
  uint32_t base = rle ? type - 8 : type;
  if (!(base == 1 || base == 2 || base == 3) || W == 0 || H == 0) { err = "unsupported TGA type"; return RT_ERR_IO; }
  if (base == 1 && (cmaptype != 1 || bpp != 8 || !(cm_bpp == 24 || cm_bpp == 32))) { err = "unsupported TGA palette"; return RT_ERR_IO; }
  if (base == 2 && !(bpp == 24 || bpp == 32)) { err = "unsupported TGA depth"; return RT_ERR_IO; }
  if (base == 3 && bpp != 8) { err = "unsupported TGA depth"; return RT_ERR_IO; }
  size_t off = 18 + idlen;
  const uint8_t* cmap = nullptr;
  if (cmaptype == 1) {
    cmap = b + off;
    off += (size_t)cm_len * (cm_bpp / 8);
  }
  if (off > len) { err = "TGA truncated"; return RT_ERR_IO; }
  uint32_t bytespp = bpp / 8;
  size_t npix = (size_t)W * H;
  std::vector<uint8_t> raw(npix * bytespp);
  if (!rle) {
    if (off + raw.size() > len) { err = "TGA truncated"; return RT_ERR_IO; }
    std::memcpy(raw.data(), b + off, raw.size());
  } else {
    size_t o = 0, p = off;
    while (o < raw.size()) {
      if (p >= len) { err = "TGA truncated"; return RT_ERR_IO; }
      uint8_t hd = b[p++];
      uint32_t cnt = (hd & 0x7F) + 1;
      if (hd & 0x80) {
        if (p + bytespp > len) { err = "TGA truncated"; return RT_ERR_IO; }
        for (uint32_t i = 0; i < cnt && o < raw.size(); ++i, o += bytespp) std::memcpy(&raw[o], b + p, bytespp);
        p += bytespp;
      } else {
        size_t nb = (size_t)cnt * bytespp;
        if (p + nb > len) { err = "TGA truncated"; return RT_ERR_IO; }
        nb = std::min(nb, raw.size() - o);
        std::memcpy(&raw[o], b + p, nb);
        o += nb;
        p += (size_t)cnt * bytespp;
      }
    }
  }
  uint8_t* out = (uint8_t*)std::malloc(npix * 3);
  if (!out) { err = "out of memory"; return RT_ERR_IO; }
  bool top = (desc & 0x20) != 0, right = (desc & 0x10) != 0;
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      uint32_t sy = top ? y : H - 1 - y, sx = right ? W - 1 - x : x;
      const uint8_t* s = &raw[((size_t)sy * W + sx) * bytespp];
      uint8_t* d = &out[((size_t)y * W + x) * 3];
      if (base == 2) {
        d[0] = s[2]; d[1] = s[1]; d[2] = s[0];
      } else if (base == 3) {
        d[0] = d[1] = d[2] = s[0];
      } else {
        uint32_t ci = s[0];
        if (ci < cm_first || ci - cm_first >= cm_len) { d[0] = d[1] = d[2] = 0; continue; }
        const uint8_t* e = cmap + (size_t)(ci - cm_first) * (cm_bpp / 8);
        d[0] = e[2]; d[1] = e[1]; d[2] = e[0];
      }
    }
  *rgb = out;
  *w = W;
  *h = H;
  return RT_OK;
}

int tga_encode_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t** bytes, size_t* len) {
  if (!rgb || !w || !h || w > 65535 || h > 65535) return RT_ERR_INVALID;
  size_t n = 18 + (size_t)w * h * 3;
  uint8_t* o = (uint8_t*)std::malloc(n);
  if (!o) return RT_ERR_IO;
  std::memset(o, 0, 18);
  o[2] = 2;
  o[12] = w & 255; o[13] = w >> 8; o[14] = h & 255; o[15] = h >> 8;
  o[16] = 24;
  o[17] = 0x20;  // top-left origin
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    o[18 + 3 * i] = rgb[3 * i + 2];
    o[18 + 3 * i + 1] = rgb[3 * i + 1];
    o[18 + 3 * i + 2] = rgb[3 * i];
  }
  *bytes = o;
  *len = n;
  return RT_OK;
}

}  // namespace rt
