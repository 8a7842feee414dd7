// rt_shade.cuh — one bounce of Scene::shade_ray (tracing.rs:300-324) for a ray whose closest hit is known:
// hit frame, material parameters, emission into the accumulator, Material::scatter, the next segment.
// Shared by the wavefront engine (k_shade) and the megakernel (k_path), so both run the same arithmetic.
// Part of the single translation unit rt_kernels.cu (everything here is __forceinline__ device code).
#ifndef RT_SHADE_CUH
#define RT_SHADE_CUH

#include "rt_materials.cuh"
#include "rt_traverse.cuh"

namespace rt {

// a path between two bounces: the ray about to be traced, the throughput it carries, and its RNG coordinates
struct PathState {
  f3 o, d, T;
  uint32_t pixel;
  uint32_t sb;    // sample | bounce << 24
  uint32_t tree;  // position in the sample's scatter tree (Camera::path_samples > 1), else 0
};

// Shades the hit `best` of path `p` and replaces p by the scattered path.  Returns false when the path ends here
// (depth reached, tracing.rs:301, or nothing left to carry).  `cls` is the material class of the hit when the
// caller already knows it (k_shade: warp-uniform, from its queue), or -1.
// MULTI (Camera::path_samples > 1): the caller runs once per child index fr.branch over the same hit; a child's
// random numbers are keyed by its position in the sample's path tree, its throughput is divided by path_samples
// (tracing.rs:319), and the hit's emission is added by the first pass only.
template <bool COUNT, bool MULTI>
__device__ __forceinline__ bool shade_hit(const rt_dev_scene& sc, const rt_frame& fr, PathState& p, const Best& best, int cls,
                                          long long* __restrict__ accum, unsigned long long* counters) {
  const f3 o = p.o, d = p.d, T = p.T;
  const uint32_t pixel = p.pixel;
  uint32_t tree = 0;
  if (MULTI) tree = p.tree * fr.path_samples + fr.branch;
  const uint32_t sample = p.sb & 0xFFFFFFu;
  uint32_t bounce = p.sb >> 24;
  // hit resolution: what the reference attaches to its RayHit
  Surface sf;
  resolve_hit<COUNT>(sc, o, d, best, sf, counters);
  const f3 hp = sf.hp, n = sf.n;
  const uint32_t meta = sf.meta;
  if (cls < 0) cls = (int)(meta & 7u);
  const bool front = (meta >> 3) & 1u;
  const uint32_t id = meta >> 4;

  // material parameters
  f3 albedo, emission;
  float roughness, metallic;
  if (cls == RT_CLASS_PARAM_TEX) {
    // StaticMesh::get_material_at_uv, geometry.rs:259-269 (Q7 defaults)
    uint32_t q = id * RT_OBJ_QUADS;
    float4 m7 = ldq(sc.objects, q + 7), m8 = ldq(sc.objects, q + 8);
    int ta = (int)fbits(m7.z), te = (int)fbits(m7.w), tm = (int)fbits(m8.x), tr = (int)fbits(m8.y);
    float u = sf.u, v = sf.v;
    albedo = ta >= 0 ? tex_sample(sc, ta, u, v) : mk(0.f, 0.f, 0.f);
    emission = te >= 0 ? tex_sample(sc, te, u, v) : mk(0.f, 0.f, 0.f);
    metallic = tm >= 0 ? tex_sample(sc, tm, u, v).x : 0.0f;
    roughness = tr >= 0 ? tex_sample(sc, tr, u, v).x : 1.0f;
    if (COUNT) atomicAdd(&counters[5], (unsigned long long)((ta >= 0) + (te >= 0) + (tm >= 0) + (tr >= 0)));
  } else {
    float4 m0 = ldq(sc.mats, id * RT_MAT_QUADS), m1 = ldq(sc.mats, id * RT_MAT_QUADS + 1);
    if (COUNT) atomicAdd(&counters[6], 1ull);
    albedo = mk(m0.x, m0.y, m0.z);
    emission = mk(m1.x, m1.y, m1.z);
    roughness = m0.w;
    metallic = m1.w;
  }

  // emitted light reaches the pixel attenuated by the path throughput (tracing.rs:321)
  if ((!MULTI || fr.branch == 0u) && (emission.x != 0.0f || emission.y != 0.0f || emission.z != 0.0f))
    accum_add(accum, pixel, mulv(T, emission));

  // scatter (materials.rs:33-166)
  u4 r = philox4x32_10(pixel, sample, bounce, 0u, fr.k0, MULTI ? fr.k1 ^ (tree * 0x9E3779B9u) : fr.k1);
  float u_choice = u01(r.x);
  f3 ball = ball_from(r.y, r.z, r.w);
  f3 dir, brdf;
  float pdf;
  const float PI = RT_PI;
  if (cls == RT_CLASS_LAMBERT) {
    dir = sample_hemisphere(n, ball);
    brdf = albedo / PI;
    pdf = 1.0f / (2.0f * PI);
  } else if (cls == RT_CLASS_METAL) {
    dir = reflectv(d, n) + roughness * ball;
    brdf = albedo;
    pdf = 1.0f;
  } else if (cls == RT_CLASS_DIELECTRIC) {
    float ior = roughness;  // stored in the roughness slot
    float eta = front ? 1.0f / ior : ior;
    float cth = fminf(-dot(d, n), 1.0f);
    bool critical = eta * sqrtf(1.0f - cth * cth) > 1.0f;
    float fres = fresnelf(d, n, ior);
    bool will_refract = !critical && u_choice >= fres;
    dir = will_refract ? refractv(d, n, eta) : reflectv(d, n);
    brdf = mk(1.0f, 1.0f, 1.0f);
    pdf = 1.0f;
  } else if (cls == RT_CLASS_ISOTROPIC) {
    dir = ball;
    brdf = albedo;
    pdf = 1.0f;
  } else {  // PARAM / PARAM_TEX, materials.rs:114-145
    float fres = fresnelf(d, n, 1.5f);
    float k_s = fres * (1.0f - roughness);
    float k_d = (1.0f - k_s) * (1.0f - metallic);
    if (u_choice < k_d) {
      dir = sample_hemisphere(n, ball);
      brdf = albedo / PI;
      pdf = 1.0f / (2.0f * PI);
    } else {
      dir = reflectv(d, n) + roughness * ball;
      brdf = (1.0f - metallic) * mk(1.0f, 1.0f, 1.0f) + metallic * albedo;  // lerpvec(1, albedo, metallic)
      pdf = 1.0f;
    }
  }
  // tracing.rs:313-316
  float dot_term = mag2(n) > 0.0f ? clampf(fabsf(dot(dir, n)), 0.0f, 1.0f) : 1.0f;
  f3 w = (dot_term * brdf) / pdf;
  f3 nT = mulv(T, w);
  if (MULTI) nT = nT / (float)fr.path_samples;
  bounce += 1;
  p.o = hp;
  p.d = dir;
  p.T = nT;
  p.sb = sample | (bounce << 24);
  p.tree = tree;
  // the next segment exists only below path_depth (tracing.rs:301); a path whose throughput is
  // exactly zero can add nothing any more
  return bounce < fr.path_depth && !(nT.x == 0.0f && nT.y == 0.0f && nT.z == 0.0f);
}

}  // namespace rt
#endif
