"""Pins for the CPU oracle (oracle/oracle.cpp).  The reference ships no tests or golden vectors
(SURVEY.md §4, §8c), so the oracle is pinned by known-answer tests derived from the reference's own
formulas, by the published Philox4x32-10 known-answer vectors, and by its two independent
closest-hit modes (reference-tree replay vs brute force) agreeing.  All CPU-only."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_ffi as O
import cs397raytracingsp22_b200 as rt
from cs397raytracingsp22_b200 import _ffi, scenes


# Random123's kat_vectors for philox4x32 with 10 rounds
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(ctr, key, want):
    assert tuple(O.philox(ctr, key)) == want


def test_ball_and_disk_are_uniform():
    """rand_sphere_vec / rand_disk_vec (tracing.rs:71-89) are uniform in the unit ball / disk; the
    direct maps that replace the rejection loops must have the same moments."""
    ball, disk = O.sample_ball_disk(1234, 200_000)
    r = np.linalg.norm(ball, axis=1)
    assert r.max() <= 1.0 + 1e-6
    assert abs(r.mean() - 0.75) < 3e-3            # E|x| = 3/4
    assert abs((r ** 2).mean() - 0.6) < 3e-3      # E|x|^2 = 3/5
    assert np.abs(ball.mean(axis=0)).max() < 5e-3
    assert np.abs((ball ** 2).mean(axis=0) - 0.2).max() < 3e-3   # E x_i^2 = 1/5
    assert abs(np.abs(ball[:, 1]).mean() - 3.0 / 8.0) < 3e-3      # E|y| = 3/8 (drives the 0.75*albedo of Q2)
    rd = np.linalg.norm(disk, axis=1)
    assert rd.max() <= 1.0 + 1e-6
    assert abs((rd ** 2).mean() - 0.5) < 3e-3
    assert np.abs(disk.mean(axis=0)).max() < 5e-3


def test_hemisphere_rotation_maps_y_to_normal():
    """sample_hemisphere (materials.rs:171-178): between_vectors(unit_y, n) is a rotation taking +y to n."""
    rng = np.random.RandomState(0)
    for _ in range(200):
        n = rng.normal(size=3).astype(np.float32)
        n /= np.linalg.norm(n)
        b = rng.uniform(-0.5, 0.5, size=3).astype(np.float32)
        d = O.sample_hemisphere(n, b)
        assert abs(np.linalg.norm(d) - np.linalg.norm(b)) < 1e-5      # rotation keeps length (not normalised, Q1)
        assert abs(float(d @ n) - abs(float(b[1]))) < 1e-5            # the |y| component ends up along n
    # n = +y: identity; n = -y: half turn about z (cgmath's from_arc fallback axis)
    b = np.array([0.3, -0.2, 0.1], np.float32)
    assert np.allclose(O.sample_hemisphere((0, 1, 0), b), [0.3, 0.2, 0.1], atol=1e-7)
    assert np.allclose(O.sample_hemisphere((0, -1, 0), b), [-0.3, -0.2, 0.1], atol=1e-6)


def test_furnace():
    """Camera inside a Sphere with Lambertian{albedo a, emission e}: every pixel converges to
    e*(1-(0.75a)^D)/(1-0.75a) (SURVEY.md §4 KAT 1; per-bounce weight 2a|d.n| with E = 0.75a, Q2)."""
    a, e, depth = 0.5, 1.0, 8
    cam = rt.Camera(eyepoint=(0, 0, 0), screen_width=24, screen_height=24, aa_sample_count=1024, path_depth=depth)
    sc = rt.Scene(camera=cam, objects=[rt.Sphere((0, 0, 0), 10.0, rt.Lambertian(albedo=(a,) * 3, emission=(e,) * 3))])
    lin, _, st = O.lower_to_oracle(sc).render(cam.to_c(), seed=7)
    want = e * (1 - (0.75 * a) ** depth) / (1 - 0.75 * a)
    assert abs(want - 1.59937) < 1e-4
    assert abs(lin.mean() - want) < 4e-3
    # closed scene: paths run to path_depth, except the few whose scattered ray is so close to tangent that
    # the far root falls below t_min = 0.001 (the sphere rule of Q9 then reports a miss)
    assert 0.99 * st.samples * depth < st.rays <= st.samples * depth


def test_volume_transmittance():
    """A ray along a diameter of ConvexVolume{Sphere r, density s} scatters with probability 1-exp(-2 r s)
    (SURVEY.md §4 KAT 2, Q10).  Directions are unit here so parametric t is distance."""
    r, sigma = 1.0, 0.6
    vol = rt.ConvexVolume(boundary=rt.Sphere((0, 0, 0), r, rt.Dielectric(1.5)),
                          phase_function=rt.Isotropic(albedo=(1, 1, 1)), density=sigma)
    sc = rt.Scene(camera=rt.Camera(), objects=[vol])
    n = 200_000
    rays = np.tile(np.array([0, 0, 5, 0, 0, -1], np.float32), (n, 1))
    res = O.lower_to_oracle(sc).intersect_rays(rays, 0.001, 100.0, seed=11)
    p = (res["obj"] == 0).mean()
    assert abs(p - (1 - math.exp(-2 * r * sigma))) < 4e-3
    hit = res["obj"] == 0
    assert res["t"][hit].min() >= 4.0 - 1e-4 and res["t"][hit].max() <= 6.0 + 1e-4
    assert np.all(res["normal"][hit] == 0.0) and np.all(res["frontface"][hit] == 0)   # zero normal, frontface false
    # half-density direction: an un-normalised direction of length 2 halves the parametric distances (Q1)
    rays2 = rays.copy()
    rays2[:, 3:] *= 2
    res2 = O.lower_to_oracle(sc).intersect_rays(rays2, 0.001, 100.0, seed=11)
    p2 = (res2["obj"] == 0).mean()
    assert abs(p2 - (1 - math.exp(-1 * r * sigma))) < 4e-3


def test_mesh_bounded_volume_transmittance():
    """ConvexVolume with a StaticMesh boundary.  A ray through a cube of side 2 scatters with
    probability 1-exp(-2 s); entry / exit come from the mesh BVH queried over all t, so a ray starting INSIDE sees the
    negative entry distance and only the part ahead of it."""
    sigma = 0.6
    cube = rt.StaticMesh.load_from_file(scenes.obj_path("cube"), material=rt.Lambertian())
    vol = rt.ConvexVolume(boundary=cube, phase_function=rt.Isotropic(albedo=(1, 1, 1)), density=sigma)
    b = O.lower_to_oracle(rt.Scene(camera=rt.Camera(), objects=[vol]))
    n = 200_000
    outside = np.tile(np.array([0.3, 0.2, 5, 0, 0, -1], np.float32), (n, 1))
    inside = np.tile(np.array([0.3, 0.2, 0.5, 0, 0, -1], np.float32), (n, 1))     # 1.5 units of fog ahead
    for mode in (O.MODE_REF_TREE, O.MODE_BRUTE):
        r = b.intersect_rays(outside, 0.001, 100.0, seed=11, mode=mode)
        assert abs((r["obj"] == 0).mean() - (1 - math.exp(-2 * sigma))) < 4e-3
        hit = r["obj"] == 0
        assert r["t"][hit].min() >= 4.0 - 1e-4 and r["t"][hit].max() <= 6.0 + 1e-4
        assert np.all(r["normal"][hit] == 0.0) and np.all(r["frontface"][hit] == 0)
        r = b.intersect_rays(inside, 0.001, 100.0, seed=12, mode=mode)
        assert abs((r["obj"] == 0).mean() - (1 - math.exp(-1.5 * sigma))) < 4e-3
        assert r["t"][r["obj"] == 0].max() <= 1.5 + 1e-4


def test_pixel_filter_footprint():
    """Multi-jitter offsets span [-1, 1) pixel with mean -(1/(2 sqrt n) + 1/(2n)) (SURVEY.md §4 KAT 3, Q8)."""
    cam = rt.Camera(screen_width=32, screen_height=32, aa_sample_count=64).to_c()
    offs = np.concatenate([O.camera_offsets(cam, 99, x, y) for x in range(8) for y in range(8)])
    assert offs.min() >= -1.0 - 1e-6 and offs.max() < 1.0
    n = 64
    want = -(1 / (2 * math.sqrt(n)) + 1 / (2 * n))
    assert abs(want - (-0.0703)) < 1e-4
    assert abs(offs.mean() - want) < 1e-2
    # the stratum part alone: cell index (i / rootn, i % rootn)
    one = O.camera_offsets(cam, 99, 3, 4)
    cells = np.floor((one + 1.0) * 4 + 1e-6)   # not asserting exact cells (jitter spans a full pixel) - only the range
    assert cells.min() >= 0 and cells.max() <= 7


def test_texture_addressing():
    """Texture::sample (texture.rs:28-31, Q7): (0,0) -> texel (0, H-1); (1,1) -> (min(floor(.999 W), W-1), floor(.001 H))."""
    w, h = 37, 21
    img = np.zeros((h, w, 3), np.uint8)
    img[..., 0] = np.arange(w)[None, :]
    img[..., 1] = np.arange(h)[:, None]
    b = O.OracleBackend()
    tid = b.add_texture(img)
    out = np.empty(3, np.float32)

    def tap(u, v):
        O.load().orc_texture_sample(b.handle, tid, u, v, _ffi.fptr(out))
        return int(round(out[0] * 255)), int(round(out[1] * 255))

    assert tap(0.0, 0.0) == (0, h - 1)
    assert tap(1.0, 1.0) == (min(int(0.999 * w), w - 1), int(0.001 * h))
    assert tap(-3.0, 7.0) == (0, int(0.001 * h))            # clamp, no wrap
    assert tap(0.5, 0.5) == (int(0.5 * w), int(0.5 * h))
    t = rt.Texture(img)                                     # host restatement agrees
    for u, v in ((0.0, 0.0), (1.0, 1.0), (0.5, 0.5), (0.123, 0.877)):
        assert tuple(int(round(c * 255)) for c in t.sample((u, v))[:2]) == tap(u, v)


def test_output_transform():
    """tracing.rs:243-256 (Q12): saturate from the pre-update copy, clamp, gamma, *255.9999, truncate; NaN -> 0."""
    assert list(O.output_transform((0.25, 0.25, 0.25), 2.0)) == [127, 127, 127]      # sqrt(.25)*255.9999 = 127.99
    assert list(O.output_transform((1.0, 0.0, 0.5), 2.0)) == [255, 0, int(math.sqrt(0.5) * 255.9999)]
    # excess of 0.5 in red spills +0.5 into green and blue; nothing spills back (reads the copy)
    assert list(O.output_transform((1.5, 0.25, 0.0), 1.0)) == [255, int(0.75 * 255.9999), int(0.5 * 255.9999)]
    # two channels in excess: each receives the other's excess but is clamped anyway
    assert list(O.output_transform((1.2, 1.4, 0.0), 1.0)) == [255, 255, int(round(0.6 * 255.9999 - 0.5))]
    assert list(O.output_transform((float("nan"), -1.0, 2.0), 2.0)) == [0, 0, 255]


def test_reachability_mask_of_the_reference_tree():
    """Q3: 51 of drone.obj's 1736 triangles sit under flat interior nodes of the index-order tree and can never
    be hit; the other meshes have none.  The product's mask must equal the oracle's replay."""
    want = (list(range(36, 43)) + list(range(50, 60)) + [1202, 1203] + list(range(1213, 1218)) + list(range(1226, 1236))
            + [1679, 1680] + list(range(1688, 1694)) + list(range(1702, 1711)))
    b = O.OracleBackend()
    for name, n_dead in (("drone", 51), ("cube", 0), ("teapot", 0), ("sphere", 0)):
        m = rt.load_obj(scenes.obj_path(name))
        mid = b.add_mesh(m.pos, m.nrm, m.uv, m.idx)
        orc = b.mesh_reachability(mid, m.ntris)
        assert int((orc == 0).sum()) == n_dead
        assert np.array_equal(orc, _ffi.mesh_reachability(m.pos, m.idx))
        if name == "drone":
            assert np.where(orc == 0)[0].tolist() == want


@pytest.mark.parametrize("name", ["c2", "c4"])
def test_tree_and_brute_force_agree(small_scenes, name):
    """SURVEY.md §4 KAT 4: the reference tree (left-then-right, t_max narrowing, strict slab test) gives the same
    closest hit as brute force over the reachable triangles with the tie rule of Q4."""
    sc = small_scenes(name)
    b = O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    a = b.trace_primary(cam, 0x5EED, 3, mode=O.MODE_REF_TREE)
    c = b.trace_primary(cam, 0x5EED, 3, mode=O.MODE_BRUTE)
    assert (a["obj"] >= 0).sum() > 1000
    assert np.array_equal(a["obj"], c["obj"])
    assert np.array_equal(a["prim"], c["prim"])
    assert np.array_equal(a["t"], c["t"])
    assert np.array_equal(a["normal"], c["normal"])
    # secondary-ray-like queries: un-normalised directions from points on the surfaces
    hit = a["obj"] >= 0
    rng = np.random.RandomState(5)
    org = a["ray"][hit, :3] + a["ray"][hit, 3:] * a["t"][hit, None]
    d = rng.uniform(-1, 1, size=org.shape).astype(np.float32)
    rays = np.concatenate([org, d], axis=1).astype(np.float32)
    r1 = b.intersect_rays(rays, 0.001, 100.0, seed=3, mode=O.MODE_REF_TREE)
    r2 = b.intersect_rays(rays, 0.001, 100.0, seed=3, mode=O.MODE_BRUTE)
    assert np.array_equal(r1["obj"], r2["obj"]) and np.array_equal(r1["prim"], r2["prim"])
    assert np.array_equal(r1["t"], r2["t"])


def test_tree_vs_brute_force_on_tiny_instances(small_scenes):
    """Where the static mask of Q3 stops being the whole story.  C5 scales drone.obj by 6e-4, so object-space ray
    origins are ~1e4 units out and (min - o) == (max - o) in f32 for the mesh's 2-ulp-thick interior boxes
    (x = 498.367676 vs 498.367737): the reference's strict slab test then rejects those nodes for THAT ray.  The
    effect is ray dependent, bounded, and one-sided (brute force only ever finds a closer hit).  The CUDA path keeps
    such boxes as GUARDS and replays the reference's test on them, so it follows the reference tree here too
    (tests/test_gpu_parity.py::test_guards_follow_the_reference_tree_where_brute_force_does_not)."""
    sc = small_scenes("c5")
    b = O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    a = b.trace_primary(cam, 0x5EED, 1, mode=O.MODE_REF_TREE)
    c = b.trace_primary(cam, 0x5EED, 1, mode=O.MODE_BRUTE)
    diff = (a["obj"] != c["obj"]) | (a["prim"] != c["prim"])
    assert diff.mean() < 2e-3
    both = diff & (a["obj"] >= 0) & (c["obj"] >= 0)
    assert (c["t"][both] <= a["t"][both]).all()
    assert not (diff & (c["obj"] < 0)).any()      # brute force never loses a hit the tree finds


def test_object_order_breaks_ties():
    """Q4: across top-level objects strict '<' keeps the EARLIEST object on equal t (tracing.rs:335)."""
    m1, m2 = rt.Lambertian(albedo=(1, 0, 0)), rt.Lambertian(albedo=(0, 1, 0))
    tri = dict(a=(-1, -1, -3), b=(1, -1, -3), c=(0, 1, -3))
    rays = np.array([[0, 0, 0, 0, 0, -1]], np.float32)
    for first, second in ((m1, m2), (m2, m1)):
        sc = rt.Scene(camera=rt.Camera(), objects=[rt.Triangle(material=first, **tri), rt.Triangle(material=second, **tri)])
        res = O.lower_to_oracle(sc).intersect_rays(rays, 0.001, 100.0)
        assert res["obj"][0] == 0 and res["t"][0] == 3.0


def test_sphere_and_plane_rules():
    """Q9: sphere takes the far root when the near one is below t_min; plane is two-sided and always front-face."""
    mat = rt.Lambertian()
    sc = rt.Scene(camera=rt.Camera(), objects=[rt.Sphere((0, 0, 0), 1.0, mat), rt.Plane((0, -2, 0), (0, 1, 0), mat)])
    b = O.lower_to_oracle(sc)
    rays = np.array([[0, 0, 3, 0, 0, -1],      # outside -> near root t=2, front face
                     [0, 0, 0, 0, 0, -2],      # inside, |d|=2 -> far root t=0.5 (parametric), back face
                     [5, 0, 0, 0, -1, 0],      # misses sphere, hits plane from above t=2
                     [5, -4, 0, 0, 1, 0]],     # plane from below: still a hit, normal flipped toward the ray
                    np.float32)
    r = b.intersect_rays(rays, 0.001, 100.0)
    assert r["obj"].tolist() == [0, 0, 1, 1]
    assert np.allclose(r["t"], [2.0, 0.5, 2.0, 2.0])
    assert r["frontface"].tolist() == [1, 0, 1, 1]
    assert np.allclose(r["normal"], [[0, 0, 1], [0, 0, 1], [0, 1, 0], [0, -1, 0]])


def test_mesh_hit_frame(small_scenes):
    """Q6: world normal is unit, faces the ray unless normal-mapped, hit point = M * object-space point."""
    sc = small_scenes("c2")
    b = O.lower_to_oracle(sc)
    r = b.trace_primary(sc.camera.to_c(), 1, 0)
    mesh = r["obj"] >= 8      # the two teapots come after the 8 Cornell objects
    assert mesh.sum() > 200
    n = r["normal"][mesh]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    d = r["ray"][mesh, 3:]
    assert ((n * d).sum(axis=1) < 1e-6).all()


def test_furnace_with_branching_paths():
    """Camera::path_samples = S (tracing.rs:308-319) changes the variance, not the expectation: the furnace value
    is the same, and a closed scene costs 1 + S + S^2 + ... closest-hit queries per camera sample."""
    a, e, depth, S = 0.5, 1.0, 4, 3
    cam = rt.Camera(eyepoint=(0, 0, 0), screen_width=16, screen_height=16, aa_sample_count=256, path_depth=depth, path_samples=S)
    sc = rt.Scene(camera=cam, objects=[rt.Sphere((0, 0, 0), 10.0, rt.Lambertian(albedo=(a,) * 3, emission=(e,) * 3))])
    lin, _, st = O.lower_to_oracle(sc).render(cam.to_c(), seed=9)
    want = e * (1 - (0.75 * a) ** depth) / (1 - 0.75 * a)
    assert abs(lin.mean() - want) < 4e-3
    per_sample = sum(S ** k for k in range(depth))
    assert 0.99 * st.samples * per_sample < st.rays <= st.samples * per_sample
    cam.path_samples = 1
    lin1, _, st1 = O.lower_to_oracle(sc).render(cam.to_c(), seed=9)
    assert lin.std() < lin1.std()                      # more scattered rays per hit: less noise per camera sample


def _phong_value(objects, light=(0.0, 2.0, 0.0), ambient=(0.05, 0.1, 0.15)):
    """Looks at the origin of the plane y = 0 from (3, 1, 0) through a very long lens: every pixel shades (nearly) the
    same point, n = (0,1,0), to_light = (0,1,0), and the mirror direction is 71.6 degrees away from the camera."""
    v = np.array([-3.0, -1.0, 0.0]) / math.sqrt(10.0)
    up = np.array([-1.0, 3.0, 0.0]) / math.sqrt(10.0)
    cam = rt.Camera(eyepoint=(3, 1, 0), view_dir=tuple(v), up=tuple(up), focal_length=400.0, screen_width=9, screen_height=9,
                    aa_sample_count=4, shading_mode=rt.ShadingMode.Phong)
    sc = rt.Scene(camera=cam, objects=objects, point_light_pos=light, ambient=ambient)
    lin, _, st = O.lower_to_oracle(sc).render(cam.to_c(), seed=5)
    assert lin.reshape(-1, 3).std(axis=0).max() < 2e-3
    return lin.reshape(-1, 3).mean(axis=0), st


def test_phong_shading_known_answers():
    """Scene::phong_shade_ray (tracing.rs:277-297): ambient + diffuse * brdf + specular^40 * 0.4, times 0.3 when the
    shadow ray is blocked - but only by an occluder that is nearer to the shaded point than to the light (`:292`)."""
    a = (0.8, 0.5, 0.2)
    amb = np.array([0.05, 0.1, 0.15])
    floor = rt.Plane(point=(0, 0, 0), normal=(0, 1, 0), material=rt.Lambertian(albedo=a))
    lit = amb + np.array(a) / math.pi                  # diffuse weight 1, specular weight (1/sqrt(10))^40 ~ 1e-20
    got, st = _phong_value([floor])
    assert np.allclose(got, lit, atol=1e-3)
    assert st.rays == 2 * st.samples                   # one camera ray + one shadow ray
    # an occluder close to the surface: shadow weight 0.3
    got, _ = _phong_value([floor, rt.Sphere((0, 0.5, 0), 0.2, rt.Lambertian())])
    assert np.allclose(got, 0.3 * lit, atol=1e-3)
    # the same occluder close to the light: hit.distance^2 > |light - hit|^2, so the reference does not darken
    got, _ = _phong_value([floor, rt.Sphere((0, 1.5, 0), 0.2, rt.Lambertian())])
    assert np.allclose(got, lit, atol=1e-3)
    # metal: the brdf term of scatter() is the albedo itself (materials.rs:64)
    got, _ = _phong_value([rt.Plane(point=(0, 0, 0), normal=(0, 1, 0), material=rt.Metal(albedo=a, roughness=0.3))])
    assert np.allclose(got, amb + np.array(a), atol=1e-3)
    # light in the mirror direction of the camera: specular weight 1
    got, _ = _phong_value([floor], light=(-6.0, 2.0, 0.0))
    n_dot_l = 2.0 / math.sqrt(40.0)
    assert np.allclose(got, amb + n_dot_l * np.array(a) / math.pi + 0.4, atol=2e-3)
    # nothing hit: black, and no shadow ray
    got, st = _phong_value([rt.Sphere((0, 50, 0), 1.0, rt.Lambertian())])
    assert np.all(got == 0) and st.rays == st.samples


def test_orthographic_projection_as_the_reference_writes_it():
    """CameraProjectionMode::Orthographic (tracing.rs:196,200): the ray starts at the camera-space pixel centre taken
    as a WORLD point (z = 0, the eyepoint is ignored) and runs along rotation * view_dir."""
    cam = rt.Camera(eyepoint=(7, 8, 9), screen_width=32, screen_height=16, aa_sample_count=4,
                    projection_mode=rt.CameraProjectionMode.Orthographic)
    sc = rt.Scene(camera=cam, objects=[rt.Sphere((0, 0, -3), 0.25, rt.Lambertian())])
    o = O.lower_to_oracle(sc)
    p = o.trace_primary(cam.to_c(), 1, 0)
    ray = p["ray"].reshape(16, 32, 6)
    assert np.all(ray[..., 2] == 0.0)
    assert np.all(ray[..., 3:] == np.array([0, 0, -1], np.float32))    # rotation * view_dir = view_dir for this basis
    ps = 1.0 / 16
    assert abs(ray[0, 0, 0] - (-1.0 + 0.5 * ps)) <= ps and abs(ray[0, 31, 0] - (1.0 - 0.5 * ps)) <= ps
    assert abs(ray[0, 0, 1] - 0.5) <= 1.5 * ps and abs(ray[15, 0, 1] + 0.5) <= 1.5 * ps
    # the sphere of radius 0.25 covers pi r^2 of the 2 x 1 window, wherever the eyepoint is
    frac = (p["obj"] == 0).mean()
    assert abs(frac - math.pi * 0.25 ** 2 / 2.0) < 0.02


# ---- Material::scatter (materials.rs:33-166), one surface point, many draws
N_UP = (0.0, 1.0, 0.0)


def _unit(v):
    v = np.asarray(v, np.float64)
    return tuple(v / np.linalg.norm(v))


def test_lambertian_and_isotropic_scatter():
    """Lambertian: uniform in the half BALL above the normal (never normalised, Q1), brdf = albedo/pi, pdf = 1/(2 pi), so
    the mean weight |d.n| * brdf / pdf is 0.75 * albedo (Q2).  Isotropic: the whole ball, brdf = albedo, pdf = 1."""
    n = _unit((1, 2, -1))
    d, brdf, pdf = O.scatter(_ffi.RT_MAT_LAMBERTIAN, n, (0, -1, 0), n=100_000, albedo=(0.2, 0.5, 0.8), seed=3)
    dn = d @ np.asarray(n, np.float32)
    assert dn.min() >= -1e-6 and np.linalg.norm(d, axis=1).max() <= 1 + 1e-6
    assert np.allclose(brdf, np.array([0.2, 0.5, 0.8]) / math.pi, rtol=1e-6) and np.allclose(pdf, 1 / (2 * math.pi), rtol=1e-6)
    assert abs(dn.mean() - 3.0 / 8.0) < 3e-3                       # E|y| of the unit ball
    weight = (np.clip(np.abs(dn), 0, 1)[:, None] * brdf / pdf[:, None]).mean(axis=0)
    assert np.allclose(weight, 0.75 * np.array([0.2, 0.5, 0.8]), atol=4e-3)
    d, brdf, pdf = O.scatter(_ffi.RT_MAT_ISOTROPIC, (0, 0, 0), (0, -1, 0), n=100_000, albedo=(0.3, 0.3, 0.9), seed=4)
    assert np.abs(d.mean(axis=0)).max() < 5e-3 and abs((d ** 2).sum(axis=1).mean() - 0.6) < 4e-3
    assert np.allclose(brdf, [0.3, 0.3, 0.9]) and np.all(pdf == 1.0)


def test_metal_scatter():
    """reflect(d, n) + roughness * ball (materials.rs:57-67); the incoming direction is used as it is (length 2 here)."""
    d_in = np.array([1.0, -1.0, 0.5], np.float32) * 2
    d, brdf, pdf = O.scatter(_ffi.RT_MAT_METAL, N_UP, d_in, n=50_000, albedo=(0.9, 0.8, 0.7), roughness=0.0)
    assert np.allclose(d, [2.0, 2.0, 1.0], atol=1e-6) and np.allclose(brdf, [0.9, 0.8, 0.7]) and np.all(pdf == 1.0)
    d, _, _ = O.scatter(_ffi.RT_MAT_METAL, N_UP, d_in, n=100_000, roughness=0.25, seed=8)
    off = d - np.array([2.0, 2.0, 1.0], np.float32)
    assert np.linalg.norm(off, axis=1).max() <= 0.25 + 1e-6 and np.abs(off.mean(axis=0)).max() < 2e-3
    assert abs(np.linalg.norm(off, axis=1).mean() - 0.25 * 0.75) < 2e-3


def test_dielectric_scatter():
    """materials.rs:80-98 + tracing.rs:58-69: Schlick with r0 = ((ior-1)/(ior+1))^2 on |d.n|, reflection with that
    probability, Snell refraction otherwise, total internal reflection when eta * sin(theta) > 1 (Q11)."""
    ior = 1.5
    r0 = ((ior - 1) / (ior + 1)) ** 2
    # normal incidence from outside: F = r0 = 0.04; the refracted ray goes straight on
    d, brdf, pdf = O.scatter(_ffi.RT_MAT_DIELECTRIC, N_UP, (0, -1, 0), frontface=True, n=200_000, ior=ior, seed=5)
    refl = d[:, 1] > 0
    assert abs(refl.mean() - r0) < 2e-3
    assert np.allclose(d[refl], [0, 1, 0], atol=1e-6) and np.allclose(d[~refl], [0, -1, 0], atol=1e-6)
    assert np.all(brdf == 1.0) and np.all(pdf == 1.0)
    # 60 degrees from outside: F = r0 + (1-r0) * (1 - cos)^5, refraction obeys sin(t) = sin(i) / ior
    th = math.radians(60.0)
    d_in = (math.sin(th), -math.cos(th), 0.0)
    fres = r0 + (1 - r0) * (1 - math.cos(th)) ** 5
    d, _, _ = O.scatter(_ffi.RT_MAT_DIELECTRIC, N_UP, d_in, frontface=True, n=200_000, ior=ior, seed=6)
    refl = d[:, 1] > 0
    assert abs(refl.mean() - fres) < 3e-3
    assert np.allclose(d[refl], [math.sin(th), math.cos(th), 0], atol=1e-6)
    sin_t = math.sin(th) / ior
    assert np.allclose(d[~refl], [sin_t, -math.sqrt(1 - sin_t ** 2), 0], atol=1e-6)
    # from inside (the stored normal faces the ray, frontface = false, eta = ior): beyond asin(1/1.5) = 41.8 deg everything reflects
    for deg, tir in ((30.0, False), (45.0, True), (80.0, True)):
        th = math.radians(deg)
        d, _, _ = O.scatter(_ffi.RT_MAT_DIELECTRIC, N_UP, (math.sin(th), -math.cos(th), 0.0), frontface=False, n=20_000, ior=ior, seed=7)
        refl = d[:, 1] > 0
        assert refl.all() if tir else (0.0 < refl.mean() < 0.2)
        if not tir:
            sin_t = math.sin(th) * ior
            assert np.allclose(d[~refl], [sin_t, -math.sqrt(1 - sin_t ** 2), 0], atol=1e-5)


def test_parameterized_scatter():
    """materials.rs:114-145: k_s = F(ior 1.5) * (1 - roughness), k_d = (1 - k_s)(1 - metallic); with probability k_d the
    Lambertian lobe (albedo/pi, 1/(2 pi)), otherwise the metal lobe with brdf lerp(white, albedo, metallic), pdf 1."""
    albedo = np.array([0.01, 0.02, 0.5])
    th = math.radians(50.0)
    d_in = (math.sin(th), -math.cos(th), 0.0)
    r0 = 0.04
    fres = r0 + (1 - r0) * (1 - math.cos(th)) ** 5
    for rough, metallic in ((0.0, 0.0), (0.25, 0.5), (1.0, 0.0), (0.5, 1.0)):
        d, brdf, pdf = O.scatter(_ffi.RT_MAT_PARAMETERIZED, N_UP, d_in, n=200_000, albedo=albedo, roughness=rough, metallic=metallic,
                                 seed=int(10 + 10 * rough + metallic * 100))
        k_d = (1 - fres * (1 - rough)) * (1 - metallic)
        diffuse = pdf < 0.5
        assert abs(diffuse.mean() - k_d) < 3e-3, (rough, metallic)
        if diffuse.any():
            assert np.allclose(brdf[diffuse], albedo / math.pi, rtol=1e-5) and np.allclose(pdf[diffuse], 1 / (2 * math.pi), rtol=1e-6)
            assert d[diffuse][:, 1].min() >= -1e-6
        if (~diffuse).any():
            assert np.allclose(brdf[~diffuse], (1 - metallic) * np.ones(3) + metallic * albedo, atol=1e-6)
            off = d[~diffuse] - np.array([math.sin(th), math.cos(th), 0.0], np.float32)
            assert np.linalg.norm(off, axis=1).max() <= rough + 1e-5


def test_defocus_rays_meet_on_the_focus_sphere():
    """Camera::generate_rays (tracing.rs:176-203): the lens sample is uniform on a disk of lens_radius around the eye, and
    every ray of a pixel sample passes through normalize(pixel centre) * focus_dist - a focus SPHERE, not a plane (Q8)."""
    kw = dict(eyepoint=(1.0, 2.0, 3.0), screen_width=24, screen_height=16, aa_sample_count=16, focal_length=0.8, focus_dist=4.0)
    sc = rt.Scene(camera=rt.Camera(**kw), objects=[rt.Sphere((0, 0, -50), 1.0, rt.Lambertian())])
    o = O.lower_to_oracle(sc)
    eye = np.array(kw["eyepoint"], np.float32)
    origins = []
    for sample in (0, 5, 11):
        pin = o.trace_primary(rt.Camera(lens_radius=0.0, **kw).to_c(), 21, sample)["ray"]
        lens = o.trace_primary(rt.Camera(lens_radius=0.3, **kw).to_c(), 21, sample)["ray"]
        assert np.allclose(pin[:, :3], eye) and np.allclose(np.linalg.norm(pin[:, 3:], axis=1), 1.0, atol=1e-6)
        focus = pin[:, :3] + pin[:, 3:] * 4.0                       # on the sphere of radius focus_dist around the eye
        lo, ld = lens[:, :3], lens[:, 3:]
        assert np.allclose(np.linalg.norm(ld, axis=1), 1.0, atol=1e-6)
        miss = np.linalg.norm(np.cross(focus - lo, ld), axis=1)     # distance from the focus point to the lens ray
        assert miss.max() < 2e-5
        assert np.allclose(lo[:, 2], eye[2], atol=1e-6)             # the lens lies in the camera's xy plane
        origins.append(lo - eye)
    r = np.linalg.norm(np.concatenate(origins), axis=1)
    assert r.max() <= 0.3 + 1e-6 and abs((r ** 2).mean() - 0.3 ** 2 / 2) < 3e-3   # uniform on the disk


def test_triangle_determinant_epsilon_is_in_parametric_units():
    """Triangle::intersect_ray (geometry.rs:433-438): |det| < 1e-4 is a miss, with det = e1 . (d x e2) for the direction AS
    GIVEN - so a small triangle that a unit-length ray cannot hit is hit by the same ray with a longer direction (Q1, Q5)."""
    tri = rt.Triangle((0, 0, 0), (0.009, 0, 0), (0, 0.009, 0), rt.Lambertian())
    o = O.lower_to_oracle(rt.Scene(camera=rt.Camera(), objects=[tri]))
    ray = np.array([[0.002, 0.002, 1.0, 0, 0, -1.0]], np.float32)
    assert o.intersect_rays(ray, 0.001, 100.0)["obj"][0] == -1       # det = 8.1e-5
    ray2 = ray.copy()
    ray2[:, 3:] *= 2.0                                               # det = 1.62e-4, t halves
    res = o.intersect_rays(ray2, 0.001, 100.0)
    assert res["obj"][0] == 0 and abs(res["t"][0] - 0.5) < 1e-6
    # both faces are hit (no culling), the reported normal faces the ray
    back = np.array([[0.002, 0.002, -1.0, 0, 0, 2.0]], np.float32)
    res = o.intersect_rays(back, 0.001, 100.0)
    assert res["obj"][0] == 0 and res["normal"][0, 2] < 0 and res["frontface"][0] == 0
