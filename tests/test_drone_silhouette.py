"""Pins Q3 (the 51 drone triangles the reference's own BVH can never reach) and the OBJ triangle order it depends on
against the REFERENCE'S OWN OUTPUT.

The render the reference ships (`render.png`, its `run()` scene at 800x800) shows the drone's three emissive lens
discs with wedge-shaped black holes: fan-triangulated, axis-aligned disc caps whose interior BVH boxes are flat, which
`AABB::intersect_ray` (geometry.rs:63-67, `tmax <= tmin`) rejects.  Which triangles fall into flat boxes depends on the
triangle ORDER that tobj produces (face order of the file, (0, i, i+1) fans) and on the index-order median split of
geometry.rs:190-217 - the one crate behaviour SURVEY.md §8c flags as index-defining.  If the product's OBJ reader or the
oracle's tree replay ordered the triangles differently, the predicted holes would sit elsewhere on the discs.

`tests/golden/shipped_render_lenses.npz` is the lens region of render.png (made by tests/golden/make_lens_fixture.py).
Prediction, per pixel of that region: HOLE = the camera ray would hit one of the unreachable triangles before whatever
it actually hits; CAP = it hits a reachable triangle of the same disc planes.  The shipped pixels must be dark on the
holes and lens-blue only on the caps.  CPU: the oracle's primary hits.  GPU: the CUDA path's.
"""
import os

import numpy as np
import pytest

import oracle_ffi as O
from cs397raytracingsp22_b200 import _ffi, scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 0x5EED
SAMPLES = (0, 333, 770, 1023)   # sample indices whose jittered camera rays must agree on a pixel's class
CAP_X = (488.36768, 498.36768)  # object-space planes of the disc caps (drone.obj)


def _moller_trumbore(o, d, tri):
    a, e1, e2 = tri[0], tri[1] - tri[0], tri[2] - tri[0]
    q = np.cross(d, e2)
    g = q @ e1
    ok = np.abs(g) > 1e-12
    f = np.where(ok, 1.0 / np.where(ok, g, 1.0), 0.0)
    s = o - a
    u = f * (s * q).sum(axis=1)
    r = np.cross(s, e1)
    v = f * (d * r).sum(axis=1)
    t = f * (r @ e2)
    return np.where(ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 1e-3), t, np.inf)


def _classify(trace):
    """trace(scene, sample) -> dict(ray, obj, prim, t) over the 800x800 frame.  Returns (hole in every sample, cap in
    every sample, cap in any sample, shipped_rgb)."""
    f = np.load(os.path.join(GOLD, "shipped_render_lenses.npz"))
    img = f["rgb"].astype(np.float32)
    x0, y0 = (int(v) for v in f["origin"])
    h, w = img.shape[:2]
    sc = scenes.make_scene("c4", width=800, height=800, spp=1024, depth=10, map_size=64)
    drone = sc.objects[0]
    md = drone.mesh
    reach = _ffi.mesh_reachability(md.pos, md.idx)
    assert (reach == 0).sum() == 51
    tris = md.pos[md.idx].astype(np.float64)
    xs = tris[:, :, 0]
    on_cap = np.zeros(len(tris), bool)
    for cx in CAP_X:
        on_cap |= (np.abs(xs - cx) < 1e-3).all(axis=1)
    assert on_cap[reach == 0].all(), "every unreachable triangle lies on a disc cap"
    M = np.asarray(drone.transform, np.float64)
    world = tris @ M[:3, :3].T + M[:3, 3]
    hole_votes = np.zeros((h, w), int)
    cap_votes = np.zeros((h, w), int)
    for s in SAMPLES:
        p = trace(sc, s)
        cut = lambda a: a.reshape(800, 800, -1)[y0:y0 + h, x0:x0 + w].reshape(h * w, -1)
        R = cut(p["ray"]).astype(np.float64)
        obj, prim = cut(p["obj"])[:, 0], cut(p["prim"])[:, 0]
        t_hit = np.where(obj >= 0, cut(p["t"])[:, 0], np.inf)
        t_un = np.full(h * w, np.inf)
        for k in np.where(reach == 0)[0]:
            t_un = np.minimum(t_un, _moller_trumbore(R[:, :3], R[:, 3:], world[k]))
        dprim = np.where(obj == 0, prim, 0)
        assert (reach[dprim[obj == 0]] == 1).all(), "a hit on an unreachable triangle"
        hole_votes += (t_un < t_hit).reshape(h, w)
        cap_votes += ((obj == 0) & on_cap[dprim]).reshape(h, w)
    return hole_votes == len(SAMPLES), cap_votes == len(SAMPLES), cap_votes > 0, img


def _check(hole, cap, cap_any, img):
    r, g, b = img[..., 0], img[..., 1], img[..., 2]
    lens_blue = (b > 150) & (b > r + 40) & (b > g + 40)       # the emissive disc as the reference rendered it
    assert hole.sum() > 500 and cap.sum() > 1000 and lens_blue.sum() > 1000
    dark_holes = (b[hole] < 100).mean()
    blue_on_cap = cap_any[lens_blue].mean()   # a pixel is anti-aliased over a 2-pixel footprint (Q8): any sample counts
    blue_on_hole = hole[lens_blue].mean()
    print(f"hole pixels {hole.sum()} (dark in render.png: {dark_holes:.3f}); lens-blue pixels {lens_blue.sum()}: "
          f"{blue_on_cap:.3f} on predicted caps, {blue_on_hole:.4f} on predicted holes")
    assert dark_holes > 0.95          # the reference shows nothing where the 51 triangles would be
    assert blue_on_cap > 0.95         # and its lit disc pixels are reachable cap triangles
    assert blue_on_hole < 0.01


def test_oracle_predicts_the_holes_of_the_shipped_render():
    def trace(sc, s):
        if not hasattr(trace, "o"):
            trace.o = O.lower_to_oracle(sc)
        return trace.o.trace_primary(sc.camera.to_c(), SEED, s, mode=O.MODE_REF_TREE)
    _check(*_classify(trace))


@pytest.mark.gpu
def test_cuda_path_predicts_the_holes_of_the_shipped_render(gpu):
    def trace(sc, s):
        if not hasattr(trace, "g"):
            trace.g = sc.commit(0)
        return trace.g.trace_primary(sc.camera.to_c(), SEED, s)
    _check(*_classify(trace))
