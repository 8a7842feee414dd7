#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/oracle.cpp, reference-tree mode).

The reference itself ships no golden vectors and cannot run here (Rust toolchain absent, OS-seeded RNG), so these
fixtures freeze the ORACLE's output for the small variants of the five BASELINE configurations:
  - closest hit of camera sample 0 for every pixel: object id, triangle id, distance (bit patterns), world normal
  - the 16-spp linear-radiance image and the RGB8 image
They serve two purposes: the CPU suite notices if the oracle ever drifts, and the GPU suite has committed vectors to
compare the CUDA path with.  Usage (from the repo root):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_ffi as O  # noqa: E402
from cs397raytracingsp22_b200 import scenes  # noqa: E402
from conftest import SMALL  # noqa: E402

SEED = 0x5EED


def build(name):
    sc = scenes.make_scene(name, **SMALL[name])
    b = O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    p = b.trace_primary(cam, SEED, 0, mode=O.MODE_REF_TREE)
    lin, rgb, st = b.render(cam, seed=SEED, mode=O.MODE_REF_TREE)
    return dict(obj=p["obj"].astype(np.int16), prim=p["prim"].astype(np.int32), t=p["t"], normal=p["normal"].astype(np.float32),
                linear=lin.astype(np.float32), rgb8=rgb, samples=np.int64(st.samples), rays=np.int64(st.rays))


if __name__ == "__main__":
    for name in sorted(SMALL):
        data = build(name)
        path = os.path.join(HERE, f"{name}_small.npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path), "bytes", "rays", int(data["rays"]))
