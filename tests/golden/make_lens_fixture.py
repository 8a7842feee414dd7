#!/usr/bin/env python
"""Makes tests/golden/shipped_render_lenses.npz: the part of the render the reference ships (render.png, 800x800, its
`run()` scene) that shows the drone's three lens discs, full resolution, 8 bit.  The wedge-shaped holes in those discs
are the triangles the reference's own BVH can never reach: tests/test_drone_silhouette.py checks that
they are where the oracle and the CUDA path predict them.
Usage (from the repo root, in the container that has the reference checkout):
    python tests/golden/make_lens_fixture.py /root/reference/render.png"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cs397raytracingsp22_b200 import _ffi  # noqa: E402

X0, X1, Y0, Y1 = 440, 640, 440, 640   # covers the lens cluster with a margin

if __name__ == "__main__":
    rgb = _ffi.png_decode(open(sys.argv[1], "rb").read())
    assert rgb.shape == (800, 800, 3), rgb.shape
    out = os.path.join(HERE, "shipped_render_lenses.npz")
    np.savez_compressed(out, rgb=rgb[Y0:Y1, X0:X1].copy(), origin=np.array([X0, Y0], np.int32))
    print(out, os.path.getsize(out), "bytes")
