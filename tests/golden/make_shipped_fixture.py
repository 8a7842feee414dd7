#!/usr/bin/env python
"""Makes tests/golden/shipped_render_200.npz from the render the reference ships (render.png, 800x800, its `run()`
scene): decoded with the library's PNG reader, 4x4 box filter on the 8-bit values, stored in quarter LSBs (uint16).
Usage (from the repo root, in the container that has the reference checkout):
    python tests/golden/make_shipped_fixture.py /root/reference/render.png"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cs397raytracingsp22_b200 import _ffi  # noqa: E402

if __name__ == "__main__":
    rgb = _ffi.png_decode(open(sys.argv[1], "rb").read())
    assert rgb.shape == (800, 800, 3), rgb.shape
    box = rgb.reshape(200, 4, 200, 4, 3).astype(np.float32).mean(axis=(1, 3))
    out = os.path.join(HERE, "shipped_render_200.npz")
    np.savez_compressed(out, rgb_box4=np.round(box * 4).astype(np.uint16))
    print(out, os.path.getsize(out), "bytes")
