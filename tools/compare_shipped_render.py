#!/usr/bin/env python
"""Renders the reference's run() scene (C4) at the size of the render the reference ships (render.png, 800x800) and
prints how far the two images are apart, region by region.  The shipped render was made with the real Drone_*.tga maps,
which are not in the checkout, so only regions out of reach of the drone's emission can be expected to agree.

usage: python tools/compare_shipped_render.py <reference render.png | tests/golden/shipped_render_200.npz> [spp]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cs397raytracingsp22_b200 import _ffi, scenes  # noqa: E402

REGIONS = {  # (y0, y1, x0, x1) in the 800x800 frame
    "material grid (15 spheres)": (30, 340, 110, 690),
    "cube top": (580, 600, 60, 140),
    "cube front": (660, 760, 40, 160),
    "magenta ball": (640, 720, 640, 760),
    "cyan emitter": (380, 420, 690, 750),
    "fog (left volume)": (480, 540, 10, 60),
    "floor bottom right": (740, 800, 700, 800),
    "floor bottom mid": (760, 800, 300, 500),
}


def box(a, k):
    h, w, c = a.shape
    return a.reshape(h // k, k, w // k, k, c).astype(np.float32).mean(axis=(1, 3))


def main():
    src = sys.argv[1]
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    if src.endswith(".npz"):
        ref = np.load(src)["rgb_box4"].astype(np.float32) / 4.0  # stored in quarter LSBs
    else:
        ref = box(_ffi.png_decode(open(src, "rb").read()), 4)
    sc = scenes.make_scene("c4", width=800, height=800, spp=spp, depth=10)
    lin, rgb, st = sc.render()
    mine = box(rgb, 4)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    open(os.path.join(out, f"c4_800_{spp}spp.png"), "wb").write(_ffi.png_encode(rgb))
    for name, (y0, y1, x0, x1) in REGIONS.items():
        a = ref[y0 // 4:y1 // 4, x0 // 4:x1 // 4]
        b = mine[y0 // 4:y1 // 4, x0 // 4:x1 // 4]
        d = np.abs(a - b)
        print(f"{name:28s} mean ref {a.mean(axis=(0, 1)).round(1)} ours {b.mean(axis=(0, 1)).round(1)}  |d| mean {d.mean():.2f} "
              f"p95 {np.percentile(d, 95):.1f} p99 {np.percentile(d, 99):.1f} max {d.max():.1f}")
    print("ms", st.ms_total, "samples", st.samples)


if __name__ == "__main__":
    main()
