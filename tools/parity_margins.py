"""How much room do the image-parity tolerances of tests/test_gpu_parity.py leave?  (development only, needs a GPU)

Prints, per configuration, the quantities test_low_spp_images_track_the_oracle_sample_for_sample and
test_converged_images_rmse assert on, so the thresholds can be set from measurements instead of guesses.
usage: python tools/parity_margins.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle_ffi as O  # noqa: E402
from cs397raytracingsp22_b200 import _ffi, scenes  # noqa: E402
from conftest import SMALL  # noqa: E402

SEED = 0x5EED
CONVERGED = {"c1": dict(width=48, height=48, spp=1024), "c2": dict(width=48, height=48, spp=1024),
             "c3": dict(width=48, height=48, spp=1024), "c4": dict(width=64, height=36, spp=1024, map_size=256),
             "c5": dict(width=64, height=36, spp=1024, map_size=128, grid=6)}


def pair(sc):
    g, o = sc.commit(0), O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    opts = _ffi.rt_render_opts()
    opts.seed = SEED
    lin_g, rgb_g, st_g = g.render(cam, opts)
    lin_o, rgb_o, st_o = o.render(cam, seed=SEED, mode=O.MODE_REF_TREE)
    g.close(); o.close()
    return lin_g, rgb_g, st_g, lin_o, rgb_o, st_o


def main():
    for name in ("c1", "c2", "c3", "c4", "c5"):
        sc = scenes.make_scene(name, **SMALL[name])
        lin_g, rgb_g, st_g, lin_o, rgb_o, st_o = pair(sc)
        diff = np.abs(lin_g - lin_o)
        scale = max(float(lin_o.mean()), 1e-6)
        bad = (diff.max(axis=2) > 1e-3 * np.maximum(lin_o.max(axis=2), scale)).mean()
        d8 = np.abs(rgb_g.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=2)
        print(f"{name} 16 spp: rays gpu/oracle {st_g.rays / st_o.rays:.5f}  median |d| / scale {np.median(diff) / max(scale, 1.0):.2e}  "
              f"decorrelated pixels {bad:.5f}  |mean diff| / scale {abs(float(lin_g.mean()) - float(lin_o.mean())) / scale:.2e}  "
              f"u8 identical {(d8 == 0).mean():.5f}  within 1 LSB {(d8 <= 1).mean():.5f}  bit-identical linear pixels "
              f"{(diff.max(axis=2) == 0).mean():.5f}", flush=True)
    for name, kw in CONVERGED.items():
        args = dict(SMALL[name]); args.update(kw)
        sc = scenes.make_scene(name, **args)
        lin_g, rgb_g, st_g, lin_o, rgb_o, st_o = pair(sc)
        finite = np.isfinite(lin_o).all(axis=2) & np.isfinite(lin_g).all(axis=2)
        err = (lin_g - lin_o)[finite]
        rmse = float(np.sqrt((err ** 2).mean()))
        mean = float(lin_o[finite].mean())
        d8 = np.abs(rgb_g.astype(np.int32) - rgb_o.astype(np.int32)).max(axis=2)
        print(f"{name} 1024 spp: finite {finite.mean():.5f}  rmse / mean {rmse / mean:.3e}  psnr {10 * math.log10(1.0 / max(rmse ** 2, 1e-20)):.1f} dB  "
              f"u8 identical {(d8 == 0).mean():.5f}  within 1 LSB {(d8 <= 1).mean():.5f}  max u8 diff {d8.max()}", flush=True)


if __name__ == "__main__":
    main()
