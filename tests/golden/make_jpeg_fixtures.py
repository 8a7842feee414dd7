#!/usr/bin/env python
"""Makes tests/golden/jpeg/*.jpg and expected.npz (what Pillow / libjpeg-turbo decodes them to): small files that
cover the decoder's branches - grey, 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1 sampling, odd sizes, progressive with
successive-approximation scans, restart intervals, optimised Huffman tables.  Needs Pillow; the tests do not."""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "jpeg")


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(x / 7.0 + seed) * np.cos(y / 9.0), 128 + 90 * np.cos((x + y) / 11.0),
                    255 * ((x // 8 + y // 8) % 2)], axis=2)
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = {
    "grey_baseline_37x29": dict(size=(37, 29), mode="L", quality=85),
    "rgb444_optimized_40x24": dict(size=(40, 24), subsampling=0, quality=95, optimize=True),
    "rgb422_progressive_50x33": dict(size=(50, 33), subsampling=1, quality=80, progressive=True),
    "rgb420_progressive_45x51": dict(size=(45, 51), subsampling=2, quality=60, progressive=True),
    "rgb420_restart_67x35": dict(size=(67, 35), subsampling=2, quality=75, restart_marker_blocks=3),
    "rgb444_restart_rows_33x41": dict(size=(33, 41), subsampling=0, quality=70, restart_marker_rows=1, progressive=True),
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    expected = {}
    for seed, (name, kw) in enumerate(CASES.items()):
        kw = dict(kw)
        w, h = kw.pop("size")
        mode = kw.pop("mode", "RGB")
        px = picture(w, h, seed)
        im = Image.fromarray(px if mode == "RGB" else px[..., 0], mode)
        buf = io.BytesIO()
        im.save(buf, "JPEG", **kw)
        path = os.path.join(OUT, name + ".jpg")
        open(path, "wb").write(buf.getvalue())
        expected[name] = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
        print(name, len(buf.getvalue()), "bytes")
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    # the two JPEG textures of the reference's run() scene (assets/texture/), every 8th pixel
    root = os.path.dirname(os.path.dirname(HERE))
    for n in ("magenta", "normal_test"):
        a = np.asarray(Image.open(os.path.join(root, "assets", "texture", n + ".jpg")).convert("RGB"), dtype=np.uint8)
        expected[f"asset_{n}_stride8"] = a[::8, ::8].copy()
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
