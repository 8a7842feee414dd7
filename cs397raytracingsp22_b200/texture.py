"""Texture (src/util/texture.rs): decode on the host, nearest-neighbour taps on the device.

`Texture.load_from_file` mirrors texture.rs:16-25: any decodable image -> RGB8, and **None on
failure** (the reference silently renders with the Q7 defaults when a map is missing, as it does
for the five Drone_*.tga maps absent from the checkout).  PNG, JPEG and TGA are decoded by the library's own
readers (csrc/rt_png.cpp, rt_jpeg.cpp, rt_lower.cpp) - what the `image` crate does in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _ffi


class Texture:
    def __init__(self, rgb8: np.ndarray):
        rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
        assert rgb8.ndim == 3 and rgb8.shape[2] == 3
        self.rgb8 = rgb8  # row 0 = top, like image::DynamicImage

    @property
    def width(self) -> int:
        return self.rgb8.shape[1]

    @property
    def height(self) -> int:
        return self.rgb8.shape[0]

    @staticmethod
    def load_from_file(file_name: str) -> "Texture | None":
        try:
            if file_name.lower().endswith(".tga"):
                with open(file_name, "rb") as f:
                    return Texture(_ffi.tga_decode(f.read()))
            if file_name.lower().endswith(".png"):
                with open(file_name, "rb") as f:
                    return Texture(_ffi.png_decode(f.read()))
            if file_name.lower().endswith((".jpg", ".jpeg")):
                with open(file_name, "rb") as f:
                    return Texture(_ffi.jpeg_decode(f.read()))
            return None  # like image::open on a format it does not know: the caller gets no texture (texture.rs:17-22)
        except Exception:
            return None

    def sample(self, uv) -> np.ndarray:
        """Host restatement of texture.rs:26-32, for tests of the addressing rule only."""
        f32 = np.float32
        u = min(max(f32(uv[0]), f32(0.0)), f32(0.999))
        v = min(max(f32(uv[1]), f32(0.0)), f32(0.999))
        x = min(int(f32(u * f32(self.width))), self.width - 1)
        y = min(int(f32((f32(1.0) - v) * f32(self.height))), self.height - 1)
        return self.rgb8[y, x].astype(np.float32) / np.float32(255.0)

    def lower(self, b) -> int:
        return b.add_texture(self.rgb8)
