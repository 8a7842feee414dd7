"""Camera and Scene of the reference's scene API (src/util/tracing.rs:138-155, 213-263), host side.

`Scene.render_to_image()` keeps the reference's call shape (tracing.rs:221) but does
lower -> C ABI -> sm_100a kernels instead of the rayon row loop (tracing.rs:228).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field

import numpy as np

from . import _ffi


class CameraProjectionMode(enum.IntEnum):  # tracing.rs:27-30
    Orthographic = _ffi.RT_PROJ_ORTHOGRAPHIC
    Perspective = _ffi.RT_PROJ_PERSPECTIVE


class ShadingMode(enum.IntEnum):  # tracing.rs:32-35
    Phong = _ffi.RT_SHADE_PHONG
    PathTrace = _ffi.RT_SHADE_PATHTRACE


@dataclass
class Camera:  # tracing.rs:138-155, same field names
    eyepoint: tuple = (0.0, 2.0, 5.5)
    view_dir: tuple = (0.0, 0.0, -1.0)
    up: tuple = (0.0, 1.0, 0.0)
    projection_mode: CameraProjectionMode = CameraProjectionMode.Perspective
    shading_mode: ShadingMode = ShadingMode.PathTrace
    path_depth: int = 10
    path_samples: int = 1
    screen_width: int = 100
    screen_height: int = 100
    focal_length: float = 0.6
    focus_dist: float = 5.0
    lens_radius: float = 0.0
    aa_sample_count: int = 100
    max_trace_dist: float = 100.0
    gamma: float = 2.0

    def to_c(self) -> _ffi.rt_camera:
        import ctypes as C
        c = _ffi.rt_camera()
        c.eyepoint = (C.c_float * 3)(*self.eyepoint)
        c.view_dir = (C.c_float * 3)(*self.view_dir)
        c.up = (C.c_float * 3)(*self.up)
        c.projection_mode = int(self.projection_mode)
        c.shading_mode = int(self.shading_mode)
        c.path_depth = self.path_depth
        c.path_samples = self.path_samples
        c.screen_width = self.screen_width
        c.screen_height = self.screen_height
        c.focal_length = self.focal_length
        c.focus_dist = self.focus_dist
        c.lens_radius = self.lens_radius
        c.aa_sample_count = self.aa_sample_count
        c.max_trace_dist = self.max_trace_dist
        c.gamma = self.gamma
        return c


class LowerContext:
    """De-duplicates shared `Arc`s (materials, textures, meshes) while a scene is lowered."""

    def __init__(self, backend):
        self.b = backend
        self._mat, self._tex, self._mesh = {}, {}, {}

    def material(self, m) -> int:
        k = id(m)
        if k not in self._mat:
            self._mat[k] = (m.lower(self.b), m)
        return self._mat[k][0]

    def texture(self, t) -> int:
        k = id(t)
        if k not in self._tex:
            self._tex[k] = (t.lower(self.b), t)
        return self._tex[k][0]

    def mesh(self, md) -> int:
        k = id(md)
        if k not in self._mesh:
            self._mesh[k] = (self.b.add_mesh(md.pos, md.nrm, md.uv, md.idx), md)
        return self._mesh[k][0]


@dataclass
class Scene:  # tracing.rs:213-218
    camera: Camera
    objects: list
    point_light_pos: tuple = (0.0, 1.0, 5.0)   # ShadingMode::Phong only (tracing.rs:216)
    ambient: tuple = (0.1, 0.1, 0.1)           # ShadingMode::Phong only (tracing.rs:217)
    seed: int = 0x5EED
    _backend: object = field(default=None, repr=False, compare=False)

    def lower(self, backend):
        """Insertion order is kept: it is the reference's tie-break between objects (tracing.rs:330-341)."""
        ctx = LowerContext(backend)
        for obj in self.objects:
            obj.lower(backend, ctx)
        return backend

    def commit(self, device: int = 0):
        if self._backend is None:
            b = _ffi.GpuBackend()
            self.lower(b)
            b.commit(device)
            self._backend = b
        return self._backend

    def render(self, opts: _ffi.rt_render_opts | None = None, device: int = 0, want_linear=True, want_rgb8=True):
        """-> (linear mean radiance HxWx3 f32, RGB8 image HxWx3, stats)."""
        b = self.commit(device)
        o = opts if opts is not None else _ffi.rt_render_opts()
        if opts is None:
            o.seed = self.seed
        o.point_light_pos[:] = [float(v) for v in self.point_light_pos]
        o.ambient[:] = [float(v) for v in self.ambient]
        return b.render(self.camera.to_c(), o, want_linear=want_linear, want_rgb8=want_rgb8)

    def render_to_image(self, device: int = 0) -> np.ndarray:
        """Scene::render_to_image (tracing.rs:221-263): the RGB8 image, row 0 = top."""
        return self.render(device=device, want_linear=False)[1]

    def close(self):
        if self._backend is not None:
            self._backend.close()
            self._backend = None
