"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see the header of oracle/oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
`OracleBackend` has the same builder methods as the product's GpuBackend, so one Scene description
lowers into either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cs397raytracingsp22_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
MODE_REF_TREE, MODE_BRUTE = 0, 1

_LIB = None
_P, _F, _I, _U8 = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "oracle.cpp")
    hdr = os.path.join(ROOT, "include", "rt_b200.h")
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(p) > os.path.getmtime(LIB) for p in (src, hdr))
    if force or stale:
        r = subprocess.run(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB):
        build()
    else:
        try:
            build()
        except Exception:
            pass  # no compiler on this box: use the prebuilt library
    lib = C.CDLL(LIB)
    cam, st = C.POINTER(_ffi.rt_camera), C.POINTER(_ffi.rt_stats)
    sig = {
        "orc_scene_create": (C.c_int, [C.POINTER(_P)]),
        "orc_scene_destroy": (None, [_P]),
        "orc_add_texture": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32]),
        "orc_add_material": (C.c_int, [_P, C.POINTER(_ffi.rt_material_desc)]),
        "orc_add_mesh": (C.c_int, [_P, _F, _F, _F, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]),
        "orc_mesh_reachability": (C.c_int, [_P, C.c_int, _U8]),
        "orc_add_instance": (C.c_int, [_P, C.c_int, _F, _F, C.c_int, _I]),
        "orc_add_sphere": (C.c_int, [_P, _F, C.c_float, C.c_int]),
        "orc_add_triangle": (C.c_int, [_P, _F, _F, _F, C.c_int]),
        "orc_add_plane": (C.c_int, [_P, _F, _F, C.c_int]),
        "orc_add_volume_sphere": (C.c_int, [_P, _F, C.c_float, C.c_float, C.c_int]),
        "orc_add_volume_mesh": (C.c_int, [_P, C.c_int, _F, _F, C.c_float, C.c_int]),
        "orc_render": (C.c_int, [_P, cam, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_int, _F, _U8, st]),
        "orc_set_lights": (C.c_int, [_P, _F, _F]),
        "orc_trace_primary": (C.c_int, [_P, cam, C.c_uint64, C.c_int, C.c_uint32, _I, _I, _F, _F, _F]),
        "orc_intersect_rays": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_uint32, _F, C.c_float, C.c_float, _I, _I, _F,
                                         _F, _F, _F, _I]),
        "orc_camera_offsets": (C.c_int, [cam, C.c_uint64, C.c_uint32, C.c_uint32, _F]),
        "orc_sample_ball_disk": (C.c_int, [C.c_uint64, C.c_uint32, _F, _F]),
        "orc_philox": (C.c_int, [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]),
        "orc_sample_hemisphere": (C.c_int, [_F, _F, _F]),
        "orc_scatter": (C.c_int, [C.POINTER(_ffi.rt_material_desc), _F, _F, C.c_int, C.c_uint64, C.c_uint32, _F, _F, _F]),
        "orc_texture_sample": (C.c_int, [_P, C.c_int, C.c_float, C.c_float, _F]),
        "orc_output_transform": (C.c_int, [_F, C.c_float, _U8]),
        "orc_num_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def _check(rc: int) -> int:
    if rc < 0:
        raise RuntimeError(f"oracle error {rc}")
    return rc


class OracleBackend:
    name = "oracle"

    def __init__(self):
        self.lib = load()
        h = _P()
        _check(self.lib.orc_scene_create(C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.lib.orc_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_texture(self, rgb8) -> int:
        rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
        return _check(self.lib.orc_add_texture(self.handle, _ffi.u8ptr(rgb8.reshape(-1)), rgb8.shape[1], rgb8.shape[0]))

    def add_material(self, tag, albedo=(0, 0, 0), emission=(0, 0, 0), roughness=0.0, metallic=0.0, ior=1.0) -> int:
        d = _ffi.rt_material_desc(tag, (C.c_float * 3)(*albedo), (C.c_float * 3)(*emission), roughness, metallic, ior)
        return _check(self.lib.orc_add_material(self.handle, C.byref(d)))

    def add_mesh(self, pos, nrm, uv, idx) -> int:
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        return _check(self.lib.orc_add_mesh(self.handle, _ffi.fptr(pos), _ffi.fptr(nrm), _ffi.fptr(uv), pos.size // 3,
                                            idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size // 3))

    def mesh_reachability(self, mesh: int, ntris: int) -> np.ndarray:
        mask = np.zeros(ntris, np.uint8)
        _check(self.lib.orc_mesh_reachability(self.handle, mesh, _ffi.u8ptr(mask)))
        return mask

    def add_instance(self, mesh, xform_colmajor, inv_colmajor, material, tex) -> int:
        x = np.ascontiguousarray(xform_colmajor, dtype=np.float32).reshape(16)
        inv = np.ascontiguousarray(inv_colmajor, dtype=np.float32).reshape(16)
        t = np.ascontiguousarray(tex, dtype=np.int32).reshape(5)
        return _check(self.lib.orc_add_instance(self.handle, mesh, _ffi.fptr(x), _ffi.fptr(inv), material, _ffi.iptr(t)))

    def add_sphere(self, center, radius, material) -> int:
        return _check(self.lib.orc_add_sphere(self.handle, _ffi.fptr(_ffi.f3(center)), float(radius), material))

    def add_triangle(self, a, b, c, material) -> int:
        return _check(self.lib.orc_add_triangle(self.handle, _ffi.fptr(_ffi.f3(a)), _ffi.fptr(_ffi.f3(b)),
                                                _ffi.fptr(_ffi.f3(c)), material))

    def add_plane(self, point, normal, material) -> int:
        return _check(self.lib.orc_add_plane(self.handle, _ffi.fptr(_ffi.f3(point)), _ffi.fptr(_ffi.f3(normal)), material))

    def add_volume_sphere(self, center, radius, density, material) -> int:
        return _check(self.lib.orc_add_volume_sphere(self.handle, _ffi.fptr(_ffi.f3(center)), float(radius),
                                                     float(density), material))

    def add_volume_mesh(self, mesh, xform_colmajor, inv_colmajor, density, material) -> int:
        x = np.ascontiguousarray(xform_colmajor, dtype=np.float32).reshape(16)
        inv = np.ascontiguousarray(inv_colmajor, dtype=np.float32).reshape(16)
        return _check(self.lib.orc_add_volume_mesh(self.handle, mesh, _ffi.fptr(x), _ffi.fptr(inv), float(density), material))

    def commit(self, device: int = 0):
        pass

    def set_lights(self, point_light_pos, ambient):
        """Scene::point_light_pos / Scene::ambient (tracing.rs:216-217), ShadingMode::Phong only."""
        p = np.asarray(point_light_pos, np.float32); a = np.asarray(ambient, np.float32)
        _check(self.lib.orc_set_lights(self.handle, _ffi.fptr(p), _ffi.fptr(a)))

    def render(self, cam, seed=0x5EED, mode=MODE_REF_TREE, sample_begin=0, sample_end=0, nthreads=0,
               want_linear=True, want_rgb8=True):
        w, h = cam.screen_width, cam.screen_height
        lin = np.empty((h, w, 3), np.float32) if want_linear else None
        rgb = np.empty((h, w, 3), np.uint8) if want_rgb8 else None
        st = _ffi.rt_stats()
        _check(self.lib.orc_render(self.handle, C.byref(cam), seed, mode, sample_begin, sample_end, nthreads,
                                   _ffi.fptr(lin.reshape(-1)) if lin is not None else None,
                                   _ffi.u8ptr(rgb.reshape(-1)) if rgb is not None else None, C.byref(st)))
        return lin, rgb, st

    def trace_primary(self, cam, seed: int, sample: int, mode=MODE_REF_TREE):
        n = cam.screen_width * cam.screen_height
        obj = np.empty(n, np.int32); prim = np.empty(n, np.int32)
        t = np.empty(n, np.float32); nrm = np.empty((n, 3), np.float32); ray = np.empty((n, 6), np.float32)
        _check(self.lib.orc_trace_primary(self.handle, C.byref(cam), seed, mode, sample, _ffi.iptr(obj), _ffi.iptr(prim),
                                          _ffi.fptr(t), _ffi.fptr(nrm.reshape(-1)), _ffi.fptr(ray.reshape(-1))))
        return dict(obj=obj, prim=prim, t=t, normal=nrm, ray=ray)

    def intersect_rays(self, rays, t_min, t_max, seed=0, mode=MODE_REF_TREE):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj = np.empty(n, np.int32); prim = np.empty(n, np.int32); front = np.empty(n, np.int32)
        t = np.empty(n, np.float32); nrm = np.empty((n, 3), np.float32); hp = np.empty((n, 3), np.float32)
        uv = np.empty((n, 2), np.float32)
        _check(self.lib.orc_intersect_rays(self.handle, seed, mode, n, _ffi.fptr(rays.reshape(-1)), t_min, t_max,
                                           _ffi.iptr(obj), _ffi.iptr(prim), _ffi.fptr(t), _ffi.fptr(nrm.reshape(-1)),
                                           _ffi.fptr(hp.reshape(-1)), _ffi.fptr(uv.reshape(-1)), _ffi.iptr(front)))
        return dict(obj=obj, prim=prim, t=t, normal=nrm, hitpoint=hp, uv=uv, frontface=front)


def lower_to_oracle(scene) -> OracleBackend:
    b = OracleBackend()
    scene.lower(b)
    b.set_lights(scene.point_light_pos, scene.ambient)
    return b


def camera_offsets(cam, seed, x, y) -> np.ndarray:
    out = np.empty((cam.aa_sample_count, 2), np.float32)
    _check(load().orc_camera_offsets(C.byref(cam), seed, x, y, _ffi.fptr(out.reshape(-1))))
    return out


def sample_ball_disk(seed, n):
    ball = np.empty((n, 3), np.float32); disk = np.empty((n, 2), np.float32)
    _check(load().orc_sample_ball_disk(seed, n, _ffi.fptr(ball.reshape(-1)), _ffi.fptr(disk.reshape(-1))))
    return ball, disk


def philox(c, k):
    out = (C.c_uint32 * 4)()
    load().orc_philox(c[0], c[1], c[2], c[3], k[0], k[1], out)
    return list(out)


def sample_hemisphere(normal, ball):
    out = np.empty(3, np.float32)
    load().orc_sample_hemisphere(_ffi.fptr(_ffi.f3(normal)), _ffi.fptr(_ffi.f3(ball)), _ffi.fptr(out))
    return out


def output_transform(mean, gamma):
    out = np.empty(3, np.uint8)
    load().orc_output_transform(_ffi.fptr(_ffi.f3(mean)), gamma, _ffi.u8ptr(out))
    return out


def scatter(tag, normal, direction, frontface=True, seed=1, n=1000, albedo=(1, 1, 1), roughness=0.0, metallic=0.0, ior=1.0):
    """Material::scatter for n draws at one surface point -> (directions, brdf terms, pdfs)."""
    d = _ffi.rt_material_desc()
    d.tag = tag
    d.albedo[:] = [float(v) for v in albedo]
    d.roughness, d.metallic, d.ior = float(roughness), float(metallic), float(ior)
    nrm = np.asarray(normal, np.float32); dr = np.asarray(direction, np.float32)
    out = np.empty((n, 3), np.float32); brdf = np.empty((n, 3), np.float32); pdf = np.empty(n, np.float32)
    _check(load().orc_scatter(C.byref(d), _ffi.fptr(nrm), _ffi.fptr(dr), int(bool(frontface)), seed, n,
                              _ffi.fptr(out.reshape(-1)), _ffi.fptr(brdf.reshape(-1)), _ffi.fptr(pdf)))
    return out, brdf, pdf
