"""ctypes binding of librt_b200.so — the C ABI declared in include/rt_b200.h.

This is the stub a reference-side maintainer would write (see INTEGRATION.md for the Rust
`extern "C"` version).  The product path FAILS LOUDLY when the CUDA library is missing or no GPU
is present; there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

RT_OK = 0
RT_ERR_INVALID, RT_ERR_UNSUPPORTED, RT_ERR_CUDA, RT_ERR_IO, RT_ERR_NOT_COMMITTED = -1, -2, -3, -4, -5
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_PARAMETERIZED, RT_MAT_ISOTROPIC = range(5)
RT_B200_ABI_VERSION = 3  # include/rt_b200.h
RT_PROJ_ORTHOGRAPHIC, RT_PROJ_PERSPECTIVE = 0, 1
RT_SHADE_PHONG, RT_SHADE_PATHTRACE = 0, 1
RT_SHARD_ALL, RT_SHARD_SAMPLES, RT_SHARD_TILES = 0, 1, 2
RT_OPT_COUNTERS, RT_OPT_NO_EVENTS = 1, 2
RT_ENGINE_AUTO, RT_ENGINE_WAVEFRONT, RT_ENGINE_MEGAKERNEL = 0, 1, 2
RT_RAYSORT_AUTO, RT_RAYSORT_OFF, RT_RAYSORT_ON = 0, 1, 2
RT_ORDER_AUTO, RT_ORDER_PIXEL_MAJOR, RT_ORDER_SAMPLE_MAJOR, RT_ORDER_GROUPED = 0, 1, 2, 3
ENGINES = {"auto": RT_ENGINE_AUTO, "wavefront": RT_ENGINE_WAVEFRONT, "megakernel": RT_ENGINE_MEGAKERNEL}


class RtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"librt_b200 error {code}: {msg}")
        self.code = code


class rt_material_desc(C.Structure):
    _fields_ = [("tag", C.c_uint32), ("albedo", C.c_float * 3), ("emission", C.c_float * 3),
                ("roughness", C.c_float), ("metallic", C.c_float), ("ior", C.c_float)]


class rt_camera(C.Structure):
    _fields_ = [("eyepoint", C.c_float * 3), ("view_dir", C.c_float * 3), ("up", C.c_float * 3),
                ("projection_mode", C.c_uint32), ("shading_mode", C.c_uint32), ("path_depth", C.c_uint32),
                ("path_samples", C.c_uint32), ("screen_width", C.c_uint32), ("screen_height", C.c_uint32),
                ("focal_length", C.c_float), ("focus_dist", C.c_float), ("lens_radius", C.c_float),
                ("aa_sample_count", C.c_uint32), ("max_trace_dist", C.c_float), ("gamma", C.c_float)]


class rt_render_opts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("shard_mode", C.c_uint32), ("shard_rank", C.c_uint32),
                ("shard_count", C.c_uint32), ("tile_size", C.c_uint32), ("sample_begin", C.c_uint32),
                ("sample_end", C.c_uint32), ("wavefront", C.c_uint32), ("flags", C.c_uint32),
                ("point_light_pos", C.c_float * 3), ("ambient", C.c_float * 3),
                ("engine", C.c_uint32), ("ray_sort", C.c_uint32), ("work_order", C.c_uint32),
                ("blocks_per_sm", C.c_uint32), ("reserved", C.c_uint32 * 4)]


class rt_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "samples", "rays", "iterations", "kernel_launches", "extend_launches", "shade_launches", "nodes_visited",
        "tlas_nodes_visited",
        "tris_tested", "instances_entered", "prims_tested", "mesh_hits", "texel_taps", "extend_texel_taps",
        "material_fetches", "warp_node_slots")] + [
        (n, C.c_double) for n in ("ms_total", "ms_extend", "ms_shade", "ms_resolve")] + [
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("engine", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class rt_lower_info(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("nodes", C.c_uint32), ("tris", C.c_uint32), ("objects", C.c_uint32),
                ("unbounded", C.c_uint32), ("tlas_depth", C.c_uint32), ("max_blas_depth", C.c_uint32),
                ("guard_boxes", C.c_uint32), ("guarded_tris", C.c_uint32)]


class rt_obj_mesh(C.Structure):
    _fields_ = [("nverts", C.c_uint32), ("ntris", C.c_uint32), ("pos", C.POINTER(C.c_float)),
                ("nrm", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)), ("idx", C.POINTER(C.c_uint32)),
                ("has_normals", C.c_uint32), ("has_texcoords", C.c_uint32)]


_P = C.c_void_p
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)
_U8 = C.POINTER(C.c_uint8)

# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rt_abi_version": (C.c_int, []),
    "rt_last_error": (C.c_char_p, []),
    "rt_device_count": (C.c_int, []),
    "rt_scene_create": (C.c_int, [C.POINTER(_P)]),
    "rt_scene_destroy": (None, [_P]),
    "rt_add_texture": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32]),
    "rt_add_material": (C.c_int, [_P, C.POINTER(rt_material_desc)]),
    "rt_add_mesh": (C.c_int, [_P, _F, _F, _F, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]),
    "rt_add_instance": (C.c_int, [_P, C.c_int, _F, _F, C.c_int, _I]),
    "rt_add_sphere": (C.c_int, [_P, _F, C.c_float, C.c_int]),
    "rt_add_triangle": (C.c_int, [_P, _F, _F, _F, C.c_int]),
    "rt_add_plane": (C.c_int, [_P, _F, _F, C.c_int]),
    "rt_add_volume_sphere": (C.c_int, [_P, _F, C.c_float, C.c_float, C.c_int]),
    "rt_add_volume_mesh": (C.c_int, [_P, C.c_int, _F, _F, C.c_float, C.c_int]),
    "rt_scene_lower": (C.c_int, [_P, C.POINTER(rt_lower_info)]),
    "rt_commit": (C.c_int, [_P, C.c_int]),
    "rt_scene_device_bytes": (C.c_uint64, [_P]),
    "rt_scene_upload": (C.c_int, [_P]),
    "rt_render": (C.c_int, [_P, C.POINTER(rt_camera), C.POINTER(rt_render_opts), _F, _U8, C.POINTER(rt_stats)]),
    "rt_render_accum": (C.c_int, [_P, C.POINTER(rt_camera), C.POINTER(rt_render_opts), _P, _P, C.POINTER(rt_stats)]),
    "rt_resolve": (C.c_int, [_P, C.POINTER(rt_camera), _P, C.c_uint32, _P, _P, _P]),
    "rt_accum_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32]),
    "rt_render_progressive": (C.c_int, [_P, C.POINTER(rt_camera), C.POINTER(rt_render_opts), C.POINTER(C.c_int64), C.c_uint32,
                                        _F, _U8, C.POINTER(rt_stats)]),
    "rt_trace_primary": (C.c_int, [_P, C.POINTER(rt_camera), C.c_uint64, C.c_uint32, _I, _I, _F, _F, _F]),
    "rt_intersect_rays": (C.c_int, [_P, C.c_uint64, C.c_uint32, _F, C.c_float, C.c_float, _I, _I, _F, _F, _F, _F, _I]),
    "rt_obj_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(rt_obj_mesh)]),
    "rt_obj_load": (C.c_int, [C.c_char_p, C.POINTER(rt_obj_mesh)]),
    "rt_obj_free": (None, [C.POINTER(rt_obj_mesh)]),
    "rt_tga_decode": (C.c_int, [_U8, C.c_size_t, C.POINTER(_U8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rt_tga_encode_rgb8": (C.c_int, [_U8, C.c_uint32, C.c_uint32, C.POINTER(_U8), C.POINTER(C.c_size_t)]),
    "rt_png_decode": (C.c_int, [_U8, C.c_size_t, C.POINTER(_U8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rt_jpeg_decode": (C.c_int, [_U8, C.c_size_t, C.POINTER(_U8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rt_png_encode_rgb8": (C.c_int, [_U8, C.c_uint32, C.c_uint32, C.POINTER(_U8), C.POINTER(C.c_size_t)]),
    "rt_free": (None, [_P]),
    "rt_mesh_reachability": (C.c_int, [_F, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, _U8]),
}


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load librt_b200.so (building it in-tree first if it is missing or stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RtError(RT_ERR_CUDA, f"{_build.LIB} is missing: the CUDA extension must be built; "
                                    "there is no CPU fallback")
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise RtError(rc, load().rt_last_error().decode("utf-8", "replace"))
    return rc


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_F)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_I)


def u8ptr(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_U8)


def f3(v) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))
    return a


def parse_obj(text: bytes):
    """tobj::load_obj(single_index, triangulate) stand-in -> (pos[n,3], nrm[n,3], uv[n,2], idx[m,3])."""
    lib = load()
    m = rt_obj_mesh()
    check(lib.rt_obj_parse(text, len(text), C.byref(m)))
    try:
        nv, nt = m.nverts, m.ntris
        pos = np.ctypeslib.as_array(m.pos, shape=(nv * 3,)).copy().reshape(nv, 3)
        nrm = np.ctypeslib.as_array(m.nrm, shape=(nv * 3,)).copy().reshape(nv, 3)
        uv = np.ctypeslib.as_array(m.uv, shape=(nv * 2,)).copy().reshape(nv, 2)
        idx = np.ctypeslib.as_array(m.idx, shape=(nt * 3,)).copy().reshape(nt, 3)
        return pos, nrm, uv, idx, bool(m.has_normals), bool(m.has_texcoords)
    finally:
        lib.rt_obj_free(C.byref(m))


def tga_decode(data: bytes) -> np.ndarray:
    lib = load()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    out = _U8()
    w, h = C.c_uint32(), C.c_uint32()
    check(lib.rt_tga_decode(u8ptr(buf), len(data), C.byref(out), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(out, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.rt_free(out)


def png_decode(data: bytes) -> np.ndarray:
    lib = load()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    out = _U8()
    w, h = C.c_uint32(), C.c_uint32()
    check(lib.rt_png_decode(u8ptr(buf), len(data), C.byref(out), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(out, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.rt_free(out)


def jpeg_decode(data: bytes) -> np.ndarray:
    lib = load()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    out = _U8()
    w, h = C.c_uint32(), C.c_uint32()
    check(lib.rt_jpeg_decode(u8ptr(buf), len(data), C.byref(out), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(out, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.rt_free(out)


def png_encode(rgb: np.ndarray) -> bytes:
    lib = load()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = _U8()
    n = C.c_size_t()
    check(lib.rt_png_encode_rgb8(u8ptr(rgb), rgb.shape[1], rgb.shape[0], C.byref(out), C.byref(n)))
    try:
        return bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    finally:
        lib.rt_free(out)


def tga_encode(rgb: np.ndarray) -> bytes:
    lib = load()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = _U8()
    n = C.c_size_t()
    check(lib.rt_tga_encode_rgb8(u8ptr(rgb), rgb.shape[1], rgb.shape[0], C.byref(out), C.byref(n)))
    try:
        return bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    finally:
        lib.rt_free(out)


def mesh_reachability(pos: np.ndarray, idx: np.ndarray) -> np.ndarray:
    lib = load()
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    mask = np.zeros(idx.shape[0], dtype=np.uint8)
    check(lib.rt_mesh_reachability(fptr(pos.reshape(-1)), pos.shape[0], idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   idx.shape[0], u8ptr(mask)))
    return mask


class GpuBackend:
    """Scene builder over the C ABI: what each reference type's `lower()` talks to."""

    name = "b200"

    def __init__(self):
        self.lib = load()
        h = _P()
        check(self.lib.rt_scene_create(C.byref(h)))
        self.handle = h
        self.device = None

    def close(self):
        if self.handle:
            self.lib.rt_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene construction
    def add_texture(self, rgb8: np.ndarray) -> int:
        rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
        return check(self.lib.rt_add_texture(self.handle, u8ptr(rgb8.reshape(-1)), rgb8.shape[1], rgb8.shape[0]))

    def add_material(self, tag, albedo=(0, 0, 0), emission=(0, 0, 0), roughness=0.0, metallic=0.0, ior=1.0) -> int:
        d = rt_material_desc(tag, (C.c_float * 3)(*albedo), (C.c_float * 3)(*emission), roughness, metallic, ior)
        return check(self.lib.rt_add_material(self.handle, C.byref(d)))

    def add_mesh(self, pos, nrm, uv, idx) -> int:
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1)
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        return check(self.lib.rt_add_mesh(self.handle, fptr(pos), fptr(nrm), fptr(uv), pos.size // 3,
                                          idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size // 3))

    def add_instance(self, mesh, xform_colmajor, inv_colmajor, material, tex) -> int:
        x = np.ascontiguousarray(xform_colmajor, dtype=np.float32).reshape(16)
        inv = np.ascontiguousarray(inv_colmajor, dtype=np.float32).reshape(16)
        t = np.ascontiguousarray(tex, dtype=np.int32).reshape(5)
        return check(self.lib.rt_add_instance(self.handle, mesh, fptr(x), fptr(inv), material, iptr(t)))

    def add_sphere(self, center, radius, material) -> int:
        return check(self.lib.rt_add_sphere(self.handle, fptr(f3(center)), float(radius), material))

    def add_triangle(self, a, b, c, material) -> int:
        return check(self.lib.rt_add_triangle(self.handle, fptr(f3(a)), fptr(f3(b)), fptr(f3(c)), material))

    def add_plane(self, point, normal, material) -> int:
        return check(self.lib.rt_add_plane(self.handle, fptr(f3(point)), fptr(f3(normal)), material))

    def add_volume_sphere(self, center, radius, density, material) -> int:
        return check(self.lib.rt_add_volume_sphere(self.handle, fptr(f3(center)), float(radius), float(density),
                                                   material))

    def add_volume_mesh(self, mesh, xform_colmajor, inv_colmajor, density, material) -> int:
        x = np.ascontiguousarray(xform_colmajor, dtype=np.float32).reshape(16)
        inv = np.ascontiguousarray(inv_colmajor, dtype=np.float32).reshape(16)
        return check(self.lib.rt_add_volume_mesh(self.handle, mesh, fptr(x), fptr(inv), float(density), material))

    def lower_info(self) -> dict:
        """Host-only lowering (no GPU needed): sizes and depths of what rt_commit would upload."""
        info = rt_lower_info()
        check(self.lib.rt_scene_lower(self.handle, C.byref(info)))
        return {n: getattr(info, n) for n, _ in info._fields_}

    # -- device
    def commit(self, device: int = 0):
        check(self.lib.rt_commit(self.handle, device))
        self.device = device

    def upload(self):
        check(self.lib.rt_scene_upload(self.handle))

    def device_bytes(self) -> int:
        return int(self.lib.rt_scene_device_bytes(self.handle))

    def render(self, cam: rt_camera, opts: rt_render_opts | None = None, want_linear=True, want_rgb8=True,
               out_linear: np.ndarray | None = None, out_rgb8: np.ndarray | None = None):
        w, h = cam.screen_width, cam.screen_height
        lin = out_linear if out_linear is not None else (np.empty((h, w, 3), np.float32) if want_linear else None)
        rgb = out_rgb8 if out_rgb8 is not None else (np.empty((h, w, 3), np.uint8) if want_rgb8 else None)
        st = rt_stats()
        o = opts if opts is not None else rt_render_opts()
        check(self.lib.rt_render(self.handle, C.byref(cam), C.byref(o), fptr(lin.reshape(-1)) if lin is not None else None,
                                 u8ptr(rgb.reshape(-1)) if rgb is not None else None, C.byref(st)))
        return lin, rgb, st

    def render_progressive(self, cam: rt_camera, opts: rt_render_opts, accum: np.ndarray, spp_in_accum: int,
                           want_linear=True, want_rgb8=True):
        """Adds opts.sample_begin..sample_end to the HOST accumulator `accum` (int64, H*W*4; the checkpoint) and
        resolves the image of the spp_in_accum samples per pixel it then holds."""
        w, h = cam.screen_width, cam.screen_height
        assert accum.dtype == np.int64 and accum.size == h * w * 4 and accum.flags["C_CONTIGUOUS"]
        lin = np.empty((h, w, 3), np.float32) if want_linear else None
        rgb = np.empty((h, w, 3), np.uint8) if want_rgb8 else None
        st = rt_stats()
        check(self.lib.rt_render_progressive(self.handle, C.byref(cam), C.byref(opts), accum.ctypes.data_as(C.POINTER(C.c_int64)),
                                             spp_in_accum, fptr(lin.reshape(-1)) if lin is not None else None,
                                             u8ptr(rgb.reshape(-1)) if rgb is not None else None, C.byref(st)))
        return lin, rgb, st

    def render_accum(self, cam: rt_camera, opts: rt_render_opts, d_accum_ptr: int, stream_ptr: int = 0) -> rt_stats:
        st = rt_stats()
        check(self.lib.rt_render_accum(self.handle, C.byref(cam), C.byref(opts), _P(d_accum_ptr), _P(stream_ptr),
                                       C.byref(st)))
        return st

    def resolve(self, cam: rt_camera, d_accum_ptr: int, total_spp: int, d_linear_ptr: int = 0, d_rgb8_ptr: int = 0,
                stream_ptr: int = 0):
        check(self.lib.rt_resolve(self.handle, C.byref(cam), _P(d_accum_ptr), total_spp, _P(d_linear_ptr or None),
                                  _P(d_rgb8_ptr or None), _P(stream_ptr)))

    def trace_primary(self, cam: rt_camera, seed: int, sample: int):
        n = cam.screen_width * cam.screen_height
        obj = np.empty(n, np.int32); prim = np.empty(n, np.int32)
        t = np.empty(n, np.float32); nrm = np.empty((n, 3), np.float32); ray = np.empty((n, 6), np.float32)
        check(self.lib.rt_trace_primary(self.handle, C.byref(cam), seed, sample, iptr(obj), iptr(prim), fptr(t),
                                        fptr(nrm.reshape(-1)), fptr(ray.reshape(-1))))
        return dict(obj=obj, prim=prim, t=t, normal=nrm, ray=ray)

    def intersect_rays(self, rays: np.ndarray, t_min: float, t_max: float, seed: int = 0):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj = np.empty(n, np.int32); prim = np.empty(n, np.int32); front = np.empty(n, np.int32)
        t = np.empty(n, np.float32); nrm = np.empty((n, 3), np.float32); hp = np.empty((n, 3), np.float32)
        uv = np.empty((n, 2), np.float32)
        check(self.lib.rt_intersect_rays(self.handle, seed, n, fptr(rays.reshape(-1)), t_min, t_max, iptr(obj),
                                         iptr(prim), fptr(t), fptr(nrm.reshape(-1)), fptr(hp.reshape(-1)),
                                         fptr(uv.reshape(-1)), iptr(front)))
        return dict(obj=obj, prim=prim, t=t, normal=nrm, hitpoint=hp, uv=uv, frontface=front)
