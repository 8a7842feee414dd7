"""Geometry of the reference's scene API (src/util/geometry.rs), host side.

The `Intersectable` trait objects (tracing.rs:42-47) become descriptions that `lower()` into the
back end; `intersect_ray` itself runs in the CUDA extend kernel.  Field names follow the Rust
structs: Sphere (geometry.rs:389-393), Triangle (:424-429), Plane (:468-472), ConvexVolume
(:495-500), StaticMesh (:127-134) with StaticMesh::load_from_file (:138-172).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np

from . import _ffi, cgmath
from .texture import Texture


@dataclass(eq=False)
class Sphere:
    center: tuple
    radius: float
    material: object

    def lower(self, b, ctx) -> int:
        return b.add_sphere(self.center, self.radius, ctx.material(self.material))


@dataclass(eq=False)
class Triangle:
    a: tuple
    b: tuple
    c: tuple
    material: object

    def lower(self, b, ctx) -> int:
        return b.add_triangle(self.a, self.b, self.c, ctx.material(self.material))


@dataclass(eq=False)
class Plane:
    point: tuple
    normal: tuple
    material: object

    def lower(self, b, ctx) -> int:
        return b.add_plane(self.point, self.normal, ctx.material(self.material))


@dataclass(eq=False)
class ConvexVolume:
    boundary: object        # a Sphere or a StaticMesh; other Intersectables can never give an exit hit (geometry.rs:508-509)
    phase_function: object
    density: float

    def lower(self, b, ctx) -> int:
        # the boundary's own material / textures are ignored by the reference as well (geometry.rs:505-510)
        if isinstance(self.boundary, Sphere):
            return b.add_volume_sphere(self.boundary.center, self.boundary.radius, self.density,
                                       ctx.material(self.phase_function))
        if isinstance(self.boundary, StaticMesh):
            m = self.boundary
            return b.add_volume_mesh(ctx.mesh(m.mesh), cgmath.colmajor(m.transform), cgmath.colmajor(m.inv_transform),
                                     self.density, ctx.material(self.phase_function))
        raise _ffi.RtError(_ffi.RT_ERR_UNSUPPORTED, "ConvexVolume: the boundary must be a Sphere or a StaticMesh")


class MeshData:
    """What tobj hands to StaticMesh: single-index arrays of one model (geometry.rs:157)."""

    def __init__(self, pos, nrm, uv, idx, name=""):
        self.pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.nrm = np.ascontiguousarray(nrm, dtype=np.float32)
        self.uv = np.ascontiguousarray(uv, dtype=np.float32)
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self.name = name

    @property
    def ntris(self) -> int:
        return self.idx.shape[0]


_MESH_CACHE: dict = {}


def load_obj(file_name: str) -> MeshData:
    """tobj::load_obj(single_index, triangulate) via the library's parser; `.obj.gz` is accepted."""
    key = os.path.abspath(file_name)
    if key in _MESH_CACHE:
        return _MESH_CACHE[key]
    path = file_name
    if not os.path.exists(path) and os.path.exists(path + ".gz"):
        path = path + ".gz"
    with open(path, "rb") as f:
        data = f.read()
    if path.endswith(".gz"):
        data = gzip.decompress(data)
    pos, nrm, uv, idx, has_n, has_t = _ffi.parse_obj(data)
    if not (has_n and has_t):
        # the reference indexes normals/texcoords unconditionally and would panic (geometry.rs:230-243)
        raise _ffi.RtError(_ffi.RT_ERR_IO, f"{file_name}: the mesh needs vn and vt on every face corner")
    m = MeshData(pos, nrm, uv, idx, name=os.path.basename(file_name))
    _MESH_CACHE[key] = m
    return m


class StaticMesh:
    def __init__(self, mesh: MeshData, textures, material, transform):
        self.mesh = mesh
        self.textures = list(textures)  # 0 albedo, 1 emission, 2 metallic, 3 roughness, 4 normal (geometry.rs:130)
        self.material = material
        self.transform = np.asarray(transform, dtype=np.float32)
        inv = cgmath.inverse_transform(self.transform)
        if inv is None:
            raise _ffi.RtError(_ffi.RT_ERR_INVALID, "StaticMesh: singular transform (the reference panics, geometry.rs:168)")
        self.inv_transform = inv

    @staticmethod
    def load_from_file(file_name, albedo_path=None, emission_path=None, metallic_path=None, roughness_path=None,
                       normal_path=None, material=None, transform=None) -> "StaticMesh":
        mesh = load_obj(file_name)
        texs = [Texture.load_from_file(p) if p is not None else None
                for p in (albedo_path, emission_path, metallic_path, roughness_path, normal_path)]
        return StaticMesh(mesh, texs, material, cgmath.identity() if transform is None else transform)

    def with_transform(self, transform, material="same") -> "StaticMesh":
        """`#[derive(Clone)]` + a new transform: shares the Arc<Mesh> (and so the BLAS)."""
        return StaticMesh(self.mesh, self.textures, self.material if material == "same" else material, transform)

    def lower(self, b, ctx) -> int:
        mesh_id = ctx.mesh(self.mesh)
        tex = [ctx.texture(t) if t is not None else -1 for t in self.textures]
        mat = ctx.material(self.material) if self.material is not None else -1
        return b.add_instance(mesh_id, cgmath.colmajor(self.transform), cgmath.colmajor(self.inv_transform), mat, tex)
