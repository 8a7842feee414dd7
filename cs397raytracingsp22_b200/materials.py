"""Materials of the reference's scene API (src/util/materials.rs), as plain descriptions.

The reference's `Material::scatter` / `emission` (materials.rs:12-15) run per ray on the CPU; here a
material is data: `lower()` registers it in the back end's tagged-union material table and the
scatter code lives in the CUDA shade kernel (csrc/rt_kernels.cu, k_shade).
"""
from __future__ import annotations

from dataclasses import dataclass, field

from . import _ffi


def _v(x):
    return tuple(float(c) for c in x)


@dataclass(eq=False)
class Lambertian:  # materials.rs:20-32
    albedo: tuple = (1.0, 1.0, 1.0)
    emission: tuple = (0.0, 0.0, 0.0)

    def lower(self, b) -> int:
        return b.add_material(_ffi.RT_MAT_LAMBERTIAN, albedo=_v(self.albedo), emission=_v(self.emission))


@dataclass(eq=False)
class Metal:  # materials.rs:51-55
    albedo: tuple = (1.0, 1.0, 1.0)
    emission: tuple = (0.0, 0.0, 0.0)
    roughness: float = 0.0

    def lower(self, b) -> int:
        return b.add_material(_ffi.RT_MAT_METAL, albedo=_v(self.albedo), emission=_v(self.emission),
                              roughness=float(self.roughness))


@dataclass(eq=False)
class Dielectric:  # materials.rs:74-76
    idx_of_refraction: float = 1.5

    def lower(self, b) -> int:
        return b.add_material(_ffi.RT_MAT_DIELECTRIC, ior=float(self.idx_of_refraction))


@dataclass(eq=False)
class ParameterizedMaterial:  # materials.rs:107-112
    albedo: tuple = (1.0, 1.0, 1.0)
    emission: tuple = (0.0, 0.0, 0.0)
    roughness: float = 1.0
    metallic: float = 0.0

    def lower(self, b) -> int:
        return b.add_material(_ffi.RT_MAT_PARAMETERIZED, albedo=_v(self.albedo), emission=_v(self.emission),
                              roughness=float(self.roughness), metallic=float(self.metallic))


@dataclass(eq=False)
class Isotropic:  # materials.rs:152-157
    albedo: tuple = (1.0, 1.0, 1.0)
    emission: tuple = (0.0, 0.0, 0.0)

    def lower(self, b) -> int:
        return b.add_material(_ffi.RT_MAT_ISOTROPIC, albedo=_v(self.albedo), emission=_v(self.emission))
