"""cs397raytracingsp22_b200 — B200-native (sm_100a) path tracer behind the scene-construction API
of mbk6/CS397RayTracingSP22.  Only the hot path is here: camera rays -> BVH / primitive
intersection -> material scattering -> accumulate / output transform.

Layout
  csrc/           CUDA kernels, host lowering and the C ABI (include/rt_b200.h) -> librt_b200.so
  _ffi.py         ctypes binding of that ABI
  tracing.py, geometry.py, materials.py, texture.py, cgmath.py
                  host-side mirror of the reference's modules of the same names
  scenes.py       the five BASELINE.json configurations
  distributed.py  one-process-per-GPU sharding + the framebuffer reduce
"""
from . import _ffi, build, cgmath  # noqa: F401
from .geometry import ConvexVolume, MeshData, Plane, Sphere, StaticMesh, Triangle, load_obj  # noqa: F401
from .materials import Dielectric, Isotropic, Lambertian, Metal, ParameterizedMaterial  # noqa: F401
from .texture import Texture  # noqa: F401
from .tracing import Camera, CameraProjectionMode, Scene, ShadingMode  # noqa: F401

__all__ = ["Camera", "CameraProjectionMode", "ShadingMode", "Scene", "Sphere", "Triangle", "Plane", "ConvexVolume",
           "StaticMesh", "MeshData", "load_obj", "Lambertian", "Metal", "Dielectric", "ParameterizedMaterial",
           "Isotropic", "Texture", "cgmath"]
