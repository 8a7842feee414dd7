"""The five BASELINE.json configurations, written in the reference's scene vocabulary.

Only C4 exists in the reference tree (the literal in `run()`, tracing.rs:356-543); C1-C3 and C5 are
authored here from the same building blocks and frozen.  HEAD's camera
(tracing.rs:357-373) is the template.  The five Drone_*.tga maps are missing from the reference
checkout (.MISSING_LARGE_BLOBS), so seeded procedural stand-ins are synthesised; the OBJ meshes and
the small textures under assets/ are verbatim copies of the reference's data files (gzip'ed OBJ).
"""
from __future__ import annotations

import os

import numpy as np

from . import cgmath as cg
from .geometry import ConvexVolume, Plane, Sphere, StaticMesh, Triangle, load_obj
from .materials import Dielectric, Isotropic, Lambertian, Metal, ParameterizedMaterial
from .texture import Texture
from .tracing import Camera, Scene

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def obj_path(name: str) -> str:
    return os.path.join(ASSETS, "obj", name + ".obj")


def tex_path(name: str) -> str:
    return os.path.join(ASSETS, "texture", name)


# ----------------------------------------------------------------------------- synthetic drone maps
def _hash2(x: np.ndarray, y: np.ndarray, seed: int) -> np.ndarray:
    """Integer hash -> uint32, pure integer arithmetic so the maps are identical on every machine."""
    h = (x.astype(np.uint64) * np.uint64(0x9E3779B1) + y.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h.astype(np.uint32)


_MAP_CACHE: dict = {}


def drone_maps(size: int = 2048, seed: int = 397) -> list:
    """[albedo, emission, metallic, roughness, normal] as Texture objects (RGB8, size x size).

    Panelled hull: 64-texel plates with per-plate colour / metalness / roughness, thin emissive
    seams on one plate in sixteen, and a normal map that bevels the plate borders.
    """
    key = (size, seed)
    if key in _MAP_CACHE:
        return _MAP_CACHE[key]
    yy, xx = np.meshgrid(np.arange(size, dtype=np.uint32), np.arange(size, dtype=np.uint32), indexing="ij")
    plate = max(size // 32, 4)
    px, py = xx // plate, yy // plate
    lx, ly = (xx % plate).astype(np.int32), (yy % plate).astype(np.int32)
    hp = _hash2(px, py, seed)
    ht = _hash2(xx, yy, seed + 1)
    # albedo: three hull paints + per-texel grain
    palette = np.array([[150, 156, 164], [60, 70, 88], [176, 96, 40], [205, 205, 198]], dtype=np.int32)
    base = palette[(hp % 4).astype(np.int64)]
    grain = ((ht >> 8) % 25).astype(np.int32)[..., None] - 12
    albedo = np.clip(base + grain, 0, 255).astype(np.uint8)
    # emission: seams on 1/16 of the plates
    glow_plate = ((hp >> 4) % 16) == 0
    seam = (np.abs(lx - plate // 2) <= max(plate // 16, 1)) | (np.abs(ly - plate // 2) <= max(plate // 16, 1))
    emission = np.zeros((size, size, 3), dtype=np.uint8)
    gm = glow_plate & seam
    emission[gm] = (40, 200, 255)
    # metallic / roughness: per plate, grey
    metallic_v = np.where(((hp >> 9) % 3) == 0, 230, 25).astype(np.uint8)
    rough_v = (40 + ((hp >> 13) % 160)).astype(np.uint8)
    metallic = np.repeat(metallic_v[..., None], 3, axis=2)
    roughness = np.repeat(rough_v[..., None], 3, axis=2)
    # normal map: flat (128,128,255) with bevelled plate borders
    bevel = max(plate // 10, 1)
    nx = np.where(lx < bevel, -70, np.where(lx >= plate - bevel, 70, 0))
    ny = np.where(ly < bevel, 70, np.where(ly >= plate - bevel, -70, 0))
    normal = np.stack([np.clip(128 + nx, 0, 255), np.clip(128 + ny, 0, 255), np.full_like(nx, 235)], axis=2).astype(np.uint8)
    maps = [Texture(albedo), Texture(emission), Texture(metallic), Texture(roughness), Texture(normal)]
    _MAP_CACHE[key] = maps
    return maps


# ----------------------------------------------------------------------------- building blocks
def head_camera(width, height, spp, depth, lens_radius=0.0, **kw) -> Camera:
    """tracing.rs:357-373 with the resolution / sample count of the configuration."""
    return Camera(eyepoint=(0.0, 2.0, 5.5), view_dir=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), focal_length=0.6,
                  focus_dist=5.0, lens_radius=lens_radius, screen_width=width, screen_height=height,
                  aa_sample_count=spp, path_depth=depth, path_samples=1, max_trace_dist=100.0, gamma=2.0, **kw)


def cornell_walls() -> list:
    """Six Planes (grey / red / green Lambertian) and an emissive quad of two Triangles, the light of
    tracing.rs:527-538 moved under the ceiling."""
    white = Lambertian(albedo=(0.73, 0.73, 0.73))
    red = Lambertian(albedo=(0.65, 0.05, 0.05))
    green = Lambertian(albedo=(0.12, 0.45, 0.15))
    light = Lambertian(albedo=(0.0, 0.6, 0.0), emission=(7.0, 7.0, 7.0))
    return [
        Plane(point=(0.0, 0.0, 0.0), normal=(0.0, 1.0, 0.0), material=white),     # floor
        Plane(point=(0.0, 4.5, 0.0), normal=(0.0, -1.0, 0.0), material=white),    # ceiling
        Plane(point=(0.0, 0.0, -2.5), normal=(0.0, 0.0, 1.0), material=white),    # back
        Plane(point=(-2.75, 0.0, 0.0), normal=(1.0, 0.0, 0.0), material=red),     # left
        Plane(point=(2.75, 0.0, 0.0), normal=(-1.0, 0.0, 0.0), material=green),   # right
        Plane(point=(0.0, 0.0, 6.5), normal=(0.0, 0.0, -1.0), material=white),    # behind the camera
        Triangle(a=(-1.0, 4.49, -0.5), b=(1.0, 4.49, -0.5), c=(1.0, 4.49, 1.5), material=light),
        Triangle(a=(-1.0, 4.49, -0.5), b=(-1.0, 4.49, 1.5), c=(1.0, 4.49, 1.5), material=light),
    ]


def c1_cornell(width=512, height=512, spp=64, depth=8) -> Scene:
    return Scene(camera=head_camera(width, height, spp, depth), objects=cornell_walls())


def c2_teapots(width=1024, height=1024, spp=256, depth=10) -> Scene:
    objs = cornell_walls()
    up = cg.from_angle_x(-90.0)  # teapot.obj is z-up
    objs.append(StaticMesh.load_from_file(
        obj_path("teapot"), material=Lambertian(albedo=(0.7, 0.6, 0.3)),
        transform=cg.chain(cg.from_translation((-1.2, 0.75, 0.8)), cg.from_angle_y(30.0), up, cg.from_scale(1.5))))
    objs.append(StaticMesh.load_from_file(
        obj_path("teapot"), material=Metal(albedo=(0.9, 0.9, 0.95), roughness=0.05),
        transform=cg.chain(cg.from_translation((1.2, 0.9, -0.4)), cg.from_angle_y(-40.0), up, cg.from_scale(1.8))))
    return Scene(camera=head_camera(width, height, spp, depth), objects=objs)


def c3_materials(width=1024, height=1024, spp=1024, depth=10) -> Scene:
    objs = cornell_walls()
    objs += [
        Sphere(center=(-1.7, 0.7, 0.2), radius=0.7, material=Metal(albedo=(0.85, 0.75, 0.4), roughness=0.15)),
        Sphere(center=(0.0, 0.6, 1.6), radius=0.6, material=Dielectric(idx_of_refraction=1.5)),
        Sphere(center=(1.7, 0.5, 1.0), radius=0.5, material=Dielectric(idx_of_refraction=2.5)),
        Sphere(center=(1.4, 2.6, -1.2), radius=0.35, material=Lambertian(albedo=(0.3, 0.3, 0.3), emission=(0.0, 1.0, 1.0))),
        Sphere(center=(0.2, 0.8, -1.0), radius=0.8, material=Lambertian(albedo=(0.2, 0.3, 0.7))),
    ]
    # "subsurface": a Dielectric sphere plus a ConvexVolume in the same place (README.md:68-69, Q10)
    sub = Sphere(center=(-0.9, 2.3, 0.6), radius=0.55, material=Dielectric(idx_of_refraction=1.5))
    objs.append(sub)
    objs.append(ConvexVolume(boundary=Sphere(center=sub.center, radius=sub.radius, material=sub.material),
                             phase_function=Isotropic(albedo=(0.9, 0.5, 0.4)), density=4.0))
    return Scene(camera=head_camera(width, height, spp, depth, lens_radius=0.04), objects=objs)


def c4_drone(width=1920, height=1080, spp=1024, depth=10, map_size=2048) -> Scene:
    """`run()` verbatim (tracing.rs:374-540); only resolution / spp differ and the drone maps are synthetic
    (map_size=0: no maps at all, which is what the reference itself gets from its checkout)."""
    maps = drone_maps(map_size) if map_size else [None] * 5
    drone = StaticMesh(load_obj(obj_path("drone")), maps, None,
                       cg.chain(cg.from_translation((0.0, 1.3, 1.7)), cg.from_angle_y(-60.0), cg.from_angle_x(180.0),
                                cg.from_scale(0.0030)))
    cube = StaticMesh.load_from_file(obj_path("cube"), tex_path("green.png"), None, None, None, tex_path("normal_test.jpg"),
                                     None, cg.chain(cg.from_translation((-1.7, 0.5, 2.7)), cg.from_angle_y(45.0),
                                                    cg.from_scale(0.4)))
    ball = StaticMesh.load_from_file(obj_path("sphere"), tex_path("magenta.jpg"), None, None, None,
                                     tex_path("normal_test.png"), None,
                                     cg.chain(cg.from_translation((1.7, 0.5, 2.7)), cg.from_angle_y(45.0), cg.from_scale(0.6)))
    objs = [drone, cube, ball]
    # demo of the parameterized material: 3 rows (metallic 0, .5, 1) x 5 columns (roughness 0..1)
    for y, metallic in ((3.3, 0.0), (4.4, 0.5), (5.5, 1.0)):
        for x, rough in ((-2.6, 0.0), (-1.3, 0.25), (0.0, 0.5), (1.3, 0.75), (2.6, 1.0)):
            objs.append(Sphere(center=(x, y, 0.0), radius=0.5,
                               material=ParameterizedMaterial(albedo=(0.01, 0.02, 0.5), emission=(0.0, 0.0, 0.0),
                                                              roughness=rough, metallic=metallic)))
    objs += [
        Sphere(center=(-2.3, 2.0, 2.0), radius=0.4, material=Dielectric(idx_of_refraction=2.5)),
        Sphere(center=(2.3, 2.0, 2.0), radius=0.4, material=Lambertian(albedo=(0.3, 0.3, 0.3), emission=(0.0, 1.0, 1.0))),
        ConvexVolume(boundary=Sphere(center=(-3.0, 1.0, 1.0), radius=1.0, material=Dielectric(idx_of_refraction=1.5)),
                     phase_function=Isotropic(albedo=(1.0, 1.0, 1.0), emission=(0.0, 0.0, 0.0)), density=0.6),
        ConvexVolume(boundary=Sphere(center=(3.0, 1.0, 1.0), radius=1.0, material=Dielectric(idx_of_refraction=1.5)),
                     phase_function=Isotropic(albedo=(0.0, 0.0, 0.0), emission=(0.0, 0.0, 0.0)), density=0.8),
        Plane(point=(0.0, 0.0, 0.0), normal=(0.0, 1.0, 0.0),
              material=ParameterizedMaterial(albedo=(0.33, 0.33, 0.33), emission=(0.0, 0.0, 0.0), metallic=0.3, roughness=0.7)),
    ]
    light = Lambertian(albedo=(0.0, 0.6, 0.0), emission=(7.0, 7.0, 7.0))
    objs += [
        Triangle(a=(-2.5, 7.5, -0.5), b=(2.5, 7.5, -0.5), c=(2.5, 7.5, 3.5), material=light),
        Triangle(a=(-2.5, 7.5, -0.5), b=(-2.5, 7.5, 3.5), c=(2.5, 7.5, 3.5), material=Lambertian(
            albedo=(0.0, 0.6, 0.0), emission=(7.0, 7.0, 7.0))),
    ]
    return Scene(camera=head_camera(width, height, spp, depth), objects=objs)


def c5_instances(width=3840, height=2160, spp=4096, depth=10, grid=16, map_size=2048, seed=5) -> Scene:
    """grid x grid transformed instances alternating drone (textured) / teapot (Lambertian, Metal): shared BLASes,
    one TLAS; floor Plane, light quad, defocus."""
    maps = drone_maps(map_size)
    drone_md, teapot_md = load_obj(obj_path("drone")), load_obj(obj_path("teapot"))
    lam = [Lambertian(albedo=a) for a in ((0.7, 0.3, 0.25), (0.25, 0.6, 0.3), (0.3, 0.35, 0.75), (0.8, 0.75, 0.3))]
    met = [Metal(albedo=(0.9, 0.9, 0.95), roughness=r) for r in (0.0, 0.1, 0.3)]
    up = cg.from_angle_x(-90.0)
    objs = []
    ii, jj = np.meshgrid(np.arange(grid, dtype=np.uint32), np.arange(grid, dtype=np.uint32), indexing="ij")
    h = _hash2(ii, jj, seed)
    span_x, span_z = 11.0, 16.0
    for i in range(grid):
        for j in range(grid):
            hv = int(h[i, j])
            x = (i + 0.5) / grid * span_x - span_x / 2 + ((hv & 255) / 255.0 - 0.5) * 0.2
            z = 2.5 - (j + 0.5) / grid * span_z + (((hv >> 8) & 255) / 255.0 - 0.5) * 0.2
            yaw = ((hv >> 16) & 255) / 255.0 * 360.0
            if (i + j) % 2 == 0:
                y = 0.55 + ((hv >> 24) & 15) / 15.0 * 1.6
                t = cg.chain(cg.from_translation((x, y, z)), cg.from_angle_y(yaw), cg.from_angle_x(180.0), cg.from_scale(0.0006))
                objs.append(StaticMesh(drone_md, maps, None, t))
            else:
                mat = lam[(hv >> 4) % 4] if (hv >> 2) % 2 == 0 else met[(hv >> 6) % 3]
                t = cg.chain(cg.from_translation((x, 0.2, z)), cg.from_angle_y(yaw), up, cg.from_scale(0.4))
                objs.append(StaticMesh(teapot_md, [None] * 5, mat, t))
    objs.append(Plane(point=(0.0, 0.0, 0.0), normal=(0.0, 1.0, 0.0),
                      material=ParameterizedMaterial(albedo=(0.33, 0.33, 0.33), metallic=0.3, roughness=0.7)))
    light = Lambertian(albedo=(0.0, 0.6, 0.0), emission=(7.0, 7.0, 7.0))
    objs += [Triangle(a=(-7.0, 7.5, -15.0), b=(7.0, 7.5, -15.0), c=(7.0, 7.5, 4.0), material=light),
             Triangle(a=(-7.0, 7.5, -15.0), b=(-7.0, 7.5, 4.0), c=(7.0, 7.5, 4.0), material=light)]
    return Scene(camera=head_camera(width, height, spp, depth, lens_radius=0.03), objects=objs)


CONFIGS = {
    "c1": c1_cornell,
    "c2": c2_teapots,
    "c3": c3_materials,
    "c4": c4_drone,
    "c5": c5_instances,
}

DESCRIPTIONS = {
    "c1": "Cornell box (Lambertian planes + emissive quad) 512x512, 64 spp, depth 8",
    "c2": "Cornell + 2x teapot.obj (Lambertian, Metal) 1024x1024, 256 spp, depth 10",
    "c3": "Cornell + metal/glass/emissive spheres + subsurface (Dielectric + Isotropic volume), defocus, 1024x1024, 1024 spp",
    "c4": "reference run() scene: textured drone + cube + sphere.obj, 15 parameterized spheres, 2 volumes, 1920x1080, 1024 spp, depth 10",
    "c5": "256 transformed drone/teapot instances, 3840x2160, 4096 spp, depth 10",
}


def make_scene(name: str, **overrides) -> Scene:
    return CONFIGS[name](**overrides)
