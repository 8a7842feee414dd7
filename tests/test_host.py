"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the asset readers
(tobj / image stand-ins), the cgmath helpers, shard planning, and loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cs397raytracingsp22_b200 as rt
from cs397raytracingsp22_b200 import _ffi, cgmath as cg, distributed as D, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(rtlib):
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    for name in sorted(declared):
        assert hasattr(rtlib, name), f"librt_b200.so does not export {name}"
    assert declared == set(_ffi.SIGNATURES), "ctypes table and header disagree"
    assert rtlib.rt_abi_version() == _ffi.RT_B200_ABI_VERSION == 3


def test_struct_layouts_match_the_header():
    # sizes the C side static-asserts implicitly through use; a drift here corrupts every call
    assert C.sizeof(_ffi.rt_material_desc) == 40
    assert C.sizeof(_ffi.rt_camera) == 84
    assert C.sizeof(_ffi.rt_render_opts) == 96
    assert C.sizeof(_ffi.rt_stats) == 16 * 8 + 4 * 8 + 3 * 8


def test_no_gpu_is_a_loud_error_not_a_fallback(rtlib):
    """The product path must fail when CUDA is unavailable (this container has no GPU)."""
    if rtlib.rt_device_count() > 0:
        pytest.skip("a GPU is present")
    sc = rt.Scene(camera=rt.Camera(), objects=[rt.Sphere((0, 0, -3), 1.0, rt.Lambertian())])
    with pytest.raises(_ffi.RtError) as e:
        sc.render_to_image()
    assert e.value.code == _ffi.RT_ERR_CUDA
    b = _ffi.GpuBackend()
    with pytest.raises(_ffi.RtError) as e:
        b.render(rt.Camera().to_c())
    assert e.value.code == _ffi.RT_ERR_NOT_COMMITTED


def test_bad_arguments_return_codes_not_crashes(rtlib):
    b = _ffi.GpuBackend()
    with pytest.raises(_ffi.RtError):
        b.add_sphere((0, 0, 0), 1.0, 5)                    # material id does not exist
    m = b.add_material(_ffi.RT_MAT_LAMBERTIAN)
    with pytest.raises(_ffi.RtError):
        b.add_instance(3, np.eye(4, dtype=np.float32).reshape(-1), np.eye(4, dtype=np.float32).reshape(-1), m, [-1] * 5)
    with pytest.raises(_ffi.RtError):
        b.add_mesh(np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 2)), np.array([[0, 1, 7]]))   # index out of range
    proj = np.eye(4, dtype=np.float32)
    proj[3, 2] = 0.5                                       # projective transform: refused, not approximated
    mesh = b.add_mesh(np.eye(3), np.eye(3), np.zeros((3, 2)), np.array([[0, 1, 2]]))
    with pytest.raises(_ffi.RtError) as e:
        b.add_instance(mesh, cg.colmajor(proj), cg.colmajor(np.eye(4)), m, [-1] * 5)
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED
    assert rtlib.rt_scene_create(None) == _ffi.RT_ERR_INVALID
    assert b"out is NULL" in rtlib.rt_last_error()


OBJ_FACTS = {  # SURVEY.md §4, measured on the reference's files
    "cube": (24, 12), "teapot": (293, 240), "drone": (1606, 1736), "sphere": (16422, 32512),
}


@pytest.mark.parametrize("name", sorted(OBJ_FACTS))
def test_obj_loader_matches_tobj_counts(name):
    m = rt.load_obj(scenes.obj_path(name))
    assert (m.pos.shape[0], m.ntris) == OBJ_FACTS[name]
    assert m.idx.max() == m.pos.shape[0] - 1
    # first-seen order: vertex ids appear in increasing order of first use
    first_use = np.full(m.pos.shape[0], -1, np.int64)
    flat = m.idx.reshape(-1)
    for i, v in enumerate(flat):
        if first_use[v] < 0:
            first_use[v] = i
    assert (np.diff(first_use) > 0).all()


def test_obj_parser_semantics():
    text = b"""
# quad + pentagon, shared corners, negative indices, second group ignored
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 1.5 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
f 1/1/1 2/2/1 3/3/1 4/4/1
f -5/1/1 -4/2/1 -3/3/1 -1/3/1 -2/4/1
g other
f 1/1/1 2/2/1 3/3/1
"""
    pos, nrm, uv, idx, has_n, has_t = _ffi.parse_obj(text)
    assert has_n and has_t
    # fan triangulation (0,i,i+1): quad -> 2, pentagon -> 3; the second group starts a new model
    assert idx.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2], [0, 2, 4], [0, 4, 3]]
    assert pos.shape == (5, 3) and np.allclose(pos[4], [0.5, 1.5, 0]) and np.allclose(uv[4], [1, 1])
    with pytest.raises(_ffi.RtError):
        _ffi.parse_obj(b"v 0 0 0\nf 1 2 3\n")
    pos, nrm, uv, idx, has_n, has_t = _ffi.parse_obj(b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    assert not has_n and not has_t and idx.tolist() == [[0, 1, 2]]


def test_tga_roundtrip_and_variants():
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(13, 17, 3)).astype(np.uint8)
    data = _ffi.tga_encode(img)
    assert np.array_equal(_ffi.tga_decode(data), img)
    # bottom-up origin and RLE packets, hand-built
    w, h = 4, 2
    hdr = bytes([0, 0, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, w, 0, h, 0, 24, 0])
    rle = bytes([0x83, 10, 20, 30]) + bytes([0x03, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13])   # run of 4, raw 4 (BGR)
    out = _ffi.tga_decode(hdr + rle)
    assert out.shape == (2, 4, 3)
    assert out[1].tolist() == [[30, 20, 10]] * 4                      # first file row is the bottom row
    assert out[0].tolist() == [[3, 2, 1], [6, 5, 4], [9, 8, 7], [13, 12, 11]]
    grey = bytes([0, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 1, 0, 8, 0x20]) + bytes([7, 200])
    assert _ffi.tga_decode(grey).tolist() == [[[7, 7, 7], [200, 200, 200]]]
    with pytest.raises(_ffi.RtError):
        _ffi.tga_decode(b"\x00" * 10)
    assert rt.Texture.load_from_file("/nonexistent/Drone_Albedo.tga") is None     # texture.rs:22-24: silently None


def test_png_reader_and_writer_against_pil(tmp_path):
    import io
    from PIL import Image
    for name in ("green.png", "white.png", "normal_test.png"):          # palette and RGB files of the reference
        data = open(scenes.tex_path(name), "rb").read()
        want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.uint8)
        assert np.array_equal(_ffi.png_decode(data), want), name
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, size=(37, 91, 3)).astype(np.uint8)
    for mode, kw in (("RGB", {}), ("RGBA", {}), ("L", {}), ("P", {}), ("1", {}), ("I;16", {}), ("RGB", dict(compress_level=0))):
        im = Image.fromarray(img).convert(mode) if mode != "I;16" else Image.fromarray((img[..., 0].astype(np.uint16) * 257))
        buf = io.BytesIO()
        im.save(buf, format="PNG", **kw)
        want = np.asarray(im.convert("RGB") if mode != "I;16" else Image.fromarray(img[..., 0]).convert("RGB"), dtype=np.uint8)
        assert np.array_equal(_ffi.png_decode(buf.getvalue()), want), mode
    enc = _ffi.png_encode(img)                                           # our writer, PIL as the reader
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(enc)).convert("RGB")), img)
    assert np.array_equal(_ffi.png_decode(enc), img)
    big = rng.randint(0, 256, size=(300, 301, 3)).astype(np.uint8)       # more than one 64 KiB stored block
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(_ffi.png_encode(big))).convert("RGB")), big)
    with pytest.raises(_ffi.RtError):
        _ffi.png_decode(b"\x89PNG\r\n\x1a\n" + b"\x00" * 40)


def test_reference_textures_decode():
    for name, size in (("green.png", (225, 225)), ("magenta.jpg", (615, 615)), ("normal_test.png", (512, 512))):
        t = rt.Texture.load_from_file(scenes.tex_path(name))
        assert t is not None and (t.width, t.height) == size and t.rgb8.shape[2] == 3   # palette PNG expanded to RGB


def test_cgmath_helpers():
    t = cg.chain(cg.from_translation((0.0, 1.3, 1.7)), cg.from_angle_y(-60.0), cg.from_angle_x(180.0), cg.from_scale(0.003))
    inv = cg.inverse_transform(t)
    assert np.allclose(cg.mul(t, inv), np.eye(4), atol=1e-4)
    assert inv[3].tolist() == [0, 0, 0, 1]
    ry = cg.from_angle_y(90.0)
    assert np.allclose(ry @ np.array([1, 0, 0, 0], np.float32), [0, 0, -1, 0], atol=1e-6)   # right-handed, like cgmath
    rx = cg.from_angle_x(90.0)
    assert np.allclose(rx @ np.array([0, 1, 0, 0], np.float32), [0, 0, 1, 0], atol=1e-6)
    assert cg.inverse_transform(np.zeros((4, 4), np.float32)) is None
    assert cg.colmajor(cg.from_translation((1, 2, 3)))[12:15].tolist() == [1, 2, 3]


def test_shard_plans_partition_the_frame():
    for spp, world in ((1024, 8), (64, 3), (5, 8), (4096, 4)):
        ranges = [D.sample_range(r, world, 0, spp) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == spp
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    for (w, h, ts, world) in ((1920, 1080, 64, 8), (100, 37, 16, 3), (64, 64, 64, 2)):
        seen = np.zeros((h, w), np.int32)
        for r in range(world):
            for tx, ty in D.tiles_of_rank(r, world, w, h, ts):
                seen[ty * ts:(ty + 1) * ts, tx * ts:(tx + 1) * ts] += 1
        assert (seen == 1).all()


def test_drone_maps_are_deterministic():
    a = scenes.drone_maps(64, seed=397)
    scenes._MAP_CACHE.clear()
    b = scenes.drone_maps(64, seed=397)
    assert all(np.array_equal(x.rgb8, y.rgb8) for x, y in zip(a, b))
    assert [int(x.rgb8.astype(np.int64).sum()) for x in a] == [int(x.rgb8.astype(np.int64).sum()) for x in b]
    assert a[1].rgb8.max() > 0 and (a[1].rgb8 == 0).mean() > 0.5      # emission: mostly black, some seams


def test_lowering_on_the_host(small_scenes):
    """rt_scene_lower is the CPU half of rt_commit: reachability mask, binned/sweep SAH BLASes, TLAS, tables."""
    b = _ffi.GpuBackend()
    small_scenes("c4").lower(b)
    info = b.lower_info()
    assert info["objects"] == 25 and info["unbounded"] == 1                 # the floor Plane is the only unbounded object
    assert info["tris"] == (1736 - 51) + 12 + 32512                        # Q3: 51 drone triangles are unreachable
    assert info["tris"] / 4 <= info["nodes"] <= 2 * info["tris"] + 200
    assert 3 <= info["tlas_depth"] <= 12 and 10 <= info["max_blas_depth"] <= 40
    assert info["tlas_depth"] + 1 + info["max_blas_depth"] <= 62            # fits the traversal stack
    assert 0 < info["guard_boxes"] < 200 and 0 < info["guarded_tris"] < 600  # drone.obj's nearly-flat disc caps
    b5 = _ffi.GpuBackend()
    small_scenes("c5").lower(b5)
    i5 = b5.lower_info()
    assert i5["objects"] == 6 * 6 + 3 and i5["tris"] == (1736 - 51) + 240   # shared BLASes: one drone, one teapot


def test_degenerate_mesh_still_builds_a_bounded_bvh():
    """All centroids coincide: SAH cannot split, the builder falls back to median splits (depth = log2 n)."""
    n = 512
    pos = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.5]], np.float32), (n, 1))
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3 * n, 1))
    uv = np.tile(np.array([[0, 0], [1, 0], [0, 1]], np.float32), (n, 1))
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    b = _ffi.GpuBackend()
    m = b.add_mesh(pos, nrm, uv, idx)
    b.add_instance(m, np.eye(4, dtype=np.float32).reshape(-1), np.eye(4, dtype=np.float32).reshape(-1),
                   b.add_material(_ffi.RT_MAT_LAMBERTIAN), [-1] * 5)
    info = b.lower_info()
    assert info["tris"] == n and info["max_blas_depth"] <= 10


def test_hostile_png_header_is_refused_before_any_allocation(rtlib):
    """A 100-byte file whose IHDR promises 65536 x 65536 RGBA16 (a 32 GB image): the reader must return RT_ERR_IO, not
    reserve the memory, not throw across the ABI."""
    import struct
    import zlib

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    ihdr = struct.pack(">IIBBBBB", 65536, 65536, 16, 6, 0, 0, 0)
    bomb = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    with pytest.raises(_ffi.RtError) as e:
        _ffi.png_decode(bomb)
    assert e.value.code == _ffi.RT_ERR_IO
    # a stream that inflates to far more than the header's image is cut off, too (zip bomb behind an honest header)
    ihdr = struct.pack(">IIBBBBB", 4, 4, 8, 2, 0, 0, 0)
    big = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(b"\0" * (64 << 20), 9)) + chunk(b"IEND", b"")
    with pytest.raises(_ffi.RtError) as e:
        _ffi.png_decode(big)
    assert e.value.code == _ffi.RT_ERR_IO


def test_caller_supplied_inverse_may_be_off_by_rounding():
    """cgmath's general Matrix4::invert (geometry.rs:168) leaves w.w = 1 +- an ulp on scaled / rotated transforms; the
    Rust shim passes that matrix as it is (integration/lower.rs)."""
    b = _ffi.GpuBackend()
    mat = b.add_material(_ffi.RT_MAT_LAMBERTIAN)
    mesh = b.add_mesh(np.eye(3), np.eye(3), np.zeros((3, 2)), np.array([[0, 1, 2]]))
    x = cg.chain(cg.from_translation((0.0, 1.3, 1.7)), cg.from_angle_y(-60.0), cg.from_scale(0.003))
    inv = np.linalg.inv(np.asarray(x, np.float64)).astype(np.float32)
    inv[3, 3] = np.nextafter(np.float32(1.0), np.float32(2.0))          # 1 + ulp
    inv[3, 0] = 1e-8
    b.add_instance(mesh, cg.colmajor(x), cg.colmajor(inv), mat, [-1] * 5)   # accepted, row set to (0, 0, 0, 1)
    inv[3, 3] = 1.001                                                      # a projective matrix is still refused
    with pytest.raises(_ffi.RtError) as e:
        b.add_instance(mesh, cg.colmajor(x), cg.colmajor(inv), mat, [-1] * 5)
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED


def test_objects_that_cannot_be_bounded_are_refused_loudly():
    b = _ffi.GpuBackend()
    mat = b.add_material(_ffi.RT_MAT_LAMBERTIAN)
    b.add_sphere((0.0, float("nan"), 0.0), 1.0, mat)
    with pytest.raises(_ffi.RtError) as e:
        b.lower_info()
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED and "non-finite" in str(e.value)


def test_render_options_are_plain_fields_not_environment_variables():
    """Round-1 development knobs (RT_LANES, RT_RAY_SORT, RT_SAMPLE_MAJOR, ...) are gone: nothing in the library reads
    the environment; what remains tunable is a field of rt_render_opts or a -D switch."""
    for f in ("rt_api.cu", "rt_lower.cpp", "rt_kernels.cu"):
        assert "getenv" not in open(os.path.join(ROOT, "cs397raytracingsp22_b200", "csrc", f)).read(), f
    names = [n for n, _ in _ffi.rt_render_opts._fields_]
    assert names[-5:] == ["engine", "ray_sort", "work_order", "blocks_per_sm", "reserved"]
    o = D.shard_opts(1, 4, 7, "tiles", tile=16, engine=_ffi.RT_ENGINE_MEGAKERNEL, ray_sort=_ffi.RT_RAYSORT_OFF)
    assert (o.shard_mode, o.shard_rank, o.shard_count, o.tile_size, o.engine, o.ray_sort) == (2, 1, 4, 16, 2, 1)
