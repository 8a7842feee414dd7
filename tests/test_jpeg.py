"""The library's JPEG reader (csrc/rt_jpeg.cpp) against what libjpeg-turbo (through Pillow) decodes the same files to;
the expectations are committed (tests/golden/jpeg/expected.npz, made by tests/golden/make_jpeg_fixtures.py), so the
tests need no Pillow.  Decoders may differ by a few LSBs (IDCT rounding, chroma upsampling), never by more."""
import glob
import os

import numpy as np
import pytest

from cs397raytracingsp22_b200 import _ffi, scenes
from cs397raytracingsp22_b200.texture import Texture

HERE = os.path.dirname(os.path.abspath(__file__))
JPEG = os.path.join(HERE, "golden", "jpeg")
FILES = sorted(glob.glob(os.path.join(JPEG, "*.jpg")))


def _close(a, b):
    assert a.shape == b.shape and a.dtype == np.uint8
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert d.max() <= 4 and d.mean() <= 0.1 and (d > 1).mean() <= 0.01, (d.max(), d.mean(), (d > 1).mean())


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_decoder_matches_libjpeg(path):
    expected = np.load(os.path.join(JPEG, "expected.npz"))[os.path.basename(path)[:-4]]
    with open(path, "rb") as f:
        _close(_ffi.jpeg_decode(f.read()), expected)


def test_fixture_set_covers_the_branches():
    assert len(FILES) == 6
    data = {os.path.basename(f)[:-4]: open(f, "rb").read() for f in FILES}
    assert b"\xff\xc2" in data["rgb420_progressive_45x51"] and b"\xff\xc0" in data["rgb420_restart_67x35"]
    assert b"\xff\xdd" in data["rgb420_restart_67x35"] and b"\xff\xd0" in data["rgb420_restart_67x35"]
    assert b"\xff\xdd" in data["rgb444_restart_rows_33x41"]


@pytest.mark.parametrize("name", ["magenta", "normal_test"])
def test_textures_of_the_run_scene(name):
    """magenta.jpg is progressive 4:2:0, normal_test.jpg baseline 4:2:0 (tracing.rs:395,405 load them)."""
    tex = Texture.load_from_file(scenes.tex_path(name + ".jpg"))
    assert tex is not None
    _close(tex.rgb8[::8, ::8], np.load(os.path.join(JPEG, "expected.npz"))[f"asset_{name}_stride8"])


def test_bad_input_is_an_error_not_a_crash():
    good = open(FILES[0], "rb").read()
    for bad in (b"", b"\xff\xd8", b"not a jpeg at all", good[:40], good[:len(good) // 2].replace(b"\xff\xc0", b"\xff\xc3")):
        with pytest.raises(_ffi.RtError):
            _ffi.jpeg_decode(bad)
    # a scan cut short still decodes (missing data reads as zero bits), like other decoders do
    img = _ffi.jpeg_decode(good[:len(good) - 40])
    assert img.shape == (29, 37, 3)
    assert Texture.load_from_file(os.path.join(JPEG, "does_not_exist.jpg")) is None


def test_mutated_files_never_crash_the_readers():
    """A light version of tools/fuzz_decoders.cpp (which runs under ASan/UBSan): a corrupted file either decodes to an
    image of the announced size or raises RtError."""
    rng = np.random.default_rng(2024)
    seeds = [open(f, "rb").read() for f in FILES] + [open(scenes.tex_path("green.png"), "rb").read()]
    for data in seeds:
        png = data[:4] == b"\x89PNG"
        for _ in range(150):
            d = bytearray(data)
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, len(d)))
                op = int(rng.integers(0, 3))
                if op == 0:
                    d[pos] = int(rng.integers(0, 256))
                elif op == 1:
                    d[pos] ^= 1 << int(rng.integers(0, 8))
                elif len(d) > 32:
                    del d[int(rng.integers(16, len(d))):]
            try:
                img = (_ffi.png_decode if png else _ffi.jpeg_decode)(bytes(d))
            except _ffi.RtError:
                continue
            assert img.ndim == 3 and img.shape[2] == 3 and img.size > 0
