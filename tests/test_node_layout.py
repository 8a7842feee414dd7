"""The packed child pairs of the lowered BVH (two centres with their links, six half extents as bf16,
three 128-bit fetches per visit) against the plain (min, max) pairs the same lowering produces with -DRT_NODE_CH=0:
every packed box must CONTAIN the box it came from (the slab test only culls, so a box may grow but never shrink),
stay tight (bf16 rounds a half extent up by < 2^-7), and carry the same links.  CPU only: tools/node_dump.cpp is the
library's own rt_lower.cpp driven from a small main()."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cs397raytracingsp22_b200", "csrc")


def _dump(tmp_path, name, defines, obj):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no host C++ compiler on this box")
    exe = str(tmp_path / name)
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", *defines, "-I", CSRC, "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "tools", "node_dump.cpp"), os.path.join(CSRC, "rt_lower.cpp"),
                        os.path.join(CSRC, "rt_png.cpp"), os.path.join(CSRC, "rt_jpeg.cpp"), "-o", exe, "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, str(obj)], capture_output=True, check=True).stdout
    return np.frombuffer(out, dtype=np.float32).reshape(-1, 4, 4)   # child pairs x quads x words


def _bf16(x):
    return (x.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("mesh", ["teapot", "drone"])
def test_packed_pairs_contain_the_min_max_boxes(tmp_path, mesh):
    obj = tmp_path / f"{mesh}.obj"
    with open(os.path.join(ROOT, "assets", "obj", f"{mesh}.obj.gz"), "rb") as f:
        obj.write_bytes(gzip.decompress(f.read()))
    plain = _dump(tmp_path, "dump_minmax", ["-DRT_NODE_CH=0"], obj)
    packed = _dump(tmp_path, "dump_packed", [], obj)           # the default: RT_NODE_CH=2
    assert plain.shape == packed.shape and plain.shape[0] > 50
    pu, qu = plain.view(np.uint32), packed.view(np.uint32)
    # links: (min, max) pairs keep them in word 3 of quads 0 and 2, packed pairs in word 3 of quads 0 and 1
    assert np.array_equal(pu[:, 0, 3], qu[:, 0, 3]) and np.array_equal(pu[:, 2, 3], qu[:, 1, 3])
    # the plain counts survive in quad 3 (never read by the kernels)
    assert np.array_equal(pu[:, 1, 3], qu[:, 3, 0]) and np.array_equal(pu[:, 3, 3], qu[:, 3, 1])
    w = qu[:, 2, :3]
    h = np.stack([_bf16(w[:, 0] & 0xFFFF), _bf16(w[:, 0] >> 16), _bf16(w[:, 1] & 0xFFFF), _bf16(w[:, 1] >> 16),
                  _bf16(w[:, 2] & 0xFFFF), _bf16(w[:, 2] >> 16)], axis=1).astype(np.float64)
    for side, (lo, hi, c, hh) in enumerate(((plain[:, 0, :3], plain[:, 1, :3], packed[:, 0, :3], h[:, :3]),
                                            (plain[:, 2, :3], plain[:, 3, :3], packed[:, 1, :3], h[:, 3:]))):
        valid = (lo <= hi).all(axis=1)
        assert valid.sum() > 50
        lo, hi, c, hh = lo[valid].astype(np.float64), hi[valid].astype(np.float64), c[valid].astype(np.float64), hh[valid]
        assert (c - hh <= lo).all() and (c + hh >= hi).all(), f"side {side}: a packed box does not contain its (min, max) box"
        ext = (hi - lo) / 2
        big = ext > 1e-6 * np.abs(c).max()
        infl = (hh[big] - ext[big]) / ext[big]
        assert infl.min() >= 0 and infl.max() < 2.0 ** -7 + 1e-6, f"side {side}: half extents inflated by up to {infl.max():.4f}"
        # slots that can never be hit (padding) are marked by a negative half extent, never by garbage
        assert np.isfinite(packed[:, side, :3]).all()
