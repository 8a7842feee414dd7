"""Development tools stay buildable: tools/bvh_lab.cpp replays k_trace's walk on the CPU for the binary tree the library
builds and for 4- and 8-wide collapses of it; all three must return the same closest hits."""
import gzip
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cs397raytracingsp22_b200", "csrc")


def test_bvh_lab_trees_agree(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no host C++ compiler on this box")
    exe = str(tmp_path / "bvh_lab")
    cuda_inc = "/usr/local/cuda/include"
    r = subprocess.run([cxx, "-std=c++17", "-O2", "-I", CSRC, "-I", cuda_inc, os.path.join(ROOT, "tools", "bvh_lab.cpp"),
                        os.path.join(CSRC, "rt_lower.cpp"), os.path.join(CSRC, "rt_png.cpp"), os.path.join(CSRC, "rt_jpeg.cpp"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    obj = tmp_path / "teapot.obj"
    with open(os.path.join(ROOT, "assets", "obj", "teapot.obj.gz"), "rb") as f:
        obj.write_bytes(gzip.decompress(f.read()))
    out = subprocess.run([exe, str(obj), "20000"], capture_output=True, text=True, check=True).stdout
    assert "240 triangles, 240 reachable" in out
    mism = [int(m) for m in re.findall(r"(\d+) closest-hit mismatches", out)]
    assert mism == [0, 0, 0], out
    # a wider tree needs fewer dependent fetches and no fewer box tests
    rows = {m[0]: (float(m[1]), float(m[2])) for m in re.findall(r"^\s+(binary|4-wide|8-wide)\s+([\d.]+)\s+([\d.]+)", out, flags=re.M)}
    assert rows["4-wide"][0] < rows["binary"][0] and rows["8-wide"][0] < rows["4-wide"][0]
