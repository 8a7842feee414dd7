"""Build recipe for librt_b200.so (the C-ABI library: CUDA kernels for sm_100a + host lowering).

Everything is compiled in-tree with nvcc; the .so is git-ignored but travels to the GPU box.
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  -fmad=false                               no implicit FMA contraction: the intersection math must be
                                            the reference's sequence of single IEEE operations
                                            (explicit __fmaf_rn is used where exactness is not needed)
  -Xcompiler -ffp-contract=off              same for the host lowering (precomputed records)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")  # RT_B200_LIB: A/B builds
SOURCES = ["rt_kernels.cu", "rt_api.cu", "rt_lower.cpp", "rt_png.cpp", "rt_jpeg.cpp"]
HEADERS = ["rt_kernels.h", "rt_lower.h", "rt_types.h", "rt_config.cuh", "rt_device_math.cuh", "rt_traverse.cuh", "rt_materials.cuh", "rt_shade.cuh", os.path.join("..", "..", "include", "rt_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> list[str]:
    # the image's $CXX wrapper lacks some spec files; the system g++ is the safe host compiler
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    if os.environ.get("RT_B200_LIB"):
        return False  # an explicitly named A/B build is used as it is (it was built with its own -D switches)
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: str | None = None) -> str:
    out = out or LIB
    if not force and not needs_build() and out == LIB:
        return LIB
    tmp = f"{out}.tmp{os.getpid()}"  # written aside and renamed: a process that has the old library mapped keeps it
    cmd = [
        _nvcc(), *_host_cxx(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-O3", "-std=c++17",
        "-fmad=false",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O3",
        *[f"-D{d}" for d in defines],
        "-shared", "-o", tmp,
        *[os.path.join(CSRC, f) for f in SOURCES],
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, out)
    if verbose:
        print(r.stderr, file=sys.stderr)
    return out


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force=True, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
