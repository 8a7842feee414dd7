"""CPU-only checks of the bench / packaging contract: the reference arm prints the agreed JSON line, and the product
package never touches the oracle (which is test infrastructure)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--width", "96", "--height", "54"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--width", "32", "--height", "18"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                       env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or execute anything under oracle/."""
    offenders = []
    for base in ("cs397raytracingsp22_b200", "include", "examples"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle_ffi|liboracle|orc_[a-z_]+\(|[\"'/]oracle/", text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"import oracle_ffi", bench)]
    assert len(uses) == 1 and "def cpu_render_sample" in bench[:uses[0]][-400:]   # only inside the CPU leg


def test_header_and_binding_agree_on_the_abi():
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    from cs397raytracingsp22_b200 import _ffi
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    assert declared == set(_ffi.SIGNATURES)
    assert len(declared) == 35


def test_graft_entry_build_runs_here():
    """The driver calls __graft_entry__.build() on a box without a GPU: nvcc cross-compiles the library for sm_100a, g++
    builds the oracle, and the ABI version of the result matches the header and the binding."""
    import __graft_entry__ as g
    g.build()
    from cs397raytracingsp22_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    assert int(re.search(r"#define RT_B200_ABI_VERSION (\d+)", hdr).group(1)) == _ffi.RT_B200_ABI_VERSION
    assert _ffi.load().rt_abi_version() == _ffi.RT_B200_ABI_VERSION
