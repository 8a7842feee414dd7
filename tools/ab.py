"""A/B helper (development only): run bench.py once per (library variant, bench arguments) and print one line each.
Variants are built with `python cs397raytracingsp22_b200/build.py -DNAME=VALUE --out=build/rt_x.so`.
usage (from the repo root): python tools/ab.py [lib.so ...] [-- "bench args" "bench args" ...]
e.g.  python tools/ab.py build/rt_a.so build/rt_b.so -- "--workload c1 --engine megakernel" "--workload c1 --engine wavefront" """
import glob, json, os, shlex, subprocess, sys
args = sys.argv[1:]
argsets = [""]
if "--" in args:
    k = args.index("--")
    argsets = args[k + 1:] or [""]
    args = args[:k]
libs = args or sorted(glob.glob("build/rt_*.so"))
BASE = ["--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-configs"]
for lib in libs:
    for es in argsets:
        env = dict(os.environ, RT_B200_LIB=os.path.abspath(lib))
        extra = shlex.split(es)
        if "--spp" not in extra and "--full" not in extra:
            extra += ["--spp", "256"]
        extra = [a for a in extra if a != "--full"]
        r = subprocess.run([sys.executable, "bench.py"] + BASE + extra, env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            rf = d["roofline"]
            print(f"{os.path.basename(lib):22s} {es:50s} {d['engine']:10s} {d['value']:8.1f} Msamples/s {d['rays_per_sec_M']:8.1f} Mrays/s  "
                  f"{d['ms_per_step']:8.2f} ms/step  {rf['kernel']} {rf['ms_per_launch']*1e3:8.1f} us ({rf['share_of_step']:.3f}) shade {rf['shade_share_of_step']:.3f} "
                  f"nodes/ray {rf['nodes_per_ray']:.2f} (tlas {rf.get('tlas_nodes_per_ray',0):.2f}) simt {rf.get('traversal_simt_efficiency', 0):.3f}", flush=True)
        except Exception as e:
            print(lib, es, "FAILED", e, r.stderr[-600:], flush=True)
