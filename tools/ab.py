"""A/B helper (development only): run bench.py once per (library variant, env knobs) and print one line each.
Variants are built with `python cs397raytracingsp22_b200/build.py -DNAME=VALUE --out=build/rt_x.so`.
usage (from the repo root): python tools/ab.py [lib.so ...] [-- KEY=VAL,KEY=VAL ...]"""
import glob, json, os, subprocess, sys
args = sys.argv[1:]
envsets = [""]
if "--" in args:
    k = args.index("--")
    envsets = args[k + 1:] or [""]
    args = args[:k]
libs = args or sorted(glob.glob("build/rt_*.so"))
for lib in libs:
    for es in envsets:
        env = dict(os.environ, RT_B200_LIB=os.path.abspath(lib))
        for kv in filter(None, es.split(",")):
            a, b = kv.split("=")
            env[a] = b
        r = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "2", "--spp", "256", "--no-cpu-baseline", "--no-e2e"]
                           + (["--workload", env["WORKLOAD"]] if "WORKLOAD" in env else []) + ["--wavefront", env.get("WAVEFRONT", "8388608")] + (["--depth", env["DEPTH"]] if "DEPTH" in env else []),
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            rf = d["roofline"]
            print(f"{os.path.basename(lib):20s} {es:28s} {d['value']:8.1f} Msamples/s {d['rays_per_sec_M']:8.1f} Mrays/s  extend {rf['ms_per_launch']*1e3:6.1f} us "
                  f"({rf['share_of_step']:.3f}) shade {rf['shade_share_of_step']:.3f} nodes/ray {rf['nodes_per_ray']:.2f} (tlas {rf.get('tlas_nodes_per_ray',0):.2f}) tris/ray {rf['tris_per_ray']:.2f} "
                  f"simt {rf.get('traversal_simt_efficiency', 0):.3f}", flush=True)
        except Exception as e:
            print(lib, es, "FAILED", e, r.stderr[-400:], flush=True)
