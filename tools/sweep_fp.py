"""GPU throughput sweep of the find_path search kernel instantiations (HBN_LANE_CFG) on the C4 workload.
usage: python tools/sweep_fp.py [queries]"""
import json, os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "400000"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# lane:N = HBN_LANE_CFG (hbn_capi.cu laneKernel): 0 shipped; 60, 38, 32, 31, 30, 1 the steps back to round 1's kernel;
# 34 = 95 shared heap entries at 11 warps per SM
variants = (("lane", None), ("lane:60", None), ("lane:38", None), ("lane:32", None), ("lane:31", None), ("lane:30", None),
            ("lane:1", None), ("lane:34", None))
os.environ.setdefault("HBN_QUERY_CACHE", "/tmp/hbn_queries")  # bench.py generates the pairs once
for g, bps in variants:
    env = dict(os.environ)
    if ":" in g:
        env["HBN_LANE_CFG"] = g.split(":")[1]
    if bps:
        env["HBN_FP_BLOCKS_PER_SM"] = bps
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "3",
                          "--queries", n, "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"G={g} blocks/SM={bps}: value={j['value']:.0f} q/s e2e={j['e2e']['value']:.0f} "
              f"path_ms={j['roofline']['kernel_ms_per_step']:.2f} snap_ms={j['roofline']['snap_ms_per_step']:.2f} "
              f"launches={j['gpu_launches']}", flush=True)
    except Exception as ex:
        print(f"G={g}: failed {ex}: {out.stdout[-300:]} {out.stderr[-600:]}", flush=True)
