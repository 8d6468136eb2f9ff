"""GPU throughput sweep of the find_path search variants (HBN_FP_G) on the C4 workload.
usage: python tools/sweep_fp.py [queries]"""
import json, os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "400000"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# lane:N = HBN_LANE_CFG (hbn_capi.cu laneKernel): 0 shipped, 1-4 shared-heap / occupancy variants,
# 5-7, 11 heap code variant 2 (two heap levels per HBM round trip), 8-10 47 / 55 shared entries at 19-21 warps/SM
variants = (("lane", None), ("lane:24", None), ("lane:17", None), ("lane:22", None), ("lane:20", None), ("lane:21", None),
            ("lane:18", None), ("lane:23", None), ("lane:1", None))
if os.environ.get("SWEEP_LANE_ALL"):
    variants = (("lane", None),) + tuple((f"lane:{c}", None) for c in range(1, 25))
os.environ.setdefault("HBN_QUERY_CACHE", "/tmp/hbn_queries")  # bench.py generates the pairs once
if os.environ.get("SWEEP_ALL"):
    variants += (("8", None), ("32", None), ("4", None))
for g, bps in variants:
    env = dict(os.environ, HBN_FP_G=g.split(":")[0])
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
