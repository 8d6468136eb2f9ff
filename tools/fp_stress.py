"""Debug helper (GPU box): find_path on chunks of the bench workload, each timed."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import habitat_sim_b200  # noqa
from habitat_sim_b200.nav import PathFinder
from bench import make_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lo0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
mode = sys.argv[4] if len(sys.argv) > 4 else "dev"
image, st, en = make_queries(n, 1000)
pf = PathFinder(0); pf.load_nav_mesh_bytes(image)
pf.set_profiling(True)
for lo in range(lo0, n, chunk):
    hi = min(lo + chunk, n)
    if mode == "dev":
        s = torch.from_numpy(st[lo:hi]).cuda(); e = torch.from_numpy(en[lo:hi]).cuda()
    else:
        s, e = st[lo:hi], en[lo:hi]
    torch.cuda.synchronize(); t0 = time.time()
    r = pf.find_paths(s, e)
    torch.cuda.synchronize(); dt = time.time() - t0
    d = r['geodesic_distance']
    d = d.cpu().numpy() if mode == "dev" else d
    print(f"{mode} slice [{lo},{hi}) in {dt:.3f}s -> {(hi-lo)/dt:.0f} q/s; found {np.isfinite(d).mean():.3f} phases {pf.phase_times()}", flush=True)
    try:
        pf.work_counters()
    except Exception as ex:
        print("  FAULT", ex, flush=True)
