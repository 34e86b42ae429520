"""Debug helper (GPU box): GPU solver vs CPU oracle iterate-by-iterate."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import landing_controller_b200 as lc
from oracle_ip import solve_cpu, default_options

N = int(sys.argv[1]) if len(sys.argv) > 1 else 21
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
drops = lc.grid_sweep(1024)[:: 1024 // B][:B]
s = lc.LandingSolver(N=N)
for mi in (0, 1, 2, 3, 5, 10, 30, 3000):
    s.options.max_iter = mi
    t = time.time()
    r = s.solve(drops)
    tg = time.time() - t
    c = solve_cpu(N, drops, default_options(max_iter=mi))
    dx = np.max(np.abs(r["x"] - c["x"]), axis=1)
    print("max_iter", mi, "gpu %.3fs" % tg, "status gpu", r["status"].tolist(), "cpu", c["status"].tolist())
    print("   iters gpu", r["iters"].tolist(), "cpu", c["iters"].tolist())
    print("   max|dx|", np.array2string(dx, precision=2), "f gpu", np.array2string(r["f"], precision=4), "f cpu", np.array2string(c["f"], precision=4))
