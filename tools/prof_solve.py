"""Profiling helper (GPU box): one solve of B grid scenarios capped at MAX_ITER iterations, timed.
   usage: python tools/prof_solve.py [N] [B] [MAX_ITER] [REPS]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import landing_controller_b200 as lc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
MI = int(sys.argv[3]) if len(sys.argv) > 3 else 20
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 1
drops = lc.grid_sweep(1024)
# more scenarios than 1024: repeat the grid (throughput runs with a small max_iter have no iteration tail)
drops = np.ascontiguousarray(np.vstack([drops] * ((B + 1023) // 1024))[:B])
s = lc.LandingSolver(N=N, lib_path=os.environ.get("LANDING_LIB", lc.api.LIB_PATH))
s.options.max_iter = MI
dev = torch.device("cuda:0")
d = torch.tensor(drops, device=dev)
nx = s.dims["nx"]
x = torch.zeros(B, nx, dtype=torch.float64, device=dev)
f = torch.zeros(B, dtype=torch.float64, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev)
try:
    import subprocess
    print("clocks:", subprocess.check_output(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw",
                                              "--format=csv,noheader"], text=True).strip())
except Exception:
    pass
for r in range(REPS):
    torch.cuda.synchronize()
    t = time.perf_counter()
    s.solve_device(d, x, f, st, it)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    its = int(it.sum().item())
    print("N %d B %d max_iter %d: %.2f ms, %d iterations, %.1f us per iteration per CTA slot (%d slots), %.0f iter/s"
          % (N, B, MI, dt * 1e3, its, dt * 1e6 * min(B, 296) / max(its, 1), min(B, 296), its / dt))
