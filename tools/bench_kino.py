"""Times the kino-dynamic evaluation kernels (g, CCS Jacobian) on device-resident SoA buffers.
   usage: python tools/bench_kino.py [N] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import landing_controller_b200 as lc
import kino_ref as kr
N = int(sys.argv[1]) if len(sys.argv) > 1 else 21
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
s = lc.LandingSolver(N=N)
d = s.kino_dims()
pbo = kr.default_problem(N)
pb = s.kino_problem(pbo["dt"])
dev = torch.device("cuda:0")
x = torch.rand(d["nx"], B, dtype=torch.float64, device=dev) - 0.5
g = torch.zeros(d["m"], B, dtype=torch.float64, device=dev)
jac = torch.zeros(d["nnzJ"], B, dtype=torch.float64, device=dev)
st = torch.cuda.ExternalStream(s.stream_ptr, device=dev)
torch.cuda.synchronize()
peak = 6552.6
for name, kw, nbytes in (("kino g", dict(g=g), 8 * (d["nx"] + d["m"])), ("kino jac", dict(jac=jac), 8 * (d["nx"] + d["nnzJ"]))):
    for _ in range(3): s.kino_eval_device(x, pb, **kw)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(10): s.kino_eval_device(x, pb, **kw)
        e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s N=%d B=%d: %.3f ms, %.1f M evals/s, %.0f GB/s algorithmic = %.1f %% of the measured copy bandwidth (nx %d m %d nnz %d)"
          % (name, N, B, ms, B / ms / 1e3, nbytes * B / ms / 1e6, 100 * nbytes * B / ms / 1e6 / peak, d["nx"], d["m"], d["nnzJ"]))
