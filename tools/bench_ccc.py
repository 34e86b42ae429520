"""The CCC variant (the problem behind the reference's stored IPOPT solutions, N = 41) on a 1k grid sweep: GPU solve time and the
CPU restatement on the host cores.   usage: python tools/bench_ccc.py [B]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import landing_controller_b200 as lc
from oracle_ip import default_options, default_problem, solve_cpu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = lc.sweeps.CCC_N
drops = lc.grid_sweep(B)
drops[:, 2] = 0.6  # the stored sweeps all start at 0.6 m
s = lc.LandingSolver(N=N)
lc.apply_ccc_parameters(s.problem)
dev = torch.device("cuda:0")
d = torch.tensor(drops, device=dev)
nx = s.dims["nx"]
x = torch.zeros(B, nx, dtype=torch.float64, device=dev); f = torch.zeros(B, dtype=torch.float64, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(2):
    s.solve_device(d, x, f, st, it)
torch.cuda.synchronize()
t = time.perf_counter(); s.solve_device(d, x, f, st, it); torch.cuda.synchronize(); dt = time.perf_counter() - t
conv = int((st == 0).sum().item()); its = int(it.sum().item())
sub = drops[:: max(1, B // 128)][:128]
opt = default_options(run_Qf=list(lc.sweeps.CCC_QF), kin_box=list(lc.sweeps.CCC_KIN_BOX))
t = time.perf_counter(); c = solve_cpu(N, sub, opt=opt, pb=lc.apply_ccc_parameters(default_problem())); dtc = time.perf_counter() - t
print(json.dumps({"workload": "CCC variant, %d grid drops at 0.6 m, N=41" % B, "gpu_ms": 1e3 * dt, "gpu_nlp_per_s": conv / dt, "converged": conv,
                  "kkt_iters": its, "cpu_nlp_per_s": float((c["status"] == 0).sum()) / dtc, "cpu_cores": os.cpu_count(), "cpu_sample": len(sub)}))
