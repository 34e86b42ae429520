"""The SRB part of BASELINE configs[4] on one GPU: 64k random drops (seed 0) with the sweep callers' parameter set and
non-uniform knot spacing (generate_training_data_automated.m:28,44-102), normal and large-tilt attitude ranges; GPU solve
time, converged fraction, iteration statistics, and the CPU restatement on a bounded sample of the same drops.
   usage: python tools/bench_tilt.py [B]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import landing_controller_b200 as lc
from oracle_ip import default_options, default_problem, solve_cpu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = lc.SWEEP_N
dt = np.asarray(lc.SWEEP_DT, dtype=np.float64)
dev = torch.device("cuda:0")
for tilt in (False, True):
    drops = lc.random_sweep(B, seed=0, large_tilt=tilt, dt1=dt[0])
    s = lc.LandingSolver(N=N)
    lc.apply_sweep_parameters(s.problem)
    s.set_dt(dt)
    d = torch.tensor(drops, device=dev)
    nx = s.dims["nx"]
    x = torch.zeros(B, nx, dtype=torch.float64, device=dev); f = torch.zeros(B, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev); it = torch.zeros(B, dtype=torch.int32, device=dev)
    s.solve_device(d, x, f, st, it); torch.cuda.synchronize()
    t = time.perf_counter(); s.solve_device(d, x, f, st, it); torch.cuda.synchronize(); el = time.perf_counter() - t
    sth, ith = st.cpu().numpy(), it.cpu().numpy()
    sub = drops[:: B // 256][:256]
    pb = lc.apply_sweep_parameters(default_problem()).set_dt(dt)
    t = time.perf_counter(); c = solve_cpu(N, sub, pb=pb); elc = time.perf_counter() - t
    s.close()
    print(json.dumps({"workload": "%d random drops, seed 0, %s, sweep callers' parameters and dt_val, N=21" % (B, "large tilt" if tilt else "tilt within pi/3"),
                      "gpu_ms": 1e3 * el, "gpu_nlp_per_s": float((sth == 0).sum()) / el, "converged_fraction": float((sth == 0).mean()),
                      "status_counts": {int(k): int(v) for k, v in zip(*np.unique(sth, return_counts=True))},
                      "kkt_iters_per_s": float(ith.sum()) / el, "iters_mean": float(ith.mean()), "iters_p99": float(np.percentile(ith, 99)), "iters_max": int(ith.max()),
                      "cpu_nlp_per_s": float((c["status"] == 0).sum()) / elc, "cpu_converged_fraction": float((c["status"] == 0).mean()),
                      "cpu_cores": os.cpu_count(), "cpu_sample": len(sub)}))
