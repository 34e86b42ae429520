"""Timing of the batched TVLQR pass (landing_tvlqr_batch) on solved trajectories, device buffers.
   usage: python tools/bench_tvlqr.py [N] [B]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import landing_controller_b200 as lc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
s = lc.LandingSolver(N=N, lib_path=os.environ.get("LANDING_LIB", lc.api.LIB_PATH))
dev = torch.device("cuda:0")
drops = lc.grid_sweep(1024)
x = torch.tensor(s.solve(drops)["x"], device=dev).repeat((B + 1023) // 1024, 1)[:B].contiguous()
par = s.tvlqr_default()
P = torch.zeros(B, par.n_steps, 576, dtype=torch.float64, device=dev)
K = torch.zeros(B, par.n_steps, 288, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(s.stream_ptr, device=dev)
call = lambda: s._check(s.lib.landing_tvlqr_batch(s.ctx, B, lc.DEVICE, ctypes.byref(par), lc.api._ptr(x), lc.api._ptr(P), lc.api._ptr(K)), "tvlqr")
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
for _ in range(10):
    call()
with torch.cuda.stream(stream):
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = 8.0 * B * (par.n_steps * (576 + 288) + 36 * N - 24)
flops = 2.0 * B * par.n_steps * (24 ** 3 + 24 * 24 * 12 + 24 * 24 * 12)
print(json.dumps({"kernel": "k_tvlqr", "N": N, "B": B, "n_steps": par.n_steps, "ms": ms, "trajectories_per_s": B / (ms * 1e-3),
                  "GBs": byt / (ms * 1e-3) / 1e9, "GFLOPs": flops / (ms * 1e-3) / 1e9}))
