"""Times the kino-dynamic evaluation kernels (g, CCS Jacobian) on device-resident SoA buffers: algorithmic
bytes = 8 * (n_x read + outputs written) per scenario against the measured HBM copy bandwidth.
   usage: python tools/bench_kino.py [N] [B]        (prints one JSON line per function)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

DT_VAL = [0.05] + [0.02] * 15 + [0.05, 0.05, 0.1, 0.2]  # generate_landingCtrller_KNITRO.m (the sweep callers' dt_val, N = 21)


def measure_kino(N=21, B=16384, reps=10, solver=None, device=0):
    import numpy as np
    import torch
    import landing_controller_b200 as lc
    from bench_eval import hbm_peak
    s = solver or lc.LandingSolver(N=N, device=device, lib_path=os.environ.get("LANDING_LIB", lc.api.LIB_PATH))
    d = s.kino_dims()
    pb = s.kino_problem(np.array(DT_VAL) if N == 21 else np.full(N - 1, 0.6 / (N - 1)))
    dev = torch.device("cuda", device)
    x = torch.rand(d["nx"], B, dtype=torch.float64, device=dev) - 0.5
    g = torch.zeros(d["m"], B, dtype=torch.float64, device=dev)
    jac = torch.zeros(d["nnzJ"], B, dtype=torch.float64, device=dev)
    st = torch.cuda.ExternalStream(s.stream_ptr, device=dev)
    torch.cuda.synchronize()
    peak, peak_src = hbm_peak()
    out = []
    for name, kw, words in (("kino_g", dict(g=g), d["nx"] + d["m"]), ("kino_jac_g", dict(jac=jac), d["nx"] + d["nnzJ"])):
        for _ in range(3):
            s.kino_eval_device(x, pb, **kw)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launches
        with torch.cuda.stream(st):
            e0.record()
        for _ in range(reps):
            s.kino_eval_device(x, pb, **kw)
        with torch.cuda.stream(st):
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 8.0 * words * B / (ms * 1e-3) / 1e9
        out.append({"function": name, "N": N, "B": B, "layout": "soa", "ms": ms, "launches_per_call": (s.launches - l0) / reps,
                    "algorithmic_bytes_per_scenario": 8 * words, "achieved": gbs, "peak": peak, "unit": "GB/s",
                    "frac": gbs / peak, "peak_source": peak_src, "evals_per_s": B / (ms * 1e-3),
                    "sizes": {"nx": d["nx"], "m": d["m"], "nnzJ": d["nnzJ"]}})
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 21
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    for r in measure_kino(N, B):
        print(json.dumps(r))
