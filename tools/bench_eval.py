"""Roofline check of the batched evaluation kernels: nlp_g, nlp_jac_g, nlp_hess_l on B scenarios,
device buffers; algorithmic bytes = 8 * (inputs read + outputs written) per scenario, against the measured HBM copy
bandwidth (MEASURED_PEAKS.json).
   usage: python tools/bench_eval.py [N] [B] [reps] [soa|aos]        (prints one JSON line per function)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure(N=30, B=16384, reps=20, layout="soa", solver=None, device=0):
    import numpy as np
    import torch
    import landing_controller_b200 as lc
    s = solver or lc.LandingSolver(N=N, device=device, lib_path=os.environ.get("LANDING_LIB", lc.api.LIB_PATH))
    LAYOUT = lc.SOA if layout == "soa" else lc.AOS
    d = s.dims
    dev = torch.device("cuda", device)
    drops = torch.tensor(np.ascontiguousarray(np.vstack([lc.grid_sweep(1024)] * ((B + 1023) // 1024))[:B]), device=dev)
    shape = (lambda n: (n, B)) if LAYOUT == lc.SOA else (lambda n: (B, n))
    z = lambda n: torch.zeros(*shape(n), dtype=torch.float64, device=dev)
    p, x = z(d["np"]), z(d["nx"])
    s._check(s.lib.landing_build_batch(s.ctx, B, lc.DEVICE, LAYOUT, ctypes.byref(s.problem), lc.api._ptr(drops),
                                       lc.api._ptr(p), lc.api._ptr(x)), "landing_build_batch")
    torch.cuda.synchronize()
    x += 0.01 * torch.randn_like(x)
    lam_g = torch.randn(*shape(d["m"]), dtype=torch.float64, device=dev)
    lam_f = torch.ones(B, dtype=torch.float64, device=dev)
    g, J, H = z(d["m"]), z(d["nnzJ"]), z(d["nnzH"])
    fv, gx, gp = torch.zeros(B, dtype=torch.float64, device=dev), z(d["nx"]), z(d["np"])
    peak, peak_src = hbm_peak()
    stream = torch.cuda.ExternalStream(s.stream_ptr, device=dev)
    # algorithmic words per scenario = what the function MUST read and write: of p only dt (N-1), mu, mass, Ib, Ib_inv
    # enter g / J / H (+ QN, Xref_N for f and the terminal Hessian diagonal) -- SURVEY 8a; the rest of p is never read
    pk = (N - 1) + 8
    gf = torch.zeros(*shape(d["nx"]), dtype=torch.float64, device=dev)
    cases = {
        "nlp_f": (dict(x=x, p=p, f=fv), 12 + 24 + 1),
        "nlp_grad_f": (dict(x=x, p=p, f=fv, grad_f=gf), 12 + 24 + 1 + d["nx"]),
        "nlp_g": (dict(x=x, p=p, g=g), d["nx"] + pk + d["m"]),
        "nlp_jac_g": (dict(x=x, p=p, g=g, jac=J), d["nx"] + pk + d["m"] + d["nnzJ"]),
        "nlp_hess_l": (dict(x=x, p=p, lam_f=lam_f, lam_g=lam_g, hess=H), d["nx"] + pk + 12 + 1 + d["m"] + d["nnzH"]),
        "nlp_grad": (dict(x=x, p=p, lam_f=lam_f, lam_g=lam_g, f=fv, g=g, grad_x=gx, grad_p=gp),
                     d["nx"] + pk + 36 + 1 + d["m"] + 1 + d["m"] + d["nx"] + d["np"]),
    }
    out = []
    torch.cuda.synchronize()
    for name, (kw, words) in cases.items():
        for _ in range(3):
            s.eval(B, lc.DEVICE, LAYOUT, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launches
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(reps):
            s.eval(B, lc.DEVICE, LAYOUT, **kw)
        with torch.cuda.stream(stream):
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 8.0 * words * B / (ms * 1e-3) / 1e9
        out.append({"function": name, "N": N, "B": B, "layout": layout, "ms": ms,
                    "launches_per_call": (s.launches - l0) / reps, "algorithmic_bytes_per_scenario": 8 * words,
                    "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
                    "evals_per_s": B / (ms * 1e-3)})
    return out


def cpu_reference(B=8192, reps=3, threads=None):
    """The REFERENCE's compiled functions (oracle/_ref/landingCtrller_IPOPT.so, N = 21) on the host cores, OpenMP over
    scenarios (oracle/ref_timing.c): evaluations per second of nlp_g, nlp_jac_g, nlp_hess_l.  None if the reference
    library was not built (make -C oracle ref)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import REF_SO, Oracle, build_oracle
    import landing_controller_b200 as lc
    if not os.path.exists(REF_SO):
        return None
    lib = ctypes.CDLL(build_oracle())
    threads = threads or (os.cpu_count() or 1)
    o = Oracle(21)
    pb = o.default_problem()
    d = lc.grid_sweep(1024)[:: 1024 // 64]
    P0 = np.zeros((len(d), o.np_))
    X0 = np.zeros((len(d), o.nx))
    for b in range(len(d)):
        P0[b], X0[b] = o.build_p_x0(pb, d[b, :6], d[b, 6:])
    rng = np.random.default_rng(0)
    P = np.ascontiguousarray(np.tile(P0, (B // len(d) + 1, 1))[:B])
    X = np.ascontiguousarray(np.tile(X0, (B // len(d) + 1, 1))[:B] + 0.01 * rng.standard_normal((B, o.nx)))
    lam_f, lam_g = np.ones(B), rng.standard_normal((B, o.m))
    dp = ctypes.POINTER(ctypes.c_double)
    out = {}
    # (the first parallel regions of a process pay for the OpenMP thread team: one untimed round first)
    for name, ins in 2 * (("nlp_g", [X, P]), ("nlp_jac_g", [X, P]), ("nlp_hess_l", [X, P, lam_f, lam_g])):
        arr = (dp * 4)(*[a.ctypes.data_as(dp) for a in ins] + [None] * (4 - len(ins)))
        sec, chk = ctypes.c_double(), ctypes.c_double()
        rc = lib.ref_time_function(REF_SO.encode(), name.encode(), B, arr, reps, threads, ctypes.byref(sec), ctypes.byref(chk))
        if rc != 0 or not np.isfinite(chk.value):
            return None
        out[name] = {"value": B * reps / sec.value, "unit": "evaluations/s", "cores": threads, "kind": "reference",
                     "sample": "%d scenarios x %d passes of the reference's gcc -O3 %s (N=21), OpenMP over scenarios"
                               % (B, reps, name)}
    return out


def dropin_latency(calls=300):
    """Per-call latency of the CasADi-ABI drop-in (one scenario per call: H2D + launch + D2H on the calling thread) next
    to the reference's own compiled function (oracle/_ref, one thread): what running the UNMODIFIED reference -- IPOPT
    calling nlp_g / nlp_jac_g / nlp_hess_l once per iteration -- costs per evaluation in either library.  Both are timed
    by the same loop (oracle/ref_timing.c: dlopen + B sequential calls of F(arg, res, iw, w, mem))."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import REF_SO, Oracle, build_oracle
    import landing_controller_b200 as lc
    lib = ctypes.CDLL(build_oracle())
    o = Oracle(21)
    pb = o.default_problem()
    d = lc.grid_sweep(1024)[:: 1024 // 8]
    rng = np.random.default_rng(0)
    X = np.zeros((calls, o.nx)); P = np.zeros((calls, o.np_))
    for b in range(calls):
        P[b], X[b] = o.build_p_x0(pb, d[b % len(d), :6], d[b % len(d), 6:])
    X += 0.01 * rng.standard_normal(X.shape)
    lam_f, lam_g = np.ones(calls), rng.standard_normal((calls, o.m))
    dp = ctypes.POINTER(ctypes.c_double)
    rows = []
    libs = [("dropin (GPU, this library)", lc.DROPIN_PATH)] + ([("reference C (gcc -O3, 1 thread)", REF_SO)] if os.path.exists(REF_SO) else [])
    for name, ins in (("nlp_g", [X, P]), ("nlp_jac_g", [X, P]), ("nlp_hess_l", [X, P, lam_f, lam_g])):
        row = {"function": name, "N": 21, "unit": "us per call"}
        for tag, path in libs:
            arr = (dp * 4)(*[a.ctypes.data_as(dp) for a in ins] + [None] * (4 - len(ins)))
            sec, chk = ctypes.c_double(), ctypes.c_double()
            for _ in range(2):  # (first pass: context creation, page-in)
                rc = lib.ref_time_function(path.encode(), name.encode(), calls, arr, 1, 1, ctypes.byref(sec), ctypes.byref(chk))
            row[tag] = None if rc != 0 else 1e6 * sec.value / calls
        rows.append(row)
    return rows


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    layout = sys.argv[4] if len(sys.argv) > 4 else "soa"
    ref = cpu_reference() if N == 21 else None
    for r in measure(N, B, reps, layout):
        if ref and r["function"] in ref:
            r["cpu_baseline"] = ref[r["function"]]
        print(json.dumps(r))
