#!/usr/bin/env python
"""bench.py -- converged landing NLPs/sec on the BASELINE.json workload.

One "step" = one pass of the hot path over one batch: solve every drop condition of the sweep
(SRB landing sweep, 1k synthetic drop conditions on a height x pitch x roll x forward-velocity grid,
N = 30 knots; BASELINE.json configs[1]) with the batched interior-point kernel.  Under torchrun each
rank solves its own 1k block of a (1k x n_gpus) grid (weak scaling, no data-path collective) and the
step ends with ONE NCCL all-gather of the per-scenario result records (x*, f*, status, iters).

    python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
    python bench.py --impl reference ...                        # the CPU path on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KNOTS = 30
BATCH = 1024
METRIC = "converged landing NLPs/sec (FP64, batch)"
UNIT = "NLP/s"


def workload(n_gpus, rank, batch):
    import landing_controller_b200 as lc
    allb = lc.grid_sweep(batch * n_gpus)
    # interleaved shards: the iteration count grows along the axes of the grid, contiguous blocks would give the last
    # rank the hard end of the sweep
    return allb[lc.shard_indices(batch * n_gpus, n_gpus, rank, "interleaved")].copy()


def iter_bytes(N):
    """Algorithmic HBM bytes of one interior-point iteration of one scenario:
    read + write the primal-dual iterate once, read p: 8 * (2 (340N - 316) + 13N + 81)."""
    return 8 * (2 * (340 * N - 316) + 13 * N + 81)


def iter_flops(N):
    """Algorithmic FP64 flops of one iteration of one scenario with ONE Riccati factorisation:
    per stage 2 x MACs of  T = Pxx G (12.12.36) + lower triangle of G'T (78 tiles x 9 x 12) + G'Pxc (36.12.12)
    + partial Cholesky of the 48x48 stage matrix, 24 pivots (sum_j (48-j)(49-j)/2) + condensing (576 terms x 2)
    + forward sweep (24.24 + 300 + 12.36 + 12.24), plus evaluation (3.1k flop / knot) and row passes."""
    chol = sum((48 - j) * (49 - j) // 2 for j in range(24))
    macs = 12 * 12 * 36 + 78 * 9 * 12 + 36 * 12 * 12 + chol + 2 * 576 + (24 * 24 + 300 + 12 * 36 + 12 * 24)
    return (2 * macs + 3100 + 3000) * (N - 1)


def scratch_bytes(N, slots):
    K, nx, MR = N - 1, 36 * N - 24, 36 + 104 * (N - 1)
    n = 3 * nx + 12 * MR + K * (388 + 192 + 480) + K * (1152 + 36) + (K + 1) * 312 + 144 + 128
    return 8 * slots * ((n + 31) // 32 * 32)


def ncu_traffic():
    """DRAM bytes per launch of k_solve from the committed ncu --set full capture of this workload, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_run(N, drops, threads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ip import solve_cpu
    t = time.perf_counter()
    r = solve_cpu(N, drops, threads=threads)
    dt = time.perf_counter() - t
    return r, dt


def run_reference(args):
    """The reference's CPU path for the hot path, on the host cores.  IPOPT/MUMPS are not available
    (SURVEY.md 8c), so this is the CPU restatement (oracle/: generated-function restatement +
    interior-point "IPOPT substitute"), OpenMP over scenarios with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or args.batch  # default: the whole sweep (about 10 s on 16 host threads)
    drops = workload(1, 0, args.batch)
    sub = drops[:: max(1, len(drops) // sample)][:sample]
    for _ in range(min(args.warmup, 1)):
        cpu_run(KNOTS, sub[:cores], cores)
    tot_t, tot_c, tot_it = 0.0, 0, 0
    for _ in range(args.steps):
        r, dt = cpu_run(KNOTS, sub, cores)
        tot_t += dt
        tot_c += int((r["status"] == 0).sum())
        tot_it += int(r["iters"].sum())
    val = tot_c / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "SRB landing sweep, grid drop conditions, N=%d knots; %d-scenario strided sample of the %d-scenario sweep per step"
                   % (KNOTS, len(sub), args.batch), "knots": KNOTS, "batch": len(sub)},
        "kkt_iters_per_s": tot_it / tot_t,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of %d scenarios per step (every %d-th), %d steps; IPOPT substitute (oracle/ip_ref.c), OpenMP over scenarios"
                         % (len(sub), args.batch, max(1, len(drops) // sample), args.steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    global KNOTS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--knots", type=int, default=KNOTS, help="knots per trajectory (BASELINE configs: 30; 50 for the 16k sweep)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios in the cpu_baseline sample (0 = auto)")
    args = ap.parse_args()
    KNOTS = args.knots
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import landing_controller_b200 as lc

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this library has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.batch, KNOTS
    solver = lc.LandingSolver(N=N, device=local_rank)
    nx = solver.dims["nx"]
    drops_h = workload(world, rank, B)

    # device-resident arm: inputs already in HBM when the timed region starts
    drops_d = torch.tensor(drops_h, device=dev)
    rec = torch.zeros(B, nx + 3, dtype=torch.float64, device=dev)  # result record [x*, f*, status, iters]
    x_d = torch.zeros(B, nx, dtype=torch.float64, device=dev)
    f_d = torch.zeros(B, dtype=torch.float64, device=dev)
    st_d = torch.zeros(B, dtype=torch.int32, device=dev)
    it_d = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = torch.zeros(world * B, nx + 3, dtype=torch.float64, device=dev) if world > 1 else None
    lib_stream = torch.cuda.ExternalStream(solver.stream_ptr, device=dev)

    def step_device():
        solver.solve_device(drops_d, x_d, f_d, st_d, it_d)
        if world > 1:
            with torch.cuda.stream(lib_stream):
                rec[:, :nx] = x_d
                rec[:, nx] = f_d
                rec[:, nx + 1] = st_d.double()
                rec[:, nx + 2] = it_d.double()
                dist.all_gather_into_tensor(gathered, rec)

    # end-to-end arm: pinned host buffers through the C ABI, copies inside the timed region
    pin = lambda *s, dt=torch.float64: torch.empty(*s, dtype=dt).pin_memory()
    drops_p = pin(B, 12)
    drops_p.copy_(torch.from_numpy(drops_h))
    x_p, f_p, v_p = pin(B, nx), pin(B), pin(B)
    st_p, it_p = pin(B, dt=torch.int32), pin(B, dt=torch.int32)
    io = lc.api.SolveIO(lc.api._ptr(drops_p), None, lc.api._ptr(x_p), lc.api._ptr(f_p), None, lc.api._ptr(v_p),
                        lc.api._ptr(st_p, lc.api._ip), lc.api._ptr(it_p, lc.api._ip))
    import ctypes

    def step_e2e():
        rc = solver.lib.landing_solve_batch(solver.ctx, B, lc.HOST, ctypes.byref(solver.problem),
                                            ctypes.byref(solver.options), ctypes.byref(io))
        assert rc == 0, solver.lib.landing_last_error()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lib_stream):
            ev0.record()
        for _ in range(steps):
            fn()
        with torch.cuda.stream(lib_stream):
            ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    l0 = solver.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launches - l0
    st = st_d.cpu().numpy()
    its = it_d.cpu().numpy()
    stats = torch.tensor([float((st == 0).sum()), float(its.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats)
    conv_per_step, iters_per_step = float(stats[0].item()), float(stats[1].item())
    value = conv_per_step * args.steps / (ms * 1e-3)

    for _ in range(min(args.warmup, 1)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    st_e = st_p.numpy()
    conv_e = torch.tensor([float((st_e == 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(conv_e)
    e2e_value = float(conv_e.item()) * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        kernel_ms = ms / args.steps  # one launch per step; the all-gather (N>1) rides on the same stream
        local_iters = float(its.sum())
        achieved = iter_bytes(N) * local_iters / (kernel_ms * 1e-3) / 1e9
        fp64_peak = solver.fp64_peak_tflops()
        fp64_ach = iter_flops(N) * local_iters / (kernel_ms * 1e-3) / 1e12
        tr = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "SRB landing sweep, %d grid drop conditions per GPU (height x pitch x roll x v_x, v_z=-3), N=%d knots"
                       % (B, N), "knots": N, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": "scenario-sharded x%d (interleaved shards of the %d-scenario grid), one all-gather of results"
                                      % (world, B * world),
                       "l2": "per-scenario solver scratch of the 296 resident CTAs (%.2f GB) exceeds the 126 MB L2 and "
                             "every step rewrites all of it; no flush needed" % (1e-9 * scratch_bytes(N, 296)),
                       "options": "tol 1e-4, constr_viol_tol 1e-3, max_iter 3000 (generate_landingCtrller_IPOPT.m:232-236)"},
            "converged_per_step": conv_per_step, "scenarios_per_step": B * world,
            "kkt_iters_per_s": iters_per_step * args.steps / (ms * 1e-3),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 12 * 8,
                    "d2h_bytes_per_step": B * (nx + 2) * 8 + B * 8, "ms_per_step": ms_e2e / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                         "peak_source": peak_src,
                         "kernel": "k_solve", "units_per_launch": local_iters,
                         "bytes_per_unit": iter_bytes(N), "flops_per_unit": iter_flops(N),
                         "fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": fp64_ach / fp64_peak if fp64_peak else None,
                                  "peak_source": "measured in this run (landing_fp64_peak: DFMA micro-kernel, CUDA events)"},
                         "note": "one launch per step = all interior-point iterations of the batch; unit = one KKT "
                                 "iteration of one scenario; the kernel is FP64-latency bound (DESIGN.md 2.3), the hbm "
                                 "line uses the algorithmic bytes of SURVEY 8d-ii, the fp64 line the algorithmic flops"},
        }
        # the evaluation kernels (HBM-bound rows a-3..a-5 of SURVEY 8): 16k scenarios, SoA, a few milliseconds
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from bench_eval import cpu_reference, measure
            keys = ("function", "N", "B", "layout", "ms", "achieved", "peak", "unit", "frac", "evals_per_s")
            line["roofline"]["eval_kernels"] = [
                {k: r[k] for k in keys} for r in measure(N, 16384, 10, "soa", solver=solver, device=local_rank)]
            # the size the reference ships generated C for (N = 21), with the reference's own compiled functions
            # timed on the host cores beside it (oracle/_ref, kind "reference")
            ref = cpu_reference()
            s21 = lc.LandingSolver(N=21, device=local_rank)
            rows = [{k: r[k] for k in keys} for r in measure(21, 16384, 10, "soa", solver=s21, device=local_rank)]
            s21.close()
            for r in rows:
                r["cpu_baseline"] = ref.get(r["function"]) if ref else None
            line["roofline"]["eval_kernels_n21"] = rows
        except Exception as e:  # reported, never hidden
            line["roofline"]["eval_kernels"] = {"error": repr(e)}
        # CPU baseline on the host cores (bounded sample of the same workload)
        cores = os.cpu_count() or 1
        ns = args.cpu_sample or B  # default: the whole sweep of rank 0 (about 10 s on 16 host threads)
        sub = drops_h[:: max(1, B // ns)][:ns]
        r, dt = cpu_run(N, sub, cores)
        line["cpu_baseline"] = {
            "value": float((r["status"] == 0).sum()) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "kkt_iters_per_s": float(r["iters"].sum()) / dt,
            "sample": "%d of %d scenarios (every %d-th), %.1f s; IPOPT substitute oracle/ip_ref.c, OpenMP over scenarios"
                      % (len(sub), B, max(1, B // ns), dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
