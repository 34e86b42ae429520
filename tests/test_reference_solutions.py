"""Known-answer tests against the reference's own stored IPOPT solutions (N = 41 "CCC" landing problem,
optimizations/landing/data/*.mat; fixture tests/golden/ccc_n41.npz made by tests/golden/make_ccc_golden.py).

The stored trajectories are outputs of the real reference pipeline (CasADi + IPOPT/MUMPS, eval_SRBM_CCC.m).
Their problem shares every constraint row with the hot-path NLP (generate_quadruped_SRBM_CCC.m:118-165 vs
generate_landingCtrller_IPOPT.m:106-169: same Euler dynamics, complementarity, no-slip, friction pyramid),
with other numeric bounds (f_max 250, kinematic box 0.05/0.05/0.27).  So: evaluated with the restated
functions at N = 41, every stored solution must be feasible to IPOPT's tolerance.  This pins the restated
dynamics / contact rows and the CRBA constants (mass, Ib, Ib_inv) at an N other than the generated C's 21.
"""
import os

import numpy as np
import pytest

import landing_controller_b200 as lc
from oracle_lib import Oracle

N = 41
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ccc_n41.npz")


def _problem(o):
    pb = o.default_problem()  # analysis/eval_SRBM_CCC.m:22-56
    pb.T = 0.6
    pb.mu = 1.0
    pb.l_leg_max = 0.35
    pb.f_max = 250.0
    return pb


def _xp(o, pb, X, c, f):
    x = np.concatenate([X.T.ravel(), np.hstack([c.T, f.T]).ravel()])
    p, _ = o.build_p_x0(pb, X[:6, 0], X[6:, 0])
    return x, p


def test_stored_ipopt_solutions_are_feasible_for_the_restated_rows():
    d = np.load(FIX)
    o = Oracle(N)
    pb = _problem(o)
    K = N - 1
    assert len(d["X"]) >= 40
    for i in range(len(d["X"])):
        x, p = _xp(o, pb, d["X"][i], d["c"][i], d["f"][i])
        _, g = o.g(x, p)
        G = np.array([g[36 + 104 * k:36 + 104 * (k + 1)] for k in range(K - 1)])
        dyn = np.array([g[36 + 104 * k:36 + 104 * k + 12] for k in range(K)])
        assert np.max(np.abs(dyn[:, [0, 1, 2, 6, 7, 8]])) < 1e-9  # linear rows (pos, v): exact up to IPOPT's linear solve
        assert np.max(np.abs(dyn)) < 2e-4                        # Euler-rate / omega rows: IPOPT tol 1e-4
        fz = G[:, 12:16]
        leg = G[:, 16:64].reshape(K - 1, 4, 12)
        assert fz.min() > -2e-6 and fz.max() < 250.0 + 1e-3      # 0 <= f_z <= f_max (bound_relax_factor 1e-6)
        assert leg[:, :, 0].min() > -2e-6                        # c_z >= 0
        assert leg[:, :, 1].max() < 0.001 + 2e-6                 # f_z c_z <= 0.001
        assert np.abs(leg[:, :, 2:8]).max() < 0.01 + 5e-5        # |f_z (c+ - c)| <= 0.01
        assert np.abs(leg[:, :, 8:10]).max() < 0.05 + 1e-5       # kinematic box x, y
        assert leg[:, :, 10].max() < 1e-5 and leg[:, :, 10].min() > -0.27 - 1e-5
        assert leg[:, :, 11].max() < 0.35 ** 2 + 1e-5            # |p_rel|^2 <= l_leg_max^2
        assert G[:, 64:80].max() < 2e-6                          # friction pyramid 0.71 mu
        # touchdown index as the reference computes it (analysis/foot_positions.m:36-37: first knot with f_z > 1)
        cs = lc.contact_set(x, N)
        td = np.array([int(np.argmax(cs[:, l])) + 1 if cs[:, l].any() else 0 for l in range(4)])
        ref_td = d["td"][i].astype(int)
        assert np.array_equal(td[ref_td > 0], ref_td[ref_td > 0])


@pytest.mark.gpu
def test_gpu_evaluation_matches_oracle_on_stored_solutions():
    d = np.load(FIX)
    o = Oracle(N)
    pb = _problem(o)
    s = lc.LandingSolver(N=N)
    B = len(d["X"])
    xs, ps = zip(*[_xp(o, pb, d["X"][i], d["c"][i], d["f"][i]) for i in range(B)])
    x, p = np.array(xs), np.array(ps)
    rng = np.random.default_rng(41)
    lam = rng.normal(size=(B, o.m))
    out = s.eval_host(x, p, np.ones(B), lam)
    for b in range(0, B, 5):
        _, g, J = o.jac_g(x[b], p[b])
        _, H = o.hess_l(x[b], p[b], 1.0, lam[b])
        for a, r in ((out["g"][b], g), (out["jac"][b], J), (out["hess"][b], H)):
            assert np.max(np.abs(a - r) / np.maximum(1.0, np.abs(r))) < 1e-10
    s.close()


def test_cpu_solver_reproduces_stored_ipopt_solutions():
    """SOFT known-answer test of the interior-point restatement against real IPOPT output: the CCC problem
    (tests/ccc_problem.py) is solved from the 43 stored drop conditions with the reference's initial guess.  The problem
    is non-convex, so some runs end in a different local solution (lower cost in some cases, higher in others); on
    the majority the restatement lands on IPOPT's solution: same cost to 0.2 % (IPOPT stopped at tol = 1e-4 with relaxed
    bounds), identical touchdown knots, terminal state (z, roll, pitch, all velocities -- the quantities the cost
    weighs) within 1e-3, vertical GRF profiles within a few newtons of 250 (median 0.5 N).  The flight-phase attitude is
    not pinned by the cost and differs more."""
    import ccc_problem as ccc
    from oracle_ip import default_options, default_problem, solve_cpu
    d = np.load(FIX)
    X, F, TD = d["X"], d["f"], d["td"]
    drops = np.ascontiguousarray(X[:, :, 0])
    opt = default_options(run_Qf=list(ccc.QF), kin_box=list(ccc.KIN_BOX))
    r = solve_cpu(ccc.N, drops, opt=opt, pb=ccc.fill_problem(default_problem()))
    assert (r["status"] == 0).mean() >= 0.9
    same, dfz = 0, []
    for b in range(len(drops)):
        if r["status"][b] != 0:
            continue
        Xs, cs, fs = ccc.split(r["x"][b])
        fref = ccc.stored_cost(X[b], F[b])
        if abs(r["f"][b] - fref) <= 2e-3 * fref and ccc.touchdown(fs) == TD[b].astype(int).tolist():
            same += 1
            assert np.max(np.abs(Xs[2:5, -1] - X[b][2:5, -1])) <= 1e-3   # terminal z, roll, pitch
            assert np.max(np.abs(Xs[6:, -1] - X[b][6:, -1])) <= 1e-3     # terminal velocities
            dfz.append(np.max(np.abs(fs[2::3] - F[b][2::3])))
    print("same local solution as IPOPT on %d of %d stored runs, median max|df_z| %.2f N" % (same, len(drops), np.median(dfz)))
    assert same >= 24 and np.median(dfz) <= 2.0


@pytest.mark.gpu
def test_gpu_solver_reproduces_stored_ipopt_solutions():
    """The same soft known-answer test through the C ABI (landing_problem.Qf / kin_box), plus agreement with the CPU
    restatement on the same problem: same cost on the runs both converge on, identical touchdown knots on most."""
    import ccc_problem as ccc
    from oracle_ip import default_options, default_problem, solve_cpu
    d = np.load(FIX)
    X, F, TD = d["X"], d["f"], d["td"]
    drops = np.ascontiguousarray(X[:, :, 0])
    s = lc.LandingSolver(N=ccc.N)
    ccc.fill_problem(s.problem)
    g = s.solve(drops)
    # iterates agree with the restatement after a fixed number of iterations (pins the running-cost terms)
    s.options.max_iter = 3
    g3 = s.solve(drops[:8])
    s.close()
    opt = default_options(run_Qf=list(ccc.QF), kin_box=list(ccc.KIN_BOX))
    pb = ccc.fill_problem(default_problem())
    opt.max_iter = 3
    c3 = solve_cpu(ccc.N, drops[:8], opt=opt, pb=pb)
    assert np.max(np.abs(g3["x"] - c3["x"])) <= 1e-9 * max(1.0, np.max(np.abs(c3["x"])))
    opt.max_iter = 3000
    c = solve_cpu(ccc.N, drops, opt=opt, pb=pb)
    assert (g["status"] == 0).mean() >= 0.9
    same = 0
    for b in range(len(drops)):
        if g["status"][b] != 0:
            continue
        Xs, cs, fs = ccc.split(g["x"][b])
        fref = ccc.stored_cost(X[b], F[b])
        if abs(g["f"][b] - fref) <= 2e-3 * fref and ccc.touchdown(fs) == TD[b].astype(int).tolist():
            same += 1
            assert np.max(np.abs(Xs[2:5, -1] - X[b][2:5, -1])) <= 1e-3 and np.max(np.abs(Xs[6:, -1] - X[b][6:, -1])) <= 1e-3
    both = (g["status"] == 0) & (c["status"] == 0)
    agree = np.abs(g["f"][both] - c["f"][both]) <= 1e-3 * np.abs(c["f"][both])
    print("GPU: same local solution as IPOPT on %d of %d stored runs; same cost as the CPU restatement on %d of %d"
          % (same, len(drops), agree.sum(), both.sum()))
    assert same >= 24 and agree.mean() >= 0.7


FIX_ALL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ccc_n41_all.npz")


def _compare_with_all_stored(r):
    """counts of (converged, same local solution as IPOPT, lower cost, higher cost) over the 294 stored runs"""
    import ccc_problem as ccc
    d = np.load(FIX_ALL)
    same = lower = higher = 0
    dfz, dterm = [], []
    for b in range(len(d["drops"])):
        if r["status"][b] != 0:
            continue
        Xs, cs, fs = ccc.split(r["x"][b])
        rel = (r["f"][b] - d["cost"][b]) / d["cost"][b]
        if abs(rel) <= 2e-3 and ccc.touchdown(fs) == d["td"][b].astype(int).tolist():
            same += 1
            dfz.append(np.max(np.abs(fs[2::3] - d["fz"][b])))
            dterm.append(max(np.max(np.abs(Xs[2:5, -1] - d["term"][b][2:5])), np.max(np.abs(Xs[6:, -1] - d["term"][b][6:]))))
        elif rel < 0:
            lower += 1
        else:
            higher += 1
    return int((r["status"] == 0).sum()), same, lower, higher, float(np.median(dfz)), float(np.max(dterm))


def test_cpu_solver_on_all_294_stored_ipopt_solutions():
    """The same soft known-answer test on EVERY valid stored IPOPT run of the reference (294; fixture
    tests/golden/ccc_n41_all.npz).  Measured: 286 converge, 190 land on IPOPT's solution (cost within 0.2 %, identical
    touchdown knots, terminal state within 1e-3), 65 find a lower cost than IPOPT did, 31 a higher one."""
    import ccc_problem as ccc
    from oracle_ip import default_options, default_problem, solve_cpu
    d = np.load(FIX_ALL)
    assert len(d["drops"]) == 294
    opt = default_options(run_Qf=list(ccc.QF), kin_box=list(ccc.KIN_BOX))
    r = solve_cpu(ccc.N, np.ascontiguousarray(d["drops"]), opt=opt, pb=ccc.fill_problem(default_problem()))
    conv, same, lower, higher, dfz, dterm = _compare_with_all_stored(r)
    print("CPU: converged %d of 294, same local solution as IPOPT %d, lower cost %d, higher cost %d; median max|df_z| %.2f N, "
          "max terminal-state difference %.1e" % (conv, same, lower, higher, dfz, dterm))
    assert conv >= 280 and same >= 180 and higher <= 40 and dfz <= 1.0 and dterm <= 1e-3


@pytest.mark.gpu
def test_gpu_solver_on_all_294_stored_ipopt_solutions():
    """... and through the C ABI on the GPU: the same counts within a few runs (the non-convex problem amplifies rounding
    differences between the two implementations on a handful of runs)."""
    import ccc_problem as ccc
    d = np.load(FIX_ALL)
    s = lc.LandingSolver(N=ccc.N)
    ccc.fill_problem(s.problem)
    g = s.solve(np.ascontiguousarray(d["drops"]))
    s.close()
    conv, same, lower, higher, dfz, dterm = _compare_with_all_stored(g)
    print("GPU: converged %d of 294, same local solution as IPOPT %d, lower cost %d, higher cost %d; median max|df_z| %.2f N, "
          "max terminal-state difference %.1e" % (conv, same, lower, higher, dfz, dterm))
    assert conv >= 280 and same >= 180 and higher <= 40 and dfz <= 1.0 and dterm <= 1e-3
