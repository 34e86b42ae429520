"""Kino-dynamic ("full-body") landing NLP, the reference's KNITRO variant.

CPU: the numpy restatement oracle/kino_ref.py against the solutions of this NLP that the reference stores
(generate_solver/prevSoln.mat, main_scripts/prevSoln.mat -> tests/golden/kino_n21.npz by make_kino_golden.py):
every one of the 2844 restated rows is feasible at the stored points, and the stored multipliers make the stored point
stationary for the restated Jacobian -- which pins row order, signs, the XYZ rotation convention, the leg kinematics
and the torque rows.  GPU: the batched CUDA evaluation (g and the CCS Jacobian) against that oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kino_ref as kr  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "kino_n21.npz")


def stored(tag):
    d = np.load(GOLD)
    x, lam = d[tag + "_x"], d[tag + "_lam_g"]
    pb = kr.default_problem(21)
    X, _, U = kr.split(x, 21)
    vb = kr.rpy_to_rot_xyz(X[0, 3:6]).T @ X[0, 9:12]  # generate_landingCtrller_KNITRO.m:255-257
    pb["kin_box"] = np.array([kr.kin_box_limits(vb[0], "x"), kr.kin_box_limits(vb[1], "y")])
    return pb, x, lam, (X[0, :6].copy(), X[0, 6:].copy(), U[0, :12].copy())


def test_sizes_match_the_stored_artefacts():
    d = kr.dims(21)
    assert d["nx"] == 972 and d["m"] == 2844  # data_param.mat output_mean, prevSoln.mat lam_g_star
    assert np.load(GOLD)["ms_lam_g"].shape == (2844,)


def test_closed_form_leg_kinematics_equal_the_spatial_chain():
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = np.concatenate([rng.normal(size=3), rng.uniform(-1, 1, 3), rng.uniform(-1.5, 2.5, 12)])
        lit = kr.forward_kin_foot_literal(q)  # get_forward_kin_foot.m on the 18-body model
        R = kr.rpy_to_rot_xyz(q[3:6])
        for l in range(4):
            assert np.allclose(q[:3] + R @ kr.leg_fk_body(l, q[6 + 3 * l:9 + 3 * l]), lit[l], atol=1e-14)
    # home pose of the model (get_robot_model.m:147): feet under the hips, symmetric
    feet = kr.forward_kin_foot_literal(np.concatenate([np.zeros(6), np.tile([0, -1.45, 2.65], 4)]))
    assert np.allclose(feet[0] * [1, -1, 1], feet[1]) and np.allclose(feet[2] * [1, -1, 1], feet[3])


@pytest.mark.parametrize("tag,tol", [("ms", 1e-6), ("gs", 5e-5)])
def test_stored_knitro_solutions_are_feasible_for_the_restated_rows(tag, tol):
    pb, x, _, (q0, qd0, c0) = stored(tag)
    g = kr.eval_g(pb, x)
    lb, ub = kr.bounds(pb, q0, qd0, c0)
    assert g.shape == (2844,) and np.all(lb <= ub)
    viol = np.maximum(np.maximum(lb - g, g - ub), 0.0)
    assert viol.max() <= tol, (viol.max(), int(viol.argmax()))
    # the equality rows (dynamics) hold to the solver's tolerance, the foot positions follow the leg kinematics to 1 cm
    for k in range(20):
        assert np.abs(g[48 + 141 * k:48 + 141 * k + 12]).max() <= tol


def test_stored_knitro_point_is_stationary_for_the_restated_jacobian():
    pb, x, lam, _ = stored("ms")
    J = kr.jac_fd(pb, x)
    r = J.T @ lam
    QN = np.array([0, 0, 100, 10, 10, 0, 10, 10, 10, 10, 10, 10.0])         # :260
    ref = np.array([0, 0, 0.25, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])              # :236-237
    r[240:252] += 2 * QN * (x[240:252] - ref)                              # grad f lives on X_{N-1} only (:86-88)
    scale = np.abs(J.T * lam).max()
    assert scale > 1e-2 and np.abs(r).max() <= 1e-4 * scale, (np.abs(r).max(), scale)


def test_library_kino_pattern_covers_the_oracle_jacobian():
    """Host logic of the product library, no GPU needed: the CCS pattern of dg/dx that landing_kino_sparsity builds (by
    probing the typed knot function on the host) has the oracle's sizes, sorted rows per column, and contains every
    non-zero of the oracle's finite-difference Jacobian at random points -- while staying sparse (13 536 entries at N=21)."""
    import ctypes
    import landing_controller_b200 as lc
    lib = lc.load_library()
    N = 21
    d = (ctypes.c_longlong * 4)()
    assert lib.landing_kino_dims(N, d) == 0
    nx, m, nnz = int(d[1]), int(d[2]), int(d[3])
    assert (nx, m) == (kr.dims(N)["nx"], kr.dims(N)["m"]) and nnz == 13536
    sp = np.ctypeslib.as_array(lib.landing_kino_sparsity(N), shape=(2 + nx + 1 + nnz,))
    assert sp[0] == m and sp[1] == nx
    colind, row = sp[2:2 + nx + 1], sp[2 + nx + 1:]
    assert colind[0] == 0 and colind[-1] == nnz and np.all(np.diff(colind) > 0)  # every variable enters some row
    for c in range(nx):
        r = row[colind[c]:colind[c + 1]]
        assert np.all(np.diff(r) > 0) and r[0] >= 0 and r[-1] < m
    pbo = kr.default_problem(N)
    rng = np.random.default_rng(11)
    cols = np.sort(rng.choice(nx, size=120, replace=False))
    for rep in range(2):
        x = rng.uniform(-0.7, 0.7, size=nx)
        Jo = kr.jac_fd(pbo, x, cols=cols, literal=False)
        for c in cols:
            nzr = np.nonzero(np.abs(Jo[:, c]) > 1e-6)[0]
            assert np.isin(nzr, row[colind[c]:colind[c + 1]]).all(), (c, nzr)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [21, 30])
def test_gpu_kino_functions_match_the_oracle(N):
    import landing_controller_b200 as lc
    s = lc.LandingSolver(N=N, device=0)
    d = s.kino_dims()
    assert d["nx"] == kr.dims(N)["nx"] and d["m"] == kr.dims(N)["m"]
    pbo = kr.default_problem(N)
    pb = s.kino_problem(pbo["dt"], mu=pbo["mu"], mass=pbo["mass"], Ib=pbo["Ib"], Ib_inv=pbo["Ib_inv"])
    rng = np.random.default_rng(N)
    B = 5
    if N == 21:  # the stored solution and perturbations of it
        _, xs, _, _ = stored("ms")
        x = xs[None, :] + 0.05 * rng.normal(size=(B, d["nx"]))
        x[0] = xs
    else:
        x = rng.uniform(-0.6, 0.6, size=(B, d["nx"]))
    g, jac = s.kino_eval_host(x, pb)
    colind, row = s.kino_sparsity()
    assert colind[-1] == d["nnzJ"] == jac.shape[1] and np.all(np.diff(colind) >= 0)
    for b in range(B):
        go = kr.eval_g(pbo, x[b])
        assert np.max(np.abs(g[b] - go) / np.maximum(1.0, np.abs(go))) <= 1e-12
    # Jacobian: dense from CCS against central differences of the oracle (one scenario; FD error ~1e-7)
    b = 0
    Jd = np.zeros((d["m"], d["nx"]))
    for c in range(d["nx"]):
        Jd[row[colind[c]:colind[c + 1]], c] = jac[b, colind[c]:colind[c + 1]]
    cols = np.sort(rng.choice(d["nx"], size=240, replace=False))
    Jo = kr.jac_fd(pbo, x[b], cols=cols, literal=False)
    assert np.max(np.abs(Jd[:, cols] - Jo[:, cols])) <= 2e-6 * max(1.0, np.abs(Jo).max())
    assert np.count_nonzero(np.abs(Jo[:, cols]) > 1e-9) <= np.count_nonzero(Jd[:, cols]) + 0  # nothing outside the pattern
    # SoA layout gives the same numbers
    g2, j2 = s.kino_eval_host(np.ascontiguousarray(x.T), pb, layout=lc.SOA)
    assert np.array_equal(g2.T, g) and np.array_equal(j2.T, jac)
    if N == 21:  # the stored multipliers make the stored point stationary for the GPU Jacobian too
        _, xs, lam, _ = stored("ms")
        r = Jd.T @ lam
        QN = np.array([0, 0, 100, 10, 10, 0, 10, 10, 10, 10, 10, 10.0])
        r[240:252] += 2 * QN * (xs[240:252] - np.array([0, 0, 0.25, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]))
        assert np.abs(r).max() <= 1e-4 * np.abs(Jd.T * lam).max()


@pytest.mark.gpu
def test_gpu_kino_setup_bounds_guess_and_cost():
    """landing_kino_setup_batch / landing_kino_cost_batch: what generate_landingCtrller_KNITRO.m computes per drop before
    it calls the solver.  Bounds against the oracle (oracle/kino_ref.py:bounds) with the drop's c_init and velocity-
    dependent kinematic box; at the drop conditions of the two STORED KNITRO solutions the stored points must be feasible
    for the library's bounds (which pins c_init, kin_box and the row order against reference output); initial guess from
    an SRB solution and from the reference trajectories; terminal cost and gradient."""
    import landing_controller_b200 as lc
    N = 21
    s = lc.LandingSolver(N=N, device=0)
    d = s.kino_dims()
    ss = np.array([[1, -1, 1], [1, 1, 1], [-1, -1, 1], [-1, 1, 1.0]])
    drops = lc.random_sweep(6, seed=3)
    xs = []
    for i, tag in enumerate(("ms", "gs")):  # the stored solutions' own drop conditions
        _, x, _, (q0, qd0, _) = stored(tag)
        drops[i, :6], drops[i, 6:] = q0, qd0
        xs.append(x)
    rng = np.random.default_rng(0)
    x_srb = rng.normal(size=(6, 36 * N - 24))
    lb, ub, x0 = s.kino_setup_host(drops, x_srb=x_srb)
    for b in range(6):
        pbo = kr.default_problem(N)
        R = kr.rpy_to_rot_xyz(drops[b, 3:6])
        vb = R.T @ drops[b, 9:12]
        pbo["kin_box"] = np.array([kr.kin_box_limits(vb[0], "x"), kr.kin_box_limits(vb[1], "y")])
        c_init = np.concatenate([drops[b, :3] + R @ (ss[l] * np.array([0.2, 0.15, -0.3])) for l in range(4)])
        lo, uo = kr.bounds(pbo, drops[b, :6], drops[b, 6:], c_init)
        assert np.array_equal(np.isfinite(lo), np.isfinite(lb[b])) and np.array_equal(np.isfinite(uo), np.isfinite(ub[b]))
        fl, fu = np.isfinite(lo), np.isfinite(uo)
        assert np.max(np.abs(lb[b][fl] - lo[fl])) <= 1e-15 and np.max(np.abs(ub[b][fu] - uo[fu])) <= 1e-15
        # initial guess: X and U of the SRB solution around the constant joint-angle guess
        assert np.array_equal(x0[b, :12 * N], x_srb[b, :12 * N]) and np.array_equal(x0[b, 24 * N - 12:], x_srb[b, 12 * N:])
        assert np.allclose(x0[b, 12 * N:24 * N - 12], np.tile([0, -np.pi / 4, np.pi / 2], 4 * (N - 1)), atol=1e-15)
    for i, tag in enumerate(("ms", "gs")):
        pb, x, _, _ = stored(tag)
        g = kr.eval_g(pb, x)
        viol = np.maximum(np.maximum(lb[i] - g, g - ub[i]), 0.0)
        assert viol.max() <= (1e-6 if tag == "ms" else 5e-5), (tag, viol.max(), int(viol.argmax()))
    # no SRB solution: the reference trajectories Xref / Uref (:272-286)
    _, _, xr = s.kino_setup_host(drops, want_bounds=False)
    X, J, U = kr.split(xr[2], N)
    q_ref = np.array([0, 0, 0.25, 0, 0, 0.0])
    for k in (0, 7, N - 1):
        t = k / (N - 1)
        assert np.allclose(X[k, :6], drops[2, :6] + (q_ref - drops[2, :6]) * t, atol=1e-14)
        assert np.allclose(X[k, 6:], drops[2, 6:] * (1 - t), atol=1e-14)
    k = 5
    Rk = kr.rpy_to_rot_xyz(X[k, 3:6])
    cref = np.concatenate([X[k, :3] + Rk @ (ss[l] * np.array([0.2, 0.2, -0.3])) for l in range(4)])
    assert np.allclose(U[k, :12], cref, atol=1e-14) and np.all(U[k, 12:] == 0)
    # terminal cost (:86-88)
    xx = rng.normal(size=(4, d["nx"]))
    f, gf = s.kino_cost_host(xx)
    QN = np.array([0, 0, 100, 10, 10, 0, 10, 10, 10, 10, 10, 10.0])
    ref = np.array([0, 0, 0.25, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    e = xx[:, 12 * (N - 1):12 * N] - ref
    assert np.allclose(f, (QN * e * e).sum(1), rtol=1e-14)
    go = np.zeros_like(xx)
    go[:, 12 * (N - 1):12 * N] = 2 * QN * e
    assert np.allclose(gf, go, rtol=1e-14, atol=0)
    # device buffers, SoA layout: the same numbers
    import torch
    dev = torch.device("cuda:0")
    dT = torch.tensor(np.ascontiguousarray(drops.T), device=dev)
    xT = torch.tensor(np.ascontiguousarray(x_srb.T), device=dev)
    lbT = torch.zeros(d["m"], 6, dtype=torch.float64, device=dev)
    ubT, x0T = torch.zeros_like(lbT), torch.zeros(d["nx"], 6, dtype=torch.float64, device=dev)
    s.kino_setup_device(dT, x_srb=xT, lbg=lbT, ubg=ubT, x0=x0T, layout=lc.SOA)
    s.synchronize()
    assert np.array_equal(lbT.cpu().numpy().T, lb) and np.array_equal(ubT.cpu().numpy().T, ub)
    assert np.array_equal(x0T.cpu().numpy().T, x0)
    s.close()


@pytest.mark.gpu
def test_gpu_kino_smallest_problem_and_ragged_batch():
    """N = 3 (one interior knot + the last knot) and a batch that does not fill a CTA: g against the oracle, the Jacobian
    against central differences of the oracle on every column."""
    import landing_controller_b200 as lc
    N, B = 3, 131
    s = lc.LandingSolver(N=N, device=0)
    d = s.kino_dims()
    pbo = kr.default_problem(N)
    pb = s.kino_problem(pbo["dt"], mu=pbo["mu"], mass=pbo["mass"], Ib=pbo["Ib"], Ib_inv=pbo["Ib_inv"])
    x = np.random.default_rng(5).uniform(-0.6, 0.6, size=(B, d["nx"]))
    g, jac = s.kino_eval_host(x, pb)
    colind, row = s.kino_sparsity()
    for b in (0, 127, 128, 130):
        go = kr.eval_g(pbo, x[b])
        assert np.max(np.abs(g[b] - go) / np.maximum(1.0, np.abs(go))) <= 1e-12
    b = 130
    Jd = np.zeros((d["m"], d["nx"]))
    for c in range(d["nx"]):
        Jd[row[colind[c]:colind[c + 1]], c] = jac[b, colind[c]:colind[c + 1]]
    Jo = kr.jac_fd(pbo, x[b], literal=False)
    assert np.max(np.abs(Jd - Jo)) <= 2e-6 * max(1.0, np.abs(Jo).max())
    s.close()


@pytest.mark.gpu
def test_gpu_kino_edge_cases():
    """Empty batch, single outputs, missing arguments: B = 0 is a no-op, every output is optional and the requested one
    does not depend on which others are requested, NULL inputs are refused with an error code (no crash)."""
    import ctypes
    import landing_controller_b200 as lc
    N = 21
    s = lc.LandingSolver(N=N, device=0)
    d = s.kino_dims()
    pbo = kr.default_problem(N)
    pb = s.kino_problem(pbo["dt"])
    x = np.random.default_rng(2).uniform(-0.5, 0.5, size=(3, d["nx"]))
    g, jac = s.kino_eval_host(x, pb)
    g1, none = s.kino_eval_host(x, pb, want_jac=False)
    none2, j1 = s.kino_eval_host(x, pb, want_g=False)
    assert none is None and none2 is None and np.array_equal(g1, g) and np.array_equal(j1, jac)
    g0, j0 = s.kino_eval_host(np.zeros((0, d["nx"])), pb)
    assert g0.shape == (0, d["m"]) and j0.shape == (0, d["nnzJ"])
    lib, dp = s.lib, ctypes.POINTER(ctypes.c_double)
    null = ctypes.cast(None, dp)
    assert lib.landing_kino_eval_batch(s.ctx, 3, lc.HOST, lc.AOS, ctypes.byref(pb), null, null, null) != 0
    assert b"kino" in lib.landing_last_error()
    ks = s.kino_setup_data()
    assert lib.landing_kino_setup_batch(s.ctx, 3, lc.HOST, lc.AOS, ctypes.byref(ks), null, null, null, null, null) != 0
    drops = lc.random_sweep(3, seed=4)
    lb, ub, none3 = s.kino_setup_host(drops, want_x0=False)
    lb2, ub2, x0 = s.kino_setup_host(drops)
    assert none3 is None and np.array_equal(lb, lb2) and np.array_equal(ub, ub2) and x0.shape == (3, d["nx"])
    assert np.all(lb <= ub)
    s.close()
