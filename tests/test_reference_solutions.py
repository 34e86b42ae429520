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
