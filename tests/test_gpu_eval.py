"""GPU parity tests of the evaluation kernels, through the C ABI (CasADi symbol set and batch API)."""
import numpy as np
import pytest

import landing_controller_b200 as lc
from oracle_lib import CasadiLib, Oracle

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(1.0, np.abs(b))))


@pytest.fixture(scope="module")
def dropin():
    return CasadiLib(lc.DROPIN_PATH)


def test_casadi_abi_against_reference_golden(dropin, golden):
    """The drop-in library called exactly as CasADi's External calls the reference .so."""
    for fn in dropin.FUNCS:
        getattr(dropin.lib, fn + "_incref")()
    for i in range(int(golden["n_cases"])):
        c = {k: golden["c%d_%s" % (i, k)] for k in ("x", "p", "lam_f", "lam_g", "f", "g", "J", "H", "gf", "gx", "gp")}
        rc, (f, g) = dropin.call("nlp", [c["x"], c["p"]])
        assert rc == 0 and rel(f, c["f"]) < TOL and rel(g, c["g"]) < TOL
        rc, (f,) = dropin.call("nlp_f", [c["x"], c["p"]])
        assert rc == 0 and rel(f, c["f"]) < TOL
        rc, (g,) = dropin.call("nlp_g", [c["x"], c["p"]])
        assert rc == 0 and rel(g, c["g"]) < TOL
        rc, (g, J) = dropin.call("nlp_jac_g", [c["x"], c["p"]])
        assert rc == 0 and rel(g, c["g"]) < TOL and rel(J, c["J"]) < TOL
        rc, (H,) = dropin.call("nlp_hess_l", [c["x"], c["p"], c["lam_f"], c["lam_g"]])
        assert rc == 0 and rel(H, c["H"]) < TOL
        rc, (f, gf) = dropin.call("nlp_grad_f", [c["x"], c["p"]])
        assert rc == 0 and rel(f, c["f"]) < TOL and rel(gf, c["gf"]) < TOL
        rc, (f, g, gx, gp) = dropin.call("nlp_grad", [c["x"], c["p"], c["lam_f"], c["lam_g"]])
        assert rc == 0 and rel(gx, c["gx"]) < TOL and rel(gp, c["gp"]) < TOL and rel(g, c["g"]) < TOL


def test_casadi_abi_is_reentrant_with_checked_out_memory_objects(dropin, golden):
    """CasADi evaluates one function from several threads, each with its own memory object: checkout -> F(.., mem) ->
    release (function_internal.cpp:721-733).  Every memory object owns a CUDA stream and staging buffers."""
    import threading
    for fn in dropin.FUNCS:
        getattr(dropin.lib, fn + "_incref")()
    n = int(golden["n_cases"])
    golden = {k: np.array(golden[k]) for k in golden.files}  # (the lazy .npz reader is not thread-safe)
    errs, mems = [], []
    gate = threading.Barrier(6)

    def worker(t):
        try:
            mem = dropin.lib.nlp_jac_g_checkout()
            mems.append(mem)
            gate.wait(timeout=60)  # all six memory objects are checked out before anyone evaluates
            for rep in range(12):
                i = (t + rep) % n
                x, p = golden["c%d_x" % i], golden["c%d_p" % i]
                rc, (g, J) = dropin.call("nlp_jac_g", [x, p], mem=mem)
                assert rc == 0 and rel(g, golden["c%d_g" % i]) < TOL and rel(J, golden["c%d_J" % i]) < TOL
                rc, (H,) = dropin.call("nlp_hess_l", [x, p, golden["c%d_lam_f" % i], golden["c%d_lam_g" % i]], mem=mem)
                assert rc == 0 and rel(H, golden["c%d_H" % i]) < TOL
            dropin.lib.nlp_jac_g_release(mem)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert len(set(mems)) == 6  # six distinct memory objects were live at the same time
    again = dropin.lib.nlp_jac_g_checkout()
    assert again == min(mems)  # released slots are handed out again
    dropin.lib.nlp_jac_g_release(again)
    for fn in dropin.FUNCS:
        getattr(dropin.lib, fn + "_decref")()


def test_casadi_abi_null_handling(dropin, golden):
    """arg[i]==NULL means zeros, res[i]==NULL means skip (landingCtrller_IPOPT.c:69,151)."""
    x, p = golden["c1_x"], golden["c1_p"]
    rc, (g, J) = dropin.call("nlp_jac_g", [x, p], skip=(1,))
    assert rc == 0 and J is None and rel(g, golden["c1_g"]) < TOL
    o = Oracle(21)
    rc, (g,) = dropin.call("nlp_g", [None, p])
    _, g0 = o.g(np.zeros(o.nx), p)
    assert rc == 0 and rel(g, g0) < TOL
    rc, (H,) = dropin.call("nlp_hess_l", [x, p, None, None])
    assert rc == 0 and np.all(H == 0)


@pytest.mark.parametrize("N,B", [(21, 64), (30, 257), (50, 33)])
@pytest.mark.parametrize("layout", [lc.AOS, lc.SOA])
def test_batch_eval_matches_oracle(N, B, layout):
    s = lc.LandingSolver(N=N)
    o = Oracle(N)
    rng = np.random.default_rng(N + B)
    x = rng.normal(size=(B, o.nx)) * 0.4
    p = rng.uniform(0.5, 1.5, size=(B, o.np_))
    lam = rng.normal(size=(B, o.m))
    lf = rng.normal(size=(B, 1))
    tr = (lambda a: np.ascontiguousarray(a.T)) if layout == lc.SOA else (lambda a: a)
    out = s.eval_host(tr(x), tr(p), tr(lf), tr(lam), want=("f", "g", "grad_f", "jac", "hess", "grad_x", "grad_p"),
                      layout=layout)
    if layout == lc.SOA:
        out = {k: (v.T if v.ndim == 2 else v) for k, v in out.items()}
    assert np.all(out["status"] == 0)
    for b in range(0, B, max(1, B // 16)):
        _, g, J = o.jac_g(x[b], p[b])
        _, H = o.hess_l(x[b], p[b], lf[b, 0], lam[b])
        _, f, gf = o.grad_f(x[b], p[b])
        _, _, _, gx, gp = o.grad(x[b], p[b], lf[b, 0], lam[b])
        assert rel(out["f"][b], [f]) < TOL and rel(out["g"][b], g) < TOL
        assert rel(out["jac"][b], J) < TOL and rel(out["hess"][b], H) < TOL
        assert rel(out["grad_f"][b], gf) < TOL and rel(out["grad_x"][b], gx) < TOL
        assert rel(out["grad_p"][b], gp) < TOL
    s.close()


def test_bounds_and_build_match_oracle():
    N = 30
    s = lc.LandingSolver(N=N)
    o = Oracle(N)
    drops = lc.random_sweep(50, seed=3)
    p, x0 = s.build_host(drops)
    lb, ub = s.bounds_host(p)
    pb = o.default_problem()
    for b in range(50):
        p0, x00 = o.build_p_x0(pb, drops[b, :6], drops[b, 6:])
        assert np.array_equal(p[b], p0) and np.array_equal(x0[b], x00)
        l0, u0 = o.bounds(p0)
        assert np.array_equal(lb[b], l0) and np.array_equal(ub[b], u0)
    s.close()


def test_nan_is_reported_per_scenario():
    """pitch = pi/2 hits the Euler singularity (Binv.m:15-17); the batch must not fail."""
    N = 21
    s = lc.LandingSolver(N=N)
    drops = lc.grid_sweep(8)
    p, x0 = s.build_host(drops)
    x0[3, 4] = np.pi / 2
    x0[3, 8] = 1.0  # yaw-axis body rate -> tan(pitch) term
    out = s.eval_host(x0, p, want=("g",))
    assert out["status"][3] == -1 or np.max(np.abs(out["g"][3])) > 1e10
    assert np.all(out["status"][[0, 1, 2, 4, 5, 6, 7]] == 0)
    s.close()


def test_large_batch_linearity_property():
    """Size-independent property at bench scale: g is affine in the forces for fixed (X, c) except the
    bilinear complementarity / no-slip / torque rows, so f and grad_f are checked by the quadratic
    identity f(x) = 0.5 * grad_f(x) . (x - xref_N)."""
    N, B = 30, 4096
    s = lc.LandingSolver(N=N)
    drops = lc.grid_sweep(B)
    p, x0 = s.build_host(drops)
    rng = np.random.default_rng(5)
    x = x0 + 0.1 * rng.normal(size=x0.shape)
    out = s.eval_host(x, p, want=("f", "grad_f"))
    xo = 12 * (N - 1)
    d = x[:, xo:xo + 12] - p[:, xo:xo + 12]
    assert np.allclose(out["f"][:, 0], 0.5 * np.sum(out["grad_f"][:, xo:xo + 12] * d, axis=1), rtol=1e-12, atol=1e-12)
    assert np.count_nonzero(out["grad_f"][:, :xo]) == 0
    s.close()
