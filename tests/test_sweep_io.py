"""The caller-side data formats (landing_controller_b200/sweep_io.py) against the reference's stored sweep files
(fixture tests/golden/ccc_n41.npz, made from optimizations/landing/data/*.mat)."""
import os

import numpy as np

import landing_controller_b200 as lc
from landing_controller_b200 import sweep_io

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ccc_n41.npz")
N = 41


def _x(X, c, f):
    return np.concatenate([X.T.ravel(), np.hstack([c.T, f.T]).ravel()])


def test_opt_sol_record_reproduces_the_stored_fields():
    d = np.load(FIX)
    for b in range(len(d["X"])):
        sol = sweep_io.opt_sol(_x(d["X"][b], d["c"][b], d["f"][b]), N)
        assert np.array_equal(sol["X_star"], d["X"][b]) and np.array_equal(sol["p_star"], d["c"][b])
        assert np.array_equal(sol["f_star"], d["f"][b])
        assert np.array_equal(sol["td"], d["td"][b])  # find(f_z > 1, 1), foot_positions.m:36-37
        assert sol["q_star"].shape == (18, N) and np.array_equal(sol["q_star"][:6], d["X"][b][:6])


def test_mat_round_trip(tmp_path):
    d = np.load(FIX)
    xs = np.array([_x(d["X"][b], d["c"][b], d["f"][b]) for b in range(5)])
    path = str(tmp_path / "pitch_0_vX.mat")
    sweep_io.save_sweep_mat(path, xs, N)
    back = sweep_io.load_sweep_mat(path)
    assert len(back) == 5
    for b in range(5):
        assert np.array_equal(back[b]["X_star"], d["X"][b]) and np.array_equal(back[b]["td"], d["td"][b])
        assert np.array_equal(back[b]["f_star"], d["f"][b])


def test_reference_sweeps_start_where_the_stored_runs_start():
    d = np.load(FIX)
    names = list(d["names"])
    for fixed, sweep, idx in (("pitch_0", "vX", 0), ("pitch_0", "vX", 6), ("pitch_30", "vX", 0), ("pitch_45", "vZ", 0),
                              ("roll_60", "vZ", 9), ("pitch_45", "vY", 6)):
        name = "%s_%s.mat[%d]" % (fixed, sweep, idx)
        if name not in names:
            continue
        drops = sweep_io.reference_sweep(fixed, sweep)
        assert np.allclose(drops[idx], d["X"][names.index(name)][:, 0], atol=1e-12), name
    assert sweep_io.reference_sweep("pitch_0", "vX").shape == (13, 12)
    assert sweep_io.reference_sweep("pitch_0", "vZ").shape == (19, 12)


def test_touchdown_feet_and_training_record():
    d = np.load(FIX)
    b = list(d["names"]).index("drop_vZ.mat[0]") if "drop_vZ.mat[0]" in d["names"] else 2
    sol = sweep_io.opt_sol(_x(d["X"][b], d["c"][b], d["f"][b]), N)
    p = sweep_io.touchdown_feet_body(sol)
    # a level vertical drop lands with the feet under the hips: +-0.19 +- box in x, +-0.1 +- box in y, below the CoM
    assert np.all(np.abs(np.abs(p[0]) - 0.19) <= 0.05 + 1e-6) and np.all(np.abs(np.abs(p[1]) - 0.1) <= 0.05 + 1e-6)
    assert np.all(p[2] < -0.1)
    drop = d["X"][b][:, 0]
    inp, out = sweep_io.training_record(drop, _x(d["X"][b], d["c"][b], d["f"][b]))
    assert inp.shape == (9,) and np.array_equal(inp[:3], drop[3:6]) and out.shape == (36 * N - 24,)


def test_training_set_normalisation_round_trip():
    """data_normalization.m:42-111 / data_denormalization.m:16-41 on the stored IPOPT solutions (N = 41): z-scores with
    MATLAB's std, forces shifted to touchdown and scaled by body weight, td appended; the inverse restores the states and
    foot positions exactly and the forces from each leg's touchdown knot on (earlier columns, all < 1 N in z, become 0)."""
    d = np.load(FIX)
    n = len(d["X"])
    outs = np.array([_x(d["X"][b], d["c"][b], d["f"][b]) for b in range(n)]).T
    ins = np.array([np.concatenate([d["X"][b][3:6, 0], d["X"][b][6:12, 0]]) for b in range(n)]).T
    keep = [b for b in range(n) if all((d["f"][b][3 * l + 2] > 1).any() for l in range(4))]
    outs, ins = outs[:, keep], ins[:, keep]
    mass = 8.251999999999999  # generate_data/data/data_stats.mat
    norm, st = sweep_io.normalize_training_set(ins, outs, N, mass)
    assert norm["input"].shape == ins.shape and norm["output"].shape == (outs.shape[0] + 4, len(keep))
    assert np.allclose(norm["input"].mean(axis=1), 0, atol=1e-9)
    assert np.allclose(norm["input"].std(axis=1, ddof=1)[st["std_input"][:, 0] > 0], 1, atol=1e-9)
    assert st["mean_X"].shape == (12, N) and st["std_U"].shape == (24, N - 1) and st["td_scale"] == 1.0
    for e in range(0, len(keep), 5):
        b = keep[e]
        X, U, jpos = sweep_io.denormalize_sample(norm["output"][:, e], st, N)
        assert jpos is None
        td = norm["output"][-4:, e].astype(int)
        assert np.array_equal(td, d["td"][b].astype(int))           # same touchdown knots as the stored `td`
        ok = st["std_X"] > 0
        assert np.allclose(X[ok], d["X"][b][ok], atol=1e-9) and np.allclose(U[:12], d["c"][b], atol=1e-9)
        for leg in range(4):
            t0 = td[leg] - 1
            assert np.allclose(U[12 + 3 * leg:15 + 3 * leg, t0:], d["f"][b][3 * leg:3 * leg + 3, t0:], atol=1e-9)
            assert np.all(U[12 + 3 * leg:15 + 3 * leg, :t0] == 0)
        # normalised vertical force of a loaded leg: fractions of body weight
        assert 0 < norm["output"][12 * N:-4, e].max() < 10
