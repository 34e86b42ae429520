"""Data formats on either side of the solve, as the reference's sweep callers use them.

* `opt_sol(x, N)` -- the per-scenario record `analysis/foot_positions.m:31-44` stores: X_star [12 x N], p_star
  [12 x (N-1)] (foot positions), f_star [12 x (N-1)] (GRFs), td [4] (first knot with f_z > 1, 1-based like MATLAB's find;
  0 if the leg never loads), q_star [18 x N] (base pose + the home leg angles `eval_SRBM_CCC.m:111`, no IK).
* `save_sweep_mat / load_sweep_mat` -- `data/<fixed>_<sweep>.mat` with a 1 x n cell `opt_sol` (`foot_positions.m:47-50`).
* `reference_sweep(fixed, sweep)` -- drop conditions of the stored sweeps (`foot_positions.m:13-33`).
* `touchdown_feet_body(sol)` -- feet relative to the CoM in the body frame at touchdown (`foot_positions.m:54-62`).
* `training_record(drop, x)` -- `training_data.input = [rpy0; qd0]` / `.output = x*`
  (`generate_data/generate_training_data_automated.m:209-214`; the SRB stage has no joint angles).
"""
import numpy as np

Q_LEG_HOME = (0.0, -1.45, 2.65)  # eval_SRBM_CCC.m:15


def split(x, N):
    x = np.asarray(x, dtype=np.float64)
    X = x[:12 * N].reshape(N, 12).T
    U = x[12 * N:].reshape(N - 1, 24).T
    return X, U[:12], U[12:]


def opt_sol(x, N):
    X, p, f = split(x, N)
    td = np.array([int(np.argmax(f[3 * l + 2] > 1)) + 1 if (f[3 * l + 2] > 1).any() else 0 for l in range(4)], dtype=np.float64)
    q = np.vstack([X[:6], np.repeat(np.tile(Q_LEG_HOME, 4)[:, None], N, axis=1)])
    return {"X_star": X.copy(), "q_star": q, "f_star": f.copy(), "p_star": p.copy(), "td": td}


def save_sweep_mat(path, xs, N):
    """xs [n, nx] -> MAT v5 file with the 1 x n cell array `opt_sol` the reference's analysis scripts load."""
    import scipy.io as sio
    cell = np.empty((1, len(xs)), dtype=object)
    for i, x in enumerate(xs):
        cell[0, i] = opt_sol(x, N)
    sio.savemat(path, {"opt_sol": cell})


def load_sweep_mat(path):
    import scipy.io as sio
    sols = np.atleast_1d(sio.loadmat(path, squeeze_me=True, struct_as_record=False)["opt_sol"])
    return [{k: np.asarray(getattr(s, k), dtype=np.float64) for k in ("X_star", "q_star", "f_star", "p_star", "td")} for s in sols]


def reference_sweep(fixed="pitch_0", sweep="vX"):
    """Drop conditions [n, 12] of a stored sweep `data/<fixed>_<sweep>.mat`: base at 0.6 m, the fixed attitude in degrees
    (`pitch_30`, `roll_60`, `pitch_0`), and one swept axis: vX / vY in -1.5:0.25:1.5 at v_z = -3 (`foot_positions.m:20,32-33`),
    vZ in 0:-0.5:-9."""
    axis, deg = fixed.split("_")
    d0 = np.zeros(12)
    d0[2] = 0.6
    d0[{"roll": 3, "pitch": 4, "yaw": 5}[axis]] = np.deg2rad(float(deg))
    if sweep in ("vX", "vY"):
        vals = np.arange(-1.5, 1.5 + 1e-9, 0.25)
        d = np.tile(d0, (len(vals), 1))
        d[:, 9 if sweep == "vX" else 10] = vals
        d[:, 11] = -3.0
    elif sweep == "vZ":
        vals = np.arange(0.0, -9.0 - 1e-9, -0.5)
        d = np.tile(d0, (len(vals), 1))
        d[:, 11] = vals
    else:
        raise ValueError("sweep must be vX, vY or vZ")
    return d


def _rpy_to_rot(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])  # body -> world, ZYX (rpyToRotMat.m:2)


def touchdown_feet_body(sol):
    """[3, 4]: foot of each leg relative to the CoM, in the body frame, at that leg's touchdown knot."""
    out = np.full((3, 4), np.nan)
    for leg in range(4):
        k = int(sol["td"][leg])
        if k <= 0:
            continue
        k = min(k, sol["p_star"].shape[1]) - 1
        R = _rpy_to_rot(sol["q_star"][3:6, k])
        out[:, leg] = R.T @ (sol["p_star"][3 * leg:3 * leg + 3, k] - sol["q_star"][0:3, k])
    return out


def training_record(drop, x):
    """(input [9], output [nx]) of one sample of the NN training set."""
    drop = np.asarray(drop, dtype=np.float64)
    return np.concatenate([drop[3:6], drop[6:12]]), np.asarray(x, dtype=np.float64).copy()
