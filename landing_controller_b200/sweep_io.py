"""Data formats on either side of the solve, as the reference's sweep callers use them.

* `opt_sol(x, N)` -- the per-scenario record `analysis/foot_positions.m:31-44` stores: X_star [12 x N], p_star
  [12 x (N-1)] (foot positions), f_star [12 x (N-1)] (GRFs), td [4] (first knot with f_z > 1, 1-based like MATLAB's find;
  0 if the leg never loads), q_star [18 x N] (base pose + the home leg angles `eval_SRBM_CCC.m:111`, no IK).
* `save_sweep_mat / load_sweep_mat` -- `data/<fixed>_<sweep>.mat` with a 1 x n cell `opt_sol` (`foot_positions.m:47-50`).
* `reference_sweep(fixed, sweep)` -- drop conditions of the stored sweeps (`foot_positions.m:13-33`).
* `touchdown_feet_body(sol)` -- feet relative to the CoM in the body frame at touchdown (`foot_positions.m:54-62`).
* `training_record(drop, x)` -- `training_data.input = [rpy0; qd0]` / `.output = x*`
  (`generate_data/generate_training_data_automated.m:209-214`; the SRB stage has no joint angles).
* `normalize_training_set / denormalize_sample` -- the NN training-set normalisation of
  `generate_data/data_normalization.m:42-111` and its inverse `data_denormalization.m:16-41`: z-scores of the inputs, the
  states, the foot positions (and joint angles when present); ground-reaction forces shifted to each leg's touchdown knot
  and scaled by body weight; the four touchdown indices appended to every output column.
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


def _blocks(N, n_out):
    nX, nU = 12 * N, 24 * (N - 1)
    nJ = n_out - nX - nU
    if nJ not in (0, 12 * (N - 1)):
        raise ValueError("output columns must hold [X(:); U(:)] or [X(:); U(:); jpos(:)]")
    return nX, nU, nJ


def normalize_training_set(inputs, outputs, N, mass):
    """inputs [9, n] (`training_data.input`), outputs [nx (+ 12(N-1)), n] (`training_data.output`, columns
    [X(:); U(:); jpos(:)], MATLAB column-major reshapes) -> (normalized, stats) after data_normalization.m:42-111.
    normalized["input"] [9, n]; normalized["output"] [n_out + 4, n] (the four touchdown indices appended, 1-based);
    stats has the fields of `data_stats.mat` (mean/std of input, X, U, jpos; td_scale; mass).  std is MATLAB's std(.,0,2)
    (n-1 in the denominator); entries with zero spread (x, y of the first knot: X_norm(1:2,1) = 0, :100) give 0."""
    inputs = np.asarray(inputs, dtype=np.float64)
    outputs = np.asarray(outputs, dtype=np.float64)
    n = inputs.shape[1]
    nX, nU, nJ = _blocks(N, outputs.shape[0])
    mean_in, std_in = inputs.mean(axis=1, keepdims=True), inputs.std(axis=1, ddof=1, keepdims=True)
    mean_out, std_out = outputs.mean(axis=1), outputs.std(axis=1, ddof=1)
    col = lambda v, r, c: v.reshape(c, r).T  # MATLAB reshape(v, [r c]) of a column-major vector
    stats = {"mean_input": mean_in, "std_input": std_in,
             "mean_X": col(mean_out[:nX], 12, N), "std_X": col(std_out[:nX], 12, N),
             "mean_U": col(mean_out[nX:nX + nU], 24, N - 1), "std_U": col(std_out[nX:nX + nU], 24, N - 1),
             "td_scale": 1.0, "mass": float(mass)}
    if nJ:
        stats["mean_jpos"] = col(mean_out[nX + nU:], 12, N - 1)
        stats["std_jpos"] = col(std_out[nX + nU:], 12, N - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.zeros((outputs.shape[0] + 4, n))
        for e in range(n):
            o = outputs[:, e]
            X, U = col(o[:nX], 12, N), col(o[nX:nX + nU], 24, N - 1)
            f = U[12:]
            Un = np.zeros_like(U)
            td = np.zeros(4)
            for leg in range(4):
                fl = f[3 * leg:3 * leg + 3]
                hit = np.nonzero(fl[2] > 1)[0]
                if len(hit) == 0:
                    raise ValueError("sample %d: leg %d never loads (f_z > 1 N): the reference's find() would be empty" % (e, leg))
                t0 = int(hit[0])  # 0-based; MATLAB's td_idx = t0 + 1
                off = np.hstack([fl[:, t0:], np.repeat(fl[:, -1:], t0, axis=1)])
                Un[12 + 3 * leg:12 + 3 * leg + 3] = off / (mass * 9.81)
                td[leg] = t0 + 1
            Xn = (X - stats["mean_X"]) / stats["std_X"]
            Xn[0:2, 0] = 0.0
            Un[:12] = (U[:12] - stats["mean_U"][:12]) / stats["std_U"][:12]
            parts = [Xn.T.ravel(), Un.T.ravel()]
            if nJ:
                parts.append(((col(o[nX + nU:], 12, N - 1) - stats["mean_jpos"]) / stats["std_jpos"]).T.ravel())
            out[:, e] = np.concatenate(parts + [td])
        normalized = {"input": (inputs - mean_in) / std_in, "output": np.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0)}
    return normalized, stats


def denormalize_sample(nn_data, stats, N):
    """One normalised output column -> (X [12, N], U [24, N-1], jpos [12, N-1] or None), data_denormalization.m:16-41: the
    z-scores are undone, and each leg's normalised force profile is shifted back to its touchdown knot (zeros before
    it) and multiplied by body weight."""
    nn_data = np.asarray(nn_data, dtype=np.float64)
    nX, nU, nJ = _blocks(N, nn_data.shape[0] - 4)
    col = lambda v, r, c: v.reshape(c, r).T
    Xn, Un = col(nn_data[:nX], 12, N), col(nn_data[nX:nX + nU], 24, N - 1)
    td = np.trunc(nn_data[-4:]).astype(int)  # int8(td_nn)
    X = Xn * stats["std_X"] + stats["mean_X"]
    U = np.zeros((24, N - 1))
    U[:12] = Un[:12] * stats["std_U"][:12] + stats["mean_U"][:12]
    for leg in range(4):
        fo = Un[12 + 3 * leg:12 + 3 * leg + 3]
        t0 = max(int(td[leg]) - 1, 0)
        U[12 + 3 * leg:12 + 3 * leg + 3] = np.hstack([np.zeros((3, t0)), fo[:, :N - 1 - t0]]) * (stats["mass"] * 9.81)
    jpos = None
    if nJ:
        jpos = col(nn_data[nX + nU:nX + nU + nJ], 12, N - 1) * stats["std_jpos"] + stats["mean_jpos"]
    return X, U, jpos
