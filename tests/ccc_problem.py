"""The "CCC" landing problem whose IPOPT solutions the reference stores (optimizations/landing/data/*.mat):
generate_quadruped_SRBM_CCC.m with the parameter values of analysis/eval_SRBM_CCC.m:21-67.  Same rows as the hot-path
NLP; differences: N = 41 (dt = 0.015), running GRF cost Qf = (1e-4, 1e-4, 1e-3) (QX = Qc = 0), kinematic box
0.05 / 0.05 / 0.27, f_max = 250, q_term_ref z = 0.2, c_ref = (+-0.2, +-0.1, -0.35), velocity bounds +-40."""
import numpy as np

N = 41
QF = (1e-4, 1e-4, 1e-3)
KIN_BOX = (0.05, 0.05, 0.27)
QN = (0, 0, 100, 100, 100, 0, 10, 10, 10, 10, 10, 10)
Q_TERM_REF = (0, 0, 0.2, 0, 0, 0)


def fill_problem(pb):
    """pb: landing_problem / srb_problem ctypes structure (shared field names)."""
    def put(name, vals):
        a = getattr(pb, name)
        for i, v in enumerate(vals):
            a[i] = v
    pb.T = 0.6
    put("q_min", [-10, -10, 0.15, -10, -10, -10]); put("q_max", [10, 10, 1.0, 10, 10, 10])
    put("qd_min", [-10, -10, -10, -40, -40, -40]); put("qd_max", [10, 10, 10, 40, 40, 40])
    put("q_term_min", [-10, -10, 0.15, -0.1, -0.1, -10]); put("q_term_max", [10, 10, 5, 0.1, 0.1, 10])
    put("qd_term_min", [-10, -10, -10, -40, -40, -40]); put("qd_term_max", [10, 10, 10, 40, 40, 40])
    put("q_term_ref", Q_TERM_REF); put("qd_term_ref", [0] * 6)
    put("QN", QN)
    side = np.array([1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1], dtype=float)
    put("c_ref", side * np.tile([0.2, 0.1, -0.35], 4))
    pb.mu, pb.l_leg_max, pb.f_max = 1.0, 0.35, 250.0
    if hasattr(pb, "Qf"):  # the product's landing_problem carries the variant data itself
        put("Qf", QF); put("kin_box", KIN_BOX)
    return pb


def stored_cost(X, F):
    """The CCC objective (generate_quadruped_SRBM_CCC.m:80-91) of a stored solution X_star [12,41], f_star [12,40]."""
    h = 0.6 / (N - 1)
    term = sum(QN[i] * (X[i, -1] - (Q_TERM_REF[i] if i < 6 else 0.0)) ** 2 for i in range(12))
    return term + h * float(np.sum(np.tile(QF, 4)[:, None] * F ** 2))


def split(x):
    """x [nx] -> X [12,N], c [12,N-1], f [12,N-1]."""
    X = x[:12 * N].reshape(N, 12).T
    U = x[12 * N:].reshape(N - 1, 24).T
    return X, U[:12], U[12:]


def touchdown(f):
    """First knot (1-based, as MATLAB's find) with f_z > 1 per leg (analysis/foot_positions.m:36-37); 0 = never."""
    return [int(np.argmax(f[3 * l + 2] > 1)) + 1 if (f[3 * l + 2] > 1).any() else 0 for l in range(4)]
