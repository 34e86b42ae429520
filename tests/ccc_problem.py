"""The "CCC" landing problem whose IPOPT solutions the reference stores (optimizations/landing/data/*.mat):
generate_quadruped_SRBM_CCC.m with the parameter values of analysis/eval_SRBM_CCC.m:21-67 (the values live in
landing_controller_b200/sweeps.py:apply_ccc_parameters).  Helpers for the known-answer tests."""
import numpy as np

from landing_controller_b200.sweeps import (CCC_KIN_BOX as KIN_BOX, CCC_N as N, CCC_QF as QF, CCC_QN as QN,  # noqa: F401
                                            CCC_Q_TERM_REF as Q_TERM_REF, apply_ccc_parameters as fill_problem)


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
