"""oracle/tvlqr_ref.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the time-varying LQR pass the reference runs on a solved landing trajectory
(optimizations/landing/quadruped_SRBM_NLP.m:428-497): variational single-rigid-body dynamics A(t), B(t)
(utilities_general/srbm-utilities/generateVariationalDynamics.m:9-62) and the backward integration of the Riccati
differential equation Pdot = A'P + PA - P B R^-1 B'P + Q with explicit Euler steps
(generateRiccatiIntegrator.m:24,49-53: "P0 = Pf + dt*k1").  PARITY UNPINNED: the reference stores no output of this pass
and MATLAB/CasADi are not available; the CUDA kernel is checked against this file, and this file against the defining
properties (symmetry, positive semi-definiteness, Riccati residual, a finite-difference check of A and B).
"""
import numpy as np


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def rpy_to_rot(rpy):
    """rpyToRotMat.m:2 = rz(yaw)' ry(pitch)' rx(roll)' (body -> world)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def variational_AB(xd, ud, Ib3, mass):
    """A [24,24], B [24,12] at the reference point xd = [p; rpy; omega; v; pf(12)], ud = GRFs (12).
    generateVariationalDynamics.m:32-55, literally: R = rpyToRotMat(rpy)' and the expressions use R'."""
    p, rpy, om = xd[0:3], xd[3:6], xd[6:9]
    pf = xd[12:24].reshape(4, 3)
    f = ud.reshape(4, 3)
    Rt = rpy_to_rot(rpy)  # = R'
    Ibi = np.linalg.inv(Ib3)
    A = np.zeros((24, 24))
    B = np.zeros((24, 12))
    A[0:3, 9:12] = np.eye(3)
    A[3:6, 3:6] = -skew(om)
    A[3:6, 6:9] = np.eye(3)
    tau = sum(Rt @ skew(pf[l] - p) @ f[l] for l in range(4))
    A[6:9, 3:6] = Ibi @ skew(tau)
    A[6:9, 0:3] = Ibi @ Rt @ skew(f.sum(axis=0))
    A[6:9, 6:9] = Ibi @ (skew(Ib3 @ om) - skew(om) @ Ib3)
    for l in range(4):
        A[6:9, 12 + 3 * l:15 + 3 * l] = -Ibi @ Rt @ skew(f[l])
        B[6:9, 3 * l:3 * l + 3] = Ibi @ Rt @ skew(pf[l] - p)
        B[9:12, 3 * l:3 * l + 3] = np.eye(3) / mass
    A[12:24, 12:24] = -0.00001 * np.eye(12)
    return A, B


def sample_reference(X, U, t_star, t_int):
    """quadruped_SRBM_NLP.m:478-487: state interpolated between the bracketing knots, feet and forces of knot k_opt."""
    N = X.shape[1]
    k = 0
    while t_int > t_star[k + 1] and k < N - 2:
        k += 1
    a = (t_star[k + 1] - t_int) / (t_star[k + 1] - t_star[k])
    xd = a * np.concatenate([X[:, k], U[0:12, k]]) + (1 - a) * np.concatenate([X[:, k + 1], U[0:12, k]])
    return xd, U[12:24, k].copy()


def riccati_backward(X, U, T, Q, Rdiag, F, dt, n_steps, Ib3, mass):
    """P [n_steps,24,24] with P[n_steps-1] = F and P[k-1] = P[k] + dt * Pdot(P[k]) at t = k dt (0-based k), and the
    feedback gains K[k] = R^-1 B(t_k)' P[k] [n_steps,12,24]."""
    N = X.shape[1]
    t_star = np.arange(N) * (T / (N - 1))
    P = np.zeros((n_steps, 24, 24))
    K = np.zeros((n_steps, 12, 24))
    P[n_steps - 1] = F
    Rinv = 1.0 / np.asarray(Rdiag)
    for k in range(n_steps - 1, -1, -1):
        xd, ud = sample_reference(X, U, t_star, k * dt)
        A, B = variational_AB(xd, ud, Ib3, mass)
        S = P[k] @ B
        K[k] = (Rinv[:, None] * S.T)
        if k > 0:
            Pdot = A.T @ P[k] + P[k] @ A - (S * Rinv[None, :]) @ S.T + Q
            P[k - 1] = P[k] + dt * Pdot
    return P, K


def default_weights():
    """quadruped_SRBM_NLP.m:441-466,474-475: F = diag(1,1,1, 5,5,5, 4,4,4, 3,3,3, 0...), Q = diag(0.25 x3, 1 x3, 0.5 x3,
    1 x3, 0...), R = 90 I, dt = 0.022."""
    F = np.zeros((24, 24))
    F[np.arange(12), np.arange(12)] = [1, 1, 1, 5, 5, 5, 4, 4, 4, 3, 3, 3]
    Q = np.zeros((24, 24))
    Q[np.arange(12), np.arange(12)] = [0.25] * 3 + [1.0] * 3 + [0.5] * 3 + [1.0] * 3
    return Q, np.full(12, 90.0), F, 0.022
