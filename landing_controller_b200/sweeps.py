"""Synthetic drop-condition sweeps (the workloads of BASELINE.json `configs`).

A drop condition is the 12-vector (q_init[6] = x,y,z,roll,pitch,yaw ; qd_init[6] = omega_body, v_world),
the quantity the reference's sweep scripts draw per sample
(optimizations/landing/generate_data/generate_training_data_automated.m:44-60,
 optimizations/landing/analysis/foot_positions.m:18-44).
"""
import numpy as np


def single_drop():
    """BASELINE config 0: 0.5 m, level, 1 m/s forward."""
    d = np.zeros((1, 12))
    d[0, 2] = 0.5
    d[0, 9] = 1.0
    return d


def _factor(B):
    """B = nz*np*nr*nv with the 1:2:1:2 proportions of SURVEY 8d (1024 = 4x8x4x8, 16384 = 8x16x8x16)."""
    dims = [1, 1, 1, 1]
    order = [1, 3, 0, 2]
    i = 0
    rem = B
    while rem > 1:
        if rem % 2:
            raise ValueError("grid sweep size must be a power of two")
        dims[order[i % 4]] *= 2
        rem //= 2
        i += 1
    return dims


def grid_sweep(B=1024, vz=-3.0):
    """Tensor grid height x pitch x roll x forward velocity (BASELINE configs 1 and 2)."""
    nz, npi, nr, nv = _factor(B)
    z = np.linspace(0.4, 0.7, nz) if nz > 1 else np.array([0.5])
    pitch = np.linspace(-np.pi / 3, np.pi / 3, npi) if npi > 1 else np.array([0.0])
    roll = np.linspace(-0.25, 0.25, nr) if nr > 1 else np.array([0.0])
    vx = np.linspace(-1.75, 1.75, nv) if nv > 1 else np.array([0.0])
    Z, P, R, V = np.meshgrid(z, pitch, roll, vx, indexing="ij")
    d = np.zeros((B, 12))
    d[:, 2] = Z.ravel()
    d[:, 3] = R.ravel()
    d[:, 4] = P.ravel()
    d[:, 9] = V.ravel()
    d[:, 11] = vz
    return d


# knot spacings of the reference's sweep / MPC callers for N = 21 (generate_training_data_automated.m:28,
# main_scripts/landing_optimization.m:28): dense around touchdown, coarse at the end of the horizon
SWEEP_DT = np.array([0.05] + [0.02] * 15 + [0.05, 0.05, 0.1, 0.2])
SWEEP_N = 21


def random_sweep(B, seed=0, large_tilt=False, dt1=0.05):
    """Seeded random drops after generate_training_data_automated.m:44-60 (BASELINE config 4).
    dt1 = dt_val(1) of the caller, which enters the drop-height rule (:28,58)."""
    rng = np.random.default_rng(seed)
    d = np.zeros((B, 12))
    lim = (np.pi / 2 - 0.1) if large_tilt else np.pi / 3
    d[:, 3] = rng.uniform(-0.25, 0.25, B)
    d[:, 4] = rng.uniform(-lim, lim, B)
    d[:, 5] = rng.uniform(-0.25, 0.25, B)
    d[:, 6:9] = rng.uniform(-0.5, 0.5, (B, 3))
    d[:, 9:11] = rng.uniform(-1.75, 1.75, (B, 2))
    d[:, 11] = rng.uniform(-6.0, -3.0, B)
    # z0 = 0.35 + |min_l (R hip_l)_z| + |dt_val(1)*vz|  (:52-60); hip_z = 0
    sr, cr = np.sin(d[:, 3]), np.cos(d[:, 3])
    sp, cp = np.sin(d[:, 4]), np.cos(d[:, 4])
    hip = np.array([[0.19, -0.1], [0.19, 0.1], [-0.19, -0.1], [-0.19, 0.1]])
    hz = -sp[:, None] * hip[None, :, 0] + (sr * cp)[:, None] * hip[None, :, 1]
    d[:, 2] = 0.35 + np.abs(hz.min(axis=1)) + np.abs(dt1 * d[:, 11])
    return d


def apply_sweep_parameters(pb):
    """Set the numeric parameters the reference's SWEEP callers use (generate_training_data_automated.m:62-102), which
    differ from the generator's defaults (generate_landingCtrller_IPOPT.m:173-196): larger force / leg-length limits
    (the random drops reach v_z = -6 m/s, infeasible with f_max = 200), wider terminal box, other weights.
    `pb` is any object with the landing_problem fields (ctypes structure of the library or of the oracle)."""
    def put(name, vals):
        a = getattr(pb, name)
        for i, v in enumerate(vals):
            a[i] = v
    put("q_term_min", [-10, -10, 0.15, -0.1, -0.1, -10])
    put("q_term_max", [10, 10, 5, 0.1, 0.1, 10])
    put("qd_term_min", [-10, -10, -10, -0.5, -0.5, -0.5])
    put("qd_term_max", [10, 10, 10, 0.5, 0.5, 0.5])
    put("q_min", [-10, -10, 0.075, -10, -10, -10])
    put("q_max", [10, 10, 1.0, 10, 10, 10])
    put("q_term_ref", [0, 0, 0.25, 0, 0, 0])
    put("QN", [0, 0, 100, 10, 10, 0, 10, 10, 10, 10, 10, 10])
    side = np.array([1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1], dtype=float)
    put("c_ref", side * np.tile([0.2, 0.2, -0.3], 4))
    pb.mu = 0.75
    pb.l_leg_max = 0.4
    pb.f_max = 500.0
    return pb


def rot_xyz(rpy):
    """rpyToRotMat_xyz.m:2 -- rx(r)' ry(p)' rz(y)' of spatial_v2 = R_x(r) R_y(p) R_z(y): the convention the sweep callers
    use to place the reference feet (NOT the ZYX convention of the NLP's dynamics)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def sweep_initial_guess(drops, pb, N):
    """x0 = [Xref(:); Uref(:)] exactly as the sweep callers build it (generate_training_data_automated.m:105-136):
    Xref = per-row linspace(init, term_ref, N); reference feet = Xref_pos + R_xyz(Xref_rpy) c_ref (ROTATED with the
    reference attitude, unlike the generator's own test call); forces 0.  Returns [B, 36N-24]."""
    drops = np.asarray(drops, dtype=np.float64)
    B = drops.shape[0]
    qt = np.array([pb.q_term_ref[i] for i in range(6)])
    qdt = np.array([pb.qd_term_ref[i] for i in range(6)])
    c_ref = np.array([pb.c_ref[i] for i in range(12)])
    x0 = np.zeros((B, 36 * N - 24))
    for b in range(B):
        X = np.zeros((12, N))
        for i in range(6):
            X[i] = np.linspace(drops[b, i], qt[i], N)
            X[6 + i] = np.linspace(drops[b, 6 + i], qdt[i], N)
        U = np.zeros((24, N - 1))
        for k in range(N - 1):
            R = rot_xyz(X[3:6, k])
            for leg in range(4):
                U[3 * leg:3 * leg + 3, k] = X[0:3, k] + R @ c_ref[3 * leg:3 * leg + 3]
        x0[b] = np.concatenate([X.T.ravel(), U.T.ravel()])
    return x0


CCC_N = 41
CCC_QF = (1e-4, 1e-4, 1e-3)
CCC_KIN_BOX = (0.05, 0.05, 0.27)
CCC_QN = (0, 0, 100, 100, 100, 0, 10, 10, 10, 10, 10, 10)
CCC_Q_TERM_REF = (0, 0, 0.2, 0, 0, 0)


def apply_ccc_parameters(pb):
    """Set the numeric parameters of the "CCC" landing problem, the one behind the reference's stored IPOPT solutions
    (generate_quadruped_SRBM_CCC.m solved through analysis/eval_SRBM_CCC.m:21-67): N = 41 knots over T = 0.6 s, running
    GRF cost Qf (QX = Qc = 0), kinematic box 0.05 / 0.05 / 0.27, f_max = 250, q_term_ref z = 0.2,
    c_ref = (+-0.2, +-0.1, -0.35), velocity bounds +-40.  `pb`: landing_problem (carries Qf / kin_box itself) or the
    oracle's srb_problem (the variant data then goes into ip_options.run_Qf / kin_box)."""
    def put(name, vals):
        a = getattr(pb, name)
        for i, v in enumerate(vals):
            a[i] = v
    pb.T = 0.6
    put("q_min", [-10, -10, 0.15, -10, -10, -10])
    put("q_max", [10, 10, 1.0, 10, 10, 10])
    put("qd_min", [-10, -10, -10, -40, -40, -40])
    put("qd_max", [10, 10, 10, 40, 40, 40])
    put("q_term_min", [-10, -10, 0.15, -0.1, -0.1, -10])
    put("q_term_max", [10, 10, 5, 0.1, 0.1, 10])
    put("qd_term_min", [-10, -10, -10, -40, -40, -40])
    put("qd_term_max", [10, 10, 10, 40, 40, 40])
    put("q_term_ref", CCC_Q_TERM_REF)
    put("qd_term_ref", [0] * 6)
    put("QN", CCC_QN)
    side = np.array([1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1], dtype=float)
    put("c_ref", side * np.tile([0.2, 0.1, -0.35], 4))
    pb.mu, pb.l_leg_max, pb.f_max = 1.0, 0.35, 250.0
    if hasattr(pb, "Qf"):
        put("Qf", CCC_QF)
        put("kin_box", CCC_KIN_BOX)
    return pb
