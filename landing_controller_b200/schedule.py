"""The fixed-contact-schedule landing problem of the reference (optimizations/landing/quadruped_SRBM_NLP.m), which
BASELINE.json configs[0] names: one drop condition, N knots, the contact state of every leg prescribed per knot."""
import numpy as np

# parameter set of quadruped_SRBM_NLP.m:178-216
SCHED_QX = (10.0,) * 12
SCHED_QN = (0, 0, 100, 10, 10, 100, 10, 10, 10, 10, 10, 10)
SCHED_QF = (1e-4, 1e-4, 1e-3)
SCHED_KIN_BOX = (0.05, 0.05, 0.27)


def apply_schedule_parameters(pb, z_max=1.0):
    """Numeric parameters of quadruped_SRBM_NLP.m:178-216 (`pb`: landing_problem or the oracle's srb_problem): bounds,
    weights, mu = 1, l_leg_max = 0.3, f_max = 200, q_term_ref z = 0.2, c_ref = (+-0.2, +-0.1, -0.2).  The reference's own
    drop starts at 0.35 m under q_max z = 0.4; z_max lifts that bound for higher drops (BASELINE configs[0]: 0.5 m)."""
    def put(name, vals):
        a = getattr(pb, name)
        for i, v in enumerate(vals):
            a[i] = v
    put("q_min", [-10, -10, 0.0, -10, -10, -10])
    put("q_max", [10, 10, z_max, 10, 10, 10])
    put("qd_min", [-10, -10, -10, -40, -40, -40])
    put("qd_max", [10, 10, 10, 40, 40, 40])
    put("q_term_ref", [0, 0, 0.2, 0, 0, 0])
    put("qd_term_ref", [0] * 6)
    put("QN", SCHED_QN)
    side = np.array([1, -1, 1, 1, 1, 1, -1, -1, 1, -1, 1, 1], dtype=float)
    put("c_ref", side * np.tile([0.2, 0.1, -0.2], 4))
    pb.mu, pb.l_leg_max, pb.f_max = 1.0, 0.3, 200.0
    if hasattr(pb, "Qf"):
        put("Qf", SCHED_QF)
        put("kin_box", SCHED_KIN_BOX)
    return pb


def reference_schedule(N=16):
    """cs_val of quadruped_SRBM_NLP.m:33: two flight knots, then all four legs in stance.  [N-1, 4]."""
    cs = np.ones((N - 1, 4), dtype=np.int32)
    cs[:2] = 0
    return cs


def ballistic_schedule(N, T, z0, vz0=0.0, z_touch=0.28):
    """All four legs in flight while the body, falling freely from z0 (explicit Euler steps, as the NLP integrates it),
    is above z_touch -- the hip height at which the feet can be on the ground inside the kinematic box and the leg
    length l_leg_max = 0.3 -- then stance.  [N-1, 4]."""
    dt = T / (N - 1)
    z, v = float(z0), float(vz0)
    cs = np.ones((N - 1, 4), dtype=np.int32)
    for k in range(N - 1):
        if z > z_touch:
            cs[k] = 0
        z, v = z + v * dt, v - 9.81 * dt
    return cs
