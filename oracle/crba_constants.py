"""oracle/crba_constants.py -- TEST INFRASTRUCTURE (CPU oracle).

numpy restatement of how the reference obtains the NLP parameters mass, Ib, Ib_inv:
  generate_landingCtrller_IPOPT.m:99-104,222-224  ->  get_mass_matrix(model, q_home, 0)
  utilities_general/dynamics-utilities/get_mass_matrix.m:19-54       (composite inertia, CRBA)
  utilities_general/dynamics-utilities/get_robot_params.m:50-115     ('mc3D')
  utilities_general/dynamics-utilities/get_robot_model.m:134-244,852-889 ('quad3D', flipAlongAxis)
  utilities_general/dynamics-utilities/spatialInertia.m:20-25, spatial_v2/dynamics/jcalc.m:19-28,
  spatial_v2/spatial/plux.m:14-16, rotx/roty.m, spatial_v2/3D/{rz,skew}.m
Pinned by generate_data/data/data_stats.mat (mass = 8.251999999999999) and used by the tests to check the
constants compiled into landing_problem_default / srb_problem_default.
"""
import numpy as np


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def unskew(A):
    return 0.5 * np.array([A[2, 1] - A[1, 2], A[0, 2] - A[2, 0], A[1, 0] - A[0, 1]])


def spatial_inertia(m, com, I3):
    c = skew(com)
    return np.block([[I3 + m * (c @ c.T), m * c], [m * c.T, m * np.eye(3)]])


def plux(E, r):
    return np.block([[E, np.zeros((3, 3))], [-E @ skew(r), E]])


def rx(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])


def ry(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])


def rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])


def rot6(E):
    return np.block([[E, np.zeros((3, 3))], [np.zeros((3, 3)), E]])


def flip_along_y(I_in):
    h = unskew(I_in[0:3, 3:6])
    Ibar = I_in[0:3, 0:3]
    m = I_in[5, 5]
    P = np.zeros((4, 4))
    P[0:3, 0:3] = 0.5 * np.trace(Ibar) * np.eye(3) - Ibar
    P[0:3, 3] = h
    P[3, 0:3] = h
    P[3, 3] = m
    X = np.diag([1.0, -1.0, 1.0, 1.0])
    P = X @ P @ X
    m, h, E = P[3, 3], P[0:3, 3], P[0:3, 0:3]
    out = np.eye(6)
    out[0:3, 0:3] = np.trace(E) * np.eye(3) - E
    out[0:3, 3:6] = skew(h)
    out[3:6, 0:3] = skew(h).T
    out[3:6, 3:6] = m * np.eye(3)
    return out


def composite_inertia(q_leg=(0.0, -1.45, 2.65)):
    # get_robot_params('mc3D')
    body = spatial_inertia(3.3, [0, 0, 0], 1e-6 * np.diag([11253.0, 36203.0, 42673.0]))
    abad = spatial_inertia(0.54, [0, 0.036, 0], 1e-6 * np.array([[381, 58, 0.45], [58, 560, 0.95], [0.45, 0.95, 444]]))
    hip = spatial_inertia(0.634, [0, 0.016, -0.02], 1e-6 * np.array([[1983, 245, 13], [245, 2103, 1.5], [13, 1.5, 408]]))
    knee = spatial_inertia(0.064, [0, 0, -0.061], 1e-6 * np.diag([6.0, 248.0, 245.0]))
    abad_loc = np.array([0.19 * 2, 0.049 * 2, 0]) * 0.5
    hip_loc = np.array([0, 0.062, 0])
    knee_loc = np.array([0, 0, -0.209])
    side = np.array([[1, 1, -1, -1], [-1, 1, -1, 1], [1, 1, 1, 1]], dtype=float)
    Ic = body.copy()
    leg_side = -1
    for leg in range(4):
        sg = side[:, leg]
        inert = [abad, hip, knee]
        if leg_side < 0:
            inert = [flip_along_y(I) for I in inert]
        Xtree = [plux(np.eye(3), sg * abad_loc),
                 plux(rz(np.pi), np.zeros(3)) @ plux(np.eye(3), sg * hip_loc),
                 plux(np.eye(3), sg * knee_loc)]
        XJ = [rot6(rx(q_leg[0])), rot6(ry(q_leg[1])), rot6(ry(q_leg[2]))]
        Xup = [XJ[i] @ Xtree[i] for i in range(3)]
        IC = [I.copy() for I in inert]
        IC[1] += Xup[2].T @ IC[2] @ Xup[2]
        IC[0] += Xup[1].T @ IC[1] @ Xup[1]
        Ic += Xup[0].T @ IC[0] @ Xup[0]
        leg_side = -leg_side
    return Ic


def zero_configuration():
    """H = get_mass_matrix(model, zeros(18,1), 0) of srbm-utilities/generateVariationalDynamics.m:4-7:
    full 3x3 body inertia and mass with straight legs (the TVLQR post-pass uses these, not the q_home values)."""
    Ic = composite_inertia((0.0, 0.0, 0.0))
    return dict(mass=Ic[5, 5], Ib3=Ic[0:3, 0:3].copy())


def constants():
    Ic = composite_inertia()
    I3 = Ic[0:3, 0:3]
    return dict(mass=Ic[5, 5], Ib=np.diag(I3).copy(), Ib_inv=np.diag(np.linalg.inv(I3)).copy(), Ic=Ic)


if __name__ == "__main__":
    c = constants()
    np.set_printoptions(precision=12)
    print("mass", repr(c["mass"]), "Ib", c["Ib"], "Ib_inv", c["Ib_inv"], "Ic13", c["Ic"][0, 2])
