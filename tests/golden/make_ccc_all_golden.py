"""Generates tests/golden/ccc_n41_all.npz: ALL valid stored IPOPT solutions of the reference's N = 41 "CCC" landing problem
(/root/reference/optimizations/landing/data/*.mat, 294 runs whose stored initial height is the swept 0.6 m), reduced to
what the soft known-answer test compares: the drop condition X_star(:,1), the stored cost (generate_quadruped_SRBM_CCC.m:
80-91 evaluated on X_star / f_star), the touchdown knots td, the terminal state X_star(:,41) and the vertical GRF
profiles.  (tests/golden/ccc_n41.npz keeps 43 of the runs in full for the row-level feasibility checks.)

Run in the build container only:  python tests/golden/make_ccc_all_golden.py"""
import glob
import os
import sys

import numpy as np
import scipy.io as sio

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ccc_problem as ccc  # noqa: E402

SRC = "/root/reference/optimizations/landing/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ccc_n41_all.npz")

drops, cost, td, term, fz, names = [], [], [], [], [], []
for path in sorted(glob.glob(os.path.join(SRC, "*.mat"))):
    sols = np.atleast_1d(sio.loadmat(path, squeeze_me=True, struct_as_record=False)["opt_sol"])
    for i, s in enumerate(sols):
        if s.X_star.shape != (12, 41) or abs(s.X_star[2, 0] - 0.6) > 1e-9:
            continue  # failed solves store a different initial height
        drops.append(s.X_star[:, 0])
        cost.append(ccc.stored_cost(s.X_star, s.f_star))
        td.append(np.asarray(s.td, dtype=np.float64))
        term.append(s.X_star[:, -1])
        fz.append(s.f_star[2::3])
        names.append("%s[%d]" % (os.path.basename(path), i))
np.savez_compressed(OUT, drops=np.array(drops), cost=np.array(cost), td=np.array(td), term=np.array(term),
                    fz=np.array(fz), names=np.array(names))
print("wrote", OUT, len(names), "solutions", os.path.getsize(OUT), "bytes")
