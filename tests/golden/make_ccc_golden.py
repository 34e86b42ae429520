"""Generates tests/golden/ccc_n41.npz from the REFERENCE's stored IPOPT solutions
(/root/reference/optimizations/landing/data/*.mat: N = 41 "CCC" problem, generate_quadruped_SRBM_CCC.m,
solved through analysis/eval_SRBM_CCC.m; fields X_star[12x41], p_star[12x40] (foot positions),
f_star[12x40] (GRFs), td[4]).  These are the only outputs of real IPOPT runs the reference keeps
(SURVEY.md 8c); they pin the restated dynamics / contact rows and the CRBA constants at an N other
than the 21 knots of the generated C.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_ccc_golden.py
"""
import glob
import os

import numpy as np
import scipy.io as sio

SRC = "/root/reference/optimizations/landing/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ccc_n41.npz")


def main():
    X, C, F, TD, names = [], [], [], [], []
    for path in sorted(glob.glob(os.path.join(SRC, "*.mat"))):
        sols = np.atleast_1d(sio.loadmat(path, squeeze_me=True, struct_as_record=False)["opt_sol"])
        # two solutions per sweep file: the first and the middle one
        for i in sorted({0, len(sols) // 2}):
            s = sols[i]
            if s.X_star.shape != (12, 41) or abs(s.X_star[2, 0] - 0.6) > 1e-9:
                continue  # failed solves store a different initial height
            if path.endswith("pitch_45_vZ.mat") and i == 9:
                continue  # stored iterate of an unconverged run (linear z-row residual 1.7e-3)
            X.append(s.X_star)
            C.append(s.p_star)
            F.append(s.f_star)
            TD.append(np.asarray(s.td, dtype=np.float64))
            names.append("%s[%d]" % (os.path.basename(path), i))
    np.savez_compressed(OUT, X=np.array(X), c=np.array(C), f=np.array(F), td=np.array(TD), names=np.array(names))
    print("wrote", OUT, len(names), "solutions", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
