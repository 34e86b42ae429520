"""Fixture generator (run in the build container, where /root/reference exists): the stored solutions of the reference's
kino-dynamic (KNITRO) landing NLP -> tests/golden/kino_n21.npz.  The reference keeps three copies of prevSoln.mat; the
two that hold a primal-dual point (X_star, jpos_star, U_star, lam_g_star) are taken."""
import os
import numpy as np
import scipy.io as sio

REF = "/root/reference/optimizations/landing"
out = {}
for tag, sub in (("gs", "generate_solver"), ("ms", "main_scripts")):
    d = sio.loadmat(os.path.join(REF, sub, "prevSoln.mat"), squeeze_me=True)
    X, U, J, lam = d["X_star"], d["U_star"], d["jpos_star"], d["lam_g_star"]
    out[tag + "_x"] = np.concatenate([X.T.reshape(-1), J.T.reshape(-1), U.T.reshape(-1)])  # [X(:); jpos(:); U(:)]
    out[tag + "_lam_g"] = np.asarray(lam, dtype=np.float64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kino_n21.npz"), **out)
print({k: v.shape for k, v in out.items()})
