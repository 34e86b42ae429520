"""Generates tests/golden/ref_n21.npz from the REFERENCE's own compiled C
(/root/reference/optimizations/landing/codegen_casadi/landingCtrller_IPOPT.c built into
oracle/_ref/landingCtrller_IPOPT.so by `make -C oracle ref`, N = 21).

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
The committed .npz is what the GPU-box tests compare against.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle_lib import CasadiLib, REF_SO  # noqa: E402


def kat0():
    """SURVEY.md Appendix B.3 input."""
    x = np.zeros(732)
    p = np.zeros(354)
    x[2] = 0.5
    x[254] = 0.3
    x[264:267] = (1.0, 0.5, 2.0)
    x[276:279] = (0.1, 0.2, 0.7)
    p[252:272] = 0.03
    p[344:348] = (1, 0.35, 200, 8.252)
    p[348:351] = (0.0576, 0.234, 0.28)
    p[351:354] = (17.4, 4.27, 3.58)
    return x, p


def main():
    ref = CasadiLib(REF_SO)
    out = {}
    out["spJ"] = ref.sparsity("nlp_jac_g", 1)
    out["spH"] = ref.sparsity("nlp_hess_l", 0)
    meta = []
    for fn in ref.FUNCS:
        n_in = getattr(ref.lib, fn + "_n_in")()
        n_out = getattr(ref.lib, fn + "_n_out")()
        ins = [getattr(ref.lib, fn + "_name_in")(i).decode() for i in range(n_in)]
        outs = [getattr(ref.lib, fn + "_name_out")(i).decode() for i in range(n_out)]
        meta.append("%s|%s|%s" % (fn, ",".join(ins), ",".join(outs)))
    out["meta"] = np.array(meta)
    rng = np.random.default_rng(20261017)
    cases = []
    x, p = kat0()
    cases.append((x, p, 1.0, np.ones(2092)))
    for t in range(5):
        scale = (0.05, 0.3, 0.6, 1.0, 0.2)[t]
        x = rng.normal(size=732) * scale
        x[2::12][:21] += 0.5
        p = rng.uniform(0.5, 1.5, size=354)
        p[252:272] = rng.uniform(0.01, 0.05, size=20)
        cases.append((x, p, float(rng.normal()), rng.normal(size=2092)))
    for i, (x, p, lf, lam) in enumerate(cases):
        lfa = np.array([lf])
        _, (f, g) = ref.call("nlp", [x, p])
        _, (g2, J) = ref.call("nlp_jac_g", [x, p])
        _, (H,) = ref.call("nlp_hess_l", [x, p, lfa, lam])
        _, (f2, gf) = ref.call("nlp_grad_f", [x, p])
        _, (f3, g3, gx, gp) = ref.call("nlp_grad", [x, p, lfa, lam])
        assert np.array_equal(g, g2) and np.array_equal(g, g3)
        for k, v in dict(x=x, p=p, lam_f=lfa, lam_g=lam, f=f, g=g, J=J, H=H, gf=gf, gx=gx, gp=gp).items():
            out["c%d_%s" % (i, k)] = v
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(__file__), "ref_n21.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
