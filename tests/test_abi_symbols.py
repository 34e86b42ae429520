"""CPU tests: the C-ABI library loads, exports every declared symbol, and answers the
metadata queries (sizes, names, sparsity) exactly like the reference's generated C -- no compute
call is made (no GPU here)."""
import ctypes
import os
import re

import numpy as np

import landing_controller_b200 as lc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNCS = ("nlp", "nlp_f", "nlp_g", "nlp_grad", "nlp_grad_f", "nlp_hess_l", "nlp_jac_g")
SUFFIXES = ("", "_alloc_mem", "_init_mem", "_free_mem", "_checkout", "_release", "_incref", "_decref",
            "_n_in", "_n_out", "_default_in", "_name_in", "_name_out", "_sparsity_in", "_sparsity_out", "_work")


def test_every_declared_symbol_is_exported(product_lib):
    hdr = open(os.path.join(ROOT, "include", "landing_b200.h")).read()
    names = set(re.findall(r"\b(landing_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(product_lib, n), n
    for f in FUNCS:
        for s in SUFFIXES:
            assert hasattr(product_lib, f + s), f + s
    dropin = ctypes.CDLL(lc.DROPIN_PATH)
    assert hasattr(dropin, "nlp_jac_g") and hasattr(dropin, "nlp_hess_l_sparsity_out")


def test_metadata_equals_reference(product_lib, golden):
    from oracle_lib import CasadiLib
    os.environ.pop("LANDING_B200_KNOTS", None)
    mine = CasadiLib(lc.LIB_PATH)
    meta = {}
    for line in golden["meta"]:
        fn, ins, outs = str(line).split("|")
        meta[fn] = (ins.split(","), outs.split(","))
    for fn, (ins, outs) in meta.items():
        assert getattr(mine.lib, fn + "_n_in")() == len(ins)
        assert getattr(mine.lib, fn + "_n_out")() == len(outs)
        assert [getattr(mine.lib, fn + "_name_in")(i).decode() for i in range(len(ins))] == ins
        assert [getattr(mine.lib, fn + "_name_out")(i).decode() for i in range(len(outs))] == outs
        sz = [ctypes.c_longlong() for _ in range(4)]
        assert getattr(mine.lib, fn + "_work")(*[ctypes.byref(s) for s in sz]) == 0
        assert (sz[0].value, sz[1].value) == (len(ins), len(outs))
    assert np.array_equal(mine.sparsity("nlp_jac_g", 1), golden["spJ"])
    assert np.array_equal(mine.sparsity("nlp_hess_l", 0), golden["spH"])
    sx = mine.sparsity("nlp", 0, out=False)
    assert sx[0] == 732 and sx[1] == 1 and sx[3] == 732 and np.array_equal(sx[4:], np.arange(732))
    assert mine.sparsity("nlp", 1, out=False)[0] == 354
    assert mine.sparsity("nlp_g", 0)[0] == 2092
    assert mine.sparsity("nlp_f", 0)[0] == 1


def test_sizes_and_patterns_generic_n(product_lib):
    from oracle_lib import Oracle
    for N in (21, 30, 50):
        d = lc.dims_for(N, product_lib)
        o = Oracle(N)
        assert (d["nx"], d["np"], d["m"], d["nnzJ"], d["nnzH"]) == (o.nx, o.np_, o.m, o.nnzJ, o.nnzH)
        assert np.array_equal(lc.sparsity_for(N, 0, product_lib), o.spJ)
        assert np.array_equal(lc.sparsity_for(N, 1, product_lib), o.spH)


def test_no_gpu_fails_loudly(product_lib):
    import torch
    if torch.cuda.is_available():
        return
    ctx = ctypes.c_void_p()
    rc = product_lib.landing_create(21, 0, ctypes.byref(ctx))
    assert rc != 0 and b"no CUDA device" in product_lib.landing_last_error()
