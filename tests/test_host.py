"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, argument validation / error behaviour of the boundary, the mgcv-structured design
builder and the parameter-vector bookkeeping.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from smoothsde_b200 import _lib as L
from smoothsde_b200 import design as D
from smoothsde_b200 import synth
from smoothsde_b200.engine import Engine, model_code

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            out += re.findall(r"\b(ssde_[a-z_0-9]+)\s*\(", txt)
    return sorted(set(out))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported"
    assert sorted(L.EXPORTS) == syms, "ctypes binding list out of sync with the header"
    assert b"sm_100a" in lib.ssde_version()


def has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    dat, par, _ = synth.make_problem("BM", 1, 20, n_dim=1)
    with pytest.raises(L.EngineError) as e:
        Engine.from_data(dat)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_unknown_and_unbuilt_types():
    with pytest.raises(L.EngineError) as e:
        model_code("XYZ")
    assert e.value.code == 1 and "Unknown SDE type" in str(e.value)      # src/smoothSDE.cpp:25
    with pytest.raises(L.EngineError) as e:
        model_code("ESEAL_SSM")
    assert e.value.code == 3
    lib = L.load()
    d = L.Desc()
    d.model = 17
    d.n_dim = 1
    h = C.c_void_p()
    assert lib.ssde_create(C.byref(d), C.byref(h)) == 1
    assert b"Unknown SDE type" in lib.ssde_create_error()
    assert lib.ssde_create(None, C.byref(h)) == 2
    # null handle is tolerated everywhere
    assert lib.ssde_n_par(None) == -1
    lib.ssde_destroy(None)
    assert lib.ssde_eval(None, None, 0, None, None, None) == 2


def test_bad_shapes_are_rejected_before_touching_the_device():
    lib = L.load()
    dat, par, _ = synth.make_problem("CTCRW", 1, 20, n_dim=3)
    with pytest.raises(L.EngineError) as e:
        Engine.from_data(dat)
    assert e.value.code == 3      # CTCRW n_dim <= 2
    dat, par, _ = synth.make_problem("BM", 1, 20, n_dim=1)
    bad = dict(dat, X_fe=dat["X_fe"][:-1])
    with pytest.raises(L.EngineError) as e:
        Engine.from_data(bad)
    assert e.value.code == 2 and "n_par * n rows" in str(e.value)


def test_design_column_bookkeeping_matches_reference_test():
    """tests/testthat/test_sde.R:53-72: mu ~ s(x1,k=5,bs='ts') + x2, sigma ~ s(ID,bs='re') +
    s(x2,k=5,bs='ts') with 10 IDs -> 3 fixed, 18 random (4 + 10 + 4), 3 lambdas."""
    rng = np.random.default_rng(1)
    n = 100
    data = {"ID": np.repeat(np.arange(10), 10), "x1": rng.normal(size=n), "x2": rng.normal(size=n)}
    des = D.make_design({"mu": "~ s(x1, k = 5, bs = 'ts') + x2",
                         "sigma": "~ s(ID, bs = 're') + s(x2, k = 5, bs = 'ts')"}, data, n)
    assert des.X_fe.shape == (2 * n, 3)
    assert des.X_re.shape == (2 * n, 18)
    assert list(des.ncol_re) == [4, 10, 4]
    assert des.S.shape == (18, 18)
    # block-diagonal by parameter: rows 0..n-1 only touch mu's columns
    assert des.X_fe[:n, 2:].nnz == 0 and des.X_fe[n:, :2].nnz == 0
    assert des.X_re[:n, 4:].nnz == 0 and des.X_re[n:, :4].nnz == 0
    # sum-to-zero constraint absorbed
    assert np.allclose(np.asarray(des.X_re[:n, :4].sum(axis=0)), 0.0, atol=1e-10)


def test_missing_covariate_is_an_error():
    with pytest.raises(KeyError):
        D.make_design({"mu": "~ x9"}, {"ID": np.zeros(5)}, 5)


def test_parameter_layout_of_synthetic_problems():
    dat, par, info = synth.make_problem("CTCRW", 2, 30, n_dim=2)
    assert par.size == 1 + info["p_fe"] + info["n_s"] + info["p_re"] == 1 + 4 + 2 + 18
    dat, par, info = synth.make_problem("OU", 3, 30, n_dim=1)
    assert info["p_fe"] == 3 and info["p_re"] == 2 * (9 + 3) and info["n_s"] == 4


def test_r_shim_type_checks_against_the_c_abi():
    """The R shim cannot be built here (no R): compile it for syntax and types against stub
    declarations of the R API it uses (tests/harness/r_stub) and the real include/smoothsde_b200.h,
    and check that every routine it registers is defined and that it only calls exported symbols."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stub = os.path.join(root, "tests", "harness", "r_stub")
    inc = os.path.join(root, "include")
    shim = os.path.join(root, "r_shim", "src", "shim.cpp")
    init = os.path.join(root, "r_shim", "src", "init.c")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + stub, "-I" + inc, shim])
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I" + stub, init])
    src = open(shim).read()
    registered = re.findall(r'\{"(ssde_\w+)",\s*\(DL_FUNC\)', open(init).read())
    defined = set(re.findall(r"^SEXP (ssde_\w+)\(", src, flags=re.M))
    assert registered and set(registered) <= defined
    from smoothsde_b200 import _lib as L
    called = set(re.findall(r"\b(ssde_\w+)\(", src)) - defined
    called -= {"ssde_handle", "ssde_laplace", "ssde_desc", "ssde_triplet"}
    assert called <= set(L.EXPORTS), called - set(L.EXPORTS)
    # the R adapter only .Call()s registered routines
    rsrc = open(os.path.join(root, "r_shim", "R", "adfun.R")).read()
    assert set(re.findall(r'\.Call\("(\w+)"', rsrc)) <= set(registered)
