"""CPU tests of the device layout built by the C-ABI library (design.cuh): the warp-tile
transposed sliced-ELL design reproduces X %*% theta for every row, uniform and per-nonzero
column storage, ragged row counts, padding."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from smoothsde_b200 import _lib as L
from smoothsde_b200 import synth
from smoothsde_b200.engine import pack_host


def layout_info():
    info = (C.c_int32 * 4)()
    L.load().ssde_layout_info(info)
    return tuple(info)


def alias_of(flags, p):
    """design.cuh alias_of: bits 8+2p, 9+2p of the descriptor flags (t + 1, or 0 = own values)."""
    return ((int(flags) >> (8 + 2 * p)) & 3) - 1


def value_slots(kb, flags):
    """(value slots per row, first value slot of every parameter) -- design.cuh value_slots_of."""
    sv, off = 0, []
    for p in range(4):
        t = alias_of(flags, p)
        if t >= 0:
            off.append(off[t])
        else:
            off.append(sv)
            sv += kb[p]
    return sv, off


def eta_from_pack(pk, n, n_par, theta):
    """Evaluate the packed design exactly the way the kernels index it."""
    LC, WT, PAD, _ = layout_info()
    eta = np.zeros((n, n_par))
    for q, d in enumerate(pk["desc"]):
        kb = [(int(d["kmax"]) >> (8 * p)) & 255 for p in range(4)]
        S = sum(kb)
        SV, voff = value_slots(kb, d["flags"])
        uniform = bool(d["flags"] & 1)
        assert uniform or SV == S
        for lane in range(32):
            for k in range(LC):
                r = q * WT + lane * LC + k
                if r >= n:
                    continue
                j = 0
                for p in range(n_par):
                    for i in range(kb[p]):
                        v = pk["val"][d["val_off"] + (k * SV + voff[p] + i) * 32 + lane]
                        c = pk["col"][d["col_off"] + j] if uniform else pk["col"][d["col_off"] + (k * S + j) * 32 + lane]
                        eta[r, p] += v * theta[c]
                        j += 1
    return eta


def dense_eta(dat, theta, n, n_par):
    X = sp.hstack([sp.csr_matrix(dat["X_fe"]), sp.csr_matrix(dat["X_re"])], format="csr")
    return np.asarray(X @ theta).reshape(n_par, n).T


def test_layout_constants():
    LC, WT, PAD, dsz = layout_info()
    assert WT == 32 * LC and PAD % WT == 0 and dsz == 24
    assert L.load().ssde_padded_rows(1) == PAD and L.load().ssde_padded_rows(PAD + 1) == 2 * PAD


@pytest.mark.parametrize("model,T,m,nd", [("CTCRW", 3, 411, 2), ("OU", 5, 130, 1), ("BM", 1, 300, 2)])
def test_packed_design_reproduces_linear_predictor(model, T, m, nd):
    dat, par, info = synth.make_problem(model, T, m, n_dim=nd, seed=4)
    n = info["n"]
    n_par = dat["X_fe"].shape[0] // n
    pk = pack_host(dat)
    assert pk["n_pad"] % 1024 == 0 and pk["desc"].size == pk["n_pad"] // 256
    rng = np.random.default_rng(0)
    theta = rng.normal(size=info["p_fe"] + info["p_re"])
    assert np.allclose(eta_from_pack(pk, n, n_par, theta), dense_eta(dat, theta, n, n_par), rtol=1e-13, atol=1e-13)
    # dense smooth blocks and along-track random effects -> every warp-tile shares its columns
    assert np.all(pk["desc"]["flags"] & 1)
    # identical consecutive column lists are stored once
    assert pk["col"].size < 4 * pk["desc"].size * 64


def test_parameters_with_the_same_smooth_share_their_value_slots(monkeypatch):
    """tau ~ s(time), nu ~ s(time): identical blocks of X_re in different columns are stored once per
    warp-tile (alias flags) -- 2 x (1 + 9) + 2 column slots, 1 + 10 value slots -- except in decay
    models, whose kernel does not read the alias flags."""
    dat, par, info = synth.make_problem("CTCRW", 2, 300, n_dim=2, seed=5)
    n = info["n"]
    pk = pack_host(dat)
    live = pk["desc"][: (n + 255) // 256]
    for d in live:
        kb = [(int(d["kmax"]) >> (8 * p)) & 255 for p in range(4)]
        assert kb == [1, 1, 10, 10]
        assert [alias_of(d["flags"], p) for p in range(4)] == [-1, 0, -1, 2]
        assert value_slots(kb, d["flags"])[0] == 11
    assert pk["val"].size == 11 * 256 * live.size
    theta = np.random.default_rng(2).normal(size=info["p_fe"] + info["p_re"])
    ref = dense_eta(dat, theta, n, 4)
    assert np.allclose(eta_from_pack(pk, n, 4, theta), ref, rtol=1e-13, atol=1e-13)
    # a parameter whose block differs in one entry keeps its own values
    Xre = sp.lil_matrix(dat["X_re"])
    r, c = 3 * n + 17, np.flatnonzero(sp.csr_matrix(dat["X_re"])[3 * n + 17].toarray().ravel())[0]
    Xre[r, c] = Xre[r, c] * 1.5
    dat2 = dict(dat, X_re=sp.csr_matrix(Xre))
    pk2 = pack_host(dat2)
    assert alias_of(pk2["desc"][0]["flags"], 3) == -1 and alias_of(pk2["desc"][1]["flags"], 3) == 2
    assert np.allclose(eta_from_pack(pk2, n, 4, theta), dense_eta(dat2, theta, n, 4), rtol=1e-13, atol=1e-13)
    # SSDE_NO_ALIAS=1: the plain layout (A/B runs)
    monkeypatch.setenv("SSDE_NO_ALIAS", "1")
    pk3 = pack_host(dat)
    assert not np.any(pk3["desc"]["flags"] & 0xff00) and pk3["val"].size == 22 * 256 * live.size
    assert np.allclose(eta_from_pack(pk3, n, 4, theta), ref, rtol=1e-13, atol=1e-13)
    monkeypatch.delenv("SSDE_NO_ALIAS")
    # OU: mu and tau share s(time) + s(ID, bs = "re"); kappa ~ 1 has its own slot
    dat4, _, info4 = synth.make_problem("OU", 3, 200, n_dim=1, seed=6)
    pk4 = pack_host(dat4)
    for d in pk4["desc"][: (info4["n"] + 255) // 256]:
        kb = [(int(d["kmax"]) >> (8 * p)) & 255 for p in range(4)]
        assert alias_of(d["flags"], 1) == 0 and alias_of(d["flags"], 2) == -1
        assert value_slots(kb, d["flags"])[0] == kb[0] + 1
    th4 = np.random.default_rng(4).normal(size=info4["p_fe"] + info4["p_re"])
    assert np.allclose(eta_from_pack(pk4, info4["n"], 3, th4), dense_eta(dat4, th4, info4["n"], 3), rtol=1e-13, atol=1e-13)
    # decay terms scale the values per parameter inside the kernel: plain layout
    n4 = info4["n"]
    dat6 = dict(dat4, t_decay=np.tile(np.linspace(0.0, 1.0, n4), 3), col_decay=np.array([1]), ind_decay=np.array([1]))
    assert not np.any(pack_host(dat6)["desc"]["flags"] & 0xff00)
    dat5, _, info5 = synth.make_problem("OU_SSM", 2, 200, n_dim=1, seed=6)
    pk5 = pack_host(dat5)
    assert alias_of(pk5["desc"][0]["flags"], 1) == 0
    th5 = np.random.default_rng(3).normal(size=info5["p_fe"] + info5["p_re"])
    assert np.allclose(eta_from_pack(pk5, info5["n"], 3, th5), dense_eta(dat5, th5, info5["n"], 3), rtol=1e-13, atol=1e-13)


def test_three_parameters_sharing_one_smooth_alias_the_first():
    """BM, d = 2: mu1, mu2, sigma ~ s(time): parameters 1 and 2 both read parameter 0's value slots (the second
    one is not adjacent to its target: the kernels' unpaired path)."""
    dat, par, info = synth.make_problem("BM", 2, 300, n_dim=2, seed=9)
    n = info["n"]
    pk = pack_host(dat)
    for d in pk["desc"][: (n + 255) // 256]:
        kb = [(int(d["kmax"]) >> (8 * p)) & 255 for p in range(4)]
        assert kb[0] == kb[1] == kb[2] and kb[3] == 0
        assert [alias_of(d["flags"], p) for p in range(3)] == [-1, 0, 0]
        assert value_slots(kb, d["flags"]) == (kb[0], [0, 0, 0, kb[0]])
    theta = np.random.default_rng(5).normal(size=info["p_fe"] + info["p_re"])
    assert np.allclose(eta_from_pack(pk, n, 3, theta), dense_eta(dat, theta, n, 3), rtol=1e-13, atol=1e-13)


def test_irregular_sparsity_falls_back_to_per_nonzero_columns():
    rng = np.random.default_rng(3)
    n, n_par, p_re = 700, 2, 60
    dat, par, info = synth.make_problem("BM", 1, n, n_dim=1, seed=8)
    Xre = sp.random(n_par * n, p_re, density=0.05, random_state=5, format="csr")
    # keep the block-diagonal-by-parameter contract: parameter 0 uses columns < 30, parameter 1 the rest
    Xre = sp.vstack([sp.hstack([Xre[:n, :30], sp.csr_matrix((n, 30))]),
                     sp.hstack([sp.csr_matrix((n, 30)), Xre[n:, 30:]])], format="csr")
    dat = dict(dat, X_re=Xre, S=sp.identity(p_re, format="csr"), ncol_re=np.array([30, 30]))
    pk = pack_host(dat)
    assert not np.all(pk["desc"]["flags"][: (n + 255) // 256] & 1)
    theta = rng.normal(size=dat["X_fe"].shape[1] + p_re)
    assert np.allclose(eta_from_pack(pk, n, n_par, theta), dense_eta(dat, theta, n, n_par), rtol=1e-13, atol=1e-13)


def test_duplicate_triplets_are_summed_and_bad_indices_rejected():
    dat, par, info = synth.make_problem("BM", 1, 50, n_dim=1, seed=2)
    n = 50
    Xfe = sp.coo_matrix(dat["X_fe"])
    dup = sp.coo_matrix((np.r_[Xfe.data, Xfe.data], (np.r_[Xfe.row, Xfe.row], np.r_[Xfe.col, Xfe.col])), shape=Xfe.shape)

    class Raw:                       # keep duplicates: scipy would sum them on conversion
        pass
    import smoothsde_b200.engine as E
    orig = E._as_triplet

    def raw_triplet(M, keep):
        if M is dup:
            i = np.ascontiguousarray(dup.row, dtype=np.int32)
            j = np.ascontiguousarray(dup.col, dtype=np.int32)
            x = np.ascontiguousarray(dup.data, dtype=np.float64)
            keep += [i, j, x]
            t = L.Triplet()
            t.nrow, t.ncol, t.nnz = dup.shape[0], dup.shape[1], x.size
            t.i, t.j, t.x = i.ctypes.data_as(L.c_int32_p), j.ctypes.data_as(L.c_int32_p), x.ctypes.data_as(L.c_double_p)
            return t
        return orig(M, keep)
    E._as_triplet = raw_triplet
    try:
        pk = pack_host(dict(dat, X_fe=dup))
    finally:
        E._as_triplet = orig
    theta = np.random.default_rng(1).normal(size=info["p_fe"] + info["p_re"])
    ref = dense_eta(dat, theta, n, 2)
    th2 = theta.copy()
    got = eta_from_pack(pk, n, 2, th2)
    fe = np.asarray(sp.csr_matrix(dat["X_fe"]) @ theta[:info["p_fe"]]).reshape(2, n).T
    assert np.allclose(got, ref + fe, rtol=1e-13, atol=1e-13)     # X_fe counted twice
