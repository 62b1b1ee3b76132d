"""ctypes wrapper of oracle/_ref/libsmoothsde_ref.so: the reference's OWN objective
(/root/reference/src/smoothSDE.cpp:9-28 + src/nllk/*.hpp, unmodified, compiled where they lie
against the TMB stand-in oracle/tmb_shim/TMB.hpp).  TEST INFRASTRUCTURE / CPU BASELINE ONLY:
imported by tests/, tests/golden/make_golden.py, __graft_entry__ and bench.py's CPU legs, never by
smoothsde_b200/.

The library can only be (re)built where /root/reference exists (this container); the built .so
travels to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).

`RefOracle(dat)` takes the same *data list* as SDE$setup() builds (R/sde.R:528-598) and evaluates
value / gradient (reverse sweep of the shim's AD tape) / Hessian-vector products (reverse over
forward) / REPORT(aest_all) on the flat joint parameter vector of SURVEY.md 8(a) A1.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsmoothsde_ref.so")
REFERENCE_SRC = "/root/reference/src"
KALMAN_TYPES = ("CTCRW", "OU_SSM", "BM_SSM")
# R's NA_real_: the NaN with low word 1954 (what R_IsNA tests; a plain NaN is NOT NA in R)
NA_REAL = np.array([0x7FF00000000007A2], dtype=np.uint64).view(np.float64)[0]

_lib = None


class _Item(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int), ("ndim", C.c_int), ("dim", C.c_long * 3),
                ("d", C.c_void_p), ("i", C.c_void_p), ("s", C.c_char_p), ("ti", C.c_void_p),
                ("tj", C.c_void_p), ("nnz", C.c_long)]


class _Par(C.Structure):
    _fields_ = [("name", C.c_char_p), ("len", C.c_long)]


def available() -> bool:
    return os.path.exists(LIB) or os.path.isdir(REFERENCE_SRC)


def build(force=False):
    """Compile the reference sources where they lie.  No-op (uses the prebuilt .so) where
    /root/reference does not exist."""
    srcs = [os.path.join(HERE, "ref_driver.cpp"), os.path.join(HERE, "tmb_shim", "TMB.hpp")]
    if not os.path.isdir(REFERENCE_SRC):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("oracle/_ref is not built and /root/reference is not present")
    srcs += [os.path.join(REFERENCE_SRC, "smoothSDE.cpp")] + \
        [os.path.join(REFERENCE_SRC, "nllk", f) for f in sorted(os.listdir(os.path.join(REFERENCE_SRC, "nllk")))]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-B", "_ref/libsmoothsde_ref.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.ssde_ref_eval.restype = C.c_int
        _lib.ssde_ref_report.restype = C.c_int
    return _lib


def _with_na(a):
    """NaN -> R's NA_real_ (the reference tests R_IsNA, nllk_ctcrw.hpp:214, tr_dens.hpp:31)."""
    a = np.array(a, dtype=np.float64, order="F", copy=True)
    a[np.isnan(a)] = NA_REAL
    return a


class RefOracle:
    def __init__(self, dat):
        self.type = str(dat["type"])
        self._keep = []
        items = []

        def add_d(name, arr, dims=None):
            arr = np.asfortranarray(np.asarray(arr, dtype=np.float64))
            self._keep.append(arr)
            it = _Item()
            it.name = name.encode()
            it.kind = 0
            dims = list(arr.shape if dims is None else dims) or [1]
            it.ndim = len(dims)
            for k, v in enumerate(dims):
                it.dim[k] = int(v)
            it.d = arr.ctypes.data
            items.append(it)

        def add_i(name, arr):
            arr = np.ascontiguousarray(np.atleast_1d(np.asarray(arr)).astype(np.int32))
            self._keep.append(arr)
            it = _Item()
            it.name = name.encode()
            it.kind = 1
            it.ndim = 1
            it.dim[0] = arr.size
            it.i = arr.ctypes.data
            items.append(it)

        def add_s(name, s):
            it = _Item()
            it.name = name.encode()
            it.kind = 2
            it.s = s.encode()
            items.append(it)

        def add_sp(name, M):
            # dgTMatrix triplets, duplicates left in place (R/utility.R:204-213)
            M = sp.coo_matrix(M)
            ti = np.ascontiguousarray(M.row, dtype=np.int32)
            tj = np.ascontiguousarray(M.col, dtype=np.int32)
            tx = np.ascontiguousarray(M.data, dtype=np.float64)
            self._keep += [ti, tj, tx]
            it = _Item()
            it.name = name.encode()
            it.kind = 3
            it.ndim = 2
            it.dim[0], it.dim[1] = M.shape
            it.d, it.ti, it.tj, it.nnz = tx.ctypes.data, ti.ctypes.data, tj.ctypes.data, tx.size
            items.append(it)

        obs = np.asarray(dat["obs"], dtype=float)
        if obs.ndim == 1:
            obs = obs[:, None]
        self.n, self.d = obs.shape
        add_s("type", self.type)
        add_d("ID", np.asarray(dat["ID"], dtype=float))
        add_d("times", np.asarray(dat["times"], dtype=float))
        add_d("obs", _with_na(obs))
        add_sp("X_fe", dat["X_fe"])
        add_sp("X_re", dat["X_re"])
        add_sp("S", dat["S"])
        ncol_re = np.atleast_1d(np.asarray(dat["ncol_re"], dtype=np.int64))
        add_i("ncol_re", ncol_re)
        self.p_fe = dat["X_fe"].shape[1]
        self.p_re = dat["X_re"].shape[1]
        self.n_s = int(ncol_re.size) if ncol_re[0] > 0 else 1
        layout = []
        self.has_decay = False
        if self.type in KALMAN_TYPES:
            add_d("a0", np.asarray(dat["a0"], dtype=float))
            add_d("P0", np.asarray(dat["P0"], dtype=float))
            H = dat.get("H_array")
            if H is None or np.size(H) <= 1:
                add_d("H_array", np.zeros(1))                  # array(0), R/sde.R:567,598
            else:
                add_d("H_array", np.asarray(H, dtype=float))   # d x d x n
            layout = [("log_sigma_obs", 1), ("coeff_fe", self.p_fe), ("log_lambda", self.n_s), ("coeff_re", self.p_re)]
        else:
            add_i("include_penalty", [int(dat.get("include_penalty", 1))])
            add_d("other_data", np.atleast_1d(np.asarray(dat.get("other_data", 0.0), dtype=float)))
            t_decay = dat.get("t_decay")
            self.has_decay = t_decay is not None and np.size(t_decay) > 1
            if self.has_decay:
                add_d("t_decay", np.asarray(t_decay, dtype=float))
                add_i("col_decay", dat["col_decay"])
                add_i("ind_decay", dat["ind_decay"])
                self.n_decay = int(np.unique(np.asarray(dat["ind_decay"])).size)
            else:
                add_d("t_decay", np.zeros(1))                  # R/sde.R:644-646
                add_i("col_decay", [0])
                add_i("ind_decay", [0])
                self.n_decay = 1                               # log_decay = log(rho) fixed by map, R/sde.R:648
            layout = [("coeff_fe", self.p_fe), ("log_lambda", self.n_s), ("log_decay", self.n_decay), ("coeff_re", self.p_re)]
        self.items = (_Item * len(items))(*items)
        self.layout_names = layout
        self.layout = (_Par * len(layout))(*[_Par(nm.encode(), ln) for nm, ln in layout])
        self.n_full = sum(ln for _, ln in layout)
        # positions of the flat vector inside the reference's full PARAMETER list:
        # the fixed log_decay of decay-free models is absent from the flat vector
        keep = []
        o = 0
        for nm, ln in layout:
            if not (nm == "log_decay" and not self.has_decay):
                keep += list(range(o, o + ln))
            o += ln
        self.keep = np.asarray(keep, dtype=np.int64)

    def _full(self, par):
        full = np.zeros(self.n_full)
        par = np.asarray(par, dtype=float)
        if par.size != self.keep.size:
            raise ValueError(f"parameter vector has {par.size} entries, expected {self.keep.size}")
        full[self.keep] = par
        return full

    def _call(self, par, order, direction=None):
        L = lib()
        full = self._full(par)
        val = C.c_double(0.0)
        grad = np.zeros(self.n_full)
        hv = np.zeros(self.n_full)
        dfull = self._full(direction) if direction is not None else np.zeros(self.n_full)
        err = C.create_string_buffer(512)
        rc = L.ssde_ref_eval(self.items, C.c_int(len(self.items)), self.layout, C.c_int(len(self.layout)),
                             full.ctypes.data_as(C.c_void_p), C.c_int(order), dfull.ctypes.data_as(C.c_void_p),
                             C.byref(val), grad.ctypes.data_as(C.c_void_p), hv.ctypes.data_as(C.c_void_p),
                             err, C.c_int(512))
        if rc != 0:
            raise RuntimeError("reference objective: " + err.value.decode())
        return val.value, grad[self.keep], hv[self.keep]

    def nllk(self, par):
        return self._call(par, 0)[0]

    def eval(self, par, want_grad=True):
        if not want_grad:
            return self.nllk(par), None
        v, g, _ = self._call(par, 1)
        return v, g

    def hvp(self, par, direction):
        v, g, h = self._call(par, 2, direction)
        return v, g, h

    def hessian(self, par):
        par = np.asarray(par, dtype=float)
        H = np.empty((par.size, par.size))
        for j in range(par.size):
            e = np.zeros(par.size)
            e[j] = 1.0
            H[:, j] = self.hvp(par, e)[2]
        return 0.5 * (H + H.T)

    def aest(self, par):
        """REPORT(aest_all), nllk_ctcrw.hpp:249 / nllk_ou_ssm.hpp:216 / nllk_bm_ssm.hpp:178."""
        if self.type not in KALMAN_TYPES:
            raise ValueError("aest_all is reported by the Kalman models only")
        L = lib()
        full = self._full(par)
        ns = 2 * self.d if self.type == "CTCRW" else self.d
        out = np.zeros((self.n, ns), order="F")
        err = C.create_string_buffer(512)
        rc = L.ssde_ref_report(self.items, C.c_int(len(self.items)), self.layout, C.c_int(len(self.layout)),
                               full.ctypes.data_as(C.c_void_p), b"aest_all", out.ctypes.data_as(C.c_void_p),
                               C.c_long(out.size), err, C.c_int(512))
        if rc != 0:
            raise RuntimeError("reference objective: " + err.value.decode())
        return np.ascontiguousarray(out)


def split_by_track(dat):
    """One data list per track (rows of every per-row object; a0 row k).  Tracks are independent
    given the parameters (nllk_ctcrw.hpp:196-200, nllk_sde.hpp:79), so the sum of the per-track
    objectives equals the stacked objective up to the penalty, which every piece would add."""
    ID = np.asarray(dat["ID"])
    n = ID.size
    cuts = np.concatenate([[0], np.nonzero(ID[1:] != ID[:-1])[0] + 1, [n]])
    X_fe = sp.csr_matrix(dat["X_fe"])
    X_re = sp.csr_matrix(dat["X_re"])
    n_par = X_fe.shape[0] // n
    out = []
    for k in range(cuts.size - 1):
        a, b = int(cuts[k]), int(cuts[k + 1])
        rows = np.concatenate([np.arange(j * n + a, j * n + b) for j in range(n_par)])
        piece = dict(dat)
        piece.update(ID=ID[a:b], times=np.asarray(dat["times"])[a:b], obs=np.asarray(dat["obs"])[a:b],
                     X_fe=X_fe[rows], X_re=X_re[rows])
        if "a0" in dat:
            piece["a0"] = np.asarray(dat["a0"])[k:k + 1]
        if dat.get("H_array") is not None and np.size(dat["H_array"]) > 1:
            piece["H_array"] = np.asarray(dat["H_array"])[:, :, a:b]
        if dat.get("t_decay") is not None and np.size(dat["t_decay"]) > 1:
            td = np.asarray(dat["t_decay"], dtype=float)
            piece["t_decay"] = np.concatenate([td[j * n + a:j * n + b] for j in range(n_par)])
        out.append(piece)
    return out


class RefOracleParallel:
    """The reference objective evaluated track by track on a thread pool (ctypes releases the GIL;
    the shim's tape is thread-local).  Penalty: pieces are built penalty-free where the model allows
    it (nllk_sde: include_penalty = 0) and otherwise the (K - 1) surplus penalties are subtracted."""

    def __init__(self, dat, nthreads):
        self.nthreads = int(nthreads)
        self.full = None
        self.dat = dat
        self.pieces = [RefOracle(p) for p in split_by_track(dat)]
        # a 2-row, single-track problem carries exactly one copy of the penalty and a zero data term
        # for the Kalman models (row 0 of a track contributes only the prior mean; the second row is
        # missing -> no likelihood term)
        self.K = len(self.pieces)
        self.pool = ThreadPoolExecutor(self.nthreads)
        pen = dict(split_by_track(dat)[0])
        n0 = np.asarray(pen["ID"]).size
        n_par = sp.csr_matrix(pen["X_fe"]).shape[0] // n0
        rows = np.concatenate([np.arange(j * n0, j * n0 + 2) for j in range(n_par)])
        obs2 = np.array(np.asarray(pen["obs"], dtype=float)[:2], copy=True)
        obs2[1, :] = np.nan
        pen.update(ID=np.asarray(pen["ID"])[:2], times=np.asarray(pen["times"])[:2], obs=obs2,
                   X_fe=sp.csr_matrix(pen["X_fe"])[rows], X_re=sp.csr_matrix(pen["X_re"])[rows])
        if pen.get("H_array") is not None and np.size(pen["H_array"]) > 1:
            pen["H_array"] = np.asarray(pen["H_array"])[:, :, :2]
        if pen.get("t_decay") is not None and np.size(pen["t_decay"]) > 1:
            td = np.asarray(pen["t_decay"], dtype=float)
            pen["t_decay"] = np.concatenate([td[j * n0:j * n0 + 2] for j in range(n_par)])
        self.pen = RefOracle(pen)

    def eval(self, par, want_grad=True):
        res = list(self.pool.map(lambda o: o.eval(par, want_grad), self.pieces))
        pv, pg = self.pen.eval(par, want_grad)
        v = sum(r[0] for r in res) - (self.K - 1) * pv
        if not want_grad:
            return v, None
        g = np.sum([r[1] for r in res], axis=0) - (self.K - 1) * pg
        return v, g

    def close(self):
        self.pool.shutdown()
