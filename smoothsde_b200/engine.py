"""Host-side handle of the CUDA likelihood engine.

`Engine` wraps one ``ssde_handle`` of the C ABI (include/smoothsde_b200.h): it takes the data
list that SDE$setup() builds for TMB::MakeADFun (R/sde.R:528-598), ships it to the GPU once and
then evaluates the joint penalised negative log-likelihood and its gradient for full parameter
vectors in the templates' PARAMETER order.  `map` / `random` handling
lives one level up, in :mod:`smoothsde_b200.adfun`, exactly as TMB's R layer sits above
EvalADFunObject.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib as L


def _as_triplet(M, keep):
    """scipy sparse / dense -> (Triplet struct, arrays kept alive).  dgTMatrix convention."""
    if M is None:
        M = sp.coo_matrix((1, 1))
    M = sp.coo_matrix(M)
    i = np.ascontiguousarray(M.row, dtype=np.int32)
    j = np.ascontiguousarray(M.col, dtype=np.int32)
    x = np.ascontiguousarray(M.data, dtype=np.float64)
    keep += [i, j, x]
    t = L.Triplet()
    t.nrow, t.ncol, t.nnz = M.shape[0], M.shape[1], x.size
    t.i = i.ctypes.data_as(L.c_int32_p)
    t.j = j.ctypes.data_as(L.c_int32_p)
    t.x = x.ctypes.data_as(L.c_double_p)
    return t


def model_code(type_):
    if type_ in L.MODEL_CODES:
        return L.MODEL_CODES[type_]
    if type_ in L.KNOWN_UNBUILT:
        raise L.EngineError(3, f"SDE type '{type_}' exists in the reference but is not built here")
    raise L.EngineError(1, "Unknown SDE type")          # src/smoothSDE.cpp:25


def make_desc(dat, device=0, shard_flags=0, t_next=0.0):
    """The reference's data list (R/sde.R:528-598) as an ``ssde_desc``; returns (desc, keep-alive)."""
    keep = []
    d = L.Desc()
    d.model = model_code(dat["type"])
    obs = np.asarray(dat["obs"], dtype=np.float64)
    if obs.ndim == 1:
        obs = obs[:, None]
    n, nd = obs.shape
    d.n_dim, d.n = nd, n
    ID = np.ascontiguousarray(dat["ID"], dtype=np.float64)
    times = np.ascontiguousarray(dat["times"], dtype=np.float64)
    obs_cm = np.asfortranarray(obs)
    keep += [ID, times, obs_cm]
    d.ID = ID.ctypes.data_as(L.c_double_p)
    d.times = times.ctypes.data_as(L.c_double_p)
    d.obs = obs_cm.ctypes.data_as(L.c_double_p)
    d.X_fe = _as_triplet(dat["X_fe"], keep)
    d.X_re = _as_triplet(dat["X_re"], keep)
    d.S = _as_triplet(dat.get("S"), keep)
    ncol_re = np.ascontiguousarray(np.atleast_1d(dat["ncol_re"]), dtype=np.int32)
    keep.append(ncol_re)
    d.n_smooth = ncol_re.size
    d.ncol_re = ncol_re.ctypes.data_as(L.c_int32_p)
    d.include_penalty = int(dat.get("include_penalty", 1))
    if dat["type"] in L.KALMAN_TYPES:
        a0 = np.asfortranarray(np.asarray(dat["a0"], dtype=np.float64).reshape(-1, 2 * nd if dat["type"] == "CTCRW" else nd))
        P0 = np.asfortranarray(np.asarray(dat["P0"], dtype=np.float64))
        keep += [a0, P0]
        d.n_ID = a0.shape[0]
        d.a0 = a0.ctypes.data_as(L.c_double_p)
        d.P0 = P0.ctypes.data_as(L.c_double_p)
        H = dat.get("H_array")
        if H is not None and np.size(H) > 1:
            H = np.asfortranarray(np.asarray(H, dtype=np.float64))
            keep.append(H)
            d.H_array = H.ctypes.data_as(L.c_double_p)
            d.H_len = H.size
    td = dat.get("t_decay")
    if td is not None and np.size(td) > 1:                  # decay terms, nllk_sde.hpp:31-33,47-59
        td = np.ascontiguousarray(td, dtype=np.float64)
        cd = np.ascontiguousarray(np.atleast_1d(dat["col_decay"]), dtype=np.int32)
        idc = np.ascontiguousarray(np.atleast_1d(dat["ind_decay"]), dtype=np.int32)
        if cd.size != idc.size:
            raise ValueError("Check length of 'other_data$ind_decay' and 'other_data$col_decay'")     # R/sde.R:174-176
        keep += [td, cd, idc]
        d.t_decay = td.ctypes.data_as(L.c_double_p)
        d.t_decay_len = td.size
        d.col_decay = cd.ctypes.data_as(L.c_int32_p)
        d.ind_decay = idc.ctypes.data_as(L.c_int32_p)
        d.n_col_decay = cd.size
    d.device = device
    d.shard_flags = shard_flags
    d.t_next = float(t_next)
    return d, keep


def pack_host(dat):
    """Design of `dat` in the device layout, packed on the host (no GPU needed).
    Returns dict(n_pad, desc [structured array], val, col)."""
    lib = L.load()
    d, keep = make_desc(dat)
    hp = L.HostPack()
    rc = lib.ssde_pack_host(C.byref(d), C.byref(hp))
    if rc != 0:
        raise L.EngineError(rc, lib.ssde_create_error().decode())
    try:
        dt = np.dtype([("val_off", "<i8"), ("col_off", "<i8"), ("kmax", "<u4"), ("flags", "<u4")])
        desc = np.frombuffer(C.string_at(hp.desc, hp.n_desc * 24), dtype=dt).copy()
        val = np.ctypeslib.as_array(hp.val, shape=(max(hp.n_val, 1),))[:hp.n_val].copy()
        col = np.ctypeslib.as_array(hp.col, shape=(max(hp.n_col, 1),))[:hp.n_col].copy()
        return {"n_pad": int(hp.n_pad), "desc": desc, "val": val, "col": col}
    finally:
        lib.ssde_pack_free(C.byref(hp))


class Engine:
    """One shard of one model on one GPU."""

    def __init__(self, handle, lib, keep=None):
        self._h = handle
        self._lib = lib
        self._keep = keep or []
        self.n_par = lib.ssde_n_par(handle)
        off = (C.c_int32 * 4)()
        siz = (C.c_int32 * 4)()
        lib.ssde_par_layout(handle, off, siz)
        names = ("log_sigma_obs", "coeff_fe", "log_lambda", "coeff_re")
        self.layout = {nm: (int(off[k]), int(siz[k])) for k, nm in enumerate(names) if siz[k] > 0 or k > 0}
        od, nd = C.c_int32(), C.c_int32()
        lib.ssde_decay_layout(handle, C.byref(od), C.byref(nd))
        if nd.value > 0:                                    # BM/OU decay models only (nllk_sde.hpp:44)
            self.layout["log_decay"] = (od.value, nd.value)
        self._grad = np.zeros(self.n_par)
        self._nllk = C.c_double()

    # ------------------------------------------------------------------------------------
    @classmethod
    def from_data(cls, dat, device=0, shard_flags=0, t_next=0.0):
        lib = L.load()
        d, keep = make_desc(dat, device, shard_flags, t_next)
        h = C.c_void_p()
        rc = lib.ssde_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise L.EngineError(rc, lib.ssde_create_error().decode())
        return cls(h, lib)       # the library copied everything it needs

    @classmethod
    def from_packed(cls, pd: "L.PackedDesc", keep):
        """Adopt device-resident arrays (see ssde_create_packed); `keep` holds their owners."""
        lib = L.load()
        h = C.c_void_p()
        rc = lib.ssde_create_packed(C.byref(pd), C.byref(h))
        if rc != 0:
            raise L.EngineError(rc, lib.ssde_create_error().decode())
        return cls(h, lib, keep)

    # ------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise L.EngineError(rc, self._lib.ssde_last_error(self._h).decode())

    def eval(self, par, order=1):
        """(nllk, grad) for the full parameter vector; host buffers in and out (obj$fn / obj$gr)."""
        par = np.ascontiguousarray(par, dtype=np.float64)
        if par.size != self.n_par:
            raise ValueError(f"parameter vector has length {par.size}, expected {self.n_par}")
        g = self._grad
        rc = self._lib.ssde_eval(self._h, par.ctypes.data_as(L.c_double_p), int(order),
                                 C.byref(self._nllk), g.ctypes.data_as(L.c_double_p), None)
        self._check(rc)
        return self._nllk.value, (g.copy() if order >= 1 else None)

    def hessian(self, par):
        """(nllk, grad, H): exact joint Hessian of the penalised objective (obj$he), order 2."""
        par = np.ascontiguousarray(par, dtype=np.float64)
        if par.size != self.n_par:
            raise ValueError(f"parameter vector has length {par.size}, expected {self.n_par}")
        g = self._grad
        H = np.zeros((self.n_par, self.n_par), order="F")
        rc = self._lib.ssde_eval(self._h, par.ctypes.data_as(L.c_double_p), 2, C.byref(self._nllk),
                                 g.ctypes.data_as(L.c_double_p), H.ctypes.data_as(L.c_double_p))
        self._check(rc)
        return self._nllk.value, g.copy(), np.ascontiguousarray(H)

    def hvp(self, par, dirs):
        """(nllk, grad, H @ dirs) by tangent passes; dirs is [n_par] or [n_par, k]."""
        par = np.ascontiguousarray(par, dtype=np.float64)
        dirs = np.asarray(dirs, dtype=np.float64)
        one = dirs.ndim == 1
        D = np.asfortranarray(dirs.reshape(self.n_par, -1))
        hv = np.zeros_like(D, order="F")
        g = self._grad
        rc = self._lib.ssde_hvp(self._h, par.ctypes.data_as(L.c_double_p), D.shape[1], D.ctypes.data_as(L.c_double_p),
                                C.byref(self._nllk), g.ctypes.data_as(L.c_double_p), hv.ctypes.data_as(L.c_double_p))
        self._check(rc)
        hv = np.ascontiguousarray(hv)
        return self._nllk.value, g.copy(), (hv[:, 0] if one else hv)

    def hvp_device(self, d_par_ptr, d_dir_ptr, d_out_ptr, d_hv_ptr, stream_ptr=None):
        rc = self._lib.ssde_hvp_device(self._h, C.c_void_p(d_par_ptr), C.c_void_p(d_dir_ptr), C.c_void_p(d_out_ptr),
                                       C.c_void_p(d_hv_ptr), C.c_void_p(stream_ptr or 0))
        self._check(rc)

    def hess_cols_device(self, d_par_ptr, first, count, d_out_ptr, d_hess_ptr, stream_ptr=None):
        rc = self._lib.ssde_hess_cols_device(self._h, C.c_void_p(d_par_ptr), int(first), int(count), C.c_void_p(d_out_ptr),
                                             C.c_void_p(d_hess_ptr), C.c_void_p(stream_ptr or 0))
        self._check(rc)

    def hess_theta_device(self, d_par_ptr, d_hess_ptr, stream_ptr=None):
        """BM / OU: Hessian of the penalised objective w.r.t. theta = [coeff_fe | coeff_re] in ONE pass over the
        design (X' W X + lambda S), into a [p_theta x p_theta] device buffer (ssde_hess_theta_device)."""
        self._check(self._lib.ssde_hess_theta_device(self._h, C.c_void_p(d_par_ptr), C.c_void_p(d_hess_ptr), C.c_void_p(stream_ptr or 0)))

    def supports_hess_theta(self):
        """True for BM / OU handles without decay terms whose design the one-pass Hessian kernel handles."""
        if "log_sigma_obs" in self.layout or "log_decay" in self.layout:
            return False
        import torch
        dev = torch.device("cuda", self._lib.ssde_device(self._h))
        p = self.layout["coeff_fe"][1] + self.layout["coeff_re"][1]
        d_par = torch.zeros(self.n_par, dtype=torch.float64, device=dev)
        H = torch.zeros((p, p), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        rc = self._lib.ssde_hess_theta_device(self._h, C.c_void_p(d_par.data_ptr()), C.c_void_p(H.data_ptr()), C.c_void_p(0))
        torch.cuda.synchronize(dev)
        return rc == 0

    def hess_theta(self, par):
        """Host convenience: the same matrix as a numpy array."""
        import torch
        dev = torch.device("cuda", self._lib.ssde_device(self._h))
        p = self.layout["coeff_fe"][1] + self.layout["coeff_re"][1]
        d_par = torch.as_tensor(np.ascontiguousarray(par, dtype=np.float64), device=dev)
        H = torch.zeros((p, p), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        self.hess_theta_device(d_par.data_ptr(), H.data_ptr())
        torch.cuda.synchronize(dev)
        self.check()
        return H.cpu().numpy().T.copy()              # column-major on the device

    def eval_device(self, d_par_ptr, d_out_ptr, order=1, stream_ptr=None):
        """Asynchronous evaluation on device buffers (raw pointers, e.g. tensor.data_ptr())."""
        rc = self._lib.ssde_eval_device(self._h, C.c_void_p(d_par_ptr), int(order),
                                        C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or 0))
        self._check(rc)

    def shard_elem_doubles(self, which):
        return int(self._lib.ssde_shard_elem_doubles(self._h, int(which)))

    def eval_stage(self, stage, d_par_ptr, d_out_ptr, d_elems_ptr=0, n_shards=1, my_shard=0, stream_ptr=None):
        """One stage of a time-sharded evaluation (see ssde_eval_stage); raw device pointers."""
        rc = self._lib.ssde_eval_stage(self._h, C.c_void_p(d_par_ptr), int(stage), C.c_void_p(d_elems_ptr or 0),
                                       int(n_shards), int(my_shard), C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr or 0))
        self._check(rc)

    def check(self):
        self._check(self._lib.ssde_check(self._h))

    def report(self, n, n_dim, state_dim=None):
        """REPORT(aest_all) at the parameters of the last eval(); state_dim = columns of aest_all
        (2 n_dim for CTCRW, the default; n_dim for BM_SSM / OU_SSM)."""
        out = np.zeros((n, state_dim or 2 * n_dim), order="F")
        self._check(self._lib.ssde_report(self._h, out.ctypes.data_as(L.c_double_p)))
        return np.ascontiguousarray(out)

    @property
    def last_eval_ms(self):
        return self._lib.ssde_last_eval_ms(self._h)

    @property
    def last_eval_launches(self):
        return self._lib.ssde_last_eval_launches(self._h)

    def launch_info(self):
        info = (C.c_int32 * 8)()
        self._lib.ssde_launch_info(self._h, info)
        keys = ("sms", "grid_fwd", "grid_bwd", "grid_sde", "tiles_fwd", "tiles_bwd", "tiles_sde", "tracks")
        return dict(zip(keys, (int(x) for x in info)))

    def set_profile(self, on=True):
        self._lib.ssde_set_profile(self._h, int(bool(on)))

    def last_kernel_times(self):
        """[(kernel name, device ms)] of the last evaluation (profiling mode only)."""
        ms = (C.c_float * 16)()
        names = (C.c_char_p * 16)()
        k = self._lib.ssde_last_kernel_times(self._h, 16, ms, names)
        return [(names[i].decode(), float(ms[i])) for i in range(k)]

    def close(self):
        if self._h is not None:
            self._lib.ssde_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
