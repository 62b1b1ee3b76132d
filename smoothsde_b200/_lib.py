"""ctypes binding of the C ABI (include/smoothsde_b200.h).  Fails loudly if the library is
missing: there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsmoothsde_b200" + os.environ.get("SSDE_LIB_SUFFIX", "") + ".so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

SSDE_BM, SSDE_OU, SSDE_CTCRW, SSDE_BM_SSM, SSDE_OU_SSM = 0, 1, 2, 3, 4
MODEL_CODES = {"BM": SSDE_BM, "OU": SSDE_OU, "CTCRW": SSDE_CTCRW, "BM_SSM": SSDE_BM_SSM, "OU_SSM": SSDE_OU_SSM}
KALMAN_TYPES = ("CTCRW", "BM_SSM", "OU_SSM")
KNOWN_UNBUILT = ("BM_t", "CIR", "ESEAL_SSM")

SHARD_CONT_PREV, SHARD_CONT_NEXT, SHARD_NO_PENALTY = 1, 2, 4

STATUS = {0: "OK", 1: "UNKNOWN_TYPE", 2: "BAD_ARG", 3: "UNSUPPORTED", 4: "CUDA", 5: "NUMERIC"}


class Triplet(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("nnz", C.c_int64),
                ("i", c_int32_p), ("j", c_int32_p), ("x", c_double_p)]


class Desc(C.Structure):
    _fields_ = [("model", C.c_int32), ("n_dim", C.c_int32), ("n", C.c_int64),
                ("ID", c_double_p), ("times", c_double_p), ("obs", c_double_p),
                ("X_fe", Triplet), ("X_re", Triplet), ("S", Triplet),
                ("n_smooth", C.c_int32), ("ncol_re", c_int32_p), ("include_penalty", C.c_int32),
                ("n_ID", C.c_int32), ("a0", c_double_p), ("P0", c_double_p),
                ("H_array", c_double_p), ("H_len", C.c_int64),
                ("device", C.c_int32), ("shard_flags", C.c_int32), ("t_next", C.c_double),
                ("t_decay", c_double_p), ("t_decay_len", C.c_int64), ("col_decay", c_int32_p),
                ("ind_decay", c_int32_p), ("n_col_decay", C.c_int32)]


class WtDesc(C.Structure):
    _fields_ = [("val_off", C.c_int64), ("col_off", C.c_int64), ("kmax", C.c_uint32), ("flags", C.c_uint32)]


class PackedDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("n_dim", C.c_int32), ("n_par", C.c_int32),
                ("n", C.c_int64), ("n_pad", C.c_int64), ("nnz", C.c_int64),
                ("d_desc", C.c_void_p), ("d_val", C.c_void_p), ("d_col", C.c_void_p),
                ("d_obs", C.c_void_p), ("d_dt", C.c_void_p), ("d_flags", C.c_void_p),
                ("p_fe", C.c_int32), ("p_re", C.c_int32), ("S", Triplet),
                ("n_smooth", C.c_int32), ("ncol_re", c_int32_p), ("include_penalty", C.c_int32),
                ("n_ID", C.c_int32), ("track_starts", c_int64_p), ("a0", c_double_p),
                ("P0", C.c_double * 3), ("device", C.c_int32), ("shard_flags", C.c_int32),
                ("mu_cols", c_int32_p), ("n_mu_cols", C.c_int32)]


class HostPack(C.Structure):
    _fields_ = [("n_pad", C.c_int64), ("n_desc", C.c_int64), ("n_val", C.c_int64), ("n_col", C.c_int64),
                ("desc", C.POINTER(WtDesc)), ("val", c_double_p), ("col", C.POINTER(C.c_uint32))]


class LaplaceOpts(C.Structure):
    _fields_ = [("max_newton", C.c_int32), ("richardson", C.c_int32), ("grad_tol", C.c_double), ("fd_step", C.c_double)]


class LaplaceInfo(C.Structure):
    _fields_ = [("joint", C.c_double), ("logdet", C.c_double), ("grad_max", C.c_double),
                ("converged", C.c_int32), ("n_newton", C.c_int32), ("n_hess", C.c_int32),
                ("n_value", C.c_int32), ("n_hvp", C.c_int32)]


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"smoothsde_b200 error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


_lib = None

# every symbol include/smoothsde_b200.h declares
EXPORTS = ["ssde_create", "ssde_create_packed", "ssde_destroy", "ssde_n_par", "ssde_par_layout", "ssde_decay_layout",
           "ssde_eval", "ssde_eval_device", "ssde_check", "ssde_report", "ssde_last_eval_ms",
           "ssde_last_eval_launches", "ssde_set_profile", "ssde_last_kernel_times", "ssde_last_error", "ssde_create_error", "ssde_version",
           "ssde_padded_rows", "ssde_layout_info", "ssde_pack_host", "ssde_pack_free",
           "ssde_simulate_ctcrw", "ssde_simulate_ou", "ssde_launch_info",
           "ssde_shard_elem_doubles", "ssde_eval_stage",
           "ssde_hvp", "ssde_hvp_device", "ssde_hess_cols_device", "ssde_hess_theta_device",
           "ssde_laplace_create", "ssde_laplace_destroy", "ssde_laplace_eval", "ssde_laplace_hessian_bb",
           "ssde_laplace_error", "ssde_device", "ssde_stream", "ssde_debug_stats",
           "ssde_debug_const_map_tol"]


def load():
    """Load libsmoothsde_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m smoothsde_b200.build` "
            "(the engine has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.ssde_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.ssde_create.restype = C.c_int
    lib.ssde_create_packed.argtypes = [C.POINTER(PackedDesc), C.POINTER(vp)]
    lib.ssde_create_packed.restype = C.c_int
    lib.ssde_destroy.argtypes = [vp]
    lib.ssde_destroy.restype = None
    lib.ssde_n_par.argtypes = [vp]
    lib.ssde_n_par.restype = C.c_int
    lib.ssde_par_layout.argtypes = [vp, c_int32_p, c_int32_p]
    lib.ssde_par_layout.restype = C.c_int
    lib.ssde_decay_layout.argtypes = [vp, c_int32_p, c_int32_p]
    lib.ssde_decay_layout.restype = C.c_int
    lib.ssde_eval.argtypes = [vp, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p]
    lib.ssde_eval.restype = C.c_int
    lib.ssde_eval_device.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.ssde_eval_device.restype = C.c_int
    lib.ssde_check.argtypes = [vp]
    lib.ssde_check.restype = C.c_int
    lib.ssde_report.argtypes = [vp, c_double_p]
    lib.ssde_report.restype = C.c_int
    lib.ssde_last_eval_ms.argtypes = [vp]
    lib.ssde_last_eval_ms.restype = C.c_double
    lib.ssde_last_eval_launches.argtypes = [vp]
    lib.ssde_last_eval_launches.restype = C.c_int
    lib.ssde_set_profile.argtypes = [vp, C.c_int]
    lib.ssde_set_profile.restype = C.c_int
    lib.ssde_last_kernel_times.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_char_p)]
    lib.ssde_last_kernel_times.restype = C.c_int
    lib.ssde_last_error.argtypes = [vp]
    lib.ssde_last_error.restype = C.c_char_p
    lib.ssde_create_error.argtypes = []
    lib.ssde_create_error.restype = C.c_char_p
    lib.ssde_padded_rows.argtypes = [C.c_int64]
    lib.ssde_padded_rows.restype = C.c_int64
    lib.ssde_layout_info.argtypes = [c_int32_p]
    lib.ssde_layout_info.restype = C.c_int
    lib.ssde_pack_host.argtypes = [C.POINTER(Desc), C.POINTER(HostPack)]
    lib.ssde_pack_host.restype = C.c_int
    lib.ssde_pack_free.argtypes = [C.POINTER(HostPack)]
    lib.ssde_pack_free.restype = None
    lib.ssde_simulate_ctcrw.argtypes = [C.c_int, C.c_int64, C.c_int64] + [vp] * 8
    lib.ssde_simulate_ctcrw.restype = C.c_int
    lib.ssde_simulate_ou.argtypes = [C.c_int, C.c_int64, C.c_int64] + [vp] * 7
    lib.ssde_simulate_ou.restype = C.c_int
    lib.ssde_launch_info.argtypes = [vp, c_int32_p]
    lib.ssde_launch_info.restype = C.c_int
    lib.ssde_shard_elem_doubles.argtypes = [vp, C.c_int]
    lib.ssde_shard_elem_doubles.restype = C.c_int
    lib.ssde_eval_stage.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    lib.ssde_eval_stage.restype = C.c_int
    lib.ssde_hvp.argtypes = [vp, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.ssde_hvp.restype = C.c_int
    lib.ssde_hvp_device.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.ssde_hvp_device.restype = C.c_int
    lib.ssde_hess_cols_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    lib.ssde_hess_cols_device.restype = C.c_int
    lib.ssde_laplace_create.argtypes = [vp, C.POINTER(LaplaceOpts), C.POINTER(vp)]
    lib.ssde_laplace_create.restype = C.c_int
    lib.ssde_laplace_destroy.argtypes = [vp]
    lib.ssde_laplace_destroy.restype = None
    lib.ssde_laplace_eval.argtypes = [vp, c_double_p, C.c_int, c_double_p, c_double_p, C.POINTER(LaplaceInfo)]
    lib.ssde_laplace_eval.restype = C.c_int
    lib.ssde_laplace_hessian_bb.argtypes = [vp, c_double_p]
    lib.ssde_laplace_hessian_bb.restype = C.c_int
    lib.ssde_laplace_error.argtypes = [vp]
    lib.ssde_laplace_error.restype = C.c_char_p
    lib.ssde_debug_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
    lib.ssde_debug_stats.restype = C.c_int
    lib.ssde_hess_theta_device.argtypes = [vp, vp, vp, vp]
    lib.ssde_hess_theta_device.restype = C.c_int
    lib.ssde_debug_const_map_tol.argtypes = [C.c_int, C.c_double]
    lib.ssde_debug_const_map_tol.restype = C.c_int
    lib.ssde_device.argtypes = [vp]
    lib.ssde_device.restype = C.c_int
    lib.ssde_stream.argtypes = [vp]
    lib.ssde_stream.restype = vp
    lib.ssde_version.argtypes = []
    lib.ssde_version.restype = C.c_char_p
    _lib = lib
    return lib
