"""Loader for the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import glob
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def c_oracle_covers(dat):
    """The C oracle restates the default-shaped models only (no user H_array, no decay terms)."""
    return dat.get("H_array") is None and dat.get("t_decay") is None


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    dat = {"type": str(z["type"]), "ID": z["ID"], "times": z["times"], "obs": z["obs"],
           "ncol_re": z["ncol_re"], "include_penalty": int(z["include_penalty"])}
    for nm in ("X_fe", "X_re", "S"):
        dat[nm] = sp.csr_matrix((z[nm + "_x"], (z[nm + "_i"], z[nm + "_j"])), shape=tuple(z[nm + "_shape"]))
    if dat["type"] in ("CTCRW", "OU_SSM", "BM_SSM"):
        dat["a0"], dat["P0"] = z["a0"], z["P0"]
    for nm in ("H_array", "t_decay", "col_decay", "ind_decay"):      # coupled filter / decay fixtures
        if nm in z:
            dat[nm] = z[nm]
    out = {"par": z["par"], "nllk": float(z["nllk"]), "grad": z["grad"],
           "known_answer": float(z["known_answer"])}
    if "nllk_mpmath" in z:
        out["nllk_mpmath"] = float(z["nllk_mpmath"])
    if "hess" in z:
        out["hess"] = z["hess"]
    if "aest_all" in z:
        out["aest_all"] = z["aest_all"]
    # outputs of the reference's own objective (tests/golden/make_ref_golden.py)
    for nm in ("ref_nllk", "ref_grad", "ref_hess", "ref_aest_all"):
        if nm in z:
            out[nm] = float(z[nm]) if nm == "ref_nllk" else z[nm]
    return dat, out
