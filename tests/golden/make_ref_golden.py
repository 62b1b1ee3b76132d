"""Adds outputs of the REFERENCE ITSELF to the committed fixtures tests/golden/*.npz.

    python tests/golden/make_ref_golden.py          (only where /root/reference exists)

Every fixture gains `ref_nllk`, `ref_grad`, `ref_hess` (and `ref_aest_all` for the Kalman models):
the reference's own objective -- /root/reference/src/smoothSDE.cpp + src/nllk/*.hpp, unmodified,
compiled where they lie against oracle/tmb_shim/TMB.hpp (oracle/Makefile) -- evaluated on the
fixture's data list and parameter vector; gradient = reverse sweep of the shim's AD tape, Hessian
= reverse-over-forward columns.  These are the vectors that pin oracle_np / oracle_c / the CUDA
engine to the reference; they travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle_ref            # noqa: E402
import golden_util as G                  # noqa: E402


def main():
    if not os.path.isdir(oracle_ref.REFERENCE_SRC):
        raise SystemExit("/root/reference is not present: the reference vectors can only be made where it is")
    oracle_ref.build(force=True)
    for name in G.names():
        dat, out = G.load(name)
        R = oracle_ref.RefOracle(dat)
        v, g = R.eval(out["par"])
        H = R.hessian(out["par"])
        z = dict(np.load(os.path.join(G.GOLDEN_DIR, name + ".npz")))
        z["ref_nllk"], z["ref_grad"], z["ref_hess"] = np.array(v), g, H
        if dat["type"] in oracle_ref.KALMAN_TYPES:
            z["ref_aest_all"] = R.aest(out["par"])
        np.savez_compressed(os.path.join(G.GOLDEN_DIR, name + ".npz"), **z)
        rel = abs(v - out["nllk"]) / abs(out["nllk"])
        gs = np.maximum(np.abs(out["grad"]), 1e-3 * np.abs(out["grad"]).max())
        print(f"{name:28s} ref_nllk={v:.15g} vs oracle_np {rel:.1e}  grad {np.max(np.abs(g - out['grad']) / gs):.1e}"
              + (f"  hess {np.max(np.abs(H - out['hess'])) / np.abs(out['hess']).max():.1e}" if "hess" in out else ""))


if __name__ == "__main__":
    main()
