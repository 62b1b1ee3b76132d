"""Generates the committed golden fixtures tests/golden/*.npz.

    python tests/golden/make_golden.py

The reference cannot be executed here (no R / TMB / Eigen in the image) and its
own test-suite holds no numeric vector for this path, so these fixtures are NOT outputs of TMB:
they are small seeded problems evaluated with the numpy restatement oracle/oracle_np.py (value)
and its complex-step derivative (gradient), cross-checked at generation time against

  * the dense multivariate-normal identity (CTCRW)  /  scipy.stats.norm.logpdf sums (BM, OU),
  * a 40-digit mpmath evaluation of the same recursion (bounds the oracle's own rounding).

They pin the oracle, the C oracle and the CUDA engine to one another across code changes and
travel to the GPU box (the -m gpu tests read them; nothing reads /root/reference at run time).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle_np as O            # noqa: E402
from oracle import known_answers as KA       # noqa: E402
from oracle import oracle_c                  # noqa: E402
from smoothsde_b200 import synth             # noqa: E402

CASES = [
    # name, model, tracks, steps, missing, n_dim, seed, mu override
    ("ctcrw_d2_3x40", "CTCRW", 3, 40, 0.15, 2, 101, [0.3, -0.2]),
    ("ctcrw_d1_2x60", "CTCRW", 2, 60, 0.10, 1, 102, [0.25]),
    ("ctcrw_d2_1x300", "CTCRW", 1, 300, 0.0, 2, 103, None),
    ("bm_d1_1x200", "BM", 1, 200, 0.1, 1, 104, None),
    ("bm_d2_3x50", "BM", 3, 50, 0.1, 2, 105, None),
    ("ou_d1_4x60", "OU", 4, 60, 0.1, 1, 106, None),
    ("ou_d2_2x80", "OU", 2, 80, 0.0, 2, 107, None),
    ("ou_ssm_d2_2x50", "OU_SSM", 2, 50, 0.1, 2, 108, [0.3, -0.2]),
    ("bm_ssm_d1_3x40", "BM_SSM", 3, 40, 0.1, 1, 109, [0.25]),
]
ONLY = sys.argv[1:]          # optional: names of the fixtures to (re)generate


def pack(dat, par, nllk, grad, extra):
    out = {"type": np.array(dat["type"]), "ID": dat["ID"], "times": dat["times"], "obs": dat["obs"],
           "ncol_re": np.asarray(dat["ncol_re"], dtype=np.int64),
           "include_penalty": np.array(dat["include_penalty"]),
           "par": par, "nllk": np.array(nllk), "grad": grad}
    for nm in ("X_fe", "X_re", "S"):
        M = sp.coo_matrix(dat[nm])
        out[nm + "_i"], out[nm + "_j"], out[nm + "_x"] = M.row.astype(np.int32), M.col.astype(np.int32), M.data
        out[nm + "_shape"] = np.array(M.shape)
    if dat["type"] in ("CTCRW", "OU_SSM", "BM_SSM"):
        out["a0"], out["P0"] = dat["a0"], dat["P0"]
    out.update(extra)
    return out


def main():
    for name, model, T, m, miss, nd, seed, mu in CASES:
        if ONLY and name not in ONLY:
            continue
        dat, par, info = synth.make_problem(model, T, m, missing_frac=miss, n_dim=nd, seed=seed,
                                            **({"k": 5} if model.endswith("_SSM") else {}))
        par = par.copy()
        if mu is not None:
            par[1:1 + nd] = mu
        v = O.nllk(dat, par)
        g = O.grad_complex_step(dat, par)
        # independent known answers
        ka = KA.known_answer(dat, par)
        assert abs(ka - v) <= 1e-12 * abs(v), (name, ka, v)
        mp = KA.nllk_mpmath(dat, par) if (info["n"] <= 200 and not model.endswith("_SSM")) else None
        if mp is not None:
            assert abs(mp - v) <= 1e-12 * abs(v), (name, mp, v)
        extra = {"known_answer": np.array(ka)}
        if mp is not None:
            extra["nllk_mpmath"] = np.array(mp)
        if model == "CTCRW":
            p = O.split_par(dat, par)
            extra["aest_all"] = O.nllk_ctcrw(dat, **p, return_aest=True)[1]
        if model.endswith("_SSM"):
            p = O.split_par(dat, par)
            extra["aest_all"] = O._nllk_ssm(dat, **p, model=model, return_aest=True)[1]
            extra["hess"] = O.hess_complex_fd(dat, par)      # the C oracle does not cover these models
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(dat, par, v, g, extra))
            print(f"{name}: n={info['n']} npar={par.size} nllk={v:.15g} known={ka:.15g}")
            continue
        # joint Hessian (obj$he): Richardson differences of the C oracle's analytic gradient; on the
        # smallest cases cross-checked against complex-step + differences of the numpy restatement
        Hc = oracle_c.COracle(dat).hessian(par)
        if info["n"] <= 120:
            Hn = O.hess_complex_fd(dat, par)
            assert np.max(np.abs(Hn - Hc)) <= 1e-7 * np.max(np.abs(Hc)), (name, np.max(np.abs(Hn - Hc)))
        extra["hess"] = Hc
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(dat, par, v, g, extra))
        print(f"{name}: n={info['n']} npar={par.size} nllk={v:.15g} known={ka:.15g}"
              + (f" mpmath={mp:.15g}" if mp is not None else ""))


if __name__ == "__main__":
    main()
