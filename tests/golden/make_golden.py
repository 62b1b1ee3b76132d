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
# user H_array / general P0 (coupled filter) and decay terms: problems from tests/test_dense.py, tests/test_decay.py
FEATURE_CASES = [
    ("ctcrw_d2_userH_2x40", "dense", ("CTCRW", 2, 40, 2, 0.1, 111, True, True)),
    ("ou_ssm_d2_userH_2x40", "dense", ("OU_SSM", 2, 40, 2, 0.1, 112, True, False)),
    ("ou_d1_decay_3x50", "decay", ("OU", 3, 50, 1, 113)),
    ("bm_d2_decay_2x60", "decay", ("BM", 2, 60, 2, 114)),
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
    for nm in ("H_array", "t_decay", "col_decay", "ind_decay"):
        if dat.get(nm) is not None:
            out[nm] = np.asarray(dat[nm])
    out.update(extra)
    return out


def feature_cases():
    """Fixtures of the coupled filter (user H_array / P0) and of the decay terms.  Known answers:
    one dense multivariate normal per track (CTCRW), the objective of the pre-scaled design (decay)."""
    import warnings
    sys.path.insert(0, os.path.dirname(HERE))
    from test_dense import dense_problem
    from test_decay import decay_problem
    for name, kind, args in FEATURE_CASES:
        if ONLY and name not in ONLY:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if kind == "dense":
                dat, par = dense_problem(*args)
                ka = KA.known_ctcrw(dat, par) if dat["type"] == "CTCRW" else O.nllk(dat, par)
            else:
                dat, par, _ = decay_problem(*args)
                p = O.split_par(dat, par)
                plain = {k: v for k, v in dat.items() if k not in ("t_decay", "col_decay", "ind_decay")}
                plain["X_re"] = O.decayed_X_re(dat, p["log_decay"])
                ka = O.nllk(plain, np.r_[p["coeff_fe"], p["log_lambda"], p["coeff_re"]])
            v = O.nllk(dat, par)
            assert abs(ka - v) <= 1e-11 * abs(v), (name, ka, v)
            g = O.grad_complex_step(dat, par)
            extra = {"known_answer": np.array(ka), "hess": O.hess_complex_fd(dat, par)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(dat, par, v, g, extra))
        print(f"{name}: n={dat['obs'].shape[0]} npar={par.size} nllk={v:.15g} known={ka:.15g}")


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
    feature_cases()
