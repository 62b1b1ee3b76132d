"""TEST INFRASTRUCTURE ONLY: an object with the interface of smoothsde_b200.engine.Engine whose
numbers come from the CPU oracle, so that the host-side logic above the C ABI (map / random /
fit bookkeeping in adfun.py and sde.py) can be tested on a machine without a GPU.  The product
never imports this."""
import numpy as np

from oracle import oracle_c
from oracle import oracle_np as O


class OracleEngine:
    def __init__(self, dat, add_penalty=True):
        self.dat = dat
        self.co = oracle_c.COracle(dat, nthreads=2)
        # a shard that leaves the penalty to rank 0 (SSDE_SHARD_NO_PENALTY): subtract the
        # penalty-only objective, i.e. the same model on a single row (no transition at all)
        self.pen = None
        if not add_penalty:
            from smoothsde_b200.sharded import shard_rows
            self.pen = oracle_c.COracle(shard_rows(dat, 0, 1)[0], nthreads=1)
        p_fe, p_re = dat["X_fe"].shape[1], dat["X_re"].shape[1]
        ncol_re = np.atleast_1d(np.asarray(dat["ncol_re"]))
        n_s = ncol_re.size
        o = 0
        self.layout = {}
        if dat["type"] == "CTCRW":
            self.layout["log_sigma_obs"] = (0, 1)
            o = 1
        self.layout["coeff_fe"] = (o, p_fe); o += p_fe
        self.layout["log_lambda"] = (o, n_s); o += n_s
        self.layout["coeff_re"] = (o, p_re); o += p_re
        self.n_par = o
        self._last = None

    def eval(self, par, order=1):
        self._last = np.asarray(par, dtype=float).copy()
        v, g = self.co.eval(self._last, order >= 1)
        if self.pen is not None:
            pv, pg = self.pen.eval(self._last, order >= 1)
            v = v - pv
            g = g - pg if order >= 1 else None
        return v, g

    def hvp(self, par, dirs, k=1e-3):
        par, dirs = np.asarray(par, dtype=float), np.asarray(dirs, dtype=float)
        v, g = self.eval(par, 1)
        D = dirs.reshape(par.size, -1)
        hv = np.zeros_like(D)
        for j in range(D.shape[1]):
            d = D[:, j]
            d1 = (self.eval(par + k * d)[1] - self.eval(par - k * d)[1]) / (2 * k)
            d2 = (self.eval(par + 0.5 * k * d)[1] - self.eval(par - 0.5 * k * d)[1]) / k
            hv[:, j] = (4 * d2 - d1) / 3
        return v, g, (hv[:, 0] if dirs.ndim == 1 else hv)

    def hessian(self, par):
        v, g = self.eval(par, 1)
        return v, g, self.co.hessian(np.asarray(par, dtype=float))

    def report(self, n, n_dim, state_dim=None):
        return self.co.aest(self._last)

    def close(self):
        pass


def oracle_adfun(data, parameters, map=None, random=None, device=0):
    from smoothsde_b200.adfun import ADFun
    return ADFun(data, parameters, map=map, random=random, engine=OracleEngine(data))


def oracle_shard_factory(dat, device, shard_flags, t_next):
    """engine_factory for smoothsde_b200.sharded on a machine without a GPU (whole-track shards)."""
    assert not (shard_flags & 3), "the oracle-backed fake engine has no time shards"
    return OracleEngine(dat, add_penalty=not (shard_flags & 4))
