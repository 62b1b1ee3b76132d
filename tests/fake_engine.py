"""TEST INFRASTRUCTURE ONLY: an object with the interface of smoothsde_b200.engine.Engine whose
numbers come from the CPU oracle, so that the host-side logic above the C ABI (map / random /
fit bookkeeping in adfun.py and sde.py) can be tested on a machine without a GPU.  The product
never imports this."""
import numpy as np

from oracle import oracle_c
from oracle import oracle_np as O


class OracleEngine:
    def __init__(self, dat):
        self.dat = dat
        self.co = oracle_c.COracle(dat, nthreads=2)
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
        return v, g

    def report(self, n, n_dim):
        return self.co.aest(self._last)

    def close(self):
        pass


def oracle_adfun(data, parameters, map=None, random=None, device=0):
    from smoothsde_b200.adfun import ADFun
    return ADFun(data, parameters, map=map, random=random, engine=OracleEngine(data))
