"""The on-device CTCRW simulator (ssde_simulate_ctcrw, csrc/kernels_sim.cuh) against the host
simulator smoothsde_b200/simulate.py -- both follow SDE$simulate (R/sde.R:1448-1478) with
CTCRW_cov (R/utility.R:188-196) -- on the SAME standard normal draws."""
import numpy as np
import pytest

from smoothsde_b200 import simulate

pytestmark = pytest.mark.gpu


class _FixedNormals:
    """rng stand-in: simulate_ctcrw asks for e1 then e2 (one value per track) at every step."""

    def __init__(self, e1, e2):
        self.e1, self.e2, self.calls = e1, e2, 0

    def standard_normal(self, T):
        step, which = 1 + self.calls // 2, self.calls % 2
        self.calls += 1
        return (self.e1 if which == 0 else self.e2)[:, step]


@pytest.mark.parametrize("T,m,with_mu", [(7, 300, True), (130, 60, False)])
def test_device_simulator_reproduces_host_simulator(T, m, with_mu):
    import torch
    from smoothsde_b200 import _lib as L
    rng = np.random.default_rng(5)
    t = simulate.make_times(T, m, rng, irregular=True)
    s = t / t[:, -1:]
    tau = np.exp(0.5 * np.sin(2 * np.pi * s))
    nu = np.exp(0.3 * np.cos(2 * np.pi * s))
    mu = 0.3 * np.cos(2 * np.pi * s) if with_mu else np.zeros_like(t)
    e1, e2 = rng.standard_normal((T, m)), rng.standard_normal((T, m))
    z0 = rng.standard_normal(T)
    z_host = np.stack([simulate.simulate_ctcrw(t[k:k + 1], mu[k:k + 1], tau[k:k + 1], nu[k:k + 1],
                                               _FixedNormals(e1[k:k + 1], e2[k:k + 1]), z0=z0[k])[0] for k in range(T)])
    dev = torch.device("cuda", 0)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    z = torch.zeros((T, m), dtype=torch.float64, device=dev)
    z[:, 0] = d(z0)
    tt, ta, nn, mm, d1, d2 = d(t), d(tau), d(nu), d(mu), d(e1), d(e2)
    lib = L.load()
    torch.cuda.synchronize()
    rc = lib.ssde_simulate_ctcrw(0, T, m, tt.data_ptr(), ta.data_ptr(), nn.data_ptr(), mm.data_ptr() if with_mu else None,
                                 d1.data_ptr(), d2.data_ptr(), z.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    z_dev = z.cpu().numpy()
    assert np.max(np.abs(z_dev - z_host)) <= 1e-12 * max(1.0, np.abs(z_host).max())
