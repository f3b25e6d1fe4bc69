"""Device chain diagnostics (bfx_chain_diagnostics) against the oracle restatement, on host arrays and on a device ring
filled by the sampler itself."""
import numpy as np
import pytest

from oracle import bfx_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


def _close(dev, ora, ess_rtol=2e-3):
    assert np.allclose(dev["mean"], ora["mean"], rtol=1e-5, atol=1e-6)
    assert np.allclose(dev["var"], ora["var"], rtol=1e-4)
    assert np.allclose(dev["rhat"], ora["rhat"], rtol=1e-4, equal_nan=True)
    assert np.allclose(dev["ess"], ora["ess"], rtol=ess_rtol, equal_nan=True)


@pytest.mark.parametrize("P,n,C", [(7, 64, 3), (300, 401, 2), (1, 1000, 1), (40, 16, 33)])
def test_matches_the_oracle_on_autocorrelated_chains(bf, P, n, C):
    rng = np.random.default_rng(P + n + C)
    phi = rng.uniform(0.0, 0.95, size=(P, 1))
    x = np.zeros((P, n, C))
    e = rng.standard_normal((P, n, C))
    x[:, 0] = e[:, 0]
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    x += rng.standard_normal((P, 1, C)) * 0.3 * (np.arange(P) % 3 == 0)[:, None, None]  # some chains disagree
    x = x.astype(np.float32)
    _close(bf.chain_diagnostics(x), O.chain_diagnostics(x))
    _close(bf.chain_diagnostics(x, max_lag=5), O.chain_diagnostics(x, max_lag=5))
    if C == 1:
        _close(bf.chain_diagnostics(x[:, :, 0]), O.chain_diagnostics(x))


def test_device_ring_of_the_sampler(bf):
    """The samples never leave HBM: bfx_mcmc_advance records into a device ring, the diagnostics read it in place
    and agree with the oracle run on a host copy of the same ring."""
    import torch
    from tests.test_gpu_parity import LIKES, PRIORS, build
    rng = np.random.default_rng(11)
    k, n = 5, 2000
    x = rng.standard_normal((k, n)).astype(np.float32)
    y = (x.T @ rng.standard_normal(k) + rng.standard_normal(n)).astype(np.float32)
    bnn, m = build(bf, (O.Dense(5, 1),), LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    P, C, nd = m.n_total, 8, 200
    th0 = np.asfortranarray((0.1 * rng.standard_normal((P, C))).astype(np.float32))
    s = bf.SGLD(stepsize_a=0.01, stepsize_b=0.0, stepsize_gamma=0.0)
    s.initialise(bnn, th0)
    ring = torch.zeros((C, nd, P), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    bf.advance(s, 100, 300, seed=2, stream=st.cuda_stream)  # burn-in, nothing recorded
    bf.advance(s, 100, nd, keep_every=1, seed=2, stream=st.cuda_stream, samples_dev=ring.data_ptr(), ring_slots=nd)
    torch.cuda.synchronize()
    dev = bf.chain_diagnostics(None, samples_dev=ring.data_ptr(), shape=(P, nd, C))
    host = np.ascontiguousarray(ring.cpu().numpy().transpose(2, 1, 0))  # P x n x C
    assert np.isfinite(host).all()
    _close(dev, O.chain_diagnostics(host))
    assert np.all(dev["ess"] > 10) and np.all(dev["rhat"] < 1.5)


def test_bad_arguments(bf):
    from bayesflux_b200 import _lib as L
    with pytest.raises(L.BfxError):
        bf.chain_diagnostics(np.zeros((3, 2, 2), np.float32))  # fewer than 4 draws
