"""Predictive helpers: bfx_predict vs the oracle forward pass, and the reference's
predictive-quantile test (test/likelihoods.jl:33-38,59-64) and shape test (test/posterior_predict.jl:5-27)."""
import numpy as np
import pytest
from scipy import stats

from oracle import bfx_oracle as O
from tests.test_gpu_parity import LIKES, PRIORS, build

pytestmark = pytest.mark.gpu
SEED = 6150533

NETS = {
    "mlp": (O.Dense(5, 16, O.TANH), O.Dense(16, 8, O.RELU), O.Dense(8, 1)),
    "cfg2": (O.Dense(10, 64, O.RELU), O.Dense(64, 64, O.RELU), O.Dense(64, 1)),
    "lstm": (O.LSTMLayer(3, 8), O.Dense(8, 1)),
    "rnn": (O.RNNLayer(3, 6), O.Dense(6, 1)),
    "stream": (O.Dense(12, 300, O.RELU), O.Dense(300, 260, O.TANH), O.Dense(260, 1)),
}


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("newdata", [False, True])
def test_predict_matches_oracle_forward(bf, net, newdata):
    rng = np.random.default_rng(SEED)
    layers = NETS[net]
    seq = layers[0].kind != O.DENSE
    d, n, T = layers[0].nin, 150, 6
    x = rng.standard_normal((T, d, n) if seq else (d, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    C = 3
    theta = (0.3 * rng.standard_normal((m.n_total, C))).astype(np.float32)
    xq = rng.standard_normal((T, d, 77) if seq else (d, 77)).astype(np.float32) if newdata else None
    yhat, sig = bf.predict(bnn, theta, xq)
    xr = x if xq is None else xq
    assert yhat.shape == (xr.shape[-1], C)
    for c in range(C):
        ref = O.net_forward(layers, theta[:m.n_net, c].astype(np.float64), xr.astype(np.float64))
        assert np.max(np.abs(yhat[:, c] - ref)) <= 1e-4 * max(1.0, float(np.max(np.abs(ref))))
        assert sig[c] == pytest.approx(float(np.exp(theta[-1, c])), rel=1e-6)


@pytest.mark.parametrize("like", ["normal", "tdist"])
def test_predictive_quantiles_like_the_reference(bf, like):
    """theta_net = 0 => yhat = 0, sigma = 1: the predictive draws follow N(0,1) / t_10; quantile check, tol 0.05."""
    rng = np.random.default_rng(SEED + 1)
    layers = (O.Dense(10, 10, O.SIGMOID), O.Dense(10, 1))
    n = 100_000
    x = rng.standard_normal((10, n)).astype(np.float32)
    y = np.zeros(n, np.float32)
    lk = O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=2.0) if like == "normal" else O.Likelihood(O.TDIST, nu=10.0, prior_a=2.0, prior_b=2.0)
    bnn, m = build(bf, layers, lk, PRIORS["gauss"], x, y)
    th = np.zeros((m.n_total, 1), np.float32)
    ypp = bf.sample_posterior_predict(bnn, th, rng=rng)[:, 0]
    dec = np.arange(1, 10) / 10.0
    ref = stats.norm.ppf(dec) if like == "normal" else stats.t(10.0).ppf(dec)
    assert np.max(np.abs(np.quantile(ypp, dec) - ref)) < 0.05


def test_sample_posterior_predict_shapes(bf):
    # test/posterior_predict.jl:5-27
    rng = np.random.default_rng(SEED + 2)
    n = 200
    x = rng.standard_normal((5, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, (O.Dense(5, 1),), LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    ch = bf.mcmc(bnn, 20, 50, bf.SGLD(stepsize_a=1e-3), theta_start=np.zeros(m.n_total, np.float32), seed=1)
    ypp = bf.sample_posterior_predict(bnn, ch, rng=rng)
    assert ypp.shape == (n, 50)
    xnew = rng.standard_normal((5, 33)).astype(np.float32)
    assert bf.sample_posterior_predict(bnn, ch, x=xnew, rng=rng).shape == (33, 50)
    assert len(bf.get_posterior_networks(bnn, ch)) == 50
