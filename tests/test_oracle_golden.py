"""Pins oracle/bfx_oracle.py against every known-answer value the reference's
own test-suite holds for the hot path, and against
torch-CPU autograd as an independent AD.  CPU only."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import bfx_oracle as O
from tests import torch_ref

SEED = 6150533  # test/runtests.jl:10


# ---------------------------------------------------------------- layout ---
@pytest.mark.parametrize("nin", [1, 2, 3])
@pytest.mark.parametrize("nout", [1, 2, 3])
def test_param_counts(nin, nout):
    # test/deconstruct.jl:19,48,80
    assert O.layer_nparams(O.Dense(nin, nout)) == nin * nout + nout
    assert O.layer_nparams(O.RNNLayer(nin, nout)) == nin * nout + nout * nout + nout
    assert O.layer_nparams(O.LSTMLayer(nin, nout)) == 4 * (nin * nout + nout * nout + nout)
    # stacked nets test/deconstruct.jl:105,129,153
    assert O.net_nparams([O.Dense(nin, nin), O.Dense(nin, nout)]) == (nin * nin + nin) + (nin * nout + nout)
    assert O.net_nparams([O.RNNLayer(nin, nin), O.Dense(nin, nout)]) == (2 * nin * nin + nin) + (nin * nout + nout)
    assert O.net_nparams([O.LSTMLayer(nin, nin), O.Dense(nin, nout)]) == 4 * (2 * nin * nin + nin) + (nin * nout + nout)


def test_layer_bounds_cfg2():
    # cfg2 offsets (0-based)
    layers = [O.Dense(10, 64, O.RELU), O.Dense(64, 64, O.RELU), O.Dense(64, 1)]
    s, e = O.layer_bounds(layers)
    assert s == [0, 704, 4864] and e == [704, 4864, 4929]
    assert O.net_nparams(layers) == 4929


def test_dense_layout_column_major():
    # src/layers/dense.jl:14-24: theta = vcat(vec(W), vec(b)); W[o,i] = theta[o + out*i]
    l = O.Dense(3, 2)
    th = np.arange(8, dtype=np.float64)
    W, b = O.unpack_layer(l, th)
    assert W.shape == (2, 3)
    for o in range(2):
        for i in range(3):
            assert W[o, i] == th[o + 2 * i]
    assert list(b) == [6.0, 7.0]


def test_split_params():
    # test/bnn.jl:24-30
    m = O.Model((O.Dense(5, 10, O.RELU), O.Dense(10, 1)))
    th = np.arange(m.n_total, dtype=np.float32)
    tn, thyp, tl = O.split_params(m, th)
    assert tn.size == m.n_net and np.all(tn == th[:m.n_net])
    assert thyp.size == 0
    assert tl.size == 1 and tl[0] == th[-1]


# ------------------------------------------------------- likelihood goldens ---
DECILES = np.arange(1, 10) / 10.0
GOLD_NORMAL = -12.846623   # SURVEY.md section 4 (test/likelihoods.jl:24-31)
GOLD_TDIST = -13.489955    # test/likelihoods.jl:51-58, nu = 10


def _sigma_prior_at_sigma1():
    # logpdf(transformed(Gamma(2,2)), 0) = logpdf(Gamma(2, scale 2), 1) + 0
    return stats.gamma(a=2.0, scale=2.0).logpdf(1.0)


@pytest.mark.parametrize("net", ["dense", "rnn", "lstm"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_likelihood_normal_known_answer(net, dtype):
    rng = np.random.default_rng(SEED)
    if net == "dense":
        layers = (O.Dense(10, 10, O.SIGMOID), O.Dense(10, 1))
        x = rng.standard_normal((10, 9)).astype(dtype)
    else:
        first = O.RNNLayer(10, 10) if net == "rnn" else O.LSTMLayer(10, 10)
        layers = (first, O.Dense(10, 1))
        x = rng.standard_normal((10, 10, 9)).astype(dtype)   # test/likelihoods.jl:88
    m = O.Model(layers, O.Likelihood(O.NORMAL, prior_kind=O.PRIOR_GAMMA, prior_a=2.0, prior_b=2.0))
    y = stats.norm.ppf(DECILES).astype(dtype)
    th = np.zeros(m.n_net, dtype)
    val = O.loglike(m, x, y, th, np.zeros(1, dtype))
    expect = stats.norm.logpdf(y.astype(np.float64)).sum() + _sigma_prior_at_sigma1()
    assert val == pytest.approx(expect, rel=1e-6)
    assert val == pytest.approx(GOLD_NORMAL, abs=2e-5)


@pytest.mark.parametrize("net", ["dense", "rnn", "lstm"])
def test_likelihood_tdist_known_answer(net):
    rng = np.random.default_rng(SEED)
    dtype = np.float32
    if net == "dense":
        layers = (O.Dense(10, 10, O.SIGMOID), O.Dense(10, 1))
        x = rng.standard_normal((10, 9)).astype(dtype)
    else:
        first = O.RNNLayer(10, 10) if net == "rnn" else O.LSTMLayer(10, 10)
        layers = (first, O.Dense(10, 1))
        x = rng.standard_normal((10, 10, 9)).astype(dtype)
    m = O.Model(layers, O.Likelihood(O.TDIST, nu=10.0, prior_kind=O.PRIOR_GAMMA, prior_a=2.0, prior_b=2.0))
    y = stats.t(10.0).ppf(DECILES).astype(dtype)
    val = O.loglike(m, x, y, np.zeros(m.n_net, dtype), np.zeros(1, dtype))
    expect = stats.t(10.0).logpdf(y.astype(np.float64)).sum() + _sigma_prior_at_sigma1()
    assert val == pytest.approx(expect, rel=1e-6)
    assert val == pytest.approx(GOLD_TDIST, abs=2e-5)


# ------------------------------------------------------------ prior goldens ---
@pytest.mark.parametrize("sigma0,gold", [(0.5, -7.732122), (1.0, -9.695447),
                                         (3.0, -18.316291), (10.0, -29.007963)])
def test_gaussian_prior_known_answer(sigma0, gold):
    # test/networkpriors.jl:18-20
    th = (np.arange(1, 10) / 10.0).astype(np.float32)
    val = O.logprior(O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=sigma0), th)
    expect = stats.norm(0, sigma0).logpdf(np.arange(1, 10) / 10.0).sum()
    assert val == pytest.approx(expect, rel=1e-6)
    assert val == pytest.approx(gold, abs=2e-5)


def test_mixture_prior_is_weighted_sum_of_logs():
    # src/netpriors/mixturescale.jl:32  (parity unpinned by the reference's tests)
    th = np.linspace(-1, 1, 7)
    pr = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.3)
    expect = 0.3 * stats.norm(0, 1.0).logpdf(th).sum() + 0.7 * stats.norm(0, 0.1).logpdf(th).sum()
    assert O.logprior(pr, th) == pytest.approx(expect, rel=1e-12)


# -------------------------------------------- analytic gradient, test/bnn.jl ---
def test_analytic_gradient_2_2_1_sigmoid():
    """Line-by-line port of test/bnn.jl:33-107 (Float64, one observation)."""
    rng = np.random.default_rng(SEED)
    x = rng.standard_normal((2, 1))
    y = rng.standard_normal()
    layers = (O.Dense(2, 2, O.SIGMOID), O.Dense(2, 1))
    m = O.Model(layers, O.Likelihood(O.NORMAL, prior_kind=O.PRIOR_GAMMA, prior_a=2.0, prior_b=2.0),
                O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=1.0))
    th = rng.standard_normal(m.n_total)
    tn, _, tl = O.split_params(m, th)
    (W1, b1), (W2, b2) = (O.unpack_layer(l, tn[s:e]) for l, (s, e) in
                          zip(layers, zip(*O.layer_bounds(layers))))
    sig = lambda v: 1 / (1 + np.exp(-v))
    z = (W1 @ x + b1[:, None]).reshape(-1)
    a = sig(z)
    yhat = float((W2 @ a + b2)[0])
    dW2 = a.reshape(1, 2)
    db2 = 1.0
    da = W2.reshape(-1)
    dz = da * sig(z) * (1 - sig(z))
    dW1 = np.vstack([dz[0] * x.reshape(1, 2), dz[1] * x.reshape(1, 2)])
    db1 = dz
    dyhat_dnet = np.concatenate([dW1.reshape(-1, order="F"), db1, dW2.reshape(-1, order="F"), [db2]])
    sigma = math.exp(tl[0])
    dlike_sigma = -1 / sigma + (y - yhat) ** 2 / sigma ** 3
    dlike_tl = dlike_sigma * sigma
    dlike_yhat = (y - yhat) / sigma ** 2
    # sigma-prior term: d/ds [logGamma(2, scale 2)(e^s) + s]  (reference "trusts Zygote" here)
    eps = 1e-6
    f = lambda s: stats.gamma(a=2.0, scale=2.0).logpdf(math.exp(s)) + s
    dprior_tl = (f(tl[0] + eps) - f(tl[0] - eps)) / (2 * eps)
    expect = np.concatenate([dlike_yhat * dyhat_dnet - tn, [dlike_tl + dprior_tl]])
    v, g = O.grad_loglikeprior(m, th, x, np.array([y]))
    np.testing.assert_allclose(g, expect, rtol=1e-6, atol=1e-8)
    assert v == pytest.approx(O.loglikeprior(m, th, x, np.array([y])), rel=1e-12)


# ------------------------------------------ autograd cross-check (all rows) ---
NETS = {
    "linreg": ((O.Dense(5, 1),), 2),
    "mlp_relu": ((O.Dense(10, 16, O.RELU), O.Dense(16, 16, O.RELU), O.Dense(16, 1)), 2),
    "mlp_tanh_sig": ((O.Dense(4, 7, O.TANH), O.Dense(7, 5, O.SIGMOID), O.Dense(5, 1)), 2),
    "rnn": ((O.RNNLayer(3, 6), O.Dense(6, 1)), 3),
    "lstm": ((O.LSTMLayer(4, 8), O.Dense(8, 1)), 3),
    "lstm_only": ((O.LSTMLayer(1, 1),), 3),      # test/derivatives.jl:23-38
    "rnn_only": ((O.RNNLayer(1, 1),), 3),
    "lstm_mlp_head": ((O.LSTMLayer(2, 5), O.Dense(5, 4, O.TANH), O.Dense(4, 1)), 3),
}
LIKES = {
    "normal_gamma": O.Likelihood(O.NORMAL, prior_kind=O.PRIOR_GAMMA, prior_a=2.0, prior_b=0.5),
    "normal_invgamma": O.Likelihood(O.NORMAL, prior_kind=O.PRIOR_INVGAMMA, prior_a=2.0, prior_b=0.5),
    "tdist_gamma": O.Likelihood(O.TDIST, nu=5.0, prior_kind=O.PRIOR_GAMMA, prior_a=2.0, prior_b=2.0),
}
PRIORS = {
    "gauss": O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=0.5),
    "mix": O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5),
}


def make_case(net, like, prior, B=7, T=6, seed=SEED, dtype=np.float64):
    layers, nd = NETS[net]
    m = O.Model(layers, LIKES[like], PRIORS[prior])
    rng = np.random.default_rng(seed)
    d = layers[0].nin
    x = rng.standard_normal((d, B) if nd == 2 else (T, d, B)).astype(dtype)
    y = rng.standard_normal(B).astype(dtype)
    th = (0.5 * rng.standard_normal(m.n_total)).astype(dtype)
    return m, th, x, y


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("like", list(LIKES))
@pytest.mark.parametrize("prior", list(PRIORS))
def test_gradient_matches_autograd(net, like, prior):
    m, th, x, y = make_case(net, like, prior)
    nb = 13.0
    v, g = O.grad_loglikeprior(m, th, x, y, nb)
    vt, gt = torch_ref.value_and_grad(m, th, x, y, nb)
    assert v == pytest.approx(vt, rel=1e-10)
    np.testing.assert_allclose(g, gt, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("net", ["mlp_relu", "lstm"])
def test_float32_path_close_to_float64(net):
    m, th, x, y = make_case(net, "normal_gamma", "gauss", B=32)
    v64, g64 = O.grad_loglikeprior(m, th, x, y, 4.0)
    v32, g32 = O.grad_loglikeprior(m, th.astype(np.float32), x.astype(np.float32), y.astype(np.float32), 4.0)
    assert g32.dtype == np.float32
    assert v32 == pytest.approx(v64, rel=1e-5)
    assert np.linalg.norm(g32 - g64) / np.linalg.norm(g64) < 1e-5


def test_clip_gradient():
    # src/utils/gradient_utils.jl:6-10
    g = np.array([3.0, 4.0], np.float32)
    np.testing.assert_allclose(O.clip_gradient(g, 2.5), [1.5, 2.0], rtol=1e-6)
    assert O.clip_gradient(g, 5.0) is g and O.clip_gradient(g, 6.0) is g


def test_make_rnn_tensor():
    mat = np.arange(12, dtype=np.float32).reshape(6, 2)
    t = O.make_rnn_tensor(mat, 4)
    assert t.shape == (4, 2, 3)
    np.testing.assert_array_equal(t[:, :, 1], mat[1:5])


# ------------------------------------------------------ driver bookkeeping ---
def test_calculate_epochs_pinned():
    # test/sgld.jl:33 and test/hmc.jl:50
    s = O.SGLD()
    s.nsampled = 20_000
    assert s.calculate_epochs(100, 25_000, continue_sampling=True) == 50
    for cls in (lambda: O.SGNHT(1e-3), lambda: O.SGNHTS(1e-3), lambda: O.GGMC()):
        q = cls()
        q.nsampled = 20_000
        assert q.calculate_epochs(100, 25_000, continue_sampling=True) == 50
    h = O.HMC(1e-4, 5)
    h.nsampled = 20_000
    assert h.calculate_epochs(100, 25_000, continue_sampling=True) == 50 * 5


def test_num_batches():
    assert O.num_batches_of(500, 50) == 10
    assert O.num_batches_of(505, 50) == 11
    assert O.num_batches_of(505, 50, partial=False) == 10
    assert O.batch_slices(105, 50)[-1] == (100, 105)


def _linreg(n=400, seed=SEED):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((5, n)).astype(np.float32)
    beta = rng.standard_normal(5).astype(np.float32)
    y = (x.T @ beta + rng.standard_normal(n)).astype(np.float32)
    m = O.Model((O.Dense(5, 1),), O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5),
                O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=10.0))
    return m, x, y, beta


@pytest.mark.parametrize("mk", [
    lambda: O.SGLD(stepsize_a=1.0),
    lambda: O.SGNHT(2e-4, 3.6, xi=3.6),
    lambda: O.SGNHTS(1e-3, 6.0, madapter=O.RMSPropMassAdapter(30)),
    lambda: O.GGMC(beta=0.1, l=1e-2, steps=2, sadapter=O.DualAveragingStepSize(1e-2, target_accept=0.25, adapt_steps=40),
                   madapter=O.DiagCovMassAdapter(40, 10, kappa=0.1, epsilon=1e-8)),
    lambda: O.HMC(1e-3, 3, sadapter=O.DualAveragingStepSize(1e-3, target_accept=0.5, adapt_steps=20),
                  madapter=O.DiagCovMassAdapter(20, 8, kappa=0.1)),
])
def test_resume_prefix_equality(mk):
    """test/sgld.jl:34-35: continuing a run keeps the already-drawn prefix."""
    m, x, y, _ = _linreg(200)
    th0 = np.zeros(m.n_total, np.float32)
    s = mk()
    ch = O.mcmc(m, x, y, 50, 60, s, th0, np.random.default_rng(1)).copy()
    ch2 = O.mcmc(m, x, y, 50, 90, s, th0, np.random.default_rng(2), continue_sampling=True)
    assert ch2.shape == (m.n_total, 90)
    np.testing.assert_array_equal(ch2[:, :60], ch)
    assert np.all(np.isfinite(ch2))


def test_sgld_posterior_linear_regression():
    """Statistical row of test/sgld.jl:6-31, scaled down (n=2000, 6000 steps)."""
    rng = np.random.default_rng(SEED)
    n = 2000
    x = rng.standard_normal((5, n)).astype(np.float32)
    beta = rng.standard_normal(5).astype(np.float32)
    y = (x.T @ beta + rng.standard_normal(n)).astype(np.float32)
    m = O.Model((O.Dense(5, 1),), O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5),
                O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=10.0))
    s = O.SGLD(stepsize_a=1.0)
    th0 = (0.5 * rng.standard_normal(m.n_total)).astype(np.float32)
    ch = O.mcmc(m, x, y, 200, 6000, s, th0, rng)[:, -3000:]
    mean = ch.mean(axis=1)
    bhat = np.linalg.lstsq(x.T.astype(np.float64), y.astype(np.float64), rcond=None)[0]
    assert np.max(np.abs(mean[:5] - bhat)) < 0.05
    assert abs(mean[5]) < 0.1
    assert 0.9 <= np.exp(ch[6]).mean() <= 1.1
