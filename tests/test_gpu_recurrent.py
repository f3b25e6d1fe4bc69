"""Recurrent path (RNN / LSTM + Dense head, SeqToOne likelihoods) against the oracle, through the C ABI.
Reference rows: SURVEY.md section 8a a3, a4, a10, a11 (src/layers/recurrent.jl:13-68,
src/likelihoods/seq_to_one.jl:31-44,89-100)."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import bfx_oracle as O
from tests.test_gpu_parity import LIKES, PRIORS, build, engine_sampler, oracle_sampler

pytestmark = pytest.mark.gpu
SEED = 6150533
TOL = 1e-4

NETS = {
    "rnn_small": (O.RNNLayer(3, 6), O.Dense(6, 1)),
    "rnn_sigmoid_head2": (O.RNNLayer(2, 5, O.SIGMOID), O.Dense(5, 4, O.TANH), O.Dense(4, 1)),
    "lstm_small": (O.LSTMLayer(4, 8), O.Dense(8, 1)),
    "lstm_odd": (O.LSTMLayer(3, 7), O.Dense(7, 3, O.RELU), O.Dense(3, 1)),
    "lstm_cfg4": (O.LSTMLayer(4, 64), O.Dense(64, 1)),
    "derivatives_rnn": (O.RNNLayer(1, 1), O.Dense(1, 1)),     # test/derivatives.jl:23-29 (+ the head the likelihood needs)
    "derivatives_lstm": (O.LSTMLayer(1, 1), O.Dense(1, 1)),   # test/derivatives.jl:31-37
}


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("like", ["normal_gamma", "tdist_gamma"])
@pytest.mark.parametrize("B", [5, 16, 40])
def test_recurrent_grad_and_value_parity(bf, net, like, B):
    rng = np.random.default_rng(SEED)
    layers = NETS[net]
    T, n = (10 if net.startswith("derivatives") else 7), 100
    x = rng.standard_normal((T, layers[0].nin, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES[like], PRIORS["gauss" if B != 40 else "mix"], x, y)
    C = 2
    theta = (0.4 * rng.standard_normal((m.n_total, C))).astype(np.float32)
    idx = np.stack([rng.permutation(n)[:B] for _ in range(C)], axis=1)
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=3.0)
    v2 = bf.loglikeprior(bnn, theta, idx, num_batches=3.0)
    for c in range(C):
        vo, go = O.grad_loglikeprior(m, theta[:, c].astype(np.float64), x[:, :, idx[:, c]].astype(np.float64),
                                     y[idx[:, c]].astype(np.float64), 3.0)
        assert np.linalg.norm(g[:, c] - go) <= TOL * np.linalg.norm(go), f"chain {c}"
        assert abs(v[c] - vo) <= TOL * abs(vo)
        assert abs(v2[c] - vo) <= TOL * abs(vo)


def test_recurrent_full_data_matches(bf):
    rng = np.random.default_rng(SEED + 1)
    layers = NETS["lstm_small"]
    x = rng.standard_normal((9, 4, 70)).astype(np.float32)
    y = rng.standard_normal(70).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    th = (0.3 * rng.standard_normal(m.n_total)).astype(np.float32)
    v, g = bf.grad_loglikeprior(bnn, th, None, 1.0)
    vo, go = O.grad_loglikeprior(m, th.astype(np.float64), x.astype(np.float64), y.astype(np.float64), 1.0)
    assert abs(v - vo) <= TOL * abs(vo)
    assert np.linalg.norm(g - go) <= TOL * np.linalg.norm(go)


@pytest.mark.parametrize("cell", ["rnn", "lstm"])
def test_reference_known_answers_recurrent(bf, cell):
    """test/likelihoods.jl:69-133: Chain(RNN(10,10), Dense(10,1)) / Chain(LSTM(10,10), Dense(10,1)), x 10x10x9,
    theta = 0 => yhat = 0, sigma = 1."""
    rng = np.random.default_rng(SEED)
    first = O.RNNLayer(10, 10) if cell == "rnn" else O.LSTMLayer(10, 10)
    layers = (first, O.Dense(10, 1))
    dec = np.arange(1, 10) / 10.0
    x = rng.standard_normal((10, 10, 9)).astype(np.float32)
    prior = O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=1.0)
    for like, y, gold in ((O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=2.0), stats.norm.ppf(dec), -12.846623),
                          (O.Likelihood(O.TDIST, nu=10.0, prior_a=2.0, prior_b=2.0), stats.t(10.0).ppf(dec), -13.489955)):
        bnn, m = build(bf, layers, like, prior, x, y.astype(np.float32))
        v = bf.loglikeprior(bnn, np.zeros(m.n_total, np.float32))
        assert v - m.n_net * (-0.5 * math.log(2 * math.pi)) == pytest.approx(gold, abs=5e-4)


@pytest.mark.parametrize("kind", ["sgld", "sgnhts_rms", "ggmc", "hmc"])
@pytest.mark.parametrize("net", ["rnn_small", "lstm_small"])
def test_recurrent_injected_step_parity(bf, kind, net):
    rng = np.random.default_rng(SEED + 3)
    layers = NETS[net]
    T, n, B, C, nsteps = 6, 96, 24, 2, 16
    x = rng.standard_normal((T, layers[0].nin, n)).astype(np.float32)
    y = (x[-1, 0] * 0.5 + 0.3 * rng.standard_normal(n)).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    P = m.n_total
    theta0 = (0.1 * rng.standard_normal((P, C))).astype(np.float32)
    nb = n // B
    oracles = [O.SGLD(stepsize_a=0.01) if kind == "sgld" else oracle_sampler(kind, n) for _ in range(C)]
    eng = engine_sampler(bf, oracles[0])
    eng.initialise(bnn, theta0)
    th_o = [theta0[:, c].copy() for c in range(C)]
    for so in oracles:
        so.initialise(theta0[:, 0].copy(), nsteps)
    full_lp = lambda th: O.loglikeprior(m, th, x, y, 1.0)
    nd = 2 if kind.startswith("ggmc") else 1
    for t in range(nsteps):
        idx = np.stack([rng.permutation(n)[:B] for _ in range(C)], axis=1).astype(np.int64)
        normals = rng.standard_normal((nd, P, C)).astype(np.float32)
        uni = rng.random(C).astype(np.float32)
        th_e = eng.step_injected(idx, float(nb), normals, uni)
        for c in range(C):
            xb, yb = x[:, :, idx[:, c]], y[idx[:, c]]
            gf = lambda th, xb=xb, yb=yb: O.grad_loglikeprior(m, th, xb, yb, float(nb))
            th_o[c] = oracles[c].update(th_o[c], gf, full_lp, [normals[d, :, c] for d in range(nd)], [uni[c]])
            scale = max(1.0, float(np.max(np.abs(th_o[c]))))
            err = float(np.max(np.abs(th_e[:, c] - th_o[c]))) / scale
            assert err < 2e-4, f"{kind}/{net}: step {t} chain {c} err {err}"


def test_cfg4_ggmc_run_and_resume(bf):
    """configs[3]: LSTM(4,64)+Dense(64,1), SeqToOneNormal, GGMC with the reference test's settings
    (test/ggmc.jl:22-33), shortened; the run is finite and resuming reproduces the prefix exactly."""
    rng = np.random.default_rng(SEED + 4)
    T, n, B, C = 50, 256, 64, 3
    series = np.cumsum(rng.standard_normal((T + n, 4)), axis=0).astype(np.float32) * 0.05
    x = np.stack([series[i:i + T] for i in range(n)], axis=2).astype(np.float32)   # T x d x n (make_rnn_tensor layout)
    y = series[T:T + n, 0].astype(np.float32)
    layers = NETS["lstm_cfg4"]
    bnn, m = build(bf, layers, O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5), O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=0.5), x, y)
    theta0 = (0.05 * rng.standard_normal((m.n_total, C))).astype(np.float32)

    def sampler():
        return bf.GGMC(beta=0.1, l=1e-2, steps=1, sadapter=bf.DualAveragingStepSize(1e-2, target_accept=0.25, adapt_steps=1000),
                       madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1, epsilon=1e-8))
    s1 = sampler()
    ch = bf.mcmc(bnn, B, 12, s1, theta_start=theta0, seed=5)
    assert ch.shape == (m.n_total, 12, C) and np.isfinite(ch).all()
    s2 = sampler()
    ch8 = bf.mcmc(bnn, B, 8, s2, theta_start=theta0, seed=5)
    np.testing.assert_array_equal(ch8, ch[:, :8])
    assert s2.calculate_epochs(n // B, 12, continue_sampling=True) == 1
