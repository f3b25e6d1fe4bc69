"""BASELINE.json configs[1] at its full size (n = 10^6 observations, 1 024 chains, B = 64), checked through
size-independent properties instead of an O(n) oracle run: the epoch schedule is a permutation, the full-data
gradient is the sum of its partitions (tile accumulation over 15 625 tiles), the two engines agree, chains are
independent of how they are sharded over samplers (the multi-GPU contract), and a resumed run equals the
uninterrupted one bit for bit."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, CHAINS, B = 1_000_000, 1024, 64
SIGMA0, GA, GB = 1.0, 2.0, 0.5


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


@pytest.fixture(scope="module")
def models(bf):
    rng = np.random.default_rng(20261019)
    x = rng.standard_normal((10, N), dtype=np.float32)
    w = rng.standard_normal(10).astype(np.float32)
    y = (np.tanh(x.T @ w) + 0.1 * rng.standard_normal(N)).astype(np.float32)
    out = {}
    for compute in ("fp32", "tc"):
        nc = bf.destruct(bf.Chain(bf.Dense(10, 64, bf.relu), bf.Dense(64, 64, bf.relu), bf.Dense(64, 1)))
        out[compute] = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(GA, GB)), bf.GaussianPrior(nc, SIGMA0),
                              compute=compute)
    return out


def test_epoch_schedule_is_a_permutation_of_a_million(bf, models):
    from bayesflux_b200 import _lib as L
    bnn = models["fp32"]
    eng = bf.SGLD()
    eng.initialise(bnn, np.zeros((bnn.num_total_params, 2), np.float32))
    big = 125_000  # the position -> observation map does not depend on the batch size: read it in big slices
    run = L.RunDesc(big, 1, 1, 1, 0, 1, L.BATCH_PER_CHAIN, 0, 6150533, 0)
    perms = []
    for chain, epoch in ((0, 0), (0, 1), (1, 0)):
        parts = []
        for k in range(N // big):
            idx = np.empty(big, np.int64)
            bk = C.c_int64()
            L.check(L.lib.bfx_debug_batch_indices(eng._h, C.byref(run), chain, epoch * (N // big) + k,
                                                  idx.ctypes.data_as(C.c_void_p), C.byref(bk)))
            assert bk.value == big
            parts.append(idx)
        p = np.concatenate(parts)
        assert np.array_equal(np.sort(p), np.arange(N))  # every observation exactly once per epoch
        perms.append(p)
    assert not np.array_equal(perms[0], perms[1]) and not np.array_equal(perms[0], perms[2])
    assert np.mean(perms[0] == np.arange(N)) < 1e-3  # not (close to) the identity


def _prior_terms(theta):
    """net prior + sigma prior of one flat theta, closed forms (netpriors/gaussian.jl:25-28, feedforward.jl:36-42)."""
    net, sl = theta[:-1].astype(np.float64), float(theta[-1])
    p = net.size * (-0.5 * math.log(2 * math.pi) - math.log(SIGMA0)) - 0.5 * np.sum(net * net) / SIGMA0 ** 2
    s = -math.lgamma(GA) - GA * math.log(GB) + (GA - 1) * sl - math.exp(sl) / GB + sl
    return p + s


@pytest.mark.parametrize("compute", ["fp32", "tc"])
def test_full_data_gradient_is_the_sum_of_its_partitions(bf, models, compute):
    bnn = models[compute]
    P = bnn.num_total_params
    theta = (0.1 * np.random.default_rng(5).standard_normal(P)).astype(np.float32)
    v_full, g_full = bf.grad_loglikeprior(bnn, theta)
    parts = [np.arange(k * N // 4, (k + 1) * N // 4, dtype=np.int64) for k in range(4)]
    vs, gs = zip(*[bf.grad_loglikeprior(bnn, theta, idx=ix) for ix in parts])
    # each partition carries the priors once: v_full = sum_k v_k - 3 (prior terms)
    assert np.isclose(v_full, sum(vs) - 3.0 * _prior_terms(theta), rtol=2e-7)
    g_sum = np.sum(np.asarray(gs, np.float64), axis=0)
    g_sum[:-1] -= 3.0 * (-theta[:-1].astype(np.float64) / SIGMA0 ** 2)       # net prior gradient, closed form
    g_sum[-1] -= 3.0 * ((GA - 1) - math.exp(float(theta[-1])) / GB + 1.0)     # d/ds of the sigma prior + Jacobian
    err = np.linalg.norm(g_full - g_sum) / np.linalg.norm(g_sum)
    assert err < 1e-4, err  # tolerance of the gradient bar (north star): 1e-4 relative
    # num_batches is linear in the likelihood part
    v1 = bf.loglikeprior(bnn, theta, idx=parts[0], num_batches=1.0)
    v2 = bf.loglikeprior(bnn, theta, idx=parts[0], num_batches=2.0)
    v3 = bf.loglikeprior(bnn, theta, idx=parts[0], num_batches=3.0)
    assert np.isclose(v3 - v2, v2 - v1, rtol=1e-9)


def test_engines_agree_on_the_full_data_gradient(bf, models):
    P = models["tc"].num_total_params
    theta = (0.1 * np.random.default_rng(6).standard_normal((P, 3))).astype(np.float32)
    v32, g32 = bf.grad_loglikeprior(models["fp32"], theta)
    vtc, gtc = bf.grad_loglikeprior(models["tc"], theta)
    assert np.allclose(vtc, v32, rtol=1e-5)
    for c in range(3):
        assert np.linalg.norm(gtc[:, c] - g32[:, c]) / np.linalg.norm(g32[:, c]) < 1e-4


def test_chains_do_not_depend_on_how_they_are_sharded(bf, models):
    """1 024 chains in one sampler == two samplers of 512 chains with chain_offset 0 / 512 (what two ranks run)."""
    bnn = models["tc"]
    P = bnn.num_total_params
    th0 = np.asfortranarray((0.1 * np.random.default_rng(7).standard_normal((P, CHAINS))).astype(np.float32))
    whole = np.array(bf.mcmc(bnn, B, 8, bf.SGLD(), theta_start=th0, keep_every=8, seed=11), copy=True)
    halves = [np.array(bf.mcmc(bnn, B, 8, bf.SGLD(), theta_start=np.asfortranarray(th0[:, o:o + 512]), keep_every=8,
                               seed=11, chain_offset=o), copy=True) for o in (0, 512)]
    assert whole.shape == (P, 1, CHAINS) and np.isfinite(whole).all()
    assert np.array_equal(np.concatenate(halves, axis=2), whole)
    assert not np.array_equal(whole[:, 0, 0], whole[:, 0, 1])  # chains differ from each other


def test_resume_equals_the_uninterrupted_run(bf, models):
    bnn = models["tc"]
    P = bnn.num_total_params
    th0 = np.asfortranarray((0.1 * np.random.default_rng(8).standard_normal((P, CHAINS))).astype(np.float32))
    one = np.array(bf.mcmc(bnn, B, 16, bf.SGLD(), theta_start=th0, keep_every=8, seed=3), copy=True)
    s = bf.SGLD()
    bf.mcmc(bnn, B, 8, s, theta_start=th0, keep_every=8, seed=3)
    two = np.array(bf.mcmc(bnn, B, 16, s, continue_sampling=True, keep_every=8, seed=3), copy=True)
    assert one.shape == two.shape == (P, 2, CHAINS)
    assert np.array_equal(one, two)


# ---- configs[2] and configs[3] at their full shapes: additivity of the gradient over data partitions -------------
def _additivity(bf, bnn, parts, n, rtol=1e-4):
    """On theta_net:  g(all data) = sum_k g(partition k) - (K - 1) g(network prior).  The network prior part comes
    from the engine itself: g(k; nb) = nb * (likelihood + sigma prior)_k + prior (src/model/BNN.jl:50-56), hence
    prior = 2 g(0; nb = 1) - g(0; nb = 2).  (The sigma prior sits inside the likelihood term and only touches
    theta_like, which is left out of the comparison.)"""
    P = bnn.num_total_params
    theta = (0.05 * np.random.default_rng(9).standard_normal(P)).astype(np.float32)
    v_full, g_full = bf.grad_loglikeprior(bnn, theta)
    K = len(parts)
    vs, gs = zip(*[bf.grad_loglikeprior(bnn, theta, idx=ix) for ix in parts])
    v2, g2 = bf.grad_loglikeprior(bnn, theta, idx=parts[0], num_batches=2.0)
    p_net = 2.0 * np.asarray(gs[0], np.float64) - np.asarray(g2, np.float64)
    g_sum = np.sum(np.asarray(gs, np.float64), axis=0) - (K - 1) * p_net
    err = np.linalg.norm((g_full - g_sum)[:-1]) / np.linalg.norm(g_sum[:-1])
    assert err < rtol, err
    assert np.isfinite(v_full) and np.all(np.isfinite(vs)) and np.isfinite(v2)


@pytest.mark.parametrize("compute", ["fp32", "tc"])
def test_cfg3_wide_mlp_gradient_is_additive_over_chunks(bf, compute):
    """configs[2]: Dense(64,512,relu) x 3 + Dense(512,1), FeedforwardTDist + MixtureScalePrior, B = 8 192 per GPU:
    the large-P (stream) regime accumulates weight gradients over 8 192-column chunks."""
    rng = np.random.default_rng(12)
    n = 16384
    x = rng.standard_normal((64, n), dtype=np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(64, 512, bf.relu), bf.Dense(512, 512, bf.relu), bf.Dense(512, 512, bf.relu),
                              bf.Dense(512, 1)))
    bnn = bf.BNN(x, y, bf.FeedforwardTDist(nc, bf.Gamma(2.0, 0.5), 5.0), bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5),
                 compute=compute)
    assert bf.model_regime(bnn) == "stream"
    parts = [np.arange(0, 8192, dtype=np.int64), np.arange(8192, n, dtype=np.int64)]
    _additivity(bf, bnn, parts, n)


def test_cfg4_lstm_gradient_is_additive_over_partitions(bf):
    """configs[3]: LSTM(4,64) + Dense(64,1), T = 50, SeqToOneNormal, n = 16 384 sequences (the full-data passes of
    GGMC run exactly this loop over 1 024 tiles of 16 sequences)."""
    rng = np.random.default_rng(13)
    T, n = 50, 16384
    x = (0.5 * rng.standard_normal((T, 4, n))).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1)))
    bnn = bf.BNN(x, y, bf.SeqToOneNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 0.5))
    parts = [np.arange(k * n // 4, (k + 1) * n // 4, dtype=np.int64) for k in range(4)]
    _additivity(bf, bnn, parts, n)
