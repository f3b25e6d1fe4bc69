"""Committed golden vectors (tests/golden/, generator: tests/golden/make_golden.py).

CPU: the oracle reproduces the frozen vectors and the reference's own known-answer values.
GPU: the CUDA path, called through the C ABI, matches the same vectors (1e-4 relative)."""
import json
import os

import numpy as np
import pytest

from oracle import bfx_oracle as O
from tests.golden import make_golden as G

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4  # FP32 relative tolerance of BASELINE.json's north star


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(HERE, "hotpath_v1.npz"))


@pytest.fixture(scope="module")
def known():
    return json.load(open(os.path.join(HERE, "reference_known_answers.json")))


ALL = [(c, l, p) for c in G.CASES for l in G.LIKES for p in G.PRIORS]


@pytest.mark.parametrize("case,like,prior", ALL)
def test_oracle_reproduces_frozen_vectors(vec, case, like, prior):
    layers, T, d = G.CASES[case]
    m = O.Model(tuple(layers), G.LIKES[like], G.PRIORS[prior])
    key = f"{case}/{like}/{prior}"
    x, y = vec[f"{case}/x"].astype(np.float64), vec[f"{case}/y"].astype(np.float64)
    v, g = O.grad_loglikeprior(m, vec[key + "/theta"].astype(np.float64), x, y, 7.0)
    assert v == pytest.approx(float(vec[key + "/value"]), rel=1e-12)
    np.testing.assert_allclose(g, vec[key + "/grad"], rtol=1e-10, atol=1e-12)


def test_reference_known_answers(known):
    # test/likelihoods.jl:24-31,51-58 through the oracle
    layers = (O.Dense(10, 10, O.SIGMOID), O.Dense(10, 1))
    rng = np.random.default_rng(G.SEED)
    x = rng.standard_normal((10, 9))
    for kind, nu, ykey, gkey in ((O.NORMAL, 0.0, "y_deciles", "likelihood_normal"),
                                 (O.TDIST, 10.0, "y_tdist10_deciles", "likelihood_tdist_nu10")):
        m = O.Model(layers, O.Likelihood(kind, nu=nu, prior_a=2.0, prior_b=2.0))
        val = O.loglike(m, x, np.array(known[ykey]), np.zeros(m.n_net), np.zeros(1))
        assert val == pytest.approx(known[gkey], rel=1e-9)
    # test/networkpriors.jl:18-20
    for s, gold in known["gaussian_prior"].items():
        assert O.logprior(O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=float(s)), np.arange(1, 10) / 10.0) == \
            pytest.approx(gold, rel=1e-9)
    # test/sgld.jl:33, test/hmc.jl:50
    s = O.SGLD()
    s.nsampled = 20000
    assert s.calculate_epochs(100, 25000, continue_sampling=True) == known["calculate_epochs"]["sgld_100_25000_after_20000"]


# --------------------------------------------------------------------------- GPU ---
def _bf_model(bf, case, like, prior, x, y):
    layers, T, d = G.CASES[case]
    conv = []
    for l in layers:
        if l.kind == O.DENSE:
            conv.append(bf.Dense(l.nin, l.nout, l.act))
        elif l.kind == O.RNN:
            conv.append(bf.RNN(l.nin, l.nout, l.act))
        else:
            conv.append(bf.LSTM(l.nin, l.nout))
    nc = bf.destruct(bf.Chain(*conv))
    lk = G.LIKES[like]
    sp = bf.Gamma(lk.prior_a, lk.prior_b) if lk.prior_kind == O.PRIOR_GAMMA else bf.InverseGamma(lk.prior_a, lk.prior_b)
    if T:
        L = bf.SeqToOneNormal(nc, sp) if lk.kind == O.NORMAL else bf.SeqToOneTDist(nc, sp, lk.nu)
    else:
        L = bf.FeedforwardNormal(nc, sp) if lk.kind == O.NORMAL else bf.FeedforwardTDist(nc, sp, lk.nu)
    pr = G.PRIORS[prior]
    P = bf.GaussianPrior(nc, pr.sigma0) if pr.kind == O.NETPRIOR_GAUSSIAN else \
        bf.MixtureScalePrior(nc, pr.sigma1, pr.sigma2, pr.pi1)
    return bf.BNN(x, y, L, P)


@pytest.mark.gpu
@pytest.mark.parametrize("case,like,prior", ALL)
def test_cuda_matches_frozen_vectors(vec, case, like, prior):
    import bayesflux_b200 as bf
    key = f"{case}/{like}/{prior}"
    x, y = vec[f"{case}/x"], vec[f"{case}/y"]
    bnn = _bf_model(bf, case, like, prior, x, y)
    th = vec[key + "/theta"]
    v, g = bf.grad_loglikeprior(bnn, th, None, num_batches=7.0)
    gold_v, gold_g = float(vec[key + "/value"]), vec[key + "/grad"]
    assert abs(v - gold_v) <= TOL * abs(gold_v)
    assert np.linalg.norm(g - gold_g) <= TOL * np.linalg.norm(gold_g)
    v2 = bf.loglikeprior(bnn, th, None, num_batches=7.0)
    assert abs(v2 - gold_v) <= TOL * abs(gold_v)
