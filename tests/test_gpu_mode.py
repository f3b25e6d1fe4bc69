"""find_mode: FluxModeFinder + Flux optimisers on the device vs the oracle.
src/inference/mode/abstract.jl:36-86, src/inference/mode/flux.jl:8-54, test/modes.jl:6-30."""
import numpy as np
import pytest

from oracle import bfx_oracle as O
from tests.test_gpu_parity import LIKES, PRIORS, build, engine_draws

pytestmark = pytest.mark.gpu
SEED = 6150533


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


def _opts(bf):
    return {"adam": (bf.ADAM(0.01), O.FluxOptimiser(O.OPT_ADAM, 0.01)),
            "rmsprop": (bf.RMSProp(0.005), O.FluxOptimiser(O.OPT_RMSPROP, 0.005)),
            "momentum": (bf.Momentum(1e-5, 0.9), O.FluxOptimiser(O.OPT_MOMENTUM, 1e-5, 0.9)),
            "descent": (bf.Descent(1e-5), O.FluxOptimiser(O.OPT_DESCENT, 1e-5))}


@pytest.mark.parametrize("opt", ["adam", "rmsprop", "momentum", "descent"])
@pytest.mark.parametrize("net", ["linreg", "mlp"])
def test_find_mode_replays_the_oracle(bf, opt, net):
    """The device run of find_mode equals the oracle's FluxModeFinder loop on the same batch schedule."""
    import ctypes as C
    from bayesflux_b200 import _lib as L
    from bayesflux_b200 import api as A
    rng = np.random.default_rng(SEED)
    layers = (O.Dense(5, 1),) if net == "linreg" else (O.Dense(5, 16, O.TANH), O.Dense(16, 1))
    n, B, epochs = 400, 100, 6
    x = rng.standard_normal((5, n)).astype(np.float32)
    beta = rng.standard_normal(5).astype(np.float32)
    y = (x.T @ beta + 0.1 * rng.standard_normal(n)).astype(np.float32)
    bnn, m = build(bf, layers, O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=2.0), O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=10.0), x, y)
    th0 = (0.3 * rng.standard_normal(m.n_total)).astype(np.float32)
    eng_opt, ora_opt = _opts(bf)[opt]
    fm = bf.FluxModeFinder(bnn, eng_opt, windowlength=7, eps=1e-6)
    th_e = bf.find_mode(bnn, B, epochs, fm, theta_start=th0, seed=3)
    # oracle on the engine's own batch schedule (debug hook)
    run = A._run_desc(B, 0, True, True, False, 1, 3, "per_chain", 0)
    nb = n // B
    perms = {}

    def perm_of(e):
        if e not in perms:
            parts = [engine_draws(bf, fm, run, 0, e * nb + k, m.n_total)[0] for k in range(nb)]
            perms[e] = np.concatenate(parts)
        return perms[e]
    th_o, conv = O.find_mode(m, x, y, B, epochs, O.FluxModeFinder(ora_opt, windowlength=7, eps=1e-6), th0, perms=perm_of)
    err = float(np.max(np.abs(th_e - th_o))) / max(1.0, float(np.max(np.abs(th_o))))
    assert err < 2e-4, err
    assert bool(fm.converged[0]) == conv


def test_find_mode_convergence_window_freezes_the_chain(bf):
    rng = np.random.default_rng(SEED + 1)
    layers = (O.Dense(3, 1),)
    n = 64
    x = rng.standard_normal((3, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    th0 = (0.1 * rng.standard_normal((m.n_total, 3))).astype(np.float32)
    # a huge tolerance: the first window holds |Inf - theta| (prevθ starts at Inf, flux.jl:27), so convergence is
    # declared at the second window check (step 2*windowlength + 1) and theta is frozen from then on
    fm = bf.FluxModeFinder(bnn, bf.ADAM(0.01), windowlength=4, eps=1e3)
    th_a = bf.find_mode(bnn, 64, 12, fm, theta_start=th0, nchains=3)
    assert fm.converged.all()
    fm2 = bf.FluxModeFinder(bnn, bf.ADAM(0.01), windowlength=4, eps=1e3)
    th_b = bf.find_mode(bnn, 64, 50, fm2, theta_start=th0, nchains=3)
    np.testing.assert_array_equal(th_a, th_b)


def test_mode_linear_regression_like_the_reference(bf):
    """test/modes.jl:6-30: k = 5, n = 100 000, noise 0.1, ADAM, batch 10 000, <= 1000 epochs, windowlength 50;
    maximum(abs, beta - betahat) < 0.01.  Several restarts at once."""
    rng = np.random.default_rng(SEED + 2)
    k, n = 5, 100_000
    x = rng.standard_normal((k, n)).astype(np.float32)
    beta = rng.standard_normal(k).astype(np.float32)
    y = (x.T @ beta + 0.1 * rng.standard_normal(n)).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(k, 1)))
    like = bf.FeedforwardNormal(nc, bf.Gamma(2.0, 2.0))
    prior = bf.GaussianPrior(nc, 10.0)
    init = bf.InitialiseAllSame(bf.Normal(0.0, 1.0), like, prior, seed=1)
    bnn = bf.BNN(x, y, like, prior, init)
    opt = bf.FluxModeFinder(bnn, bf.ADAM(), windowlength=50)
    th = bf.find_mode(bnn, 10_000, 1000, opt, nchains=4, seed=5)
    for c in range(4):
        assert np.max(np.abs(beta - th[:k, c])) < 0.01
