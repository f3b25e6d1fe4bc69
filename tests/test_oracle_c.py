"""The C restatement (oracle/bfx_oracle_c.c, the CPU-baseline arm of bench.py) against the numpy
oracle, which is itself pinned to the reference's known-answer tests (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import bfx_oracle as O
from oracle import cbind as OC

SEED = 6150533  # the reference's test seed, test/runtests.jl:10

NETS = {
    "cfg1_linreg": ([5, 1], [O.IDENTITY]),
    "cfg2_mlp": ([10, 64, 64, 1], [O.RELU, O.RELU, O.IDENTITY]),
    "sigmoid_tanh": ([7, 9, 5, 1], [O.SIGMOID, O.TANH, O.IDENTITY]),
}


def models(dims, acts, like, prior):
    layers = tuple(O.Dense(dims[i], dims[i + 1], acts[i]) for i in range(len(acts)))
    if like == "normal":
        lk, ck = O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5), dict(like_kind=0)
    else:
        lk, ck = O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5), dict(like_kind=1, nu=5.0)
    if prior == "gaussian":
        pr, cp = O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=0.7), dict(prior="gaussian", sigma0=0.7)
    else:
        pr = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.4)
        cp = dict(prior="mixture", sigma1=1.0, sigma2=0.1, pi1=0.4)
    return O.Model(layers, lk, pr), OC.make_model(dims, acts, sprior_a=2.0, sprior_b=0.5, **ck, **cp)


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("like", ["normal", "tdist"])
@pytest.mark.parametrize("prior", ["gaussian", "mixture"])
def test_c_gradient_matches_numpy_oracle(net, like, prior):
    dims, acts = NETS[net]
    mo, mc = models(dims, acts, like, prior)
    rng = np.random.default_rng(SEED)
    n, B = 300, 37
    x = rng.standard_normal((dims[0], n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    P = O.net_nparams(mo.layers) + 1
    theta = (0.3 * rng.standard_normal(P)).astype(np.float32)
    idx = rng.permutation(n)[:B]
    v, g = OC.grad(mc, theta, x, y, idx, 9.0)
    vo, go = O.grad_loglikeprior(mo, theta.astype(np.float64), x[:, idx].astype(np.float64),
                                 y[idx].astype(np.float64), 9.0)
    assert abs(v - vo) <= 1e-5 * abs(vo)
    assert np.linalg.norm(g - go) <= 1e-5 * np.linalg.norm(go)  # float32 arithmetic vs float64


def test_c_sgld_run_matches_numpy_oracle_with_injected_draws():
    dims, acts = NETS["cfg2_mlp"]
    mo, mc = models(dims, acts, "normal", "gaussian")
    rng = np.random.default_rng(SEED)
    n, B, Cn, nsteps = 256, 32, 3, 12
    x = rng.standard_normal((dims[0], n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    P = O.net_nparams(mo.layers) + 1
    theta0 = np.asfortranarray((0.1 * rng.standard_normal((P, Cn))).astype(np.float32))
    normals = np.asfortranarray(rng.standard_normal((P, Cn, nsteps)).astype(np.float32))
    batches = np.asfortranarray(rng.integers(0, n, (B, Cn, nsteps)).astype(np.int64))
    theta = theta0.copy(order="F")
    samples = OC.sgld_run(mc, x, y, theta, B, nsteps, normals=normals, batches=batches, record=True, nthreads=2)
    nb = float(O.num_batches_of(n, B))
    for c in range(Cn):
        s = O.SGLD()
        th = theta0[:, c].copy()
        s.initialise(th, nsteps)
        for t in range(nsteps):
            idx = batches[:, c, t]
            gf = lambda q, idx=idx: O.grad_loglikeprior(mo, q, x[:, idx], y[idx], nb)
            th = s.update(th, gf, None, [normals[:, c, t]])
        err = np.abs(samples[:, :, c] - s.samples).max()
        assert err < 2e-5, err
        assert np.array_equal(samples[:, -1, c], theta[:, c])
