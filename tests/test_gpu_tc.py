"""Tensor-core (tcgen05, 3-term BF16 split) path against the float64 oracle, through the C ABI.
Same 1e-4 bar as the FP32 gate path (BASELINE.json north star)."""
import numpy as np
import pytest

from oracle import bfx_oracle as O

pytestmark = pytest.mark.gpu
SEED = 6150533
TOL = 1e-4

ACT = {"relu": O.RELU, "tanh": O.TANH, "sigmoid": O.SIGMOID, "identity": O.IDENTITY}

NETS = {
    "cfg2": [(10, 64, "relu"), (64, 64, "relu"), (64, 1, "identity")],
    "tanh3": [(5, 32, "tanh"), (32, 48, "sigmoid"), (48, 16, "tanh"), (16, 1, "identity")],
    "one_hidden": [(7, 16, "relu"), (16, 1, "identity")],
    "wide_in": [(64, 64, "tanh"), (64, 64, "relu"), (64, 1, "tanh")],
}


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


def build(bf, net, like, prior, x, y, compute="tc"):
    spec = NETS[net]
    nc = bf.destruct(bf.Chain(*[bf.Dense(i, o, ACT[a]) for (i, o, a) in spec]))
    if like == "normal":
        L = bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5))
        ol = O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5)
    else:
        L = bf.FeedforwardTDist(nc, bf.InverseGamma(2.0, 0.5), 5.0)
        ol = O.Likelihood(O.TDIST, nu=5.0, prior_kind=O.PRIOR_INVGAMMA, prior_a=2.0, prior_b=0.5)
    if prior == "gauss":
        P = bf.GaussianPrior(nc, 0.8)
        op = O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=0.8)
    else:
        P = bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5)
        op = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5)
    m = O.Model(tuple(O.Dense(i, o, ACT[a]) for (i, o, a) in spec), ol, op)
    return bf.BNN(x, y, L, P, compute=compute), m


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("like", ["normal", "tdist"])
@pytest.mark.parametrize("B", [64, 150, 17])
def test_tc_grad_and_value_parity(bf, net, like, B):
    rng = np.random.default_rng(SEED)
    d = NETS[net][0][0]
    n = 400
    x = rng.standard_normal((d, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, net, like, "gauss" if B != 150 else "mix", x, y)
    assert bnn.compute == "tc"
    C = 3
    theta = (0.3 * rng.standard_normal((bnn.num_total_params, C))).astype(np.float32)
    idx = np.stack([rng.permutation(n)[:B] for _ in range(C)], axis=1)
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=6.0)
    v2 = bf.loglikeprior(bnn, theta, idx, num_batches=6.0)
    for c in range(C):
        vo, go = O.grad_loglikeprior(m, theta[:, c].astype(np.float64), x[:, idx[:, c]].astype(np.float64),
                                     y[idx[:, c]].astype(np.float64), 6.0)
        rel = np.linalg.norm(g[:, c] - go) / np.linalg.norm(go)
        assert rel < TOL, f"chain {c}: gradient rel err {rel}"
        # the same max-abs bound the FP32 matrix holds (tests/test_gpu_parity.py): no single entry may be off
        assert np.max(np.abs(g[:, c] - go)) <= TOL * max(1.0, np.max(np.abs(go))), f"chain {c}: max-abs gradient error"
        assert abs(v[c] - vo) <= TOL * abs(vo)
        assert abs(v2[c] - vo) <= TOL * abs(vo)


def test_tc_full_data_value(bf):
    rng = np.random.default_rng(SEED + 1)
    n = 1000
    x = rng.standard_normal((10, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, "cfg2", "normal", "gauss", x, y)
    theta = (0.2 * rng.standard_normal(bnn.num_total_params)).astype(np.float32)
    v, g = bf.grad_loglikeprior(bnn, theta, None, num_batches=1.0)
    vo, go = O.grad_loglikeprior(m, theta.astype(np.float64), x.astype(np.float64), y.astype(np.float64), 1.0)
    assert abs(v - vo) <= TOL * abs(vo)
    assert np.linalg.norm(g - go) <= TOL * np.linalg.norm(go)


def test_tc_unsupported_nets_are_rejected(bf):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 50)).astype(np.float32)
    y = rng.standard_normal(50).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(5, 10, bf.relu), bf.Dense(10, 1)))  # width 10 is not a multiple of 16
    like = bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5))
    with pytest.raises(bf.BfxError) as e:
        bf.BNN(x, y, like, bf.GaussianPrior(nc, 1.0), compute="tc")
    assert e.value.code == bf._lib.ERR_UNSUPPORTED
    assert bf.BNN(x, y, like, bf.GaussianPrior(nc, 1.0), compute="auto").compute == "fp32"


@pytest.mark.parametrize("kind", ["sgld", "sgnhts", "ggmc", "hmc"])
def test_tc_sampler_steps_track_the_fp32_engine(bf, kind):
    """Same seed, same schedule: the tensor-core engine and the FP32 engine differ only by the
    rounding of the contractions, so a short run must agree closely."""
    rng = np.random.default_rng(SEED + 2)
    n = 2000
    x = rng.standard_normal((10, n)).astype(np.float32)
    y = (x[0] - 0.5 * x[1] + 0.1 * rng.standard_normal(n)).astype(np.float32)
    C = 4
    out = {}
    for compute in ("fp32", "tc"):
        bnn, _ = build(bf, "cfg2", "normal", "gauss", x, y, compute=compute)
        theta0 = (0.1 * np.random.default_rng(5).standard_normal((bnn.num_total_params, C))).astype(np.float32)
        if kind == "sgld":
            s = bf.SGLD(stepsize_a=1e-3)
        elif kind == "sgnhts":
            s = bf.SGNHTS(1e-4, 1.0)
        elif kind == "ggmc":
            s = bf.GGMC(l=1e-3, beta=0.5, steps=2, sadapter=bf.ConstantStepsize(1e-3), madapter=bf.FixedMassAdapter())
        else:
            s = bf.HMC(1e-4, 3, sadapter=bf.ConstantStepsize(1e-4), madapter=bf.FixedMassAdapter())
        out[compute] = bf.mcmc(bnn, 64, 12, s, theta_start=theta0, seed=11)
    a, b = out["fp32"], out["tc"]
    assert a.shape == b.shape
    rel = np.linalg.norm(a - b) / np.linalg.norm(a)
    assert rel < 2e-3, rel
