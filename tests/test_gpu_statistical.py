"""The reference's per-sampler statistical tests: linear regression k = 5, n = 10 000,
noise sd 1, Chain(Dense(5,1)), FeedforwardNormal(Gamma(2, 0.5)), GaussianPrior(10), batchsize 1000, 20 000 samples,
last 10 000 kept; thresholds  max|beta - betahat| < 0.05, |intercept| < 0.05, 0.9 <= mean(exp(theta_like)) <= 1.1
(test/sgnht.jl:7-65, test/sgnht-s.jl:5-65, test/ggmc.jl:5-76, test/hmc.jl:6-83), with the reference's sampler
settings and its MAP warm start through find_mode.  The reference repeats each test 10 times and asks for pass
rates > 0.9 / 0.9 / 0.8; here the repeats are independent chains of one run.  Resume: calculate_epochs and exact
prefix equality as in the reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 6150533
C = 10


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


@pytest.fixture(scope="module")
def problem(bf):
    rng = np.random.default_rng(SEED)
    k, n = 5, 10_000
    x = rng.standard_normal((k, n)).astype(np.float32)
    beta = rng.standard_normal(k).astype(np.float32)
    y = (x.T @ beta + rng.standard_normal(n)).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(k, 1)))
    like = bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5))
    prior = bf.GaussianPrior(nc, 10.0)
    init = bf.InitialiseAllSame(bf.Normal(0.0, 1.0), like, prior, seed=1)
    bnn = bf.BNN(x, y, like, prior, init)
    # MAP warm start: find_mode(bnn, 1000, 100, FluxModeFinder(bnn, ADAM()))  (test/hmc.jl:31-32)
    theta_map = bf.find_mode(bnn, 1000, 100, bf.FluxModeFinder(bnn, bf.ADAM()), nchains=C, seed=3)
    return bnn, beta, theta_map


def sampler_of(bf, kind):
    if kind == "sgnht":
        return bf.SGNHT(2e-4, 3.6, xi=3.6), False                                   # test/sgnht.jl:21-22
    if kind == "sgnhts":
        return bf.SGNHTS(1e-3, 6.0, madapter=bf.RMSPropMassAdapter(1000)), True     # test/sgnht-s.jl:21-25
    if kind.startswith("ggmc"):
        steps = int(kind[-1])
        return bf.GGMC(beta=0.1, l=1e-2, steps=steps,
                       sadapter=bf.DualAveragingStepSize(1e-2, adapt_steps=1000, target_accept=0.25),
                       madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1, epsilon=1e-8)), True   # test/ggmc.jl:22-33
    madapter = bf.DiagCovMassAdapter(1000, 100, kappa=0.1) if kind == "hmc_diag" else bf.RMSPropMassAdapter(1000)
    return bf.HMC(1e-4, 5, sadapter=bf.DualAveragingStepSize(1e-4, adapt_steps=1000, target_accept=0.5),
                  madapter=madapter), True                                          # test/hmc.jl:19-29,60-63


@pytest.mark.parametrize("kind", ["sgnht", "sgnhts", "ggmc1", "ggmc3", "hmc_diag", "hmc_rms"])
def test_sampler_recovers_the_regression_posterior(bf, problem, kind):
    bnn, beta, theta_map = problem
    s, warm = sampler_of(bf, kind)
    kw = dict(theta_start=theta_map) if warm else dict(nchains=C)
    ch = bf.mcmc(bnn, 1000, 20_000, s, seed=17, **kw)
    tail = ch[:, -10_000:, :]
    t1 = t2 = t3 = 0
    for c in range(C):
        mean = tail[:, :, c].mean(axis=1)
        t1 += bool(np.max(np.abs(beta - mean[:5])) < 0.05)
        t2 += bool(abs(mean[5]) < 0.05)
        t3 += bool(0.9 <= float(np.exp(tail[6, :, c]).mean()) <= 1.1)
    assert t1 > 0.9 * C - 1e-9 or t1 >= C - 1, f"{kind}: coefficients {t1}/{C}"
    assert t2 >= 0.9 * C - 1, f"{kind}: intercept {t2}/{C}"
    assert t3 >= 0.8 * C, f"{kind}: variance {t3}/{C}"
    # resume: calculate_epochs(sampler, 100, 25 000; continue_sampling = true) == 50 (x path_len for HMC) and the
    # continued run reproduces the first 20 000 samples exactly
    plen = getattr(s, "path_len", 1)
    assert s.calculate_epochs(100, 25_000, continue_sampling=True) == 50 * plen
    longer = bf.mcmc(bnn, 1000, 25_000, s, continue_sampling=True, seed=17)
    assert longer.shape[1] == 25_000
    np.testing.assert_array_equal(longer[:, :20_000, :], ch)
