"""Oracle chain diagnostics (split-Rhat, Geyer ESS) against closed forms: the reference has no diagnostics of its own
(README.md:191-193 defers to MCMCChains), so these known-answer cases are what pins the restatement."""
import numpy as np

from oracle import bfx_oracle as O


def test_iid_draws_have_rhat_one_and_full_ess():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((6, 2000, 4)) * np.arange(1, 7)[:, None, None] + 2.0
    d = O.chain_diagnostics(x)
    assert np.allclose(d["mean"], 2.0, atol=0.15)
    assert np.allclose(d["var"], x.reshape(6, -1).var(axis=1, ddof=1), rtol=1e-10)
    assert np.all(np.abs(d["rhat"] - 1.0) < 0.01)
    assert np.all(d["ess"] > 0.8 * 8000) and np.all(d["ess"] < 1.25 * 8000)


def test_ar1_ess_matches_the_closed_form():
    rng = np.random.default_rng(4)
    phi, n, C = 0.9, 20000, 4
    e = rng.standard_normal((1, n, C))
    x = np.zeros_like(e)
    x[:, 0] = e[:, 0] / np.sqrt(1 - phi ** 2)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    d = O.chain_diagnostics(x)
    expect = n * C * (1 - phi) / (1 + phi)
    assert 0.7 * expect < d["ess"][0] < 1.4 * expect
    assert abs(d["rhat"][0] - 1.0) < 0.02
    capped = O.chain_diagnostics(x, max_lag=400)  # the sequence terminates long before lag 400
    assert np.isclose(capped["ess"][0], d["ess"][0], rtol=1e-12)


def test_shifted_chains_are_flagged_and_rhat_matches_the_direct_formula():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 400, 3))
    x[1, :, 2] += 3.0  # parameter 1: one chain sits elsewhere
    d = O.chain_diagnostics(x)
    assert abs(d["rhat"][0] - 1.0) < 0.05 and d["rhat"][1] > 1.5
    # BDA3 11.4 written out for parameter 1
    halves = [x[1, :200, c] for c in range(3)] + [x[1, 200:, c] for c in range(3)]
    nh, m = 200, 6
    means = np.array([h.mean() for h in halves])
    W = np.mean([h.var(ddof=1) for h in halves])
    B = nh * means.var(ddof=1)
    assert np.isclose(d["rhat"][1], np.sqrt(((nh - 1) / nh * W + B / nh) / W), rtol=1e-12)
    assert d["ess"][1] < 0.1 * 1200


def test_odd_n_drops_the_middle_draw_and_2d_input_is_one_chain():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((3, 101))
    a = O.chain_diagnostics(x)
    b = O.chain_diagnostics(np.delete(x, 50, axis=1)[:, :, None])
    for k in a:
        assert np.allclose(a[k], b[k], rtol=1e-12, equal_nan=True)
