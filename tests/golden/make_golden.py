#!/usr/bin/env python
"""Generates tests/golden/*.npz / *.json -- committed fixtures for the parity tests.

Two kinds of content:
  (1) reference_known_answers.json: the closed-form values the REFERENCE's own tests assert
      (test/likelihoods.jl:24-31,51-58; test/networkpriors.jl:18-20), evaluated here with scipy the
      way those tests evaluate them with Distributions.jl.
  (2) hotpath_v1.npz: seeded inputs and the float64 oracle's outputs (value, gradient, one sampler
      update with injected draws) for the MLP / RNN / LSTM x likelihood x prior cases.  The
      reference cannot be executed in the build image (no Julia), so these vectors come from the
      oracle AFTER it has been pinned by (1) and by tests/test_oracle_golden.py; they freeze the
      oracle (any later edit that changes a number fails tests/test_golden_fixtures.py) and give the
      GPU tests inputs/outputs that do not depend on numpy's RNG stream.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bfx_oracle as O  # noqa: E402

SEED = 6150533  # test/runtests.jl:10

CASES = {
    # name: (layers, seq T or 0, d_in)
    "mlp_tanh": ((O.Dense(5, 8, O.TANH), O.Dense(8, 8, O.SIGMOID), O.Dense(8, 1)), 0, 5),
    "mlp_relu_cfg2": ((O.Dense(10, 64, O.RELU), O.Dense(64, 64, O.RELU), O.Dense(64, 1)), 0, 10),
    "linreg_cfg1": ((O.Dense(5, 1),), 0, 5),
    "rnn": ((O.RNNLayer(3, 6), O.Dense(6, 1)), 7, 3),
    "lstm": ((O.LSTMLayer(4, 8), O.Dense(8, 1)), 9, 4),
    "lstm_cfg4_small": ((O.LSTMLayer(4, 64), O.Dense(64, 1)), 12, 4),
}
LIKES = {
    "normal": O.Likelihood(O.NORMAL, prior_a=2.0, prior_b=0.5),
    "tdist": O.Likelihood(O.TDIST, nu=5.0, prior_kind=O.PRIOR_INVGAMMA, prior_a=2.0, prior_b=0.5),
}
PRIORS = {
    "gauss": O.NetPrior(O.NETPRIOR_GAUSSIAN, sigma0=0.7),
    "mix": O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5),
}


def known_answers():
    dec = stats.norm.ppf(np.arange(1, 10) / 10.0)
    tdec = stats.t(10).ppf(np.arange(1, 10) / 10.0)  # y = quantile.(TDist(10), 0.1:0.1:0.9)
    sp = stats.gamma(a=2.0, scale=2.0).logpdf(1.0)  # logpdf(transformed(Gamma(2,2)), 0)
    out = {
        "likelihood_normal": float(stats.norm.logpdf(dec).sum() + sp),       # test/likelihoods.jl:24-31
        "likelihood_tdist_nu10": float(stats.t(10).logpdf(tdec).sum() + sp),  # test/likelihoods.jl:51-58
        "gaussian_prior": {str(s): float(stats.norm(0, s).logpdf(np.arange(1, 10) / 10.0).sum())
                           for s in (0.5, 1.0, 3.0, 10.0)},                  # test/networkpriors.jl:18-20
        "y_deciles": [float(v) for v in dec],
        "y_tdist10_deciles": [float(v) for v in tdec],
        "calculate_epochs": {"sgld_100_25000_after_20000": 50, "hmc_path5": 250},  # test/sgld.jl:33, test/hmc.jl:50
    }
    return out


def main():
    rng = np.random.default_rng(SEED)
    store = {}
    for cname, (layers, T, d) in CASES.items():
        B = 12
        x = rng.standard_normal((T, d, B) if T else (d, B))
        y = rng.standard_normal(B)
        store[f"{cname}/x"] = x.astype(np.float32)
        store[f"{cname}/y"] = y.astype(np.float32)
        for lname, lk in LIKES.items():
            for pname, pr in PRIORS.items():
                m = O.Model(tuple(layers), lk, pr)
                th = (0.3 * rng.standard_normal(m.n_total)).astype(np.float32)
                x64 = store[f"{cname}/x"].astype(np.float64)
                y64 = store[f"{cname}/y"].astype(np.float64)
                v, g = O.grad_loglikeprior(m, th.astype(np.float64), x64, y64, 7.0)
                key = f"{cname}/{lname}/{pname}"
                store[key + "/theta"] = th
                store[key + "/value"] = np.float64(v)
                store[key + "/grad"] = g.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), **store)
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1, sort_keys=True)
    print("wrote", len(store), "arrays")


if __name__ == "__main__":
    main()
