"""Oracle restatement of the Flux.Optimise rules used by find_mode, pinned against torch.optim (independent
implementations of the same published update rules), and FluxModeFinder's window logic (flux.jl:33-54)."""
import numpy as np
import pytest
import torch

from oracle import bfx_oracle as O


@pytest.mark.parametrize("kind", ["adam", "rmsprop", "momentum", "descent"])
def test_optimiser_matches_torch(kind):
    rng = np.random.default_rng(0)
    th0 = rng.standard_normal(12)
    grads = [rng.standard_normal(12) for _ in range(25)]
    t = torch.tensor(th0, dtype=torch.float64, requires_grad=True)
    if kind == "adam":
        o, opt = O.FluxOptimiser(O.OPT_ADAM, 0.01, 0.9, 0.999, 1e-8), torch.optim.Adam([t], lr=0.01, betas=(0.9, 0.999), eps=1e-8)
    elif kind == "rmsprop":
        o, opt = O.FluxOptimiser(O.OPT_RMSPROP, 0.01, 0.9, eps=1e-8), torch.optim.RMSprop([t], lr=0.01, alpha=0.9, eps=1e-8)
    elif kind == "momentum":
        o, opt = O.FluxOptimiser(O.OPT_MOMENTUM, 0.01, 0.9), torch.optim.SGD([t], lr=0.01, momentum=0.9)
    else:
        o, opt = O.FluxOptimiser(O.OPT_DESCENT, 0.1), torch.optim.SGD([t], lr=0.1)
    th = th0.copy()
    for g in grads:
        th = o.apply(th, g)
        t.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
    np.testing.assert_allclose(th, t.detach().numpy(), rtol=1e-9, atol=1e-12)


def test_mode_finder_window_semantics():
    # constant zero gradient: theta never moves; the first window holds |Inf - theta| = Inf, so convergence is
    # reported at the second check, i.e. on call 2*windowlength + 1 (flux.jl:36-47)
    f = O.FluxModeFinder(O.FluxOptimiser(O.OPT_DESCENT, 0.1), windowlength=3, eps=1e-6)
    th = np.ones(4, np.float32)
    calls = 0
    while True:
        th, conv = f.step(th, lambda t: (0.0, np.zeros_like(t)))
        calls += 1
        if conv:
            break
        assert calls < 20
    assert calls == 2 * 3 + 1
