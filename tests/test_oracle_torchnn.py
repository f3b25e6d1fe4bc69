"""Independent check of the recurrent rows that does not share code with this repository:
the oracle's RNN / LSTM forward and gradient against torch.nn.RNN / torch.nn.LSTM (library-defined cells; gate order
input, forget, cell, output and weight layout [4H x d] -- the same convention Flux's LSTMCell documents).  The
hand-written cell in tests/torch_ref.py assumes the same order as the oracle; these modules do not come from here."""
import numpy as np
import pytest
import torch

from oracle import bfx_oracle as O

SEED = 6150533


def _split(theta, sizes):
    out, off = [], 0
    for s in sizes:
        out.append(theta[off:off + s])
        off += s
    return out


def _torch_forward(kind, d, H, T, B, theta, x, act="tanh"):
    """theta in the reference's flat order [vec(Wi), vec(Wh), b, vec(W_dense), b_dense] (recurrent.jl:13-68,
    dense.jl:13-26: column-major vec), x as T x d x B.  Returns yhat (B,) and the parameter tensors."""
    G = 4 * H if kind == "lstm" else H
    th = torch.tensor(theta, dtype=torch.float64)
    wi, wh, b, wd, bd = _split(th, [G * d, G * H, G, H, 1])
    cell = (torch.nn.LSTM(d, H) if kind == "lstm" else torch.nn.RNN(d, H, nonlinearity=act)).double()
    lin = torch.nn.Linear(H, 1).double()
    with torch.no_grad():
        cell.weight_ih_l0.copy_(wi.reshape(d, G).T)     # column-major vec of a G x d matrix
        cell.weight_hh_l0.copy_(wh.reshape(H, G).T)
        cell.bias_ih_l0.copy_(b)
        cell.bias_hh_l0.zero_()                          # Flux cells carry ONE bias vector
        lin.weight.copy_(wd.reshape(H, 1).T)
        lin.bias.copy_(bd)
    xin = torch.tensor(x, dtype=torch.float64).permute(0, 2, 1)   # (T, B, d)
    out, _ = cell(xin)                                            # zero initial state (recurrent.jl:30,60,64)
    yhat = lin(out[-1]).reshape(-1)                               # seq_to_one.jl:36 keeps the last output only
    return yhat, cell, lin


@pytest.mark.parametrize("kind,H,d,T,B", [("lstm", 6, 3, 7, 5), ("lstm", 16, 4, 12, 9), ("rnn", 8, 3, 9, 4)])
def test_recurrent_forward_and_gradient_match_torch_nn(kind, H, d, T, B):
    rng = np.random.default_rng(SEED + H)
    layers = ((O.LSTMLayer(d, H) if kind == "lstm" else O.RNNLayer(d, H)), O.Dense(H, 1))
    P = O.net_nparams(layers)
    theta = 0.4 * rng.standard_normal(P)
    x = rng.standard_normal((T, d, B))
    dy = rng.standard_normal(B)
    yo, cache = O.net_forward(layers, theta, x, want_cache=True)
    go = O.net_backward(layers, theta, x, dy, cache)
    yt, cell, lin = _torch_forward(kind, d, H, T, B, theta, x)
    assert np.allclose(yo, yt.detach().numpy(), rtol=1e-10, atol=1e-12)
    (yt * torch.tensor(dy)).sum().backward()
    G = 4 * H if kind == "lstm" else H
    gt = np.concatenate([cell.weight_ih_l0.grad.T.reshape(-1).numpy(), cell.weight_hh_l0.grad.T.reshape(-1).numpy(),
                         cell.bias_ih_l0.grad.numpy(), lin.weight.grad.T.reshape(-1).numpy(), lin.bias.grad.numpy()])
    assert gt.shape == go.shape == (G * d + G * H + G + H + 1,)
    assert np.linalg.norm(go - gt) <= 1e-10 * np.linalg.norm(gt)


def test_lstm_gate_order_is_detectable():
    """The check above is not vacuous: permuting the forget and cell gate blocks of the weights changes the output."""
    rng = np.random.default_rng(1)
    H, d, T, B = 5, 3, 6, 4
    layers = (O.LSTMLayer(d, H), O.Dense(H, 1))
    theta = 0.5 * rng.standard_normal(O.net_nparams(layers))
    x = rng.standard_normal((T, d, B))
    yo = O.net_forward(layers, theta, x)
    yo = yo[0] if isinstance(yo, tuple) else yo
    G = 4 * H
    th2 = theta.copy()
    b0 = G * d + G * H
    b = th2[b0:b0 + G].copy()
    b[H:2 * H], b[2 * H:3 * H] = theta[b0 + 2 * H:b0 + 3 * H], theta[b0 + H:b0 + 2 * H]   # swap the f / g bias blocks
    th2[b0:b0 + G] = b
    yt, _, _ = _torch_forward("lstm", d, H, T, B, th2, x)
    assert not np.allclose(yo, yt.detach().numpy(), rtol=1e-3, atol=1e-6)
