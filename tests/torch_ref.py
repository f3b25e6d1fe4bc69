"""Independent torch-CPU-autograd statement of loglikeprior (float64).

Stands in for Zygote (src/model/BNN.jl:66) when validating the oracle's
hand-derived reverse pass.  Uses torch.distributions for every log-density so
the formulas are not shared with oracle/bfx_oracle.py."""
import math

import torch

from oracle import bfx_oracle as O


def _act(a, z):
    return {O.IDENTITY: lambda v: v, O.RELU: torch.relu, O.TANH: torch.tanh,
            O.SIGMOID: torch.sigmoid}[a](z)


def forward(layers, th, x):
    seq = x.dim() == 3
    T = x.shape[0] if seq else 1
    B = x.shape[-1]
    off = 0
    params = []
    for l in layers:
        if l.kind == O.DENSE:
            W = th[off:off + l.nin * l.nout].reshape(l.nin, l.nout).T
            off += l.nin * l.nout
            b = th[off:off + l.nout]
            off += l.nout
            params.append((W, b))
        else:
            H = l.nout * (4 if l.kind == O.LSTM else 1)
            Wi = th[off:off + H * l.nin].reshape(l.nin, H).T
            off += H * l.nin
            Wh = th[off:off + H * l.nout].reshape(l.nout, H).T
            off += H * l.nout
            b = th[off:off + H]
            off += H
            params.append((Wi, Wh, b))
    state = [None] * len(layers)
    out = None
    for t in range(T):
        a = x[t] if seq else x
        for i, (l, p) in enumerate(zip(layers, params)):
            if l.kind == O.DENSE:
                a = _act(l.act, p[0] @ a + p[1][:, None])
            elif l.kind == O.RNN:
                h = state[i] if state[i] is not None else torch.zeros(l.nout, B, dtype=th.dtype)
                a = _act(l.act, p[0] @ a + p[1] @ h + p[2][:, None])
                state[i] = a
            else:
                h, c = state[i] if state[i] is not None else (torch.zeros(l.nout, B, dtype=th.dtype),) * 2
                g = p[0] @ a + p[1] @ h + p[2][:, None]
                H = l.nout
                ig, fg = torch.sigmoid(g[:H]), torch.sigmoid(g[H:2 * H])
                cg, og = torch.tanh(g[2 * H:3 * H]), torch.sigmoid(g[3 * H:])
                c = fg * c + ig * cg
                a = og * torch.tanh(c)
                state[i] = (a, c)
        out = a
    return out.reshape(-1)


def loglikeprior(m, th, x, y, num_batches=1.0):
    D = torch.distributions
    f64 = lambda v: torch.tensor(float(v), dtype=torch.float64)
    nnet = m.n_net
    tn, s = th[:nnet], th[nnet]
    yhat = forward(m.layers, tn, x)
    sig = torch.exp(s)
    z = (y - yhat) / sig
    if m.like.kind == O.NORMAL:
        ll = D.Normal(f64(0), f64(1)).log_prob(z).sum()
    else:
        ll = D.StudentT(f64(m.like.nu)).log_prob(z).sum()
    ll = ll - y.shape[0] * torch.log(sig)
    if m.like.prior_kind == O.PRIOR_GAMMA:   # Distributions.Gamma(shape, SCALE)
        pri = D.Gamma(f64(m.like.prior_a), f64(1.0 / m.like.prior_b)).log_prob(sig)
    else:                                     # InverseGamma(shape, scale)
        pri = D.InverseGamma(f64(m.like.prior_a), f64(m.like.prior_b)).log_prob(sig)
    ll = ll + pri + s
    if m.prior.kind == O.NETPRIOR_GAUSSIAN:
        lp = D.Normal(f64(0), f64(m.prior.sigma0)).log_prob(tn).sum()
    else:
        lp = (m.prior.pi1 * D.Normal(f64(0), f64(m.prior.sigma1)).log_prob(tn).sum()
              + (1 - m.prior.pi1) * D.Normal(f64(0), f64(m.prior.sigma2)).log_prob(tn).sum())
    return num_batches * ll + lp


def value_and_grad(m, theta, x, y, num_batches=1.0):
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    v = loglikeprior(m, th, torch.tensor(x, dtype=torch.float64),
                     torch.tensor(y, dtype=torch.float64), num_batches)
    v.backward()
    return float(v.detach()), th.grad.numpy()
