"""CPU oracle for the BayesFlux hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy restatement of the reference algorithm
(enweg/BayesFlux.jl v0.2.4, Julia) for the path

    minibatch  grad_theta[ num_batches*loglike + logprior ]  of a BNN
    + the SG-MCMC `update!` that consumes it.

It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product path (``bayesflux.jl_b200``) never does.

Parity status: the reference cannot be executed in the build container (no
Julia, no network, no Manifest).  The oracle is pinned against every
known-answer value the reference's own tests hold for this path
(tests/test_oracle_golden.py: test/likelihoods.jl, test/networkpriors.jl,
test/bnn.jl, test/deconstruct.jl, test/sgld.jl:33, test/hmc.jl:50) and
cross-checked against torch-CPU autograd as an independent AD.  Everything
the reference tests do NOT pin (see SURVEY.md section 8c "Parity unpinned":
MixtureScalePrior density, RNN/LSTM values for theta != 0, single-step
sampler updates, adapter arithmetic) rests on this restatement of the cited
source lines alone  ->  for those rows: **parity unpinned**.

All file:line citations are relative to /root/reference/.

Conventions
-----------
* theta is the flat vector  [theta_net ; theta_hyper(empty) ; theta_like]
  (src/model/BNN.jl:32-45).
* Feed-forward data: ``x`` has math shape (d, B)  (Julia ``Matrix`` d x n).
  Sequence data: ``x`` has math shape (T, d, B) (Julia ``Array{T,3}``,
  src/utils/rnn_utils.jl:11-21).  Memory order is irrelevant here.
* ``dtype`` selects the arithmetic of forward/backward (float32 mirrors the
  reference's Float32 path; float64 is used for tight analytic checks).
  Scalar reductions are accumulated in float64 (the reference promotes the
  likelihood sum to Float64: ``MvNormal(zeros(n), I)``,
  src/likelihoods/feedforward.jl:42).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# layer descriptors  (src/layers/dense.jl:13-26, src/layers/recurrent.jl)
# --------------------------------------------------------------------------
DENSE, RNN, LSTM = 0, 1, 2
IDENTITY, RELU, TANH, SIGMOID = 0, 1, 2, 3


@dataclass(frozen=True)
class Layer:
    kind: int
    nin: int
    nout: int
    act: int = IDENTITY  # Dense: any; RNN: cell activation (Flux default tanh)


def Dense(nin, nout, act=IDENTITY):
    return Layer(DENSE, nin, nout, act)


def RNNLayer(nin, nout, act=TANH):
    return Layer(RNN, nin, nout, act)


def LSTMLayer(nin, nout):
    return Layer(LSTM, nin, nout, TANH)


def layer_nparams(l: Layer) -> int:
    """Parameter counts pinned by test/deconstruct.jl:19,48,80."""
    if l.kind == DENSE:
        return l.nin * l.nout + l.nout
    if l.kind == RNN:
        return l.nin * l.nout + l.nout * l.nout + l.nout
    if l.kind == LSTM:
        return 4 * (l.nin * l.nout + l.nout * l.nout + l.nout)
    raise ValueError(l.kind)


def layer_bounds(layers: Sequence[Layer]) -> Tuple[List[int], List[int]]:
    """0-based [start, end) of each layer in theta_net.

    Follows destruct(::Chain), src/model/deconstruct.jl:45-58 (which stores
    1-based inclusive starts/ends)."""
    s, starts, ends = 0, [], []
    for l in layers:
        starts.append(s)
        s += layer_nparams(l)
        ends.append(s)
    return starts, ends


def net_nparams(layers: Sequence[Layer]) -> int:
    return sum(layer_nparams(l) for l in layers)


def _act(a: int, z):
    if a == IDENTITY:
        return z
    if a == RELU:
        return np.maximum(z, 0)
    if a == TANH:
        return np.tanh(z)
    if a == SIGMOID:
        return 1.0 / (1.0 + np.exp(-z))
    raise ValueError(a)


def _dact_from_out(a: int, out):
    """sigma'(z) expressed through the activation OUTPUT."""
    if a == IDENTITY:
        return np.ones_like(out)
    if a == RELU:
        return (out > 0).astype(out.dtype)
    if a == TANH:
        return 1 - out * out
    if a == SIGMOID:
        return out * (1 - out)
    raise ValueError(a)


def unpack_layer(l: Layer, th):
    """Views of one layer's parameters.

    Dense: theta = vcat(vec(W), vec(b)), W is out x in column-major
    (src/layers/dense.jl:14-15).  RNN/LSTM: vcat(vec(Wi), vec(Wh), vec(b))
    (src/layers/recurrent.jl:16,46); the initial state is NOT a parameter and
    is rebuilt as zeros (recurrent.jl:30,60,64)."""
    if l.kind == DENSE:
        nw = l.nin * l.nout
        W = th[:nw].reshape((l.nout, l.nin), order="F")
        b = th[nw:nw + l.nout]
        return W, b
    H = l.nout * (4 if l.kind == LSTM else 1)
    nwi, nwh = H * l.nin, H * l.nout
    Wi = th[:nwi].reshape((H, l.nin), order="F")
    Wh = th[nwi:nwi + nwh].reshape((H, l.nout), order="F")
    b = th[nwi + nwh:nwi + nwh + H]
    return Wi, Wh, b


def _sigm(z):
    return 1.0 / (1.0 + np.exp(-z))


# --------------------------------------------------------------------------
# network forward / backward
# --------------------------------------------------------------------------
def net_forward(layers, theta_net, x, want_cache=False):
    """yhat (length B) of the chain applied to x.

    Feed-forward (x.ndim == 2): ``vec(net(x))`` src/likelihoods/feedforward.jl:35.
    Seq-to-one (x.ndim == 3): the whole chain is applied to every time slice
    ``x[t,:,:]`` in order, recurrent state carried across slices starting
    from zeros, only the last output is kept
    (src/likelihoods/seq_to_one.jl:36)."""
    dt = theta_net.dtype
    starts, ends = layer_bounds(layers)
    params = [unpack_layer(l, theta_net[s:e]) for l, s, e in zip(layers, starts, ends)]
    seq = x.ndim == 3
    T = x.shape[0] if seq else 1
    B = x.shape[-1]
    # recurrent state per layer
    state = []
    for l in layers:
        if l.kind == RNN:
            state.append(np.zeros((l.nout, B), dt))
        elif l.kind == LSTM:
            state.append((np.zeros((l.nout, B), dt), np.zeros((l.nout, B), dt)))
        else:
            state.append(None)
    cache = []
    out = None
    for t in range(T):
        a = (x[t] if seq else x).astype(dt, copy=False)
        step = []
        for li, (l, p) in enumerate(zip(layers, params)):
            if l.kind == DENSE:
                W, b = p
                o = _act(l.act, W @ a + b[:, None])
                step.append((a, o))
                a = o
            elif l.kind == RNN:
                Wi, Wh, b = p
                hprev = state[li]
                h = _act(l.act, Wi @ a + Wh @ hprev + b[:, None])
                step.append((a, hprev, h))
                state[li] = h
                a = h
            else:  # LSTM, Flux LSTMCell: rows [i; f; g(cell); o]
                Wi, Wh, b = p
                hprev, cprev = state[li]
                H = l.nout
                g = Wi @ a + Wh @ hprev + b[:, None]
                ig, fg = _sigm(g[:H]), _sigm(g[H:2 * H])
                cg, og = np.tanh(g[2 * H:3 * H]), _sigm(g[3 * H:])
                c = fg * cprev + ig * cg
                tc = np.tanh(c)
                h = og * tc
                step.append((a, hprev, cprev, ig, fg, cg, og, tc))
                state[li] = (h, c)
                a = h
        cache.append(step)
        out = a
    yhat = out.reshape(-1)  # vec(net(x)); single output => length B
    if want_cache:
        return yhat, (params, cache)
    return yhat


def net_backward(layers, theta_net, x, dyhat, fwd_cache):
    """Gradient of sum_b dyhat[b]*yhat[b] w.r.t. theta_net (hand-derived
    reverse pass; stands in for Zygote at src/model/BNN.jl:66)."""
    dt = theta_net.dtype
    params, cache = fwd_cache
    starts, ends = layer_bounds(layers)
    g = np.zeros_like(theta_net)
    T = len(cache)
    B = dyhat.shape[0]
    nl = len(layers)
    # gradient flowing into recurrent state from the future
    dstate = []
    for l in layers:
        if l.kind == RNN:
            dstate.append(np.zeros((l.nout, B), dt))
        elif l.kind == LSTM:
            dstate.append((np.zeros((l.nout, B), dt), np.zeros((l.nout, B), dt)))
        else:
            dstate.append(None)
    for t in reversed(range(T)):
        # gradient w.r.t. the chain output at time t: only the last slice counts
        if t == T - 1:
            da = dyhat.reshape(layers[-1].nout, B).astype(dt)
        else:
            da = np.zeros((layers[-1].nout, B), dt)
        for li in reversed(range(nl)):
            l, p, st = layers[li], params[li], cache[t][li]
            gs = g[starts[li]:ends[li]]
            if l.kind == DENSE:
                W, b = p
                a_in, o = st
                dz = da * _dact_from_out(l.act, o)
                nw = l.nin * l.nout
                gs[:nw] += (dz @ a_in.T).reshape(-1, order="F")
                gs[nw:] += dz.sum(axis=1)
                da = W.T @ dz
            elif l.kind == RNN:
                Wi, Wh, b = p
                a_in, hprev, h = st
                dh = da + dstate[li]
                dz = dh * _dact_from_out(l.act, h)
                H = l.nout
                nwi, nwh = H * l.nin, H * H
                gs[:nwi] += (dz @ a_in.T).reshape(-1, order="F")
                gs[nwi:nwi + nwh] += (dz @ hprev.T).reshape(-1, order="F")
                gs[nwi + nwh:] += dz.sum(axis=1)
                dstate[li] = Wh.T @ dz
                da = Wi.T @ dz
            else:
                Wi, Wh, b = p
                a_in, hprev, cprev, ig, fg, cg, og, tc = st
                H = l.nout
                dh_f, dc_f = dstate[li]
                dh = da + dh_f
                do = dh * tc
                dc = dh * og * (1 - tc * tc) + dc_f
                di, df, dg = dc * cg, dc * cprev, dc * ig
                dz = np.concatenate([di * ig * (1 - ig), df * fg * (1 - fg),
                                     dg * (1 - cg * cg), do * og * (1 - og)], axis=0)
                nwi, nwh = 4 * H * l.nin, 4 * H * H
                gs[:nwi] += (dz @ a_in.T).reshape(-1, order="F")
                gs[nwi:nwi + nwh] += (dz @ hprev.T).reshape(-1, order="F")
                gs[nwi + nwh:] += dz.sum(axis=1)
                dstate[li] = (Wh.T @ dz, dc * fg)
                da = Wi.T @ dz
    return g


# --------------------------------------------------------------------------
# likelihoods (src/likelihoods/feedforward.jl, seq_to_one.jl) and priors
# --------------------------------------------------------------------------
NORMAL, TDIST = 0, 1
PRIOR_GAMMA, PRIOR_INVGAMMA = 0, 1
NETPRIOR_GAUSSIAN, NETPRIOR_MIXTURE = 0, 1

LOG2PI = math.log(2.0 * math.pi)


@dataclass(frozen=True)
class Likelihood:
    """FeedforwardNormal / FeedforwardTDist / SeqToOneNormal / SeqToOneTDist.

    Feed-forward vs seq-to-one is decided by x.ndim, exactly like Julia's
    dispatch on Matrix vs Array{T,3}.  prior_sigma is a positive-support
    univariate distribution mapped through Bijectors' log link: sigma =
    exp(theta_like[1]) and logpdf(transformed(d), s) = logpdf(d, e^s) + s
    (feedforward.jl:36-37; test/bnn.jl:75)."""
    kind: int = NORMAL
    nu: float = 0.0
    prior_kind: int = PRIOR_GAMMA
    prior_a: float = 2.0      # Gamma shape alpha   | InverseGamma shape alpha
    prior_b: float = 0.5      # Gamma SCALE theta   | InverseGamma scale beta


def logpdf_sigma_prior_transformed(lk: Likelihood, s: float) -> float:
    sig = math.exp(s)
    a, b = lk.prior_a, lk.prior_b
    if lk.prior_kind == PRIOR_GAMMA:
        lp = -math.lgamma(a) - a * math.log(b) + (a - 1) * math.log(sig) - sig / b
    else:
        lp = a * math.log(b) - math.lgamma(a) - (a + 1) * math.log(sig) - b / sig
    return lp + s


def dlogpdf_sigma_prior_transformed(lk: Likelihood, s: float) -> float:
    """d/ds of the above."""
    sig = math.exp(s)
    a, b = lk.prior_a, lk.prior_b
    if lk.prior_kind == PRIOR_GAMMA:
        return a - sig / b
    return -a + b / sig


@dataclass(frozen=True)
class NetPrior:
    """GaussianPrior (src/netpriors/gaussian.jl:25-28) or MixtureScalePrior
    (src/netpriors/mixturescale.jl:30-33: pi1*logN(0,s1^2 I) + pi2*logN(0,s2^2 I),
    a weighted SUM of log densities)."""
    kind: int = NETPRIOR_GAUSSIAN
    sigma0: float = 1.0
    sigma1: float = 1.0
    sigma2: float = 1.0
    pi1: float = 0.5


def netprior_coeffs(pr: NetPrior) -> Tuple[float, float]:
    """logprior(theta_net) = n*c0 - 0.5*c2*sum(theta^2); returns (c0, c2)."""
    if pr.kind == NETPRIOR_GAUSSIAN:
        return (-0.5 * LOG2PI - math.log(pr.sigma0), 1.0 / pr.sigma0 ** 2)
    pi2 = 1.0 - pr.pi1
    c0 = pr.pi1 * (-0.5 * LOG2PI - math.log(pr.sigma1)) + pi2 * (-0.5 * LOG2PI - math.log(pr.sigma2))
    c2 = pr.pi1 / pr.sigma1 ** 2 + pi2 / pr.sigma2 ** 2
    return c0, c2


def logprior(pr: NetPrior, theta_net) -> float:
    c0, c2 = netprior_coeffs(pr)
    return theta_net.size * c0 - 0.5 * c2 * float(np.sum(theta_net.astype(np.float64) ** 2))


@dataclass(frozen=True)
class Model:
    """BNN(x, y, like, prior, init) minus the data (src/model/BNN.jl:5-38)."""
    layers: Tuple[Layer, ...]
    like: Likelihood = Likelihood()
    prior: NetPrior = NetPrior()

    @property
    def n_net(self) -> int:
        return net_nparams(self.layers)

    @property
    def n_total(self) -> int:
        return self.n_net + 1  # num_params_like = 1, num_params_hyper = 0


def split_params(m: Model, theta):
    """src/model/BNN.jl:40-45 (theta_hyper is empty for both shipped priors)."""
    return theta[:m.n_net], theta[m.n_net:m.n_net], theta[m.n_net:]


def loglike(m: Model, x, y, theta_net, theta_like, want_grad=False):
    """(l::BNNLikelihood)(x, y, theta_net, theta_like).

    Normal: feedforward.jl:30-43 / seq_to_one.jl:31-44.
    TDist : feedforward.jl:86-97 / seq_to_one.jl:89-100."""
    lk = m.like
    dt = theta_net.dtype
    s = float(theta_like[0])
    sig = math.exp(s)
    B = y.shape[0]
    if want_grad:
        yhat, cache = net_forward(m.layers, theta_net, x, want_cache=True)
    else:
        yhat = net_forward(m.layers, theta_net, x)
    r = (y.astype(dt) - yhat)
    z = (r / dt.type(sig)).astype(np.float64)
    if lk.kind == NORMAL:
        val = -0.5 * B * LOG2PI - 0.5 * float(np.sum(z * z))
        dval_dyhat = z / sig                       # (y - yhat)/sigma^2
        dval_ds = float(np.sum(z * z))             # sigma * d/dsigma of -1/2 sum z^2
    else:
        nu = lk.nu
        cst = math.lgamma(0.5 * (nu + 1)) - math.lgamma(0.5 * nu) - 0.5 * math.log(nu * math.pi)
        val = B * cst - 0.5 * (nu + 1) * float(np.sum(np.log1p(z * z / nu)))
        w = (nu + 1) * z / (nu + z * z)
        dval_dyhat = w / sig
        dval_ds = float(np.sum(w * z))
    val += -B * s + logpdf_sigma_prior_transformed(lk, s)
    if not want_grad:
        return val
    dlike_ds = dval_ds - B + dlogpdf_sigma_prior_transformed(lk, s)
    return val, dval_dyhat.astype(dt), dlike_ds, cache


def loglikeprior(m: Model, theta, x, y, num_batches=1.0) -> float:
    """src/model/BNN.jl:50-56: num_batches*like + prior (the sigma-prior term
    sits inside `like` and is scaled by num_batches too)."""
    tn, _, tl = split_params(m, theta)
    return num_batches * loglike(m, x, y, tn, tl) + logprior(m.prior, tn)


def grad_loglikeprior(m: Model, theta, x, y, num_batches=1.0):
    """src/model/BNN.jl:61-69 -> (value, gradient[P])."""
    tn, _, tl = split_params(m, theta)
    dt = theta.dtype
    val, dyhat, dlike_ds, cache = loglike(m, x, y, tn, tl, want_grad=True)
    g = np.empty_like(theta)
    gnet = net_backward(m.layers, tn, x, (num_batches * dyhat).astype(dt), cache)
    _, c2 = netprior_coeffs(m.prior)
    g[:m.n_net] = gnet - dt.type(c2) * tn
    g[m.n_net] = num_batches * dlike_ds
    v = num_batches * val + logprior(m.prior, tn)
    return v, g


# --------------------------------------------------------------------------
# utilities
# --------------------------------------------------------------------------
def clip_gradient(g, maxnorm):
    """src/utils/gradient_utils.jl:6-10 (norm over ALL entries, incl. theta_like)."""
    ng = float(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
    if ng > maxnorm:
        return (g * g.dtype.type(maxnorm / ng)).astype(g.dtype)
    return g


def make_rnn_tensor(mat, seq_len=10):
    """src/utils/rnn_utils.jl:11-21: (timesteps, features) -> (seq_len, features, nseq)."""
    nseq = mat.shape[0] - seq_len + 1
    out = np.empty((seq_len, mat.shape[1], nseq), mat.dtype)
    for i in range(nseq):
        out[:, :, i] = mat[i:i + seq_len, :]
    return out


def num_batches_of(n, batchsize, partial=True):
    """length(DataLoader) - MLUtils: ceil(n/B) with partial batches else floor."""
    return -(-n // batchsize) if partial else n // batchsize


def batch_slices(n, batchsize, partial=True):
    nb = num_batches_of(n, batchsize, partial)
    return [(i * batchsize, min(n, (i + 1) * batchsize)) for i in range(nb)]


# --------------------------------------------------------------------------
# step-size and mass adapters
# --------------------------------------------------------------------------
@dataclass
class ConstantStepsize:
    """adapters/stepsize/constantstepsize.jl:5-8"""
    l: float

    def __call__(self, sampler, mh_prob):
        return self.l


@dataclass
class DualAveragingStepSize:
    """adapters/stepsize/dualaveragestepsize.jl:11-62 (incl. the reference's
    sqrt(t*gamma) and the counter starting at t0)."""
    l: float
    target_accept: float = 0.65
    gamma: float = 0.05
    t0: int = 10
    kappa: float = 0.75
    adapt_steps: int = 1000
    mu: float = field(init=False)
    t: int = field(init=False)
    error_sum: float = 0.0
    log_averaged_step: float = 0.0

    def __post_init__(self):
        self.mu = math.log(10 * self.l)
        self.t = self.t0

    def __call__(self, sampler, mh_prob):
        if self.t > self.adapt_steps:
            sampler.l = math.exp(self.log_averaged_step)
            return sampler.l
        self.error_sum += self.target_accept - mh_prob
        log_step = self.mu - self.error_sum / math.sqrt(self.t * self.gamma)
        eta = self.t ** (-self.kappa)
        self.log_averaged_step = eta * log_step + (1 - eta) * self.log_averaged_step
        self.t += 1
        self.l = math.exp(log_step)
        return self.l


@dataclass
class FixedMassAdapter:
    """adapters/mass/fixedmassmatrix.jl:5-14 (identity, diagonal stored)."""
    Minv: Optional[np.ndarray] = None

    def __call__(self, sampler, theta, grad_fn):
        if self.Minv is None:
            self.Minv = np.ones_like(theta)
        return self.Minv


@dataclass
class DiagCovMassAdapter:
    """adapters/mass/diagcovariancemassadapter.jl:13-50."""
    adapt_steps: int
    windowlength: int
    kappa: float = 0.5
    epsilon: float = 1e-6
    t: int = 1
    Minv: Optional[np.ndarray] = None

    def __call__(self, sampler, theta, grad_fn):
        if self.t == 1 and self.Minv is None:
            self.Minv = np.ones_like(theta)
        if self.t > self.adapt_steps:
            return self.Minv
        if self.windowlength > sampler.nsampled:
            return self.Minv
        w = sampler.samples[:, sampler.nsampled - self.windowlength:sampler.nsampled]
        var = np.var(w.astype(np.float64), axis=1, ddof=1)
        dt = theta.dtype
        self.Minv = ((1 - self.kappa) * var + self.kappa + self.epsilon).astype(dt)
        self.t += 1
        return self.Minv


@dataclass
class RMSPropMassAdapter:
    """adapters/mass/rmspropmassadapter.jl:9-46 (one extra grad-eval per call
    while adapting; g normalised by its 2-norm)."""
    adapt_steps: int = 1000
    lam: float = 1e-5
    alpha: float = 0.99
    t: int = 1
    Minv: Optional[np.ndarray] = None
    V: Optional[np.ndarray] = None

    def __call__(self, sampler, theta, grad_fn):
        if self.t > self.adapt_steps:
            return self.Minv
        if self.t == 1 and self.Minv is None:
            self.Minv = np.ones_like(theta)
        if self.t == 1:
            self.V = np.ones_like(theta)
        _, g = grad_fn(theta)
        dt = theta.dtype
        g = g / dt.type(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
        self.V = (dt.type(self.alpha) * self.V + dt.type(1 - self.alpha) * g * g).astype(dt)
        self.Minv = (1 / (dt.type(self.lam) + np.sqrt(self.V))).astype(dt)
        self.t += 1
        return self.Minv


# --------------------------------------------------------------------------
# samplers.  Every update takes its random draws explicitly so the CUDA
# engine can be compared step by step ("injected noise").
# --------------------------------------------------------------------------
class NaNError(RuntimeError):
    pass


def _epochs(n_new, nbatches):
    return -(-n_new // nbatches)


@dataclass
class SGLD:
    """src/inference/mcmc/sgld.jl:12,31-112."""
    stepsize_a: float = 0.1
    stepsize_b: float = 1.0
    stepsize_gamma: float = 0.55
    min_stepsize: float = -math.inf
    maxnorm: float = 25.0
    t: int = 1
    nsampled: int = 0
    samples: Optional[np.ndarray] = None
    draws_per_step = ("normal",)

    def initialise(self, theta, nsamples, continue_sampling=False):
        samples = np.empty((theta.size, nsamples), theta.dtype)
        if continue_sampling:
            samples[:, :self.nsampled] = self.samples[:, :self.nsampled]
        self.t = self.nsampled + 1 if continue_sampling else 1
        self.nsampled = self.nsampled if continue_sampling else 0
        self.samples = samples

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        return _epochs(nsamples - self.nsampled if continue_sampling else nsamples, nbatches)

    def stepsize(self):
        dt = np.float32
        a = dt(self.stepsize_a) * (dt(self.stepsize_b) + dt(self.t)) ** dt(-self.stepsize_gamma)
        return max(float(a), self.min_stepsize)

    def update(self, theta, grad_fn, full_lp_fn, normals, uniforms=None):
        dt = theta.dtype
        alpha = self.stepsize()
        _, g = grad_fn(theta)
        g = clip_gradient(g, self.maxnorm)
        theta = (theta + dt.type(alpha / 2) * g + dt.type(math.sqrt(alpha)) * normals[0].astype(dt)).astype(dt)
        self.samples[:, self.t - 1] = theta
        self.nsampled += 1
        self.t += 1
        return theta


@dataclass
class SGNHT:
    """src/inference/mcmc/sgnht.jl:20-90."""
    l: float
    A: float = 1.0
    xi: float = 0.0
    t: int = 1
    nsampled: int = 0
    samples: Optional[np.ndarray] = None
    kinetic: Optional[np.ndarray] = None
    p: Optional[np.ndarray] = None
    draws_per_step = ("normal",)

    def initialise(self, theta, nsamples, continue_sampling=False):
        samples = np.empty((theta.size, nsamples), theta.dtype)
        kinetic = np.empty(nsamples, theta.dtype)
        if continue_sampling:
            samples[:, :self.nsampled] = self.samples[:, :self.nsampled]
            kinetic[:self.nsampled] = self.kinetic[:self.nsampled]
        self.t = self.nsampled + 1 if continue_sampling else 1
        self.nsampled = self.nsampled if continue_sampling else 0
        self.samples, self.kinetic = samples, kinetic
        self.p = np.zeros_like(theta)

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        return _epochs(nsamples - self.nsampled if continue_sampling else nsamples, nbatches)

    def update(self, theta, grad_fn, full_lp_fn, normals, uniforms=None):
        dt = theta.dtype
        _, g = grad_fn(theta)
        g = -g
        if np.any(np.isnan(g)):
            raise NaNError("NaN in g")
        l, xi = dt.type(self.l), dt.type(self.xi)
        self.p = (self.p - xi * l * self.p - l * g
                  + dt.type(math.sqrt(self.l * 2 * self.A)) * normals[0].astype(dt)).astype(dt)
        theta = (theta + l * self.p).astype(dt)
        n = theta.size
        kin = float(np.sum(self.p.astype(np.float64) ** 2)) / n
        self.kinetic[self.t - 1] = kin
        self.xi = float(dt.type(self.xi + (kin - 1) * self.l))
        if np.any(np.isnan(theta)):
            raise NaNError("NaN in θ")
        self.samples[:, self.t - 1] = theta
        self.t += 1
        self.nsampled += 1
        return theta


@dataclass
class SGNHTS:
    """src/inference/mcmc/sgnht-s.jl:26-116 (diagonal Minv)."""
    l: float
    sigmaA: float = 1.0
    xi: float = 1.0
    mu: float = 1.0
    madapter: object = field(default_factory=FixedMassAdapter)
    t: int = 1
    nsampled: int = 0
    samples: Optional[np.ndarray] = None
    kinetic: Optional[np.ndarray] = None
    p: Optional[np.ndarray] = None
    draws_per_step = ("normal",)

    def initialise(self, theta, nsamples, continue_sampling=False):
        samples = np.empty((theta.size, nsamples), theta.dtype)
        kinetic = np.empty(nsamples, theta.dtype)
        if continue_sampling:
            samples[:, :self.nsampled] = self.samples[:, :self.nsampled]
            kinetic[:self.nsampled] = self.kinetic[:self.nsampled]
        self.t = self.nsampled + 1 if continue_sampling else 1
        self.nsampled = self.nsampled if continue_sampling else 0
        self.samples, self.kinetic = samples, kinetic
        self.p = np.zeros_like(theta)

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        return _epochs(nsamples - self.nsampled if continue_sampling else nsamples, nbatches)

    def update(self, theta, grad_fn, full_lp_fn, normals, uniforms=None):
        dt = theta.dtype
        Minv = self.madapter(self, theta, grad_fn)
        Mhalf = np.sqrt(1 / Minv).astype(dt)   # noise ~ N(0, M), M = inv(Minv)  (:84)
        _, g = grad_fn(theta)
        g = -g
        if np.any(np.isnan(g)):
            raise NaNError("NaN in g")
        n = theta.size
        l, mu = dt.type(self.l), dt.type(self.mu)
        half = dt.type(self.l / 2)

        def pMp(p):
            return float(np.sum(p.astype(np.float64) ** 2 * Minv.astype(np.float64)))

        p = (self.p - half * g).astype(dt)
        theta = (theta + half * Minv * p).astype(dt)
        xi = dt.type(self.xi) + l / (dt.type(2) * mu) * dt.type(pMp(p) - n)
        zeta = Mhalf * normals[0].astype(dt)
        if xi != 0:
            e1 = dt.type(math.exp(-float(xi) * self.l))
            sc = dt.type(self.sigmaA * math.sqrt((1 - math.exp(-2 * float(xi) * self.l)) / (2 * float(xi))))
            p = (e1 * p + sc * zeta).astype(dt)
        else:
            p = (p + dt.type(math.sqrt(self.l) * self.sigmaA) * zeta).astype(dt)
        xi = xi + l / (dt.type(2) * mu) * dt.type(pMp(p) - n)
        theta = (theta + half * Minv * p).astype(dt)
        if np.any(np.isnan(theta)):
            raise NaNError("NaN in θ")
        self.p, self.xi = p, float(xi)
        self.samples[:, self.t - 1] = theta
        self.kinetic[self.t - 1] = pMp(p) / n
        self.t += 1
        self.nsampled += 1
        return theta


@dataclass
class GGMC:
    """src/inference/mcmc/ggmc.jl:34-198, diagonal mass matrices.  Reference
    quirks kept: h <- sqrt(h) (:154); lMH += K(m34) - K(m14) (:171); momentum
    not restored on reject (:188-191)."""
    beta: float = 0.55
    l: float = 1e-4
    steps: int = 1
    maxnorm: float = 5.0
    sadapter: object = None
    madapter: object = None
    n_data: int = 0                      # length(bnn.y)
    t: int = 1
    nsampled: int = 0
    current_step: int = 1
    lMH: float = 0.0
    samples: Optional[np.ndarray] = None
    accepted: Optional[np.ndarray] = None
    Minv: Optional[np.ndarray] = None
    Mhalf: Optional[np.ndarray] = None
    momentum: Optional[np.ndarray] = None
    draws_per_step = ("normal", "normal", "uniform")

    def __post_init__(self):
        if self.sadapter is None:
            self.sadapter = DualAveragingStepSize(self.l)
        if self.madapter is None:
            self.madapter = DiagCovMassAdapter(1000, 100)

    def initialise(self, theta, nsamples, continue_sampling=False):
        samples = np.empty((theta.size, nsamples), theta.dtype)
        accepted = np.zeros(nsamples, np.int64)
        if continue_sampling:
            samples[:, :self.nsampled] = self.samples[:, :self.nsampled]
            accepted[:self.nsampled] = self.accepted[:self.nsampled]
        self.t = self.t if continue_sampling else 1
        self.nsampled = self.nsampled if continue_sampling else 0
        self.samples, self.accepted = samples, accepted
        self.current_step = 1
        if not continue_sampling:
            self.Minv = np.ones_like(theta)
            self.Mhalf = np.ones_like(theta)
        self.momentum = np.zeros_like(theta)

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        return _epochs(nsamples - self.nsampled if continue_sampling else nsamples, nbatches)

    def update(self, theta, grad_fn, full_lp_fn, normals, uniforms):
        dt = theta.dtype
        N = self.n_data
        gam = -math.sqrt(N / self.l) * math.log(self.beta)
        h = math.sqrt(self.l / N)
        a = math.exp(-gam * h)
        if self.current_step == 1:
            self.lMH = -full_lp_fn(theta)
        _, g = grad_fn(theta)
        g = clip_gradient(-g, self.maxnorm)
        h = math.sqrt(h)
        sa, s1a = dt.type(math.sqrt(a)), dt.type(math.sqrt(1 - a))
        hh = dt.type(h / 2)

        def K(m):
            return 0.5 * float(np.sum(m.astype(np.float64) ** 2 * self.Minv.astype(np.float64)))

        m14 = (sa * self.momentum + s1a * self.Mhalf * normals[0].astype(dt)).astype(dt)
        m12 = (m14 - hh * g).astype(dt)
        theta = (theta + dt.type(h) * self.Minv * m12).astype(dt)
        if np.any(np.isnan(theta)):
            raise NaNError("NaN in θ")
        _, g = grad_fn(theta)
        g = clip_gradient(-g, self.maxnorm)
        m34 = (m12 - hh * g).astype(dt)
        self.momentum = (sa * m34 + s1a * self.Mhalf * normals[1].astype(dt)).astype(dt)
        self.lMH += K(m34) - K(m14)
        self.samples[:, self.t - 1] = theta
        self.nsampled += 1
        if self.current_step == self.steps and self.t > self.steps:
            self.lMH += full_lp_fn(theta)
            acc_prob = min(math.exp(min(self.lMH, 0.0)), 1.0) if not math.isnan(self.lMH) else math.nan
            self.l = self.sadapter(self, acc_prob)
            self.Minv = self.madapter(self, theta, grad_fn)
            self.Mhalf = np.sqrt(1 / self.Minv).astype(dt)
            r = float(uniforms[0])
            if r < acc_prob:
                self.accepted[self.nsampled - self.steps:self.nsampled] = 1
            else:
                back = self.samples[:, self.nsampled - self.steps - 1].copy()
                self.samples[:, self.nsampled - self.steps:self.nsampled] = back[:, None]
                theta = back
        self.t += 1
        self.current_step = 1 if self.current_step == self.steps else self.current_step + 1
        return theta


@dataclass
class HMC:
    """src/inference/mcmc/hmc.jl:38-159, diagonal Minv.  One update() is one
    leapfrog *stage*; half kicks at every stage and no Minv in the position
    update, as in the reference."""
    l: float
    path_len: int
    maxnorm: float = 5.0
    sadapter: object = None
    madapter: object = None
    t: int = 1
    nsampled: int = 0
    current_step: int = 1
    samples: Optional[np.ndarray] = None
    accepted: Optional[np.ndarray] = None
    Minv: Optional[np.ndarray] = None
    momentum: Optional[np.ndarray] = None
    momentumold: Optional[np.ndarray] = None
    thetaold: Optional[np.ndarray] = None
    draws_per_step = ("normal", "uniform")

    def __post_init__(self):
        if self.sadapter is None:
            self.sadapter = DualAveragingStepSize(self.l)
        if self.madapter is None:
            self.madapter = DiagCovMassAdapter(1000, 100)

    def initialise(self, theta, nsamples, continue_sampling=False):
        samples = np.empty((theta.size, nsamples), theta.dtype)
        accepted = np.zeros(nsamples, bool)
        if continue_sampling:
            samples[:, :self.nsampled] = self.samples[:, :self.nsampled]
            accepted[:self.nsampled] = self.accepted[:self.nsampled]
        self.t = self.nsampled + 1 if continue_sampling else 1
        self.nsampled = self.nsampled if continue_sampling else 0
        self.samples, self.accepted = samples, accepted
        self.thetaold = theta.copy()
        self.momentum = np.zeros_like(theta)
        self.momentumold = np.zeros_like(theta)
        self.l = self.sadapter.l
        m = getattr(self.madapter, "Minv", None)
        self.Minv = np.ones_like(theta) if m is None else m

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        n_new = nsamples - self.nsampled if continue_sampling else nsamples
        return _epochs(n_new * self.path_len, nbatches)

    def _half_moment(self, theta, grad_fn):
        dt = theta.dtype
        _, g = grad_fn(theta)
        g = clip_gradient(-g, self.maxnorm)
        self.momentum = (self.momentum - dt.type(self.l) * g / dt.type(2)).astype(dt)

    def update(self, theta, grad_fn, full_lp_fn, normals, uniforms):
        dt = theta.dtype
        Minv = self.Minv  # moment_dist = N(0, Minv)  (:115)

        def neg_logpdf_mom(p):  # -logpdf up to the constant that cancels in lMH
            return 0.5 * float(np.sum(p.astype(np.float64) ** 2 / Minv.astype(np.float64)))

        if self.current_step == 1:
            self.momentum = (np.sqrt(Minv) * normals[0].astype(dt)).astype(dt)
            self.thetaold = theta.copy()
            self.momentumold = self.momentum.copy()
            self._half_moment(theta, grad_fn)
        elif self.current_step < self.path_len:
            theta = (theta + dt.type(self.l) * self.momentum).astype(dt)
            self._half_moment(theta, grad_fn)
        else:
            theta = (theta + dt.type(self.l) * self.momentum).astype(dt)
            self._half_moment(theta, grad_fn)
            self.momentum = -self.momentum
            start_mh = -full_lp_fn(self.thetaold) + neg_logpdf_mom(self.momentumold)
            end_mh = -full_lp_fn(theta) + neg_logpdf_mom(self.momentum)
            lMH = start_mh - end_mh
            lr = math.log(float(uniforms[0]))
            acc_prob = min(math.exp(min(lMH, 0.0)), 1.0) if not math.isnan(lMH) else math.nan
            self.l = self.sadapter(self, acc_prob)
            self.Minv = self.madapter(self, theta, grad_fn)
            if lr < lMH:
                self.samples[:, self.nsampled] = theta
                self.accepted[self.nsampled] = True
            else:
                self.samples[:, self.nsampled] = self.thetaold
                theta = self.thetaold.copy()
            self.nsampled += 1
        self.current_step = 1 if self.current_step == self.path_len else self.current_step + 1
        self.t += 1
        return theta


# --------------------------------------------------------------------------
# the mcmc driver (src/inference/mcmc/abstract.jl:74-127)
# --------------------------------------------------------------------------
def take_batch(x, y, idx):
    return x[..., idx], y[idx]


def mcmc(m: Model, x, y, batchsize, numsamples, sampler, theta_start, rng,
         shuffle=True, partial=True, continue_sampling=False, perms=None, noise=None):
    """Reference-semantic single-chain run.

    Batch schedule = MLUtils DataLoader: a fresh permutation of 0..n-1 per
    epoch when shuffle=True, batches are consecutive chunks of it
    (abstract.jl:102-103,117-118).  Randomness: ``rng`` (numpy Generator) or,
    for engine parity, explicit ``perms`` (callable epoch -> permutation) and
    ``noise`` (callable global_step -> (normals, uniforms))."""
    n = y.shape[0]
    dt = theta_start.dtype
    theta = sampler.samples[:, -1].copy() if continue_sampling else theta_start.copy()
    slices = batch_slices(n, batchsize, partial)
    nb = len(slices)
    if hasattr(sampler, "n_data"):
        sampler.n_data = n
    sampler.initialise(theta, numsamples, continue_sampling)
    epochs = sampler.calculate_epochs(nb, numsamples, continue_sampling)
    full_lp = lambda th: loglikeprior(m, th, x, y, 1.0)
    step = 0
    for e in range(epochs):
        if perms is not None:
            perm = perms(e)
        elif shuffle:
            perm = rng.permutation(n)
        else:
            perm = np.arange(n)
        for (s, t) in slices:
            if sampler.nsampled == numsamples:
                break
            idx = perm[s:t]
            xb, yb = take_batch(x, y, idx)
            grad_fn = lambda th, xb=xb, yb=yb: grad_loglikeprior(m, th, xb, yb, float(nb))
            if noise is not None:
                normals, uniforms = noise(step)
            else:
                nn = sum(1 for d in sampler.draws_per_step if d == "normal")
                normals = [rng.standard_normal(theta.size).astype(dt) for _ in range(nn)]
                uniforms = [rng.random()]
            theta = sampler.update(theta, grad_fn, full_lp, normals, uniforms)
            step += 1
    return sampler.samples


# --------------------------------------------------------------------------
# find_mode: src/inference/mode/abstract.jl:36-86,
# src/inference/mode/flux.jl:8-54.  The optimisers are Flux.Optimise rules
# (third party, Flux 0.13: Descent, Momentum, RMSProp, ADAM with EPS = 1e-8),
# applied to -g because Flux minimises (flux.jl:49-51).
# --------------------------------------------------------------------------
OPT_DESCENT, OPT_MOMENTUM, OPT_RMSPROP, OPT_ADAM = 0, 1, 2, 3


@dataclass
class FluxOptimiser:
    kind: int = OPT_ADAM
    eta: float = 0.001
    rho: float = 0.9       # Momentum rho / RMSProp rho / ADAM beta1
    beta2: float = 0.999   # ADAM beta2
    eps: float = 1e-8
    m: Optional[np.ndarray] = None
    v: Optional[np.ndarray] = None
    bp1: float = 0.0
    bp2: float = 0.0

    def apply(self, theta, delta):
        """Flux.update!(opt, theta, delta): theta .-= apply!(opt, theta, delta)."""
        dt = theta.dtype
        f = dt.type
        if self.m is None:
            self.m = np.zeros_like(theta)
            self.v = np.zeros_like(theta)
            self.bp1, self.bp2 = self.rho, self.beta2
        if self.kind == OPT_DESCENT:
            step = f(self.eta) * delta
        elif self.kind == OPT_MOMENTUM:
            self.m = (f(self.rho) * self.m - f(self.eta) * delta).astype(dt)
            step = -self.m
        elif self.kind == OPT_RMSPROP:
            self.v = (f(self.rho) * self.v + f(1.0 - self.rho) * delta * delta).astype(dt)
            step = delta * (f(self.eta) / (np.sqrt(self.v) + f(self.eps)))
        else:
            self.m = (f(self.rho) * self.m + f(1.0 - self.rho) * delta).astype(dt)
            self.v = (f(self.beta2) * self.v + f(1.0 - self.beta2) * delta * delta).astype(dt)
            step = self.m / f(1.0 - self.bp1) / (np.sqrt(self.v / f(1.0 - self.bp2)) + f(self.eps)) * f(self.eta)
            self.bp1 *= self.rho
            self.bp2 *= self.beta2
        return (theta - step.astype(dt)).astype(dt)


@dataclass
class FluxModeFinder:
    """src/inference/mode/flux.jl:8-54."""
    opt: FluxOptimiser
    windowlength: int = 100
    eps: float = 1e-6
    window: Optional[np.ndarray] = None
    prev: Optional[np.ndarray] = None
    i: int = 1

    def step(self, theta, grad_fn):
        if self.window is None:
            self.window = np.full(self.windowlength, -np.inf, theta.dtype)
            self.prev = np.full(theta.size, np.inf, theta.dtype)
        if self.i == 1:                                          # flux.jl:36-41
            if float(np.max(np.abs(self.window))) <= self.eps:
                return theta, True
        _, g = grad_fn(theta)                                    # :43
        self.window[self.i - 1] = np.max(np.abs(self.prev - theta))  # :44-45
        self.prev = theta.copy()                                 # :46
        self.i = 1 if self.i == self.windowlength else self.i + 1   # :47
        theta = self.opt.apply(theta, -g)                        # :51
        return theta, False


def find_mode(m: Model, x, y, batchsize, epochs, finder: FluxModeFinder, theta_start, rng=None, shuffle=True,
              partial=True, perms=None):
    """src/inference/mode/abstract.jl:36-86 (theta_start replaces the bnn.init() draw)."""
    n = y.shape[0]
    theta = theta_start.copy()
    slices = batch_slices(n, batchsize, partial)
    nb = len(slices)
    for e in range(epochs):
        perm = perms(e) if perms is not None else (rng.permutation(n) if shuffle else np.arange(n))
        for (s, t) in slices:
            idx = perm[s:t]
            xb, yb = take_batch(x, y, idx)
            theta, conv = finder.step(theta, lambda th, xb=xb, yb=yb: grad_loglikeprior(m, th, xb, yb, float(nb)))
            if conv:
                return theta, True
    return theta, False


# ---------------------------------------------------------------------------------------------------------
# Chain diagnostics.  The reference ships none -- its README hands samples' to MCMCChains
# (README.md:191-193) -- so this restates the published arithmetic: split-Rhat (Gelman et al., BDA3 11.4) and the
# effective sample size from Geyer's initial positive + monotone sequence as in the Stan reference manual
# ("Effective sample size"), without rank normalisation.  Parity unpinned against the reference (nothing to pin to);
# pinned against closed forms in tests/test_oracle_diag.py (iid draws, AR(1) chains).
# ---------------------------------------------------------------------------------------------------------
def chain_diagnostics(samples, max_lag=0):
    """samples: P x n x C (or P x n).  Returns dict(mean, var, rhat, ess), each of length P, float64.
    Every chain is split in halves (the middle draw of an odd n is dropped): m = 2C sequences of nh = n // 2."""
    x = np.asarray(samples, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    P, n, C = x.shape
    nh = n // 2
    halves = np.concatenate([x[:, :nh, :], x[:, n - nh:, :]], axis=2)  # P x nh x 2C (order does not matter)
    m = 2 * C
    mu = halves.mean(axis=1)                       # P x m
    d = halves - mu[:, None, :]
    var_s = (d * d).sum(axis=1) / (nh - 1)         # P x m, ddof 1
    gm = mu.mean(axis=1)
    W = var_s.mean(axis=1)
    bsum = ((mu - gm[:, None]) ** 2).sum(axis=1)
    var_means = bsum / (m - 1) if m > 1 else np.zeros(P)
    var_plus = W * (nh - 1) / nh + var_means
    total = m * nh
    out = {"mean": gm, "var": (W * (nh - 1) * m + bsum * nh) / (total - 1.0)}
    with np.errstate(divide="ignore", invalid="ignore"):
        out["rhat"] = np.where(W > 0, np.sqrt(var_plus / W), np.nan)
    L = nh
    if max_lag > 0 and max_lag + 1 < L:
        L = max_lag + 1
    ess = np.full(P, np.nan)
    if nh >= 4 and L >= 2:
        acov = np.empty((L, P))
        for k in range(L):                          # biased autocovariance, averaged over the sequences
            acov[k] = (d[:, : nh - k, :] * d[:, k:, :]).sum(axis=1).mean(axis=1) / nh
        for j in range(P):
            if not var_plus[j] > 0:
                continue
            rho = 1.0 - (W[j] - acov[:, j]) / var_plus[j]
            rho[0] = 1.0
            tau, prev, t, last_even = 0.0, np.inf, 0, 1.0
            while t + 1 < L:
                pe, po = rho[t], rho[t + 1]
                pair = pe + po
                if not pair > 0:
                    last_even = pe
                    break
                pair = min(pair, prev)              # initial monotone sequence
                tau += pair
                prev = pair
                t += 2
                last_even = 0.0
            if last_even > 0:
                tau += 0.5 * last_even
            tau_hat = max(-1.0 + 2.0 * tau, 1.0 / math.log10(total))
            ess[j] = total / tau_hat
    out["ess"] = ess
    return out
