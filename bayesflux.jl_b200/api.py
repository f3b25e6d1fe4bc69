"""Host-side mirror of BayesFlux's Julia API for the hot path, bound to libbfx.

Same names, argument meaning and error behaviour as the reference so the parity tests
read like the reference's own tests (all file:line relative to the reference tree):

    Chain / Dense / RNN / LSTM      the Flux layer *descriptions* destruct() walks
    destruct -> NetConstructor      src/model/deconstruct.jl:21-58
    FeedforwardNormal / FeedforwardTDist / SeqToOneNormal / SeqToOneTDist
                                    src/likelihoods/feedforward.jl, seq_to_one.jl
    GaussianPrior / MixtureScalePrior   src/netpriors/*.jl
    InitialiseAllSame               src/initialisers/basics.jl:8-27
    BNN, split_params, loglikeprior, grad_loglikeprior (∇loglikeprior)
                                    src/model/BNN.jl:5-69
    SGLD, SGNHT, SGNHTS, GGMC, HMC  src/inference/mcmc/*.jl
    ConstantStepsize, DualAveragingStepSize, FixedMassAdapter, DiagCovMassAdapter,
    RMSPropMassAdapter, FullCovMassAdapter (rejected)   src/inference/mcmc/adapters
    mcmc(bnn, batchsize, numsamples, sampler; ...)      src/inference/mcmc/abstract.jl:74-127

Arrays follow Julia's shapes: theta is (P,) or (P, C); samples are (P, numsamples) for one
chain and (P, numsamples, C) for many; x is (d, n) or (T, d, n).  All compute happens in the
CUDA engine; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L

# ------------------------------------------------------------------ layers ---
identity, relu, tanh, sigmoid = L.IDENTITY, L.RELU, L.TANH, L.SIGMOID


@dataclass(frozen=True)
class Dense:
    nin: int
    nout: int
    act: int = identity
    kind: int = L.DENSE


@dataclass(frozen=True)
class RNN:
    nin: int
    nout: int
    act: int = tanh
    kind: int = L.RNN


@dataclass(frozen=True)
class LSTM:
    nin: int
    nout: int
    act: int = tanh
    kind: int = L.LSTM


class Chain:
    def __init__(self, *layers):
        self.layers = tuple(layers)

    def __iter__(self):
        return iter(self.layers)

    def __len__(self):
        return len(self.layers)


def _layer_nparams(l) -> int:
    if l.kind == L.DENSE:
        return l.nin * l.nout + l.nout
    g = 4 if l.kind == L.LSTM else 1
    return g * (l.nin * l.nout + l.nout * l.nout + l.nout)


@dataclass
class NetConstructor:
    """src/model/deconstruct.jl:21-27; starts/ends are 1-based inclusive like the reference."""
    num_params_network: int
    layers: tuple
    starts: List[int]
    ends: List[int]


def destruct(net: Chain) -> NetConstructor:
    s, starts, ends = 1, [], []
    for l in net:
        starts.append(s)
        s += _layer_nparams(l)
        ends.append(s - 1)
    return NetConstructor(s - 1, tuple(net), starts, ends)


# ------------------------------------------------------------ distributions ---
@dataclass(frozen=True)
class Gamma:
    """Distributions.Gamma(shape α, scale θ)"""
    alpha: float = 1.0
    theta: float = 1.0


@dataclass(frozen=True)
class InverseGamma:
    """Distributions.InverseGamma(shape α, scale β)"""
    alpha: float = 1.0
    beta: float = 1.0


@dataclass(frozen=True)
class Normal:
    mu: float = 0.0
    sigma: float = 1.0

    def rand(self, rng, n):
        return self.mu + self.sigma * rng.standard_normal(n)


@dataclass(frozen=True)
class Uniform:
    a: float = 0.0
    b: float = 1.0

    def rand(self, rng, n):
        return rng.uniform(self.a, self.b, n)


# -------------------------------------------------------------- likelihoods ---
class BNNLikelihood:
    num_params_like = 1
    seq = False
    kind = L.LIKE_NORMAL
    nu = 0.0

    def __init__(self, nc: NetConstructor, prior_sigma, nu: float = 0.0):
        if not isinstance(prior_sigma, (Gamma, InverseGamma)):
            raise TypeError("prior_σ must be Gamma or InverseGamma (positive support, log link)")
        self.nc = nc
        self.prior_sigma = prior_sigma
        self.nu = float(nu)

    def _desc(self) -> L.LikeDesc:
        p = self.prior_sigma
        if isinstance(p, Gamma):
            return L.LikeDesc(self.kind, L.SIGMA_GAMMA, self.nu, p.alpha, p.theta)
        return L.LikeDesc(self.kind, L.SIGMA_INVGAMMA, self.nu, p.alpha, p.beta)


class FeedforwardNormal(BNNLikelihood):
    """src/likelihoods/feedforward.jl:21-28"""


class FeedforwardTDist(BNNLikelihood):
    """src/likelihoods/feedforward.jl:76-84"""
    kind = L.LIKE_TDIST

    def __init__(self, nc, prior_sigma, nu):
        super().__init__(nc, prior_sigma, nu)


class SeqToOneNormal(BNNLikelihood):
    """src/likelihoods/seq_to_one.jl:22-29"""
    seq = True


class SeqToOneTDist(BNNLikelihood):
    """src/likelihoods/seq_to_one.jl:79-87"""
    seq = True
    kind = L.LIKE_TDIST

    def __init__(self, nc, prior_sigma, nu):
        super().__init__(nc, prior_sigma, nu)


# ------------------------------------------------------------------- priors ---
class NetworkPrior:
    num_params_hyper = 0


class GaussianPrior(NetworkPrior):
    """src/netpriors/gaussian.jl:6-23"""

    def __init__(self, nc: NetConstructor, sigma0: float = 1.0):
        self.nc, self.sigma0 = nc, float(sigma0)

    def _desc(self):
        return L.PriorDesc(L.PRIOR_GAUSSIAN, 0, self.sigma0, 1.0, 1.0, 0.5)

    def sample_prior(self, rng):
        return (self.sigma0 * rng.standard_normal(self.nc.num_params_network)).astype(np.float32)


class MixtureScalePrior(NetworkPrior):
    """src/netpriors/mixturescale.jl:18-28; argument order (nc, σ1, σ2, π1)"""

    def __init__(self, nc: NetConstructor, sigma1: float, sigma2: float, pi1: float):
        self.nc, self.sigma1, self.sigma2, self.pi1 = nc, float(sigma1), float(sigma2), float(pi1)
        self.pi2 = 1.0 - self.pi1

    def _desc(self):
        return L.PriorDesc(L.PRIOR_MIXTURE, 0, 1.0, self.sigma1, self.sigma2, self.pi1)


class InitialiseAllSame:
    """src/initialisers/basics.jl:8-27 (host-side draw of the default θstart)"""

    def __init__(self, dist, like: BNNLikelihood, prior: NetworkPrior, seed=None):
        self.dist = dist
        self.length_net = like.nc.num_params_network
        self.length_hyper = prior.num_params_hyper
        self.length_like = like.num_params_like
        self.rng = np.random.default_rng(seed)

    def __call__(self, rng=None):
        rng = rng or self.rng
        f = lambda n: np.asarray(self.dist.rand(rng, n), np.float32)
        return f(self.length_net), f(self.length_hyper), f(self.length_like)


# ---------------------------------------------------------------------- BNN ---
def _as_f32(a, order="F"):
    return np.require(a, dtype=np.float32, requirements=["F" if order == "F" else "C", "A"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BNN:
    """BNN(x, y, like, prior, init)  src/model/BNN.jl:5-38.  Owns the device model handle."""

    def __init__(self, x, y, like: BNNLikelihood, prior: NetworkPrior, init=None, device: int = 0,
                 compute: str = "fp32"):
        """compute: "fp32" (CUDA-core FMA path, the correctness gate), "tc" (tcgen05 tensor cores with
        the 3-term BF16 split; raises BfxError(ERR_UNSUPPORTED) for nets it does not cover) or "auto"
        (tensor cores when the net is eligible, FP32 otherwise)."""
        self.like, self.prior, self.init = like, prior, init
        self.device = device
        nc = like.nc
        self.start_net = 1
        self.start_hyper = 1 + nc.num_params_network
        self.start_like = self.start_hyper
        self.num_total_params = nc.num_params_network + like.num_params_like + prior.num_params_hyper
        layers = (L.LayerDesc * len(nc.layers))(*[L.LayerDesc(l.kind, l.nin, l.nout, l.act) for l in nc.layers])
        ld, pd = like._desc(), prior._desc()
        h = C.c_void_p()
        L.check(L.lib.bfx_model_create(layers, len(nc.layers), C.byref(ld), C.byref(pd), device, C.byref(h)))
        self._h = h
        self.compute = "fp32"
        if compute in ("tc", "auto"):
            code = L.lib.bfx_model_set_compute(h, L.COMPUTE_TC_BF16X3)
            if code == L.OK:
                self.compute = "tc"
            elif compute == "tc" or code != L.ERR_UNSUPPORTED:
                L.check(code)
        elif compute != "fp32":
            raise ValueError("compute must be 'fp32', 'tc' or 'auto'")
        self.x = self.y = None
        if x is not None:
            self.set_data(x, y)

    def set_data(self, x, y):
        x = _as_f32(x)
        y = _as_f32(np.asarray(y).reshape(-1))
        if x.ndim not in (2, 3):
            raise ValueError("x must be d x n or T x d x n")
        dims = (C.c_int64 * x.ndim)(*x.shape)
        L.check(L.lib.bfx_model_set_data(self._h, _ptr(x), dims, x.ndim, _ptr(y), y.shape[0]))
        self.x, self.y = x, y

    def set_data_device(self, x_ptr: int, shape, y_ptr: int):
        """x / y already resident on this model's device (raw device pointers)."""
        dims = (C.c_int64 * len(shape))(*shape)
        L.check(L.lib.bfx_model_set_data_device(self._h, C.c_void_p(x_ptr), dims, len(shape),
                                                C.c_void_p(y_ptr), shape[-1]))
        self.x, self.y = None, None
        self._n = shape[-1]

    @property
    def n(self):
        return self.y.shape[0] if self.y is not None else self._n

    def layer_bounds(self, i):
        s, e = C.c_int64(), C.c_int64()
        L.check(L.lib.bfx_model_layer_bounds(self._h, i, C.byref(s), C.byref(e)))
        return s.value, e.value

    def close(self):
        if getattr(self, "_h", None):
            L.lib.bfx_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_params(bnn: BNN, theta):
    """src/model/BNN.jl:40-45"""
    nnet = bnn.like.nc.num_params_network
    return theta[:nnet], theta[nnet:nnet], theta[nnet:]


def _theta_cols(bnn, theta):
    th = _as_f32(theta)
    if th.ndim == 1:
        th = th.reshape(-1, 1, order="F")
    if th.shape[0] != bnn.num_total_params:
        raise ValueError(f"theta must have {bnn.num_total_params} rows")
    return th, th.shape[1]


def _idx_args(idx, nchains):
    if idx is None:
        return None, 0, 0, None
    idx = np.require(idx, dtype=np.int64, requirements=["F", "A"])
    if idx.ndim == 1:
        return _ptr(idx), idx.shape[0], 0, idx
    if idx.shape[1] != nchains:
        raise ValueError("per-chain idx must be B x C")
    return _ptr(idx), idx.shape[0], idx.shape[0], idx


def loglikeprior(bnn: BNN, theta, idx=None, num_batches=1.0):
    """loglikeprior(bnn, θ, x, y; num_batches) on the minibatch ``idx`` (0-based observation
    indices, None = all data).  src/model/BNN.jl:50-56.  Returns float (1 chain) or (C,) array."""
    th, nc = _theta_cols(bnn, theta)
    p, B, stride, keep = _idx_args(idx, nc)
    v = np.empty(nc, np.float64)
    L.check(L.lib.bfx_loglikeprior(bnn._h, _ptr(th), nc, p, B, stride, float(num_batches), _ptr(v)))
    return float(v[0]) if np.ndim(theta) == 1 else v


def grad_loglikeprior(bnn: BNN, theta, idx=None, num_batches=1.0):
    """∇loglikeprior -> (v, g).  src/model/BNN.jl:61-69"""
    th, nc = _theta_cols(bnn, theta)
    p, B, stride, keep = _idx_args(idx, nc)
    v = np.empty(nc, np.float64)
    g = np.empty(th.shape, np.float32, order="F")
    L.check(L.lib.bfx_grad_loglikeprior(bnn._h, _ptr(th), nc, p, B, stride, float(num_batches), _ptr(v), _ptr(g)))
    if np.ndim(theta) == 1:
        return float(v[0]), g[:, 0]
    return v, g


def predict(bnn: BNN, theta, x=None):
    """Network outputs vec(net(x)) for every column of theta, and sigma = exp(theta_like) per column.
    Returns (yhat (n, C), sigma (C,)); x = None uses the model's data."""
    th, nc = _theta_cols(bnn, theta)
    if x is None:
        n, xp, dims, nd = bnn.n, None, None, 0
    else:
        xx = _as_f32(x)
        n, xp, nd = xx.shape[-1], _ptr(xx), xx.ndim
        dims = (C.c_int64 * xx.ndim)(*xx.shape)
    yhat = np.empty((n, nc), np.float32, order="F")
    sig = np.empty(nc, np.float32)
    L.check(L.lib.bfx_predict(bnn._h, _ptr(th), nc, xp, dims, nd, _ptr(yhat), _ptr(sig)))
    return yhat, sig


def chain_diagnostics(samples, max_lag: int = 0, device: int = 0, samples_dev: int = 0, shape=None):
    """Pooled mean / variance, split-Rhat and effective sample size of every parameter, computed on the device
    (include/bfx.h bfx_chain_diagnostics; SURVEY.md section 8f #3 -- the reference defers this to MCMCChains,
    README.md:191-193).  ``samples``: the P x n x C array ``mcmc`` returns (or P x n for one chain); alternatively
    ``samples_dev`` = device pointer of a [C][n][P] ring (``advance(..., ring_slots=n)``) with ``shape=(P, n, C)``.
    Returns a dict of four float32 vectors of length P."""
    if samples_dev:
        P, n, Cn = (int(v) for v in shape)
        ptr, on_dev = C.c_void_p(samples_dev), 1
    else:
        a = np.asarray(samples, dtype=np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        a = np.asfortranarray(a)
        P, n, Cn = a.shape
        ptr, on_dev = _ptr(a), 0
    out = {k: np.empty(P, np.float32) for k in ("mean", "var", "rhat", "ess")}
    L.check(L.lib.bfx_chain_diagnostics(int(device), ptr, on_dev, P, n, Cn, int(max_lag), _ptr(out["mean"]),
                                        _ptr(out["var"]), _ptr(out["rhat"]), _ptr(out["ess"])))
    return out


def sample_posterior_predict(bnn: BNN, ch, x=None, rng=None):
    """sample_posterior_predict(bnn, ch; x = bnn.x)  src/model/BNN.jl:129-135: one predictive draw per posterior
    draw (columns of ch).  The forward passes run on the device; the draw around yhat is made here on the host
    (posterior_predict: Normal -> yhat + sigma * randn, TDist -> yhat + sigma * rand(TDist(nu)))."""
    rng = rng or np.random.default_rng()
    yhat, sig = predict(bnn, ch, x)
    if bnn.like.kind == L.LIKE_NORMAL:
        eps = rng.standard_normal(yhat.shape)
    else:
        eps = rng.standard_t(bnn.like.nu, yhat.shape)
    return (yhat + sig[None, :] * eps).astype(np.float32)


def get_posterior_networks(bnn: BNN, ch):
    """get_posterior_networks  src/model/BNN.jl:104-107: the network parameter vectors of the posterior draws."""
    ch = np.asarray(ch)
    return [split_params(bnn, ch[:, i])[0].astype(np.float32) for i in range(ch.shape[1])]


# ------------------------------------------------------------------ adapters ---
@dataclass
class ConstantStepsize:
    l: float


@dataclass
class DualAveragingStepSize:
    """adapters/stepsize/dualaveragestepsize.jl:25-32"""
    l: float
    target_accept: float = 0.65
    gamma: float = 0.05
    t0: int = 10
    kappa: float = 0.75
    adapt_steps: int = 1000


@dataclass
class FixedMassAdapter:
    pass


@dataclass
class DiagCovMassAdapter:
    """adapters/mass/diagcovariancemassadapter.jl:23-33"""
    adapt_steps: int
    windowlength: int
    kappa: float = 0.5
    epsilon: float = 1e-6


@dataclass
class RMSPropMassAdapter:
    """adapters/mass/rmspropmassadapter.jl:17-24"""
    adapt_steps: int = 1000
    lam: float = 1e-5
    alpha: float = 0.99


@dataclass
class FullCovMassAdapter:
    """Dense P x P mass matrices are outside the engine's scope: creating a sampler with this
    adapter raises BfxError(ERR_UNSUPPORTED)."""
    adapt_steps: int
    windowlength: int
    kappa: float = 0.5
    epsilon: float = 1e-6


# ------------------------------------------------------------------ samplers ---
class MCMCState:
    """src/inference/mcmc/abstract.jl:1-28: carries samples (P x numsamples [x C]) and nsampled."""
    kind = -1

    def __init__(self):
        self.samples = None
        self.nsampled = 0
        self.t = 1
        self._h = None
        self._bnn = None
        self._nchains = 0
        self.kinetic = None
        self.accepted = None
        self._samples3 = None

    # -- descriptor ------------------------------------------------------------
    def _desc(self) -> L.SamplerDesc:
        d = L.SamplerDesc()
        d.kind = self.kind
        d.steps = getattr(self, "steps", 1)
        d.path_len = getattr(self, "path_len", 1)
        d.stepsize_a = getattr(self, "stepsize_a", 0.0)
        d.stepsize_b = getattr(self, "stepsize_b", 0.0)
        d.stepsize_gamma = getattr(self, "stepsize_gamma", 0.0)
        d.min_stepsize = getattr(self, "min_stepsize", -math.inf)
        d.maxnorm = getattr(self, "maxnorm", 5.0)
        d.l = getattr(self, "l0", 0.0)
        d.A = getattr(self, "A", 1.0)
        d.sigmaA = getattr(self, "sigmaA", 1.0)
        d.xi = getattr(self, "xi0", 0.0)
        d.mu = getattr(self, "mu", 1.0)
        d.beta = getattr(self, "beta", 0.55)
        sa = getattr(self, "sadapter", None)
        if isinstance(sa, DualAveragingStepSize):
            d.sadapter_kind = L.STEP_DUAL_AVERAGING
            d.da_t0, d.da_adapt_steps = sa.t0, sa.adapt_steps
            d.da_target_accept, d.da_gamma, d.da_kappa = sa.target_accept, sa.gamma, sa.kappa
            d.l = sa.l  # initialise!: s.l = s.sadapter.l (hmc.jl:95)
        else:
            d.sadapter_kind = L.STEP_CONSTANT
            if isinstance(sa, ConstantStepsize):
                d.l = sa.l
        ma = getattr(self, "madapter", None)
        if isinstance(ma, DiagCovMassAdapter):
            d.madapter_kind = L.MASS_DIAGCOV
            d.ma_adapt_steps, d.ma_windowlength = ma.adapt_steps, ma.windowlength
            d.ma_kappa, d.ma_epsilon = ma.kappa, ma.epsilon
        elif isinstance(ma, RMSPropMassAdapter):
            d.madapter_kind = L.MASS_RMSPROP
            d.ma_adapt_steps, d.ma_lambda, d.ma_alpha = ma.adapt_steps, ma.lam, ma.alpha
        elif isinstance(ma, FullCovMassAdapter):
            d.madapter_kind = L.MASS_FULLCOV
        else:
            d.madapter_kind = L.MASS_FIXED
        opt = getattr(self, "opt", None)
        if opt is not None:
            d.opt_kind, d.opt_eta, d.opt_rho, d.opt_beta2, d.opt_eps = opt.kind, opt.eta, opt.rho, opt.beta2, opt.eps
            d.mode_windowlength, d.mode_abstol = self.windowlength, self.eps
        return d

    # -- device handle ------------------------------------------------------------
    def _bind(self, bnn: BNN, nchains: int):
        if self._h is not None and self._bnn is bnn and self._nchains == nchains:
            return
        self.close()
        d = self._desc()
        h = C.c_void_p()
        L.check(L.lib.bfx_sampler_create(C.byref(d), bnn._h, nchains, C.byref(h)))
        self._h, self._bnn, self._nchains = h, bnn, nchains

    def initialise(self, bnn: BNN, theta, nchains=None):
        """initialise!(sampler, θ, numsamples) for the injected-step entry."""
        th, nc = _theta_cols(bnn, theta)
        self._bind(bnn, nchains or nc)
        L.check(L.lib.bfx_sampler_init(self._h, _ptr(th)))
        self.nsampled, self.t = 0, 1

    def get_state(self, field_id, dtype, per_param):
        P, Cn = self._bnn.num_total_params, self._nchains
        out = np.empty((P, Cn) if per_param else (Cn,), dtype, order="F")
        L.check(L.lib.bfx_sampler_get_state(self._h, field_id, _ptr(out), out.nbytes))
        return out

    def set_state(self, field_id, arr, dtype):
        a = np.require(arr, dtype=dtype, requirements=["F", "A"])
        L.check(L.lib.bfx_sampler_set_state(self._h, field_id, _ptr(a), a.nbytes))

    @property
    def theta(self):
        return self.get_state(L.STATE_THETA, np.float32, True)

    @property
    def momentum(self):
        return self.get_state(L.STATE_MOMENTUM, np.float32, True)

    p = momentum

    @property
    def Minv(self):
        return self.get_state(L.STATE_MINV, np.float32, True)

    @property
    def xi(self):
        return self.get_state(L.STATE_XI, np.float32, False)

    @property
    def l(self):
        return self.get_state(L.STATE_STEPSIZE, np.float32, False)

    def step_injected(self, idx, num_batches, normals, uniforms=None):
        """update!(s, θ, bnn, ∇θ) with explicit draws.  normals: (ndraws, P, C) or (P,) / (P, C)."""
        P, Cn = self._bnn.num_total_params, self._nchains
        nd = 2 if self.kind == L.GGMC else 1
        nrm = np.asarray(normals, np.float32).reshape(nd, P, Cn) if np.ndim(normals) != 3 else np.asarray(normals, np.float32)
        # ABI layout: ndraws x P x C with P fastest  ==  C-order array [nd][C][P]
        buf = np.ascontiguousarray(np.transpose(nrm, (0, 2, 1)))
        u = None if uniforms is None else np.ascontiguousarray(np.asarray(uniforms, np.float32).reshape(Cn))
        p, B, stride, keep = _idx_args(idx, Cn)
        out = np.empty((P, Cn), np.float32, order="F")
        L.check(L.lib.bfx_sampler_step_injected(self._h, p, B, stride, float(num_batches), _ptr(buf), _ptr(u), _ptr(out)))
        return out

    def calculate_epochs(self, nbatches, nsamples, continue_sampling=False):
        n_new = nsamples - self.nsampled if continue_sampling else nsamples
        return -(-(n_new * getattr(self, "path_len", 1)) // nbatches)

    def close(self):
        for ptr in list(self.__dict__.get("_pinned", {})):
            L.lib.bfx_host_unregister(C.c_void_p(ptr))
        self.__dict__.pop("_pinned", None)
        if self._h is not None:
            L.lib.bfx_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SGLD(MCMCState):
    """src/inference/mcmc/sgld.jl:47-54"""
    kind = L.SGLD

    def __init__(self, stepsize_a=0.1, stepsize_b=1.0, stepsize_gamma=0.55, min_stepsize=-math.inf, maxnorm=25.0):
        super().__init__()
        self.stepsize_a, self.stepsize_b, self.stepsize_gamma = stepsize_a, stepsize_b, stepsize_gamma
        self.min_stepsize, self.maxnorm = min_stepsize, maxnorm


class SGNHT(MCMCState):
    """SGNHT(l, A = 1; xi = 0)  src/inference/mcmc/sgnht.jl:32"""
    kind = L.SGNHT

    def __init__(self, l, A=1.0, xi=0.0):
        super().__init__()
        self.l0, self.A, self.xi0 = l, A, xi


class SGNHTS(MCMCState):
    """SGNHTS(l, σA = 1; xi = 1, μ = 1, madapter = FixedMassAdapter())  sgnht-s.jl:41-47"""
    kind = L.SGNHTS

    def __init__(self, l, sigmaA=1.0, xi=1.0, mu=1.0, madapter=None):
        super().__init__()
        self.l0, self.sigmaA, self.xi0, self.mu = l, sigmaA, xi, mu
        self.madapter = madapter or FixedMassAdapter()


class GGMC(MCMCState):
    """GGMC(; β, l, sadapter, madapter, steps, maxnorm)  ggmc.jl:59-67"""
    kind = L.GGMC

    def __init__(self, beta=0.55, l=1e-4, sadapter=None, madapter=None, steps=1, maxnorm=5.0):
        super().__init__()
        self.beta, self.l0, self.steps, self.maxnorm = beta, l, steps, maxnorm
        self.sadapter = sadapter or DualAveragingStepSize(l)
        self.madapter = madapter or DiagCovMassAdapter(1000, 100)


class HMC(MCMCState):
    """HMC(l, path_len; sadapter, madapter, maxnorm)  hmc.jl:59-65.  The reference's default
    madapter is FullCovMassAdapter, which this engine rejects: pass a diagonal adapter."""
    kind = L.HMC

    def __init__(self, l, path_len, sadapter=None, madapter=None, maxnorm=5.0):
        super().__init__()
        self.l0, self.path_len, self.maxnorm = l, path_len, maxnorm
        self.sadapter = sadapter or DualAveragingStepSize(l)
        self.madapter = madapter if madapter is not None else FullCovMassAdapter(1000, 100)


# ---------------------------------------------------------------- find_mode ---
@dataclass
class _FluxOpt:
    kind: int
    eta: float
    rho: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8


def Descent(eta=0.1):
    """Flux.Optimise.Descent(η)"""
    return _FluxOpt(L.OPT_DESCENT, eta)


def Momentum(eta=0.01, rho=0.9):
    """Flux.Optimise.Momentum(η, ρ)"""
    return _FluxOpt(L.OPT_MOMENTUM, eta, rho)


def RMSProp(eta=0.001, rho=0.9, eps=1e-8):
    """Flux.Optimise.RMSProp(η, ρ, ϵ)"""
    return _FluxOpt(L.OPT_RMSPROP, eta, rho, eps=eps)


def ADAM(eta=0.001, beta=(0.9, 0.999), eps=1e-8):
    """Flux.Optimise.ADAM(η, β, ϵ)"""
    return _FluxOpt(L.OPT_ADAM, eta, beta[0], beta[1], eps)


class FluxModeFinder(MCMCState):
    """FluxModeFinder(bnn, opt; windowlength = 100, ϵ = 1e-6)  src/inference/mode/flux.jl:8-31"""
    kind = L.MODE

    def __init__(self, bnn, opt, windowlength=100, eps=1e-6):
        super().__init__()
        self.bnn, self.opt, self.windowlength, self.eps = bnn, opt, windowlength, eps
        self.converged = None


def find_mode(bnn: BNN, batchsize: int, epochs: int, optimiser: FluxModeFinder, shuffle=True, partial=True,
              showprogress=False, theta_start=None, nchains: int = 1, seed: int = 0):
    """find_mode(bnn, batchsize, epochs, optimiser; shuffle, partial)  src/inference/mode/abstract.jl:36-86, all
    epochs on the device.  Engine extensions: theta_start (default: draws of bnn.init, one per chain), nchains
    (many restarts at once), seed.  Returns θmode (P,) or (P, C); optimiser.converged tells which chains met the
    convergence window (the reference logs "Converged!" / "Failed to converge.")."""
    if theta_start is None:
        if bnn.init is None:
            raise ValueError("theta_start or bnn.init required")
        theta_start = np.stack([np.concatenate(bnn.init()) for _ in range(nchains)], axis=1)
    th, nc = _theta_cols(bnn, theta_start)
    optimiser._bind(bnn, nc)
    n = bnn.n
    B = min(batchsize, n)
    nb = -(-n // B) if partial else max(1, n // B)
    run = _run_desc(batchsize, epochs * nb, shuffle, partial, False, 1, seed, "per_chain", 0)
    code = L.lib.bfx_mcmc_run(optimiser._h, C.byref(run), _ptr(th), None, None, None)
    L.check(code)
    optimiser.converged = optimiser.get_state(L.STATE_CONVERGED, np.int32, False).astype(bool)
    theta = optimiser.theta
    return theta[:, 0] if np.ndim(theta_start) == 1 or nc == 1 and nchains == 1 else theta


# --------------------------------------------------------------------- mcmc ---
_SHARD = {None: L.SHARD_NONE, "none": L.SHARD_NONE, "chains": L.SHARD_CHAINS, "minibatch": L.SHARD_MINIBATCH}


def _run_desc(batchsize, numsamples, shuffle, partial, continue_sampling, keep_every, seed, batch_mode,
              chain_offset=0, ngpus=1, shard_mode=None):
    return L.RunDesc(batchsize, numsamples, int(shuffle), int(partial), int(continue_sampling), keep_every,
                     L.BATCH_SHARED if batch_mode == "shared" else L.BATCH_PER_CHAIN, 0, seed, chain_offset,
                     int(ngpus), _SHARD[shard_mode])


def mcmc(bnn: BNN, batchsize: int, numsamples: int, sampler: MCMCState, shuffle=True, partial=True,
         showprogress=False, continue_sampling=False, theta_start=None, nchains: Optional[int] = None,
         keep_every: int = 1, seed: int = 0, batch_mode: str = "per_chain", chain_offset: int = 0, out=None,
         ngpus: int = 1, shard_mode: Optional[str] = None):
    """mcmc(bnn, batchsize, numsamples, sampler; shuffle, partial, continue_sampling, θstart)
    src/inference/mcmc/abstract.jl:74-127, run entirely on the device.  Engine extensions:
    nchains (θstart may be P x C), keep_every, seed, batch_mode, and ``out``: a caller-owned Fortran-ordered
    float32 array of shape (P, nkept, nchains) that receives the new samples instead of a fresh allocation
    (repeated calls then pay no page faults under the device-to-host copy).  Returns sampler.samples."""
    if continue_sampling:
        if sampler._h is None or sampler.samples is None:
            raise ValueError("continue_sampling needs a sampler that has already sampled")
        nchains = sampler._nchains
        th = None
    else:
        if theta_start is None:
            if bnn.init is None:
                raise ValueError("theta_start or bnn.init required")
            nchains = nchains or 1
            theta_start = np.stack([np.concatenate(bnn.init()) for _ in range(nchains)], axis=1)
        th, nc = _theta_cols(bnn, theta_start)
        if nchains is not None and nchains != nc:
            if nc != 1:
                raise ValueError("theta_start must have 1 or nchains columns")
            th = np.asfortranarray(np.repeat(th, nchains, axis=1))
        nchains = th.shape[1]
        sampler._bind(bnn, nchains)
    run = _run_desc(batchsize, numsamples, shuffle, partial, continue_sampling, keep_every, seed, batch_mode,
                    chain_offset, ngpus, shard_mode)
    nnew, nkept = C.c_int64(), C.c_int64()
    L.check(L.lib.bfx_mcmc_num_kept(sampler._h, C.byref(run), C.byref(nnew), C.byref(nkept)))
    nnew, nkept = nnew.value, nkept.value
    P = bnn.num_total_params
    if out is not None:
        if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.f_contiguous
                and out.shape == (P, nkept, nchains)):
            raise ValueError(f"out must be a Fortran-ordered float32 array of shape {(P, nkept, nchains)}")
        new = out
        # a caller-owned result buffer is reused call after call: page-lock it once (DMA copies, no bounce buffer);
        # the sampler keeps a reference and unlocks it when it is closed
        pins = sampler.__dict__.setdefault("_pinned", {})
        if out.nbytes >= (1 << 20) and out.ctypes.data not in pins:
            if L.lib.bfx_host_register(C.c_void_p(out.ctypes.data), out.nbytes) == L.OK:
                pins[out.ctypes.data] = out
    else:
        new = np.empty((P, nkept, nchains), np.float32, order="F")
    kin = np.zeros((nnew, nchains), np.float32, order="F") if sampler.kind in (L.SGNHT, L.SGNHTS) else None
    acc = np.zeros((nnew, nchains), np.int32, order="F") if sampler.kind in (L.GGMC, L.HMC) else None
    code = L.lib.bfx_mcmc_run(sampler._h, C.byref(run), _ptr(th), _ptr(new), _ptr(kin), _ptr(acc))
    # initialise!: samples of a continued run keep the already drawn prefix (sgld.jl:77-87)
    if continue_sampling and sampler._samples3 is not None:
        sampler._samples3 = np.concatenate([sampler._samples3, new], axis=1)
        if kin is not None:
            sampler.kinetic = np.concatenate([sampler.kinetic, kin], axis=0)
        if acc is not None:
            sampler.accepted = np.concatenate([sampler.accepted, acc], axis=0)
    else:
        sampler._samples3, sampler.kinetic, sampler.accepted = new, kin, acc
    sampler.nsampled = int(sampler.get_state(L.STATE_NSAMPLED, np.int64, False)[0])
    sampler.t = int(sampler.get_state(L.STATE_T, np.int64, False)[0])
    sampler.samples = sampler._samples3[:, :, 0] if nchains == 1 else sampler._samples3
    L.check(code)  # raises "NaN in g" / "NaN in θ" like the reference (sgnht.jl:70-72,81-83)
    return sampler.samples


def advance(sampler: MCMCState, batchsize: int, nsteps: int, shuffle=True, partial=True, keep_every: int = 1,
            seed: int = 0, batch_mode: str = "per_chain", chain_offset: int = 0, stream: int = 0,
            samples_dev: int = 0, ring_slots: int = 0, ngpus: int = 1, shard_mode: Optional[str] = None):
    """Device-resident continuation of ``mcmc``: enqueue ``nsteps`` update! calls for every chain of an
    initialised sampler (``sampler.initialise(bnn, theta)``) on ``stream`` (a raw cudaStream_t, 0 = the
    engine's own stream) and return immediately.  Recorded samples (every ``keep_every``-th) are written
    to the caller's device ring ``samples_dev`` ([C][ring_slots][P] floats) when given.  Same loop body as
    mcmc (src/inference/mcmc/abstract.jl:117-124); nothing crosses PCIe."""
    run = _run_desc(batchsize, 0, shuffle, partial, True, keep_every, seed, batch_mode, chain_offset, ngpus,
                    shard_mode)
    L.check(L.lib.bfx_mcmc_advance(sampler._h, C.byref(run), int(nsteps), C.c_void_p(stream),
                                   C.c_void_p(samples_dev), int(ring_slots)))


def comm_unique_id() -> bytes:
    """The 128-byte NCCL id of a new communicator (rank 0 creates it, the caller distributes it)."""
    buf = C.create_string_buffer(L.COMM_ID_BYTES)
    L.check(L.lib.bfx_comm_unique_id(buf))
    return buf.raw


def comm_init(sampler: MCMCState, comm_id: bytes, nranks: int, rank: int):
    """Attach the in-library NCCL communicator of a minibatch-sharded chain to an initialised sampler replica
    (collective over the ranks).  Afterwards ``mcmc`` / ``advance`` with ``shard_mode="minibatch"`` all-reduce the
    gradient inside the engine's update graph."""
    assert len(comm_id) == L.COMM_ID_BYTES
    L.check(L.lib.bfx_sampler_comm_init(sampler._h, C.c_char_p(comm_id), int(nranks), int(rank)))


def comm_destroy(sampler: MCMCState):
    L.check(L.lib.bfx_sampler_comm_destroy(sampler._h))


def peer_ranks(sampler: MCMCState) -> int:
    """Ranks whose exchange buffers a minibatch-sharded sampler reads over peer memory (0: NCCL all-reduce)."""
    n = C.c_int32()
    L.check(L.lib.bfx_sampler_peer_ranks(sampler._h, C.byref(n)))
    return n.value


def graph_replays(sampler: MCMCState) -> int:
    n = C.c_int64()
    L.check(L.lib.bfx_sampler_graph_replays(sampler._h, C.byref(n)))
    return n.value


def launch_count(sampler: MCMCState) -> int:
    n = C.c_int64()
    L.check(L.lib.bfx_sampler_launch_count(sampler._h, C.byref(n)))
    return n.value
