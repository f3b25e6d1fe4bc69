"""ctypes binding of oracle/bfx_oracle_c.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (bayesflux.jl_b200/) never does.  The shared object is built by
__graft_entry__.build() (gcc -O3 -fopenmp) into oracle/libbfx_oracle.so; build_if_missing()
rebuilds it on a box where only the sources travelled."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "bfx_oracle_c.c")
SO = os.path.join(HERE, "libbfx_oracle.so")
GCC_FLAGS = ["-O3", "-mavx2", "-mfma", "-fopenmp", "-shared", "-fPIC"]
MAXL = 8


class CModel(C.Structure):
    _fields_ = [("nlayers", C.c_int), ("dims", C.c_int * (MAXL + 1)), ("acts", C.c_int * MAXL),
                ("like_kind", C.c_int), ("nu", C.c_double), ("sprior_kind", C.c_int),
                ("sprior_a", C.c_double), ("sprior_b", C.c_double), ("c0", C.c_double), ("c2", C.c_double)]


def build_if_missing(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc"] + GCC_FLAGS + ["-o", SO, SRC, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_if_missing()
        _lib = C.CDLL(SO)
        _lib.bfxo_num_params.restype = C.c_int
        _lib.bfxo_num_params.argtypes = [C.POINTER(CModel)]
        _lib.bfxo_grad.restype = C.c_double
        _lib.bfxo_grad.argtypes = [C.POINTER(CModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_float, C.c_void_p]
        _lib.bfxo_sgld_run.restype = C.c_int64
        _lib.bfxo_sgld_run.argtypes = [C.POINTER(CModel), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.bfxo_max_threads.restype = C.c_int
    return _lib


def make_model(dims, acts, like_kind=0, nu=0.0, sprior_kind=0, sprior_a=2.0, sprior_b=0.5, prior="gaussian",
               sigma0=1.0, sigma1=1.0, sigma2=1.0, pi1=0.5):
    """dims = [d, out_1, ..., out_L]; acts = activation ids per layer (0 id, 1 relu, 2 tanh, 3 sigmoid).
    Network prior constants as in src/netpriors/gaussian.jl:25-28 / mixturescale.jl:30-33."""
    m = CModel()
    m.nlayers = len(acts)
    for i, d in enumerate(dims):
        m.dims[i] = d
    for i, a in enumerate(acts):
        m.acts[i] = a
    m.like_kind, m.nu = like_kind, nu
    m.sprior_kind, m.sprior_a, m.sprior_b = sprior_kind, sprior_a, sprior_b
    l2pi = np.log(2 * np.pi)
    if prior == "gaussian":
        m.c0 = -0.5 * l2pi - np.log(sigma0)
        m.c2 = 1.0 / sigma0 ** 2
    else:
        pi2 = 1.0 - pi1
        m.c0 = pi1 * (-0.5 * l2pi - np.log(sigma1)) + pi2 * (-0.5 * l2pi - np.log(sigma2))
        m.c2 = pi1 / sigma1 ** 2 + pi2 / sigma2 ** 2
    return m


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grad(m, theta, x, y, idx, nb):
    """(value, gradient) of num_batches*loglike + logprior on minibatch idx (src/model/BNN.jl:50-69)."""
    theta = np.ascontiguousarray(theta, np.float32)
    x = np.asfortranarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    idx = np.ascontiguousarray(idx, np.int64)
    g = np.empty(lib().bfxo_num_params(C.byref(m)), np.float32)
    v = lib().bfxo_grad(C.byref(m), _p(theta), _p(x), _p(y), _p(idx), len(idx), float(nb), _p(g))
    return v, g


def sgld_run(m, x, y, theta, B, nsteps, a=0.1, b=1.0, gamma=0.55, maxnorm=25.0, seed=0, normals=None,
             batches=None, record=False, nthreads=0):
    """mcmc(bnn, B, nsteps, SGLD(...)) for the C chains in theta (P x C, Fortran order), in place.
    normals: (P, C, nsteps) Fortran order or None; batches: (B, C, nsteps) int64 Fortran order or None.
    Returns samples (P, nsteps, C) when record else None."""
    x = np.asfortranarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    assert theta.dtype == np.float32 and theta.flags.f_contiguous
    P, Cn = theta.shape
    n = y.shape[0]
    samples = np.empty((P, nsteps, Cn), np.float32, order="F") if record else None
    if normals is not None:
        normals = np.asfortranarray(normals, np.float32)
        assert normals.shape == (P, Cn, nsteps)
    if batches is not None:
        batches = np.asfortranarray(batches, np.int64)
        assert batches.shape == (B, Cn, nsteps)
    lib().bfxo_sgld_run(C.byref(m), _p(x), _p(y), n, Cn, B, nsteps, a, b, gamma, maxnorm, seed, _p(theta),
                        _p(normals), _p(batches), _p(samples), nthreads)
    return samples


def max_threads():
    return lib().bfxo_max_threads()
