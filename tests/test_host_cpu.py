"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/bfx.h declares, rejects bad descriptors with the documented status codes, and the
host mirror reproduces the reference's bookkeeping.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bayesflux_b200 as bf
from bayesflux_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bfx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L.lib, n), f"{n} declared in include/bfx.h but not exported by libbfx.so"
    assert sorted(L.SYMBOLS) == names, "ctypes table out of sync with include/bfx.h"


def test_version_and_error_buffer():
    assert L.lib.bfx_version() == 100
    buf = C.create_string_buffer(64)
    assert L.lib.bfx_last_error(buf, 64) == L.OK


def _mk(layers, like=None, prior=None):
    arr = (L.LayerDesc * len(layers))(*[L.LayerDesc(*l) for l in layers])
    like = like or L.LikeDesc(L.LIKE_NORMAL, L.SIGMA_GAMMA, 0.0, 2.0, 0.5)
    prior = prior or L.PriorDesc(L.PRIOR_GAUSSIAN, 0, 1.0, 1.0, 1.0, 0.5)
    h = C.c_void_p()
    rc = L.lib.bfx_model_create(arr, len(layers), C.byref(like), C.byref(prior), 0, C.byref(h))
    return rc, h


def test_model_create_layout_cfg2():
    # destruct(::Chain) offsets, SURVEY.md 8a row a1 / src/model/deconstruct.jl:45-58
    rc, h = _mk([(L.DENSE, 10, 64, L.RELU), (L.DENSE, 64, 64, L.RELU), (L.DENSE, 64, 1, L.IDENTITY)])
    assert rc == L.OK
    pt, pn = C.c_int64(), C.c_int64()
    assert L.lib.bfx_model_num_params(h, C.byref(pt), C.byref(pn)) == L.OK
    assert (pt.value, pn.value) == (4930, 4929)
    bounds = []
    for i in range(3):
        s, e = C.c_int64(), C.c_int64()
        assert L.lib.bfx_model_layer_bounds(h, i, C.byref(s), C.byref(e)) == L.OK
        bounds.append((s.value, e.value))
    assert bounds == [(0, 704), (704, 4864), (4864, 4929)]
    assert L.lib.bfx_model_destroy(h) == L.OK


def test_model_create_rejects_bad_descriptors():
    rc, _ = _mk([(L.DENSE, 10, 64, L.RELU), (L.DENSE, 32, 1, L.IDENTITY)])
    assert rc == L.ERR_BAD_ARG and "fan-in" in L.last_error()
    rc, _ = _mk([(L.DENSE, 10, 2, L.RELU)])
    assert rc == L.ERR_BAD_ARG and "single network output" in L.last_error()
    rc, _ = _mk([(L.DENSE, 5, 1, 9)])
    assert rc == L.ERR_BAD_ARG
    rc, _ = _mk([(L.DENSE, 5, 1, 0)], like=L.LikeDesc(L.LIKE_TDIST, L.SIGMA_GAMMA, 0.0, 2.0, 0.5))
    assert rc == L.ERR_BAD_ARG
    rc, _ = _mk([(L.DENSE, 5, 1, 0)], prior=L.PriorDesc(L.PRIOR_GAUSSIAN, 0, -1.0, 1.0, 1.0, 0.5))
    assert rc == L.ERR_BAD_ARG
    # beyond the shared-memory regime -> the large-P (stream) regime for Dense nets ...
    rc, h = _mk([(L.DENSE, 64, 512, L.RELU), (L.DENSE, 512, 512, L.RELU), (L.DENSE, 512, 1, 0)])
    assert rc == L.OK
    reg = C.c_int32(-1)
    assert L.lib.bfx_model_regime(h, C.byref(reg)) == L.OK and reg.value == L.REGIME_STREAM
    L.lib.bfx_model_destroy(h)
    # ... and unsupported (not a crash) for a recurrent net that large
    rc, _ = _mk([(L.LSTM, 64, 256, L.TANH), (L.DENSE, 256, 1, 0)])
    assert rc == L.ERR_UNSUPPORTED


def test_fullcov_adapter_rejected():
    rc, h = _mk([(L.DENSE, 5, 1, 0)])
    assert rc == L.OK
    d = L.SamplerDesc()
    d.kind, d.path_len, d.l, d.madapter_kind = L.HMC, 5, 1e-3, L.MASS_FULLCOV
    s = C.c_void_p()
    assert L.lib.bfx_sampler_create(C.byref(d), h, 4, C.byref(s)) == L.ERR_UNSUPPORTED
    assert "FullCov" in L.last_error()
    L.lib.bfx_model_destroy(h)


def test_compute_entries_fail_loudly_without_gpu():
    n = C.c_int32()
    rc = L.lib.bfx_device_count(C.byref(n))
    if rc == L.OK and n.value > 0:
        pytest.skip("a GPU is present")
    rc, h = _mk([(L.DENSE, 5, 1, 0)])
    assert rc == L.OK
    x = np.zeros((5, 8), np.float32, order="F")
    y = np.zeros(8, np.float32)
    dims = (C.c_int64 * 2)(5, 8)
    rc = L.lib.bfx_model_set_data(h, x.ctypes.data_as(C.c_void_p), dims, 2, y.ctypes.data_as(C.c_void_p), 8)
    assert rc == L.ERR_CUDA  # no CPU fallback
    L.lib.bfx_model_destroy(h)


def test_host_mirror_bookkeeping():
    nc = bf.destruct(bf.Chain(bf.Dense(10, 64, bf.relu), bf.Dense(64, 64, bf.relu), bf.Dense(64, 1)))
    assert nc.num_params_network == 4929
    assert nc.starts == [1, 705, 4865] and nc.ends == [704, 4864, 4929]  # 1-based inclusive like Julia
    assert bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1))).num_params_network == 17729
    assert bf.destruct(bf.Chain(bf.RNN(3, 3), bf.Dense(3, 2))).num_params_network == (9 + 9 + 3) + (6 + 2)
    # calculate_epochs: test/sgld.jl:33, test/hmc.jl:50
    s = bf.SGLD()
    s.nsampled = 20_000
    assert s.calculate_epochs(100, 25_000, continue_sampling=True) == 50
    h = bf.HMC(1e-4, 5, madapter=bf.DiagCovMassAdapter(1000, 100))
    h.nsampled = 20_000
    assert h.calculate_epochs(100, 25_000, continue_sampling=True) == 250
    th = np.arange(7, dtype=np.float32)

    class _B:
        like = bf.FeedforwardNormal(bf.destruct(bf.Chain(bf.Dense(5, 1))), bf.Gamma(2.0, 0.5))
    tn, thyp, tl = bf.split_params(_B, th)
    assert tn.size == 6 and thyp.size == 0 and tl.size == 1 and tl[0] == 6


def test_sampler_descriptor_mirrors_reference_defaults():
    d = bf.SGLD()._desc()
    assert (d.kind, round(d.stepsize_a, 6), d.stepsize_b, round(d.stepsize_gamma, 6), d.maxnorm) == (L.SGLD, 0.1, 1.0, 0.55, 25.0)
    d = bf.SGNHTS(1e-3)._desc()
    assert (d.sigmaA, d.xi, d.mu, d.madapter_kind) == (1.0, 1.0, 1.0, L.MASS_FIXED)
    d = bf.GGMC()._desc()
    assert d.sadapter_kind == L.STEP_DUAL_AVERAGING and d.madapter_kind == L.MASS_DIAGCOV
    assert (d.steps, d.maxnorm, d.da_t0, d.da_adapt_steps, d.ma_adapt_steps, d.ma_windowlength) == (1, 5.0, 10, 1000, 1000, 100)
    assert bf.HMC(1e-4, 5)._desc().madapter_kind == L.MASS_FULLCOV  # reference default, rejected by the engine


def test_mcmc_has_no_hidden_result_pool():
    """Results are fresh arrays unless the caller passes its own ``out=`` buffer (no refcount-based recycling)."""
    import inspect
    from bayesflux_b200 import api

    assert "out" in inspect.signature(api.mcmc).parameters
    assert not hasattr(api, "_result_buffer")
    assert "getrefcount" not in inspect.getsource(api)
