"""ctypes binding of libbfx.so (include/bfx.h).  There is no CPU fallback: if the
library is missing or has not been built, importing this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIBBFX") or os.path.join(_HERE, "libbfx.so")

# names mirror include/bfx.h
OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NAN_G, ERR_NAN_THETA, ERR_NO_DATA, ERR_STATE, ERR_PEER = range(9)
DENSE, RNN, LSTM = 0, 1, 2
IDENTITY, RELU, TANH, SIGMOID = 0, 1, 2, 3
LIKE_NORMAL, LIKE_TDIST = 0, 1
SIGMA_GAMMA, SIGMA_INVGAMMA = 0, 1
PRIOR_GAUSSIAN, PRIOR_MIXTURE = 0, 1
SGLD, SGNHT, SGNHTS, GGMC, HMC, MODE = range(6)
OPT_DESCENT, OPT_MOMENTUM, OPT_RMSPROP, OPT_ADAM = range(4)
STEP_CONSTANT, STEP_DUAL_AVERAGING = 0, 1
MASS_FIXED, MASS_DIAGCOV, MASS_RMSPROP, MASS_FULLCOV = range(4)
(STATE_THETA, STATE_MOMENTUM, STATE_MINV, STATE_XI, STATE_STEPSIZE, STATE_T, STATE_NSAMPLED,
 STATE_STATUS, STATE_CONVERGED) = range(9)
BATCH_PER_CHAIN, BATCH_SHARED = 0, 1
COMPUTE_FP32, COMPUTE_TC_BF16X3 = 0, 1
REGIME_SMEM, REGIME_STREAM = 0, 1
SHARD_NONE, SHARD_CHAINS, SHARD_MINIBATCH = 0, 1, 2
COMM_ID_BYTES = 128


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nin", C.c_int32), ("nout", C.c_int32), ("act", C.c_int32)]


class LikeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sigma_prior_kind", C.c_int32), ("nu", C.c_double),
                ("sigma_prior_a", C.c_double), ("sigma_prior_b", C.c_double)]


class PriorDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("sigma0", C.c_double), ("sigma1", C.c_double),
                ("sigma2", C.c_double), ("pi1", C.c_double)]


class SamplerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("steps", C.c_int32), ("path_len", C.c_int32),
                ("sadapter_kind", C.c_int32), ("madapter_kind", C.c_int32), ("da_t0", C.c_int32),
                ("da_adapt_steps", C.c_int32), ("ma_adapt_steps", C.c_int32), ("ma_windowlength", C.c_int32),
                ("_pad", C.c_int32),
                ("stepsize_a", C.c_float), ("stepsize_b", C.c_float), ("stepsize_gamma", C.c_float),
                ("min_stepsize", C.c_float), ("maxnorm", C.c_float), ("l", C.c_float), ("A", C.c_float),
                ("sigmaA", C.c_float), ("xi", C.c_float), ("mu", C.c_float), ("beta", C.c_float),
                ("da_target_accept", C.c_float), ("da_gamma", C.c_float), ("da_kappa", C.c_float),
                ("ma_kappa", C.c_float), ("ma_epsilon", C.c_float), ("ma_lambda", C.c_float),
                ("ma_alpha", C.c_float),
                ("opt_kind", C.c_int32), ("mode_windowlength", C.c_int32), ("opt_eta", C.c_float),
                ("opt_rho", C.c_float), ("opt_beta2", C.c_float), ("opt_eps", C.c_float),
                ("mode_abstol", C.c_float), ("_pad2", C.c_float)]


class RunDesc(C.Structure):
    _fields_ = [("batchsize", C.c_int64), ("numsamples", C.c_int64), ("shuffle", C.c_int32),
                ("partial", C.c_int32), ("continue_sampling", C.c_int32), ("keep_every", C.c_int32),
                ("batch_mode", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64),
                ("chain_offset", C.c_int64), ("ngpus", C.c_int32), ("shard_mode", C.c_int32)]


# every symbol include/bfx.h declares: name -> argtypes
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_pf, _pd, _pi64, _pi32 = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
SYMBOLS = {
    "bfx_version": [],
    "bfx_last_error": [C.c_char_p, C.c_size_t],
    "bfx_device_count": [_pi32],
    "bfx_model_create": [C.POINTER(LayerDesc), _i32, C.POINTER(LikeDesc), C.POINTER(PriorDesc), _i32,
                         C.POINTER(_vp)],
    "bfx_model_destroy": [_vp],
    "bfx_model_num_params": [_vp, _pi64, _pi64],
    "bfx_model_layer_bounds": [_vp, _i32, _pi64, _pi64],
    "bfx_model_set_data": [_vp, _vp, _pi64, _i32, _vp, _i64],
    "bfx_model_set_data_device": [_vp, _vp, _pi64, _i32, _vp, _i64],
    "bfx_model_set_compute": [_vp, _i32],
    "bfx_model_get_compute": [_vp, _pi32],
    "bfx_loglikeprior": [_vp, _vp, _i32, _vp, _i64, _i64, _f32, _vp],
    "bfx_grad_loglikeprior": [_vp, _vp, _i32, _vp, _i64, _i64, _f32, _vp, _vp],
    "bfx_predict": [_vp, _vp, _i32, _vp, _pi64, _i32, _vp, _vp],
    "bfx_chain_diagnostics": [_i32, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    "bfx_sampler_create": [C.POINTER(SamplerDesc), _vp, _i32, C.POINTER(_vp)],
    "bfx_sampler_destroy": [_vp],
    "bfx_sampler_init": [_vp, _vp],
    "bfx_sampler_get_state": [_vp, _i32, _vp, _i64],
    "bfx_sampler_set_state": [_vp, _i32, _vp, _i64],
    "bfx_sampler_step_injected": [_vp, _vp, _i64, _i64, _f32, _vp, _vp, _vp],
    "bfx_mcmc_num_kept": [_vp, C.POINTER(RunDesc), _pi64, _pi64],
    "bfx_mcmc_run": [_vp, C.POINTER(RunDesc), _vp, _vp, _vp, _vp],
    "bfx_progress": [_vp, _pi64],
    "bfx_mcmc_advance": [_vp, C.POINTER(RunDesc), _i64, _vp, _vp, _i64],
    "bfx_sampler_launch_count": [_vp, _pi64],
    "bfx_host_register": [_vp, C.c_size_t],
    "bfx_host_unregister": [_vp],
    "bfx_model_regime": [_vp, _pi32],
    "bfx_stream_grad_local": [_vp, C.POINTER(RunDesc), _f32, _vp, _vp, _i64],
    "bfx_stream_apply_update": [_vp, C.POINTER(RunDesc), _f32, _vp, _vp, _vp],
    "bfx_stream_allreduce": [_vp, _vp, _vp],
    "bfx_comm_unique_id": [_vp],
    "bfx_sampler_comm_init": [_vp, _vp, _i32, _i32],
    "bfx_sampler_comm_destroy": [_vp],
    "bfx_sampler_graph_replays": [_vp, _pi64],
    "bfx_sampler_peer_ranks": [_vp, C.POINTER(C.c_int32)],
    "bfx_debug_batch_indices": [_vp, C.POINTER(RunDesc), _i64, _i64, _vp, _pi64],
    "bfx_debug_noise": [_vp, C.POINTER(RunDesc), _i64, _i64, _i32, _vp, _vp],
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
        "bayesflux_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
for _name, _args in SYMBOLS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = C.c_int32


class BfxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def last_error():
    buf = C.create_string_buffer(512)
    lib.bfx_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(code):
    if code != OK:
        raise BfxError(code, last_error())
