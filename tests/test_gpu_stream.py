"""Large-P ("stream") regime and the minibatch-sharded chain against the oracle, through the C ABI.
SURVEY.md section 8e; configs[2] shape class (wide MLP, FeedforwardTDist + MixtureScalePrior, SGNHTS)."""
import numpy as np
import pytest

from oracle import bfx_oracle as O
from tests.test_gpu_parity import LIKES, PRIORS, build, engine_sampler, oracle_sampler

pytestmark = pytest.mark.gpu
SEED = 6150533
TOL = 1e-4

NETS = {
    "wide3": (O.Dense(20, 300, O.RELU), O.Dense(300, 260, O.TANH), O.Dense(260, 1)),
    "cfg3_small": (O.Dense(64, 200, O.RELU), O.Dense(200, 200, O.RELU), O.Dense(200, 200, O.RELU), O.Dense(200, 1)),
    "odd": (O.Dense(7, 333, O.SIGMOID), O.Dense(333, 129, O.RELU), O.Dense(129, 1, O.TANH)),
}


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


@pytest.mark.parametrize("net", list(NETS))
@pytest.mark.parametrize("like,prior", [("normal_gamma", "gauss"), ("tdist_gamma", "mix")])
@pytest.mark.parametrize("B", [50, 700])
def test_stream_grad_and_value_parity(bf, net, like, prior, B):
    rng = np.random.default_rng(SEED)
    layers = NETS[net]
    n = 900
    x = rng.standard_normal((layers[0].nin, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES[like], PRIORS[prior], x, y)
    assert bf.model_regime(bnn) == "stream"
    C = 2
    theta = (0.1 * rng.standard_normal((m.n_total, C))).astype(np.float32)
    idx = np.stack([rng.permutation(n)[:B] for _ in range(C)], axis=1)
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=4.0)
    v2 = bf.loglikeprior(bnn, theta, idx, num_batches=4.0)
    for c in range(C):
        vo, go = O.grad_loglikeprior(m, theta[:, c].astype(np.float64), x[:, idx[:, c]].astype(np.float64),
                                     y[idx[:, c]].astype(np.float64), 4.0)
        assert np.linalg.norm(g[:, c] - go) <= TOL * np.linalg.norm(go), f"chain {c}"
        assert abs(v[c] - vo) <= TOL * abs(vo)
        assert abs(v2[c] - vo) <= TOL * abs(vo)


def test_stream_full_data_value(bf):
    rng = np.random.default_rng(SEED + 1)
    layers = NETS["wide3"]
    n = 10_000   # > one 8192-column chunk
    x = rng.standard_normal((20, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    th = (0.1 * rng.standard_normal(m.n_total)).astype(np.float32)
    v, g = bf.grad_loglikeprior(bnn, th, None, 1.0)
    vo, go = O.grad_loglikeprior(m, th.astype(np.float64), x.astype(np.float64), y.astype(np.float64), 1.0)
    assert abs(v - vo) <= TOL * abs(vo)
    assert np.linalg.norm(g - go) <= TOL * np.linalg.norm(go)


@pytest.mark.parametrize("kind", ["sgld", "sgld_clip", "sgnht", "sgnhts", "sgnhts_rms", "sgnhts_xi0"])
def test_stream_injected_step_parity(bf, kind):
    rng = np.random.default_rng(SEED + 2)
    layers = NETS["wide3"]
    n, B, C, nsteps = 256, 64, 2, 8
    x = rng.standard_normal((20, n)).astype(np.float32)
    y = (0.3 * x[0] + 0.5 * rng.standard_normal(n)).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    P = m.n_total
    theta0 = (0.05 * rng.standard_normal((P, C))).astype(np.float32)
    nb = n // B
    oracles = [O.SGLD(stepsize_a=1e-3) if kind == "sgld" else oracle_sampler(kind, n) for _ in range(C)]
    eng = engine_sampler(bf, oracles[0])
    eng.initialise(bnn, theta0)
    th_o = [theta0[:, c].copy() for c in range(C)]
    for so in oracles:
        so.initialise(theta0[:, 0].copy(), nsteps)
    for t in range(nsteps):
        idx = np.stack([rng.permutation(n)[:B] for _ in range(C)], axis=1).astype(np.int64)
        normals = rng.standard_normal((1, P, C)).astype(np.float32)
        th_e = eng.step_injected(idx, float(nb), normals, None)
        for c in range(C):
            xb, yb = x[:, idx[:, c]], y[idx[:, c]]
            gf = lambda th, xb=xb, yb=yb: O.grad_loglikeprior(m, th, xb, yb, float(nb))
            th_o[c] = oracles[c].update(th_o[c], gf, None, [normals[0, :, c]], [0.5])
            scale = max(1.0, float(np.max(np.abs(th_o[c]))))
            err = float(np.max(np.abs(th_e[:, c] - th_o[c]))) / scale
            assert err < 2e-4, f"{kind}: step {t} chain {c} err {err}"


def test_stream_unsupported_samplers_are_rejected(bf):
    rng = np.random.default_rng(0)
    layers = NETS["wide3"]
    x = rng.standard_normal((20, 64)).astype(np.float32)
    y = rng.standard_normal(64).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    with pytest.raises(bf.BfxError) as e:
        bf.mcmc(bnn, 16, 4, bf.HMC(1e-3, 3, madapter=bf.FixedMassAdapter()), theta_start=np.zeros(m.n_total, np.float32))
    assert e.value.code == bf._lib.ERR_UNSUPPORTED


def test_stream_mcmc_run_resume_and_determinism(bf):
    rng = np.random.default_rng(SEED + 3)
    layers = NETS["cfg3_small"]
    n = 2048
    x = rng.standard_normal((64, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    like = O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5)
    prior = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5)
    bnn, m = build(bf, layers, like, prior, x, y)
    th0 = (0.05 * rng.standard_normal(m.n_total)).astype(np.float32)
    s1 = bf.SGNHTS(1e-3, 1.0)
    ch = bf.mcmc(bnn, 256, 10, s1, theta_start=th0, seed=3)
    assert ch.shape == (m.n_total, 10) and np.isfinite(ch).all()
    assert s1.kinetic.shape == (10, 1) and np.all(s1.kinetic > 0)
    s2 = bf.SGNHTS(1e-3, 1.0)
    ch6 = bf.mcmc(bnn, 256, 6, s2, theta_start=th0, seed=3)
    np.testing.assert_allclose(ch6, ch[:, :6], rtol=0, atol=1e-6)   # double atomics: summation order may differ in the last bit
    ch10 = bf.mcmc(bnn, 256, 10, s2, continue_sampling=True)
    assert ch10.shape == (m.n_total, 10)
    np.testing.assert_allclose(ch10[:, :6], ch6, rtol=0, atol=0)


def test_minibatch_sharded_chain_matches_the_unsharded_update(bf):
    """Two 'ranks' emulated on one GPU (two models holding half of the data each, the all-reduce done by hand):
    the replicated theta after every update equals a single-process update on the union minibatch."""
    import torch
    rng = np.random.default_rng(SEED + 4)
    layers = NETS["cfg3_small"]
    n, Bl = 512, 64
    x = rng.standard_normal((64, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    like = O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5)
    prior = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5)
    halves = [slice(0, n // 2), slice(n // 2, n)]
    bnns = [build(bf, layers, like, prior, np.asfortranarray(x[:, h]), y[h])[0] for h in halves]
    bnn_all, m = build(bf, layers, like, prior, x, y)
    th0 = (0.05 * rng.standard_normal(m.n_total)).astype(np.float32)
    nb = bf.global_num_batches(n, 2 * Bl)
    chains = []

    def fake_allreduce(_buf):  # sum of the two ranks' buffers, written back to both
        total = chains[0].buf + chains[1].buf
        chains[0].buf.copy_(total)
        chains[1].buf.copy_(total)

    for r in range(2):
        chains.append(bf.ShardedChain(bnns[r], bf.SGNHTS(1e-3, 1.0), th0, Bl, nb, seed=7, rank=r, allreduce=lambda b: None))
    ref = bf.SGNHTS(1e-3, 1.0)
    ref.initialise(bnn_all, th0.reshape(-1, 1))
    import ctypes as C
    from bayesflux_b200 import _lib as L
    for t in range(5):
        idx_union = []
        for r in range(2):
            chains[r].grad_local()
            # the local minibatch of rank r at this step (debug hook), mapped to global observation ids
            ib = np.empty(Bl, np.int64)
            Bk = C.c_int64()
            L.check(L.lib.bfx_debug_batch_indices(chains[r].sampler._h, C.byref(chains[r].run), 0, t,
                                                  ib.ctypes.data_as(C.c_void_p), C.byref(Bk)))
            idx_union.append(ib[:Bk.value] + halves[r].start)
        torch.cuda.synchronize()
        fake_allreduce(None)
        for r in range(2):
            chains[r].apply_update()
        torch.cuda.synchronize()
        th_a, th_b = chains[0].theta, chains[1].theta
        np.testing.assert_array_equal(th_a, th_b)      # replicas stay bit-identical
        # single-process reference: same union minibatch, same noise (Philox key: chain 0, step t)
        z = np.empty(m.n_total, np.float32)
        u = np.zeros(1, np.float32)
        L.check(L.lib.bfx_debug_noise(chains[0].sampler._h, C.byref(chains[0].run), 0 - chains[0].run.chain_offset, t, 0,
                                      z.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p)))
        th_ref = ref.step_injected(np.concatenate(idx_union), float(nb), z.reshape(1, -1, 1), None)[:, 0]
        err = float(np.max(np.abs(th_a - th_ref))) / max(1.0, float(np.max(np.abs(th_ref))))
        assert err < 1e-5, (t, err)


# ---------------------------------------------------------------------------------------------------------
# tcgen05 GEMMs of the stream regime (bfx_stream_tc.cu): same 1e-4 bar through the 3-term BF16 split
# ---------------------------------------------------------------------------------------------------------
def relu_margin(layers, theta, xb):
    """Smallest |pre-activation| of a ReLU unit relative to its magnitude scale sum|w||a| + |b|: below ~2e-5 the
    3-term BF16 split (pre-activations accurate to ~1e-5) can land on the other side of the kink, which flips that
    unit's derivative -- a discontinuity of the function, not an arithmetic error."""
    th = theta.astype(np.float64)
    a = xb.astype(np.float64)
    starts, ends = O.layer_bounds(layers)
    worst = np.inf
    for l, s0, e0 in zip(layers, starts, ends):
        W, b = O.unpack_layer(l, th[s0:e0])
        z = W @ a + b[:, None]
        if l.act == O.RELU:
            worst = min(worst, float(np.min(np.abs(z) / (np.abs(W) @ np.abs(a) + np.abs(b)[:, None]))))
        a = O._act(l.act, z)
    return worst


TC_NETS = {
    "cfg3_small": NETS["cfg3_small"],
    "cfg3_small_smooth": (O.Dense(64, 200, O.TANH), O.Dense(200, 200, O.SIGMOID), O.Dense(200, 200, O.TANH), O.Dense(200, 1)),
    "tc_mixed": (O.Dense(24, 256, O.RELU), O.Dense(256, 192, O.TANH), O.Dense(192, 1)),
    "cfg3_one_wide": (O.Dense(64, 512, O.RELU), O.Dense(512, 512, O.SIGMOID), O.Dense(512, 1)),
}


def _tc(bf, bnn):
    import ctypes as C
    from bayesflux_b200 import _lib as L
    L.check(L.lib.bfx_model_set_compute(bnn._h, L.COMPUTE_TC_BF16X3))
    mode = C.c_int32()
    L.check(L.lib.bfx_model_get_compute(bnn._h, C.byref(mode)))
    assert mode.value == L.COMPUTE_TC_BF16X3


@pytest.mark.parametrize("net", list(TC_NETS))
@pytest.mark.parametrize("like,prior", [("normal_gamma", "gauss"), ("tdist_gamma", "mix")])
@pytest.mark.parametrize("B", [50, 700])
def test_stream_tc_grad_and_value_parity(bf, net, like, prior, B):
    rng = np.random.default_rng(SEED + 10)
    layers = TC_NETS[net]
    n = 900
    x = rng.standard_normal((layers[0].nin, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES[like], PRIORS[prior], x, y)
    assert bf.model_regime(bnn) == "stream"
    _tc(bf, bnn)
    theta = (0.1 * rng.standard_normal((m.n_total, 2))).astype(np.float32)
    idx = np.stack([rng.permutation(n)[:B] for _ in range(2)], axis=1)
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=4.0)
    for c in range(2):
        vo, go = O.grad_loglikeprior(m, theta[:, c].astype(np.float64), x[:, idx[:, c]].astype(np.float64),
                                     y[idx[:, c]].astype(np.float64), 4.0)
        rel = np.linalg.norm(g[:, c] - go) / np.linalg.norm(go)
        # 1e-4 wherever the gradient is continuous; a looser bound when a ReLU pre-activation sits within the
        # split's rounding of its kink (see relu_margin)
        tol = TOL if relu_margin(layers, theta[:, c], x[:, idx[:, c]]) > 5e-5 else 1e-2
        assert rel <= tol, f"chain {c}: {rel}"
        assert abs(v[c] - vo) <= TOL * abs(vo)


def test_stream_tc_full_data_two_chunks(bf):
    rng = np.random.default_rng(SEED + 11)
    layers = TC_NETS["cfg3_small_smooth"]
    n = 9000   # > one 8192-column chunk: weight gradients accumulate over chunks
    x = rng.standard_normal((64, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    _tc(bf, bnn)
    th = (0.1 * rng.standard_normal(m.n_total)).astype(np.float32)
    v, g = bf.grad_loglikeprior(bnn, th, None, 1.0)
    vo, go = O.grad_loglikeprior(m, th.astype(np.float64), x.astype(np.float64), y.astype(np.float64), 1.0)
    assert abs(v - vo) <= TOL * abs(vo)
    assert np.linalg.norm(g - go) <= TOL * np.linalg.norm(go)


def test_stream_tc_sgnhts_tracks_fp32(bf):
    rng = np.random.default_rng(SEED + 12)
    layers = TC_NETS["cfg3_small_smooth"]
    n = 2048
    x = rng.standard_normal((64, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    like = O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5)
    prior = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5)
    th0 = (0.05 * rng.standard_normal(sum(l.nin * l.nout + l.nout for l in layers) + 1)).astype(np.float32)
    out = []
    for tcmode in (False, True):
        bnn, m = build(bf, layers, like, prior, x, y)
        if tcmode:
            _tc(bf, bnn)
        out.append(bf.mcmc(bnn, 256, 8, bf.SGNHTS(1e-3, 1.0), theta_start=th0, seed=3))
    rel = np.linalg.norm(out[0] - out[1]) / np.linalg.norm(out[0])
    assert rel < 1e-4, rel


# ---------------------------------------------------------------------------------------------------------
# in-library collective and the update graph (bfx_sampler_comm_init, run desc shard_mode / ngpus)
# ---------------------------------------------------------------------------------------------------------
def _cfg3_small(bf, rng, n=2048, tc=True):
    layers = NETS["cfg3_small"]
    x = rng.standard_normal((64, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    like = O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5)
    prior = O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5)
    bnn, m = build(bf, layers, like, prior, x, y)
    if tc:
        _tc(bf, bnn)
    return bnn, m


@pytest.mark.parametrize("tc", [False, True])
def test_update_graph_equals_direct_launches(bf, tc, monkeypatch):
    """One update! of a single stream-regime chain is replayed from a CUDA graph whose per-step quantities (batch
    position, permutation keys, Philox counter, sample slot) are resolved on the device: bit-identical to the direct
    launches with host-side step parameters, including the short last batch of every epoch (launched directly)."""
    rng = np.random.default_rng(SEED + 20)
    bnn, m = _cfg3_small(bf, rng, n=2048 + 100, tc=tc)      # 8 full batches of 256 + a short one of 100 per epoch
    th0 = (0.05 * rng.standard_normal(m.n_total)).astype(np.float32)
    s1 = bf.SGNHTS(1e-3, 1.0)
    a = np.array(bf.mcmc(bnn, 256, 22, s1, theta_start=th0, seed=3))
    assert bf.graph_replays(s1) == 20                          # steps 9 and 18 are the short batches
    monkeypatch.setenv("BFX_NO_GRAPH", "1")
    s2 = bf.SGNHTS(1e-3, 1.0)
    b = np.array(bf.mcmc(bnn, 256, 22, s2, theta_start=th0, seed=3))
    assert bf.graph_replays(s2) == 0
    assert np.array_equal(a, b)
    assert np.array_equal(s1.kinetic, s2.kinetic)
    monkeypatch.delenv("BFX_NO_GRAPH")
    # a second run on the same handle replays again and reproduces the first (deterministic reductions)
    c = np.array(bf.mcmc(bnn, 256, 22, s1, theta_start=th0, seed=3))
    assert np.array_equal(a, c)


def test_minibatch_shard_mode_needs_a_communicator(bf):
    rng = np.random.default_rng(SEED + 21)
    bnn, m = _cfg3_small(bf, rng, n=512, tc=False)
    th0 = (0.05 * rng.standard_normal(m.n_total)).astype(np.float32)
    with pytest.raises(bf.BfxError) as e:
        bf.mcmc(bnn, 64, 4, bf.SGNHTS(1e-3, 1.0), theta_start=th0, seed=1, ngpus=2, shard_mode="minibatch")
    assert e.value.code == bf._lib.ERR_STATE


def test_in_library_allreduce_single_rank(bf):
    """The sharded path end to end on ONE rank (an NCCL communicator of size 1): gradient -> pack -> ncclAllReduce
    inside the captured update graph -> update from the exchange buffer.  Equals the unsharded run up to the float
    rounding of the three likelihood sums that travel through the buffer."""
    rng = np.random.default_rng(SEED + 22)
    bnn, m = _cfg3_small(bf, rng, n=2048, tc=True)
    th0 = (0.05 * rng.standard_normal(m.n_total)).astype(np.float32)
    ref = np.array(bf.mcmc(bnn, 256, 12, bf.SGNHTS(1e-3, 1.0), theta_start=th0, seed=3))
    s = bf.SGNHTS(1e-3, 1.0)
    s.initialise(bnn, th0.reshape(-1, 1))
    try:
        cid = bf.comm_unique_id()
    except bf.BfxError as exc:
        pytest.skip(f"NCCL not loadable here: {exc}")
    bf.comm_init(s, cid, 1, 0)
    # (the handle, and with it the communicator, is kept: same model, same number of chains)
    out = np.array(bf.mcmc(bnn, 256, 12, s, theta_start=th0, seed=3, ngpus=1, shard_mode="minibatch"))
    assert bf.graph_replays(s) == 12
    bf.comm_destroy(s)
    rel = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert rel < 1e-5, rel
