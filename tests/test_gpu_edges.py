"""Edge cases of the device loop: ragged / tiny data sets, batch larger than the data, partial = false, a single
sample, more chains than resident CTAs, chunked launches -- on the FP32 and the tensor-core engines."""
import numpy as np
import pytest

from oracle import bfx_oracle as O
from tests.test_gpu_parity import LIKES, PRIORS, build, engine_draws

pytestmark = pytest.mark.gpu
SEED = 6150533
CFG2 = (O.Dense(10, 64, O.RELU), O.Dense(64, 64, O.RELU), O.Dense(64, 1))


@pytest.fixture(scope="module")
def bf():
    import bayesflux_b200 as bf
    return bf


def _bnn(bf, n, compute, layers=CFG2, seed=0):
    rng = np.random.default_rng(SEED + seed)
    x = rng.standard_normal((layers[0].nin, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    if compute == "tc":
        import ctypes as C
        from bayesflux_b200 import _lib as L
        L.check(L.lib.bfx_model_set_compute(bnn._h, L.COMPUTE_TC_BF16X3))
    return bnn, m, x, y


@pytest.mark.parametrize("compute", ["fp32", "tc"])
@pytest.mark.parametrize("n,B,partial", [(1, 64, True), (7, 64, True), (65, 64, True), (65, 64, False), (130, 50, False), (200, 200, True)])
def test_ragged_schedules_match_the_oracle_replay(bf, compute, n, B, partial):
    """Every update uses exactly the minibatch the DataLoader semantics prescribe (abstract.jl:102-105): replayed
    by the oracle through the debug hooks, incl. the short last batch and num_batches = ceil / floor."""
    from bayesflux_b200 import api as A
    bnn, m, x, y = _bnn(bf, n, compute)
    th0 = (0.1 * np.random.default_rng(1).standard_normal(m.n_total)).astype(np.float32)
    nsamp = 5
    s = bf.SGLD(stepsize_a=1e-3)
    ch = bf.mcmc(bnn, B, nsamp, s, theta_start=th0, seed=4, partial=partial)
    assert ch.shape == (m.n_total, nsamp) and np.isfinite(ch).all()
    Bc = min(B, n)
    nb = -(-n // Bc) if partial else max(1, n // Bc)
    run = A._run_desc(B, nsamp, True, partial, False, 1, 4, "per_chain", 0)
    so = O.SGLD(stepsize_a=1e-3)
    so.initialise(th0.copy(), nsamp)
    th = th0.copy()
    seen = []
    for t in range(nsamp):
        idx, z, _ = engine_draws(bf, s, run, 0, t, m.n_total)
        seen.append(len(idx))
        gf = lambda q, idx=idx: O.grad_loglikeprior(m, q, x[:, idx], y[idx], float(nb))
        th = so.update(th, gf, None, [z], [0.5])
        tol = 2e-4 if compute == "fp32" else 2e-3   # tensor-core path: ReLU kink flips possible on tiny batches
        assert np.max(np.abs(ch[:, t] - th)) <= tol * max(1.0, float(np.max(np.abs(th))))
        th = ch[:, t].copy()   # re-synchronise: this test is about the schedule, not about error growth
    if partial and n % Bc:
        assert (n % Bc) in seen or nsamp < nb


@pytest.mark.parametrize("compute", ["fp32", "tc"])
def test_more_chains_than_resident_ctas_and_odd_chunks(bf, compute):
    """700 chains (> 2 x 148 CTAs), 37 updates (not a multiple of the chunk length): every chain advances exactly
    37 steps and the result does not depend on how the launch is cut."""
    bnn, m, x, y = _bnn(bf, 512, compute, seed=2)
    C = 700
    th0 = (0.1 * np.random.default_rng(3).standard_normal((m.n_total, C))).astype(np.float32)
    a = bf.mcmc(bnn, 64, 37, bf.SGLD(), theta_start=th0, seed=9, keep_every=37)
    s = bf.SGLD()
    b1 = bf.mcmc(bnn, 64, 20, s, theta_start=th0, seed=9, keep_every=20)
    b2 = bf.mcmc(bnn, 64, 37, s, continue_sampling=True, seed=9, keep_every=37)
    assert a.shape == (m.n_total, 1, C)
    assert int(s.get_state(bf._lib.STATE_NSAMPLED, np.int64, False).min()) == 37
    np.testing.assert_array_equal(a[:, 0, :], b2[:, -1, :])


def test_keep_every_and_single_sample(bf):
    bnn, m, x, y = _bnn(bf, 256, "tc", seed=5)
    th0 = (0.1 * np.random.default_rng(6).standard_normal((m.n_total, 2))).astype(np.float32)
    one = bf.mcmc(bnn, 64, 1, bf.SGLD(), theta_start=th0, seed=2)
    assert one.shape == (m.n_total, 1, 2)
    full = bf.mcmc(bnn, 64, 12, bf.SGLD(), theta_start=th0, seed=2)
    thin = bf.mcmc(bnn, 64, 12, bf.SGLD(), theta_start=th0, seed=2, keep_every=4)
    np.testing.assert_array_equal(one[:, 0], full[:, 0])
    np.testing.assert_array_equal(thin, full[:, 3::4])


@pytest.mark.parametrize("compute,kind", [("fp32", "sgld"), ("tc", "sgld"), ("fp32", "sgnht")])
def test_chain_group_pipelining_is_bit_identical(bf, compute, kind, monkeypatch):
    """bfx_mcmc_run pipelines fresh runs with many chains in chain groups (upload / kernel / download overlap).
    Chains are independent and keyed by their global index, so the result must equal the single-launch run bit
    for bit, including a ragged last group."""
    bnn, m, x, y = _bnn(bf, 640, compute)
    C = 515
    th0 = np.asfortranarray((0.1 * np.random.default_rng(7).standard_normal((m.n_total, C))).astype(np.float32))
    out = {}
    for groups in ("1", "4", "3"):
        monkeypatch.setenv("BFX_PIPELINE_GROUPS", groups)
        s = bf.SGLD() if kind == "sgld" else bf.SGNHT(1e-3)
        ch = bf.mcmc(bnn, 64, 12, s, theta_start=th0, keep_every=4, seed=5)
        out[groups] = (np.array(ch, copy=True), None if s.kinetic is None else np.array(s.kinetic, copy=True))
    ref, kin = out["1"]
    assert ref.shape == (m.n_total, 3, C) and np.isfinite(ref).all()
    for groups in ("4", "3"):
        assert np.array_equal(out[groups][0], ref)
        if kin is not None:
            assert np.array_equal(out[groups][1], kin)


@pytest.mark.parametrize("kind", ["sgnhts_rms", "ggmc", "hmc"])
def test_sharding_invariance_of_the_mh_and_thermostat_samplers(bf, kind):
    """What two ranks run (640 chains as 320 + 320 with chain_offset) equals the one-sampler run bit for bit, also
    for the samplers with per-chain uniforms (MH accept), adapters and thermostats; the one-sampler run goes through
    the chain-group pipeline of bfx_mcmc_run."""
    bnn, m, x, y = _bnn(bf, 400, "tc")
    C = 640
    th0 = np.asfortranarray((0.1 * np.random.default_rng(17).standard_normal((m.n_total, C))).astype(np.float32))

    def mk():
        if kind == "sgnhts_rms":
            return bf.SGNHTS(1e-3, madapter=bf.RMSPropMassAdapter(5))
        if kind == "ggmc":
            return bf.GGMC(l=1e-3, steps=2, sadapter=bf.DualAveragingStepSize(1e-3, adapt_steps=4),
                           madapter=bf.DiagCovMassAdapter(6, 3))
        return bf.HMC(1e-3, 3, sadapter=bf.DualAveragingStepSize(1e-3, adapt_steps=3), madapter=bf.DiagCovMassAdapter(6, 3))

    s = mk()
    whole = np.array(bf.mcmc(bnn, 64, 8, s, theta_start=th0, seed=21), copy=True)
    acc = None if s.accepted is None else np.array(s.accepted, copy=True)
    parts, accs = [], []
    for o in (0, 320):
        sp = mk()
        parts.append(np.array(bf.mcmc(bnn, 64, 8, sp, theta_start=np.asfortranarray(th0[:, o:o + 320]), seed=21,
                                      chain_offset=o), copy=True))
        accs.append(sp.accepted)
    assert np.isfinite(whole).all()
    assert np.array_equal(np.concatenate(parts, axis=2), whole)
    if acc is not None:
        assert np.array_equal(np.concatenate(accs, axis=1), acc)


@pytest.mark.parametrize("kind", ["sgld", "sgnht", "hmc"])
def test_continue_after_device_resident_advance(bf, kind):
    """mcmc -> bfx_mcmc_advance (asynchronous, device only) -> mcmc(continue_sampling): the continued run must take
    its sample / history origins from the DEVICE counters (the host mirror knows nothing of the advance).  With
    stale origins the slots of the continued run would land past the per-chain stride of the sample buffer."""
    rng = np.random.default_rng(SEED + 77)
    bnn, m, x, y = _bnn(bf, 300, "fp32", seed=3)
    C = 3
    th0 = (0.1 * rng.standard_normal((m.n_total, C))).astype(np.float32)

    def make():
        if kind == "sgld":
            return bf.SGLD()
        if kind == "sgnht":
            return bf.SGNHT(1e-3)
        return bf.HMC(1e-3, 3, sadapter=bf.ConstantStepsize(1e-3), madapter=bf.FixedMassAdapter())

    ups = 3 if kind == "hmc" else 1  # update! calls per sample
    s = make()
    first = np.array(bf.mcmc(bnn, 50, 10, s, theta_start=th0, seed=9))
    bf.advance(s, 50, 7 * ups, seed=9)  # 7 more samples, nothing recorded, host mirror untouched
    out = np.array(bf.mcmc(bnn, 50, 25, s, continue_sampling=True, seed=9))
    assert s.nsampled == 25
    assert out.shape == (m.n_total, 10 + 8, C)  # 10 old + (25 - 17) new columns
    assert np.array_equal(out[:, :10], first)
    assert np.all(np.isfinite(out))
    if kind == "sgld":
        # counter-based noise and schedule: samples 18..25 equal those of an uninterrupted run
        ref = np.array(bf.mcmc(bnn, 50, 25, make(), theta_start=th0, seed=9))
        assert np.array_equal(out[:, 10:], ref[:, 17:])
    if s.kinetic is not None:
        assert s.kinetic.shape[0] == 18 and np.all(np.isfinite(s.kinetic))
