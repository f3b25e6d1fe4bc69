#!/usr/bin/env python
"""bench.py -- the headline benchmark of the BayesFlux hot path on B200.

Metric (BASELINE.json): grad-evals/s summed over chains (= posterior samples/s x chains for SGLD).
Workload at every N: BASELINE.json configs[1] -- MLP Chain(Dense(10,64,relu), Dense(64,64,relu),
Dense(64,1)), synthetic n = 1e6, FeedforwardNormal(Gamma(2,0.5)), GaussianPrior(1.0), SGLD defaults,
1024 chains PER GPU (chains are sharded over ranks with no collective: weak scaling), minibatch 64,
an independent shuffled batch schedule per chain (reference semantics).

One "step" = ONE launch of the fused gradient+SGLD kernel that advances every chain of the rank by
`--updates` update! calls (each = one minibatch gradient + one SGLD update + one recorded sample,
keep_every = 1, written to an HBM ring -- the "HBM-ring" egress mode of SURVEY.md section 8d).

  value   device-timed (CUDA events on the launching stream), inputs resident in HBM
  e2e     the same work through the public API  mcmc(bnn, B, U, SGLD(); theta_start = HOST array,
          nchains, keep_every = U)  with host theta in, host samples out (H2D + D2H in the timed region)
  --impl reference   the CPU restatement of the reference (oracle/bfx_oracle_c.c, OpenMP over chains) on
          the host cores: Julia is not in the image, so the reference itself cannot run.

Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_START = time.time()
SEED = 6150533  # the reference's test seed (test/runtests.jl:10)
D_IN, H, N_DATA = 10, 64, 1_000_000
P_TOTAL = D_IN * H + H + H * H + H + H + 1 + 1  # 4930


def algorithmic(B):
    """per chain-step (one update!) FLOPs and streaming-model bytes, keep_every = 1."""
    ws = [(D_IN, H), (H, H), (H, 1)]
    flops = 2 * B * (2 * sum(i * o for i, o in ws) + sum(i * o for i, o in ws[1:]))
    q = 4 * P_TOTAL * 3 + 4 * B * (D_IN + 1)
    return flops, q


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def synth_data(n, rng):
    """x ~ N(0,1)^{10 x n}, teacher of the same architecture, y = teacher(x) + N(0,1)."""
    x = rng.standard_normal((D_IN, n), dtype=np.float32)
    W1 = rng.standard_normal((H, D_IN), dtype=np.float32)
    W2 = rng.standard_normal((H, H), dtype=np.float32) / np.sqrt(H).astype(np.float32)
    w3 = rng.standard_normal(H, dtype=np.float32) / np.sqrt(H).astype(np.float32)
    y = np.empty(n, np.float32)
    for s in range(0, n, 1 << 17):
        a = np.maximum(W1 @ x[:, s:s + (1 << 17)], 0)
        a = np.maximum(W2 @ a, 0)
        y[s:s + (1 << 17)] = w3 @ a
    y += rng.standard_normal(n, dtype=np.float32)
    return np.asfortranarray(x), y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.1] or [l for (_, l) in self.lines]
        for ln in rows:
            f = [s.strip() for s in ln.split(",")]
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the C restatement of the reference on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_sgld(B, target_seconds, x=None, y=None, n=200_000):
    from oracle import cbind as OC
    # all host cores of the box: torchrun exports OMP_NUM_THREADS=1, which would otherwise throttle the CPU arm
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count() or 1
    nthreads = max(nthreads, OC.max_threads())
    rng = np.random.default_rng(SEED)
    if x is None:
        x, y = synth_data(n, rng)
    n = y.shape[0]
    m = OC.make_model([D_IN, H, H, 1], [1, 1, 0], sprior_a=2.0, sprior_b=0.5, prior="gaussian", sigma0=1.0)
    chains = 2 * nthreads

    def run(nsteps):
        theta = np.asfortranarray((0.1 * rng.standard_normal((P_TOTAL, chains))).astype(np.float32))
        # explicit batch indices: Fisher-Yates of all n per epoch would dominate a bounded sample,
        # while the reference amortises it over n/B = 15 625 steps
        batches = np.asfortranarray(rng.integers(0, n, (B, chains, nsteps)).astype(np.int64))
        t0 = time.perf_counter()
        OC.sgld_run(m, x, y, theta, B, nsteps, seed=SEED, batches=batches, nthreads=nthreads)
        return time.perf_counter() - t0

    t = run(20)
    per_step = t / 20
    nsteps = int(max(20, min(20000, target_seconds / per_step)))
    t = run(nsteps)
    return {"value": chains * nsteps / t, "unit": "grad-evals/s", "cores": nthreads, "kind": "port",
            "sample": f"{chains} chains x {nsteps} SGLD update! (B={B}) of the cfg2 MLP on n={n} synthetic rows, "
                      f"C port of the reference path with OpenMP over chains, {t:.1f} s",
            "seconds": t, "chains": chains, "nsteps": nsteps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    flops, q = algorithmic(B)
    per_step_target = min(12.0, 150.0 / max(1, args.steps))
    total_units, total_t, last = 0, 0.0, None
    rng = np.random.default_rng(SEED)
    x, y = synth_data(200_000, rng)
    for i in range(args.warmup + args.steps):
        r = cpu_sgld(B, 1.0 if i < args.warmup else per_step_target, x, y)
        if i >= args.warmup:
            total_units += r["chains"] * r["nsteps"]
            total_t += r["seconds"]
            last = r
    v = total_units / total_t
    out = {"impl": "reference", "metric": "grad-evals/s (posterior samples/s x chains)", "value": v,
           "unit": "grad-evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "cfg2 MLP 10-64-64-1, FeedforwardNormal+GaussianPrior(1.0), SGLD, B=64, "
                                  "CPU port of the reference path (Julia absent), bounded sample",
                      "batch": B, "n_data": 200_000},
           "cpu_baseline": {"value": v, "unit": "grad-evals/s", "cores": last["cores"], "kind": "port",
                            "sample": last["sample"]},
           "e2e": {"value": v, "unit": "grad-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(json.dumps(out))



# ----------------------------------------------------------------------------------------------
# secondary workloads measured in the same run (BASELINE.json configs[0], [2], [3], [4]); the headline stays configs[1]
# ----------------------------------------------------------------------------------------------
def log(msg):
    """progress on stderr (stdout carries the one JSON line only)"""
    sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')} +{time.time() - T_START:6.1f}s] {msg}\n")
    sys.stderr.flush()


def _ev(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(torch, dist, world, ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def secondary_cfg3(bf, torch, dist, world, rank, local, steps=30):
    """configs[2]: ONE chain, wide MLP 64-512-512-512-1 (ReLU), FeedforwardTDist(nu=5) + MixtureScalePrior, SGNHTS,
    the minibatch sharded over the ranks (8 192 observations per GPU per update, 2^20 per GPU in the shard), the
    gradient all-reduced over NCCL INSIDE the library's update graph (bfx_mcmc_advance, shard_mode MINIBATCH)."""
    import ctypes as C
    from bayesflux_b200 import _lib as L
    from bayesflux_b200 import api as A
    W, Bl, n_loc = 512, 8192, 1 << 20
    g = torch.Generator(device="cuda").manual_seed(SEED + rank)
    x = torch.randn((n_loc, 64), generator=g, device="cuda", dtype=torch.float32)   # = 64 x n_loc column-major
    y = torch.randn((n_loc,), generator=g, device="cuda", dtype=torch.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(64, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, 1)))
    bnn = bf.BNN(None, None, bf.FeedforwardTDist(nc, bf.Gamma(2.0, 0.5), 5.0), bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5),
                 device=local, compute="tc")
    bnn.set_data_device(x.data_ptr(), (64, n_loc), y.data_ptr())
    P = bnn.num_total_params
    th0 = (0.03 * np.random.default_rng(SEED).standard_normal((P, 1))).astype(np.float32)   # the same on every rank
    s = bf.SGNHTS(1e-3, 1.0)
    s.initialise(bnn, th0)
    shard = "minibatch" if world > 1 else None
    log("cfg3: model and sampler ready")
    if world > 1:
        idt = torch.zeros(L.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(bf.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        bf.comm_init(s, bytes(idt.cpu().numpy().tobytes()), world, rank)
        log("cfg3: communicator up")
    st = torch.cuda.Stream()
    kw = dict(seed=SEED, chain_offset=rank, stream=st.cuda_stream, ngpus=world, shard_mode=shard)
    bf.advance(s, Bl, 5, **kw)
    torch.cuda.synchronize()
    log("cfg3: warm-up updates done")
    if world > 1:
        dist.barrier()
    e0, e1 = _ev(torch)
    e0.record(st)
    bf.advance(s, Bl, steps, **kw)
    e1.record(st)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1)) / steps
    log(f"cfg3: {ms * 1e3:.1f} us per update")
    replays = bf.graph_replays(s)
    # the same update with per-layer all-reduce buckets on a side stream inside the backward pass (BFX_BUCKETS=1)
    ms_single = None
    if world > 1:
        os.environ["BFX_BUCKETS"] = "1"
        bf.advance(s, Bl, 3, **kw)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record(st)
        bf.advance(s, Bl, steps, **kw)
        e1.record(st)
        torch.cuda.synchronize()
        ms_single = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1)) / steps
        del os.environ["BFX_BUCKETS"]
        log(f"cfg3: {ms_single * 1e3:.1f} us per update with per-layer buckets")
    # ... and with the NCCL all-reduce in place of the peer-memory exchange (BFX_NO_P2P=1), when the latter is in use
    peers = bf.peer_ranks(s) if world > 1 else 0
    ms_nccl = None
    if world > 1 and peers > 1:
        os.environ["BFX_NO_P2P"] = "1"
        bf.advance(s, Bl, 3, **kw)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record(st)
        bf.advance(s, Bl, steps, **kw)
        e1.record(st)
        torch.cuda.synchronize()
        ms_nccl = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1)) / steps
        del os.environ["BFX_NO_P2P"]
        log(f"cfg3: {ms_nccl * 1e3:.1f} us per update with the NCCL all-reduce")
    # phase split through the three-call form of the same update (direct launches, one event between the phases)
    run = A._run_desc(Bl, 0, True, True, True, 1, SEED, "per_chain", rank, world, shard)
    buf = torch.empty(P + 3, dtype=torch.float32, device="cuda")
    nb = float((n_loc + Bl - 1) // Bl)
    tg = tc = tu = 0.0
    ksplit = 10
    with torch.cuda.stream(st):
        for it in range(ksplit + 2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(st)
            L.check(L.lib.bfx_stream_grad_local(s._h, C.byref(run), nb, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr()), P + 3))
            ev[1].record(st)
            L.check(L.lib.bfx_stream_allreduce(s._h, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr())))
            ev[2].record(st)
            L.check(L.lib.bfx_stream_apply_update(s._h, C.byref(run), nb, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr()), None))
            ev[3].record(st)
            torch.cuda.synchronize()
            if it >= 2:
                tg += ev[0].elapsed_time(ev[1]); tc += ev[1].elapsed_time(ev[2]); tu += ev[2].elapsed_time(ev[3])
    theta = s.theta[:, 0]
    ok = bool(np.isfinite(theta).all()) and not s.get_state(L.STATE_STATUS, np.int32, False).any()
    # replicas must be bit-identical: compare a checksum of theta's bit pattern across ranks
    chk = torch.tensor([float(np.frombuffer(theta.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 52))],
                       dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool((lo == hi).item())
    if world > 1:
        bf.comm_destroy(s)
    flop = 3279872.0 * Bl                      # per rank per update (SURVEY.md section 8 sizes table)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1355.0) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    ar_bytes = 4 * (P + 3)
    ar_ms = tc / ksplit
    out = {"workload": f"configs[2]: wide MLP 64-512-512-512-1 relu, FeedforwardTDist(5) + MixtureScalePrior(1,0.1,0.5), "
                       f"SGNHTS(l=1e-3), ONE chain, minibatch {Bl}/GPU x {world} GPUs (shard n={n_loc}/GPU), tcgen05 3-term BF16 split",
           "n_gpus": world, "updates": steps, "ms_per_update": ms, "updates_per_s": 1e3 / ms,
           "observations_per_s": Bl * world * 1e3 / ms,
           "tflops_fp32_equiv_per_gpu": flop / (ms * 1e-3) / 1e12,
           "frac_of_bf16x3_tensor_roofline": 3 * flop / (ms * 1e-3) / 1e12 / peak,
           "tensor_peak_tflops": peak,
           "one_cuda_graph_per_update": replays >= steps, "graph_replays": int(replays),
           "split_ms_direct_launches": {"grad_local": tg / ksplit, "allreduce": ar_ms, "update": tu / ksplit,
                                        "sum": (tg + tc + tu) / ksplit},
           "allreduce_bytes": ar_bytes,
           "allreduce_busbw_gbs": (2.0 * (world - 1) / world * ar_bytes / (ar_ms * 1e-3) / 1e9) if world > 1 and ar_ms > 0 else None,
           "collective": ("peer-memory exchange: every rank's finish kernel reads the P+3 floats of all ranks over NVLink "
                          "(CUDA IPC mappings, flags in peer memory) and sums them in rank order; no collective call"
                          if peers > 1 else
                          "one ncclAllReduce of P+3 floats issued by libbfx on the engine's stream inside the update graph")
                         if world > 1 else "none (single rank)",
           "peer_ranks": peers, "ms_per_update_nccl_allreduce": ms_nccl,
           "ms_per_update_bucketed": ms_single,
           "bucketed_minus_single_ms": (ms_single - ms) if ms_single is not None else None,
           "replicas_bit_identical": identical, "finite": ok}
    bnn.close() if hasattr(bnn, "close") else None
    return out


def secondary_cfg1(bf, torch, dist, world, rank, local, updates=200):
    """configs[0]: README linear regression Dense(5,1), n=500, FeedforwardNormal + GaussianPrior(0.5), SGLD B=50;
    65 536 chains per GPU (chains sharded, no collective)."""
    rng = np.random.default_rng(SEED)
    n, C = 500, 1 << 16
    x = rng.standard_normal((5, n)).astype(np.float32)
    beta = rng.standard_normal(5).astype(np.float32)
    y = (x.T @ beta + rng.standard_normal(n)).astype(np.float32)
    ncn = bf.destruct(bf.Chain(bf.Dense(5, 1)))
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(ncn, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(ncn, 0.5), device=local)
    th0 = (0.5 * np.random.default_rng(SEED + rank).standard_normal((7, C))).astype(np.float32)
    s = bf.SGLD(stepsize_a=0.1, stepsize_b=0.0, stepsize_gamma=0.55)
    s.initialise(bnn, th0)
    ring = torch.empty((C, 4, 7), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    kw = dict(seed=SEED, chain_offset=rank * C, stream=st.cuda_stream, samples_dev=ring.data_ptr(), ring_slots=4)
    for _ in range(2):
        bf.advance(s, 50, updates, **kw)
    torch.cuda.synchronize()
    e0, e1 = _ev(torch)
    e0.record(st)
    bf.advance(s, 50, updates, **kw)
    e1.record(st)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1))
    v = world * C * updates / ms * 1e3
    q = 4 * 7 * 3 + 4 * 50 * 6     # streaming-model bytes per chain-step (theta read + write + sample, one minibatch)
    hbm, _, _ = peaks()
    return {"workload": "configs[0]: README linreg Dense(5,1), n=500, B=50, SGLD(a=0.1,b=0,gamma=0.55), 65536 chains per GPU",
            "n_gpus": world, "grad_evals_per_s": v, "ms": ms, "chains_per_gpu": C, "updates": updates,
            "roofline": {"bound": "hbm", "bytes_per_chain_step": q, "achieved_gbs": v / world * q / 1e9, "peak_gbs": hbm,
                         "frac": v / world * q / 1e9 / hbm}}


def secondary_cfg4(bf, torch, dist, world, rank, local, updates=16, chains=296, ggmc_steps=8):
    """configs[3]: LSTM(4,64)+Dense(64,1) seq-to-one, T=50, n=16 384, B=64, SeqToOneNormal, GGMC with the reference
    test's adapters; 296 chains per GPU (the recurrent kernel runs one 207 KB CTA per SM: two full waves of 148)."""
    rng = np.random.default_rng(SEED)
    T, n, B, C = 50, 16384, 64, chains
    series = (np.cumsum(rng.standard_normal((T + n, 4)), axis=0) * 0.05).astype(np.float32)
    win = np.lib.stride_tricks.sliding_window_view(series[:T + n - 1], (T, 4))[:n, 0]    # n x T x 4
    x = np.asfortranarray(np.transpose(win, (1, 2, 0)))                                   # T x d x n
    y = series[T:T + n, 0].copy()
    ncn = bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1)))
    bnn = bf.BNN(x, y, bf.SeqToOneNormal(ncn, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(ncn, 0.5), device=local)
    P = bnn.num_total_params
    th0 = (0.05 * np.random.default_rng(SEED + rank).standard_normal((P, C))).astype(np.float32)
    s = bf.GGMC(beta=0.1, l=1e-2, steps=ggmc_steps,
                sadapter=bf.DualAveragingStepSize(1e-2, target_accept=0.25, adapt_steps=1000),
                madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1, epsilon=1e-8))
    s.initialise(bnn, th0)
    st = torch.cuda.Stream()
    kw = dict(seed=SEED, chain_offset=rank * C, stream=st.cuda_stream)
    bf.advance(s, B, 2 * ggmc_steps, **kw)     # two windows: the second one runs the MH test (ggmc.jl:174: t > steps)
    torch.cuda.synchronize()
    e0, e1 = _ev(torch)
    e0.record(st)
    bf.advance(s, B, updates, **kw)            # steady state: whole windows, each with its MH test
    e1.record(st)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1))
    ok = not s.get_state(bf._lib.STATE_STATUS, np.int32, False).any()
    fl = 2 * B * 5120384.0 * C * updates          # two minibatch gradients per GGMC update
    sm_max = peaks()[1]
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    return {"workload": f"configs[3]: LSTM(4,64)+Dense(64,1), T=50, n=16384, B=64, SeqToOneNormal, GGMC(beta=0.1,l=1e-2,steps={ggmc_steps}) "
                        f"+ DualAveraging + DiagCov, {C} chains per GPU",
            "n_gpus": world, "ms": ms, "updates": updates, "grad_evals_per_s": world * 2 * C * updates / ms * 1e3,
            "samples_per_s": world * C * updates / ms * 1e3,
            "minibatch_tflops_fp32_per_gpu": fl / (ms * 1e-3) / 1e12,
            "roofline": {"bound": "fp32 pipe (the recurrent kernel is CUDA-core FMA)", "peak_tflops": fp32_peak,
                         "frac_minibatch_only": fl / (ms * 1e-3) / 1e12 / fp32_peak,
                         "note": "the full-data MH forward of each GGMC window (16384 sequences per chain: the start value of "
                                 "a window is the one its predecessor's MH test left behind) is extra work not counted in "
                                 "the numerator"},
            "finite": bool(ok)}


def secondary_cfg5(bf, torch, dist, world, rank, local, nsamp=2, chains=4096):
    """configs[4]: many-chain HMC (path_len 5, DualAveraging, DiagCov) on the cfg2 MLP, full batch n = 1e5,
    4 096 chains per GPU (chains sharded, no collective)."""
    rng = np.random.default_rng(SEED)
    n, C, path = 100_000, chains, 5
    x = rng.standard_normal((10, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    ncn = bf.destruct(bf.Chain(bf.Dense(10, 64, bf.relu), bf.Dense(64, 64, bf.relu), bf.Dense(64, 1)))
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(ncn, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(ncn, 1.0), device=local, compute="tc")
    P = bnn.num_total_params
    th0 = (0.1 * np.random.default_rng(SEED + rank).standard_normal((P, C))).astype(np.float32)
    s = bf.HMC(1e-4, path, sadapter=bf.DualAveragingStepSize(1e-4, target_accept=0.5),
               madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1))
    s.initialise(bnn, th0)
    st = torch.cuda.Stream()
    bf.advance(s, n, path, seed=SEED, chain_offset=rank * C, stream=st.cuda_stream)   # first trajectory (+ BF16 copy of x)
    torch.cuda.synchronize()
    e0, e1 = _ev(torch)
    e0.record(st)
    bf.advance(s, n, nsamp * path, seed=SEED, chain_offset=rank * C, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1))
    # per trajectory: path_len full-batch gradients and ONE full-data value pass (hmc.jl:135-136 evaluates two; the start
    # value is the one the previous MH test left behind)
    flop = (path * 27520.0 + 9600.0) * n * C * nsamp
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1355.0) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    tf = flop / (ms * 1e-3) / 1e12
    return {"workload": f"configs[4]: HMC path_len 5 + DualAveraging + DiagCov, MLP 10-64-64-1, full batch n={n}, {C} chains per GPU, "
                        f"tcgen05 3-term BF16 split (trajectories 2-3 of a run)",
            "n_gpus": world, "ms": ms, "hmc_samples": nsamp, "full_batch_grad_evals_per_s": world * C * path * nsamp / ms * 1e3,
            "tflops_fp32_equiv_per_gpu": tf,
            "roofline": {"bound": "tensor", "peak_tflops": peak, "issued_bf16_tflops": 3 * tf, "frac": 3 * tf / peak}}

# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import bayesflux_b200 as bf

    C, B, U = args.chains, args.batch, args.updates
    rng = np.random.default_rng(SEED)
    x, y = synth_data(args.n, rng)
    net = bf.Chain(bf.Dense(D_IN, H, bf.relu), bf.Dense(H, H, bf.relu), bf.Dense(H, 1))
    nc = bf.destruct(net)
    like = bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5))
    prior = bf.GaussianPrior(nc, 1.0)
    bnn = bf.BNN(x, y, like, prior, device=local, compute=args.compute)
    P = bnn.num_total_params
    assert P == P_TOTAL
    rng_c = np.random.default_rng(SEED + 1 + rank)
    theta0 = np.asfortranarray((0.1 * rng_c.standard_normal((P, C))).astype(np.float32))
    chain_offset = rank * C

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (`value`) ----------------
    log("device-resident arm")
    stream = torch.cuda.Stream()
    ring = torch.empty((C, U, P), dtype=torch.float32, device="cuda")  # HBM ring of recorded samples
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    s = bf.SGLD()
    s.initialise(bnn, theta0)
    launches0 = bf.launch_count(s)

    def one_step(timed):
        with torch.cuda.stream(stream):
            flush.fill_(1)  # L2 flush between iterations, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            bf.advance(s, B, U, keep_every=1, seed=SEED, chain_offset=chain_offset, stream=stream.cuda_stream,
                       samples_dev=ring.data_ptr(), ring_slots=U)
            e1.record(stream)
        return (e0, e1) if timed else None

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    launches_w = bf.launch_count(s)
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    t0 = time.time()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    t1 = time.time()
    clk = clocks.stop(t0, t1)
    per_launch = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(per_launch)
    gpu_launches = bf.launch_count(s) - launches_w
    st = s.get_state(bf._lib.STATE_STATUS, np.int32, False)
    assert not st.any(), "a chain hit a NaN guard during the benchmark"
    last = ring[:, -1, :].float().cpu().numpy()
    assert np.isfinite(last).all()
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    units = world * C * U * args.steps  # chain-steps = grad-evals
    value = units / (dev_ms * 1e-3)

    # ---------------- end-to-end arm through the public API (`e2e`) ----------------
    log("end-to-end arm")
    e2e_steps = max(2, min(args.steps, 5))
    theta_h = theta0.copy(order="F")
    s2 = bf.SGLD()

    e2e_calls = [0]
    # caller-owned result buffers, alternated: the array of call i stays valid while call i+1 runs
    e2e_out = [np.empty((P, 1, C), np.float32, order="F") for _ in range(2)]

    def e2e_step():
        nonlocal theta_h
        e2e_calls[0] += 1
        ch = bf.mcmc(bnn, B, U, s2, theta_start=theta_h, keep_every=U, seed=SEED + e2e_calls[0],
                     chain_offset=chain_offset, out=e2e_out[e2e_calls[0] & 1])
        theta_h = np.asfortranarray(ch[:, -1, :]) if ch.ndim == 3 else np.asfortranarray(ch[:, -1:])
        return ch

    for _ in range(2):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        ch = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - w0
    assert np.isfinite(ch).all()
    tw = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_value = world * C * U * e2e_steps / float(tw.item())

    # ---------------- the headline line is complete here; everything below only adds keys to it ----------------
    out = None
    if rank == 0:
        flops, q = algorithmic(B)
        hbm, sm_max, src = peaks()
        per_gpu = value / world
        traffic = None
        tc = bnn.compute == "tc"
        tp = os.path.join(ROOT, "profiles", "traffic_k_mcmc_tc.json" if tc else "traffic_k_mcmc.json")
        if os.path.exists(tp):  # dram__bytes_read+write of one ncu --set full capture, rescaled to this launch size
            traffic = json.load(open(tp))["dram_bytes_per_chain_step"] * C * U
        kernel = "k_mcmc<256,64,SGLD,TC> (tcgen05, 3-term BF16 split)" if tc else "k_mcmc<256,64,SGLD> (FP32 FMA)"
        achieved = per_gpu * q / 1e9
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        out = {
            "metric": "grad-evals/s (posterior samples/s x chains)", "value": value, "unit": "grad-evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3->f32" if tc else "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: MLP Chain(Dense(10,64,relu),Dense(64,64,relu),Dense(64,1)), "
                                   "n=1e6 synthetic, FeedforwardNormal(Gamma(2,0.5)) + GaussianPrior(1.0), SGLD "
                                   "defaults, 1024 chains per GPU, per-chain shuffled minibatches",
                       "chains_per_gpu": C, "batch": B, "n_data": args.n, "updates_per_step": U,
                       "keep_every": 1, "sample_egress": "HBM ring (device-resident)", "compute": bnn.compute,
                       "l2": "flushed between timed steps (256 MiB fill outside the event pair)",
                       "parallelism": f"chains sharded x{world}, no collective"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "grad-evals/s", "h2d_bytes_per_step": int(theta0.nbytes),
                    "d2h_bytes_per_step": int(ch.nbytes), "steps": e2e_steps,
                    "api": "mcmc(bnn, 64, U, SGLD(); theta_start=host, nchains, keep_every=U) -> host samples"},
            "gpu_launches": int(gpu_launches),
            "ms_per_launch": {"min": float(np.min(per_launch)), "median": float(np.median(per_launch)),
                              "max": float(np.max(per_launch)), "n": len(per_launch)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "peak_source": src, "kernel": kernel,
                         "bytes_per_chain_step": q, "chain_steps_per_launch": C * U,
                         "note": "theta is resident in shared memory, so the kernel moves fewer bytes than the "
                                 "streaming model's Q; frac = algorithmic bytes / time / measured HBM peak (DESIGN.md)",
                         "fp32_equiv": {"achieved_tflops": per_gpu * flops / 1e12, "fp32_pipe_peak_tflops": fp32_peak,
                                        "frac_of_fp32_pipe": per_gpu * flops / 1e12 / fp32_peak,
                                        "tensor_flops_issued_tflops": (3 if tc else 0) * per_gpu * flops / 1e12}},
        }
        if not args.no_cpu and world == 1:
            log("cpu baseline")
            cb = cpu_sgld(B, args.cpu_seconds)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    barrier()

    # A hang in an optional leg (a collective that never completes) must not cost the headline: past the deadline a
    # watchdog thread emits the line as it stands (rank 0) and ends the process on every rank.
    emitted = threading.Event()

    def finish(note=None):
        if emitted.is_set():
            return
        emitted.set()
        if rank == 0:
            if note:
                out.setdefault("secondary", {})["aborted"] = note
            emit(json.dumps(out))

    def watchdog():
        if not emitted.wait(args.secondary_budget):
            log(f"secondary legs exceeded {args.secondary_budget:.0f} s: emitting the headline and leaving")
            finish(f"optional legs stopped after {args.secondary_budget:.0f} s")
            os._exit(0)

    if not args.no_secondary:
        threading.Thread(target=watchdog, daemon=True).start()
        # reference semantics of the egress: EVERY sample crosses to the host (sgld.jl:106 keeps all of them)
        log("e2e with every sample copied to the host")
        s3 = bf.SGLD()
        out_all = np.empty((P, U, C), np.float32, order="F")
        bf.mcmc(bnn, B, U, s3, theta_start=theta_h, keep_every=1, seed=SEED, chain_offset=chain_offset, out=out_all)
        barrier()
        w0 = time.perf_counter()
        for i in range(2):
            bf.mcmc(bnn, B, U, s3, theta_start=theta_h, keep_every=1, seed=SEED + 100 + i, chain_offset=chain_offset, out=out_all)
        barrier()
        tall = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tall, op=dist.ReduceOp.MAX)
        if rank == 0:
            out["e2e"]["every_sample_to_host"] = {
                "value": world * C * U * 2 / float(tall.item()), "unit": "grad-evals/s",
                "d2h_bytes_per_step": int(out_all.nbytes), "h2d_bytes_per_step": int(theta0.nbytes),
                "note": "keep_every = 1: all 128 samples of every chain are copied to the host each call (PCIe-bound)"}
        del out_all, s3
        # secondary workloads (every rank takes part: chains sharded, or the minibatch-sharded chain of configs[2])
        del ring
        torch.cuda.empty_cache()
        for name, fn in (("cfg3", secondary_cfg3), ("cfg1", secondary_cfg1), ("cfg5", secondary_cfg5), ("cfg4", secondary_cfg4)):
            log(f"secondary {name}")
            try:
                res = fn(bf, torch, dist, world, rank, local)
            except Exception as exc:  # a secondary leg never takes the headline down
                res = {"error": f"{type(exc).__name__}: {exc}"}
                log(f"secondary {name} failed: {res['error']}")
            if rank == 0:
                out.setdefault("secondary", {})[name] = res
            barrier()
    finish()
    log("done")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class StdoutGuard:
    """Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are sent to stderr
    for the duration of the run, the result is written to the real stdout at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


GUARD = None


def emit(line):
    if GUARD is not None:
        GUARD.emit(line)
    else:
        print(line, flush=True)


def main():
    global GUARD
    GUARD = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--updates", type=int, default=128, help="update! calls per chain per step (= per launch)")
    ap.add_argument("--n", type=int, default=N_DATA)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (configs[0], [2], [3], [4])")
    ap.add_argument("--secondary-budget", type=float, default=150.0,
                    help="seconds the optional legs may take before the headline line is emitted without them")
    ap.add_argument("--compute", default="tc", choices=["fp32", "tc", "auto"],
                    help="contraction arithmetic: tc = tcgen05 tensor cores (3-term BF16 split), fp32 = CUDA-core gate path")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
