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
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
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
            cb = cpu_sgld(B, args.cpu_seconds)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(json.dumps(out))
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
