#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs[0], [2], [3], [4] (bench.py is the headline, configs[1]).

    python tools/bench_cfg.py --cfg 1|3|4|5 [--steps K]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_cfg.py --cfg 3      (minibatch-sharded chain)

Prints one JSON line per run (rank 0).  Synthetic data of each config's shape."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 6150533


def events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def cfg1(bf, torch, args):
    rng = np.random.default_rng(SEED)
    n, C, U = 500, 1 << 16, args.steps or 200
    x = rng.standard_normal((5, n)).astype(np.float32)
    beta = rng.standard_normal(5).astype(np.float32)
    y = (x.T @ beta + rng.standard_normal(n)).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(5, 1)))
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 0.5))
    th0 = (0.5 * rng.standard_normal((7, C))).astype(np.float32)
    s = bf.SGLD(stepsize_a=0.1, stepsize_b=0.0, stepsize_gamma=0.55)
    s.initialise(bnn, th0)
    ring = torch.empty((C, 4, 7), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    for _ in range(2):
        bf.advance(s, 50, U, seed=SEED, stream=st.cuda_stream, samples_dev=ring.data_ptr(), ring_slots=4)
    torch.cuda.synchronize()
    e0, e1 = events(torch)
    e0.record(st)
    bf.advance(s, 50, U, seed=SEED, stream=st.cuda_stream, samples_dev=ring.data_ptr(), ring_slots=4)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"cfg": 1, "workload": "README linreg Dense(5,1), n=500, B=50, SGLD, 65536 chains", "grad_evals_per_s": C * U / ms * 1e3,
            "ms": ms, "chains": C, "updates": U}


def cfg3(bf, torch, args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, Bl, n_loc = 512, 8192, args.n or (1 << 20)
    g = torch.Generator(device="cuda").manual_seed(SEED + rank)
    x = torch.randn((n_loc, 64), generator=g, device="cuda", dtype=torch.float32)   # = 64 x n_loc column-major
    y = torch.randn((n_loc,), generator=g, device="cuda", dtype=torch.float32)
    net = bf.Chain(bf.Dense(64, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, 1))
    nc = bf.destruct(net)
    bnn = bf.BNN(None, None, bf.FeedforwardTDist(nc, bf.Gamma(2.0, 0.5), 5.0), bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5), device=local,
                 compute=args.compute)
    bnn.set_data_device(x.data_ptr(), (64, n_loc), y.data_ptr())
    P = bnn.num_total_params
    th0 = (0.03 * np.random.default_rng(SEED).standard_normal(P)).astype(np.float32)
    nb = bf.global_num_batches(n_loc * world, Bl * world)
    ch = bf.ShardedChain(bnn, bf.SGNHTS(1e-3, 1.0), th0, Bl, nb, seed=SEED, rank=rank)
    K = args.steps or 50
    for _ in range(5):
        ch.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tg = tc = tu = 0.0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    for _ in range(K):
        e[0].record(ch.stream); ch.grad_local(); e[1].record(ch.stream); ch._reduce(); e[2].record(ch.stream)
        ch.apply_update(); e[3].record(ch.stream)
        torch.cuda.synchronize()
        tg += e[0].elapsed_time(e[1]); tc += e[1].elapsed_time(e[2]); tu += e[2].elapsed_time(e[3])
    wall = time.perf_counter() - t0
    # un-instrumented loop for the throughput number
    if world > 1:
        dist.barrier()
    a, b = events(torch)
    a.record(ch.stream)
    for _ in range(K):
        ch.step()
    b.record(ch.stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    tm = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    flop = 3279872.0 * Bl  # per rank per step (SURVEY.md section 8 sizes table)
    theta = ch.theta
    out = {"cfg": 3, "workload": f"wide MLP 64-512-512-512-1, TDist(5)+MixtureScale, SGNHTS, 1 chain, B={Bl}/GPU x {world} GPUs, n={n_loc}/GPU, compute={bnn.compute}",
           "n_gpus": world, "steps": K, "ms_per_step": ms / K, "grad_evals_per_s": K / ms * 1e3,
           "observations_per_s": K * Bl * world / ms * 1e3,
           "tflops_per_gpu": flop * K / (ms * 1e-3) / 1e12,
           "split_ms": {"grad_local": tg / K, "allreduce": tc / K, "update": tu / K},
           "allreduce_bytes": 4 * (P + 3), "finite": bool(np.isfinite(theta).all())}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def cfg4(bf, torch, args):
    rng = np.random.default_rng(SEED)
    T, n, B, C = 50, 16384, 64, args.chains or 256
    series = (np.cumsum(rng.standard_normal((T + n, 4)), axis=0) * 0.05).astype(np.float32)
    win = np.lib.stride_tricks.sliding_window_view(series[:T + n - 1], (T, 4))[:n, 0]    # n x T x 4
    x = np.asfortranarray(np.transpose(win, (1, 2, 0)))                                   # T x d x n
    y = series[T:T + n, 0].copy()
    nc = bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1)))
    bnn = bf.BNN(x, y, bf.SeqToOneNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 0.5))
    P = bnn.num_total_params
    th0 = (0.05 * rng.standard_normal((P, C))).astype(np.float32)
    # steps > 1 keeps the full-data MH evaluation (2 x 16384 sequences per chain) off most updates; the reference
    # test's steps = 1 would make every update a full-data pass
    s = bf.GGMC(beta=0.1, l=1e-2, steps=args.ggmc_steps, sadapter=bf.DualAveragingStepSize(1e-2, target_accept=0.25, adapt_steps=1000),
                madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1, epsilon=1e-8))
    s.initialise(bnn, th0)
    U = args.steps or 8
    st = torch.cuda.Stream()
    bf.advance(s, B, 2, seed=SEED, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = events(torch)
    e0.record(st)
    bf.advance(s, B, U, seed=SEED, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = not s.get_state(bf._lib.STATE_STATUS, np.int32, False).any()
    return {"cfg": 4, "workload": f"LSTM(4,64)+Dense(64,1), T=50, n=16384, B=64, GGMC steps={args.ggmc_steps}, {C} chains",
            "ms": ms, "updates": U, "samples_per_s": C * U / ms * 1e3, "grad_evals_per_s": 2 * C * U / ms * 1e3,
            "minibatch_tflops": 2 * 64 * 5120384.0 * C * U / (ms * 1e-3) / 1e12, "finite": ok}


def cfg5(bf, torch, args):
    rng = np.random.default_rng(SEED)
    n, C, path = args.n or 100_000, args.chains or 4096, 5
    x = rng.standard_normal((10, n)).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(10, 64, bf.relu), bf.Dense(64, 64, bf.relu), bf.Dense(64, 1)))
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 1.0), compute=args.compute)
    P = bnn.num_total_params
    th0 = (0.1 * rng.standard_normal((P, C))).astype(np.float32)
    s = bf.HMC(1e-4, path, sadapter=bf.DualAveragingStepSize(1e-4, target_accept=0.5), madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1))
    s.initialise(bnn, th0)
    nsamp = args.steps or 1
    st = torch.cuda.Stream()
    e0, e1 = events(torch)
    e0.record(st)
    bf.advance(s, n, nsamp * path, seed=SEED, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = (path * 27520.0 + 2 * 9600.0) * n * C * nsamp
    return {"cfg": 5, "workload": f"HMC path_len 5, MLP 10-64-64-1, full batch n={n}, {C} chains, compute={bnn.compute}",
            "ms": ms, "hmc_samples": nsamp, "grad_evals_per_s": C * path * nsamp / ms * 1e3,
            "samples_per_s": C * nsamp / ms * 1e3, "tflops": flop / (ms * 1e-3) / 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--ggmc-steps", type=int, default=8)
    ap.add_argument("--compute", default="tc")
    args = ap.parse_args()
    import torch
    import bayesflux_b200 as bf
    out = {1: cfg1, 3: cfg3, 4: cfg4, 5: cfg5}[args.cfg](bf, torch, args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
