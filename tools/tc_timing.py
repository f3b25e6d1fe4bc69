#!/usr/bin/env python
"""Phase-cycle breakdown of the cfg2 fused tensor-core kernel (developer tool, not a bench).

Needs a library built with -DBFX_TIMING on bfx_k_tc.cu (thread 0 of every CTA adds the cycles between
phase marks to a device counter array):
    LIBBFX=tools/libbfx_timing.so python tools/tc_timing.py [--sampler sgld]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

NAMES = ["step prologue (batch keys, dispatch)", "stage: indices + copy issue", "pack weights + copy wait", "fwd MMA issue+wait",
         "hidden epilogue", "output layer+like+delta", "bwd dA MMA wait", "bwd epilogue", "last dW wait",
         "gradient dump", "eval_finish", "sampler update pass", "chunk store+hand-over", "chunk fetch+load",
         "  (update: setup before the pass)", "  (update: thread 0's own pass)",
         "  (forward MMA issue by warp 0)", "  (backward MMA issue by warp 0)",
         "  (forward issue block entry: fence, elect, descriptors)"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--updates", type=int, default=500)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--sampler", default="sgld")
    ap.add_argument("--keep-every", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bayesflux_b200 as bf
    import bench
    fn = bf._lib.lib.bfx_debug_timing
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    buf = (C.c_ulonglong * 24)()
    rng = np.random.default_rng(1)
    x, y = bench.synth_data(200_000, rng)
    net = bf.Chain(bf.Dense(bench.D_IN, bench.H, bf.relu), bf.Dense(bench.H, bench.H, bf.relu), bf.Dense(bench.H, 1))
    nc = bf.destruct(net)
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 1.0), compute="tc")
    P = bnn.num_total_params
    th0 = np.asfortranarray((0.1 * rng.standard_normal((P, args.chains))).astype(np.float32))
    if args.sampler == "hmc":  # configs[4] shape class: full-batch trajectories; the table is then per 64-row tile
        s = bf.HMC(1e-4, 5, sadapter=bf.DualAveragingStepSize(1e-4, target_accept=0.5),
                   madapter=bf.DiagCovMassAdapter(1000, 100, kappa=0.1))
        args.batch = x.shape[1]
    else:
        s = {"sgld": bf.SGLD, "sgnht": bf.SGNHT, "sgnhts": bf.SGNHTS}[args.sampler]()
    s.initialise(bnn, th0)
    ring = torch.empty((args.chains, 4, P), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()

    def run():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        bf.advance(s, args.batch, args.updates, keep_every=args.keep_every, seed=1, stream=st.cuda_stream,
                   samples_dev=ring.data_ptr(), ring_slots=4)
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    run()
    fn(buf, 1)
    ms = run()
    fn(buf, 1)
    evals = args.chains * args.updates
    if args.sampler == "hmc":  # gradient tiles + the two full-data value passes of each 5-stage trajectory
        tiles = (args.batch + 63) // 64
        evals = args.chains * (args.updates * tiles + 2 * (args.updates // 5) * tiles)
    tot = sum(buf[i] for i in range(24))
    print(f"{ms:.2f} ms for {evals} grad-evals = {evals / ms / 1e3:.2f} M/s (instrumented build)")
    print(f"cycles per grad-eval (thread 0 of the CTA): {tot / evals:.0f}")
    for i, nme in enumerate(NAMES):
        print(f"  {i:2d} {nme:40s} {buf[i] / evals:9.0f}  {100.0 * buf[i] / tot:5.1f}%")


if __name__ == "__main__":
    main()
