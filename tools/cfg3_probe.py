"""cfg3 (wide MLP, one chain, stream regime) update timing through bfx_mcmc_advance (CUDA-graph replays).
usage: python tools/cfg3_probe.py [--steps K] [--compute tc|fp32] [--batch B]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--compute", default="tc")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--n", type=int, default=1 << 20)
    args = ap.parse_args()
    import torch
    import bayesflux_b200 as bf
    W = 512
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((args.n, 64), generator=g, device="cuda", dtype=torch.float32)
    y = torch.randn((args.n,), generator=g, device="cuda", dtype=torch.float32)
    nc = bf.destruct(bf.Chain(bf.Dense(64, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, 1)))
    bnn = bf.BNN(None, None, bf.FeedforwardTDist(nc, bf.Gamma(2.0, 0.5), 5.0), bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5),
                 compute=args.compute)
    bnn.set_data_device(x.data_ptr(), (64, args.n), y.data_ptr())
    P = bnn.num_total_params
    th0 = (0.03 * np.random.default_rng(0).standard_normal((P, 1))).astype(np.float32)
    s = bf.SGNHTS(1e-3, 1.0)
    s.initialise(bnn, th0)
    st = torch.cuda.Stream()
    bf.advance(s, args.batch, 5, seed=1, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    bf.advance(s, args.batch, args.steps, seed=1, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    th = s.theta
    flop = 3279872.0 * args.batch
    print(f"cfg3 {args.compute} B={args.batch}: {ms * 1e3:.1f} us per update, {flop / ms / 1e9:.1f} TFLOP/s FP32-equivalent, "
          f"graph replays {bf.graph_replays(s)}, launches {bf.launch_count(s)}, finite {bool(np.isfinite(th).all())}, "
          f"status {s.get_state(bf._lib.STATE_STATUS, np.int32, False)}")


if __name__ == "__main__":
    main()
