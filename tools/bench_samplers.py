#!/usr/bin/env python
"""Device-resident throughput of the five samplers on the cfg2 model (secondary measurement; bench.py is the headline).
    python tools/bench_samplers.py [--compute tc|fp32] [--chains 1024] [--updates 128]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--compute", default="tc")
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--updates", type=int, default=128)
    ap.add_argument("--n", type=int, default=200_000)
    args = ap.parse_args()
    import torch
    import bench
    import bayesflux_b200 as bf
    rng = np.random.default_rng(1)
    x, y = bench.synth_data(args.n, rng)
    net = bf.Chain(bf.Dense(bench.D_IN, bench.H, bf.relu), bf.Dense(bench.H, bench.H, bf.relu), bf.Dense(bench.H, 1))
    nc = bf.destruct(net)
    bnn = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 1.0), compute=args.compute)
    P, C, U = bnn.num_total_params, args.chains, args.updates
    th0 = np.asfortranarray((0.1 * rng.standard_normal((P, C))).astype(np.float32))
    ring = torch.empty((C, 4, P), dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream()
    samplers = {
        "SGLD": lambda: bf.SGLD(),
        "SGNHT": lambda: bf.SGNHT(1e-4),
        "SGNHTS": lambda: bf.SGNHTS(1e-4),
        "SGNHTS+RMSProp": lambda: bf.SGNHTS(1e-4, madapter=bf.RMSPropMassAdapter()),
    }
    out = {}
    for name, mk in samplers.items():
        s = mk()
        s.initialise(bnn, th0)
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            bf.advance(s, 64, U, keep_every=1, seed=1, stream=st.cuda_stream, samples_dev=ring.data_ptr(), ring_slots=4)
            e1.record(st)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"ms": ms, "updates_per_s": C * U / ms * 1e3}
    print(json.dumps({"workload": f"cfg2 model, {C} chains x {U} updates, B=64, n={args.n}, compute={args.compute}",
                      "samplers": out}))


if __name__ == "__main__":
    main()
