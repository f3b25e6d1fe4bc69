#!/usr/bin/env python
"""Raw throughput of the recurrent kernel: LSTM(4,64)+Dense(64,1), T=50, B=64, SGLD (no MH passes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bayesflux_b200 as bf

rng = np.random.default_rng(1)
T, n, B = 50, 16384, 64
C = int(sys.argv[1]) if len(sys.argv) > 1 else 296
U = int(sys.argv[2]) if len(sys.argv) > 2 else 16
series = (np.cumsum(rng.standard_normal((T + n, 4)), axis=0) * 0.05).astype(np.float32)
win = np.lib.stride_tricks.sliding_window_view(series[:T + n - 1], (T, 4))[:n, 0]
x = np.asfortranarray(np.transpose(win, (1, 2, 0)))
y = series[T:T + n, 0].copy()
nc = bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1)))
bnn = bf.BNN(x, y, bf.SeqToOneNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 0.5))
P = bnn.num_total_params
th0 = np.asfortranarray((0.05 * rng.standard_normal((P, C))).astype(np.float32))
s = bf.SGLD(stepsize_a=1e-4)
s.initialise(bnn, th0)
st = torch.cuda.Stream()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    bf.advance(s, B, U, seed=1, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flop = 2 * 64 * 5120384.0
print(json.dumps({"workload": f"LSTM(4,64)+Dense(64,1) T=50 B=64 SGLD, {C} chains x {U} updates", "P": int(P), "ms": ms,
                  "grad_evals_per_s": C * U / ms * 1e3, "tflops": flop * C * U / (ms * 1e-3) / 1e12}))
