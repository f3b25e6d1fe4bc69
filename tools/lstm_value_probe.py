#!/usr/bin/env python
"""The forward-only LSTM tile alone: bfx_loglikeprior of LSTM(4,64)+Dense(64,1), T=50, over n sequences for C theta
columns (one CTA per column: the full-data passes of the GGMC / HMC acceptance tests run exactly this)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bayesflux_b200 as bf

rng = np.random.default_rng(1)
T, n = 50, 16384
C = int(sys.argv[1]) if len(sys.argv) > 1 else 148
series = (np.cumsum(rng.standard_normal((T + n, 4)), axis=0) * 0.05).astype(np.float32)
win = np.lib.stride_tricks.sliding_window_view(series[:T + n - 1], (T, 4))[:n, 0]
x = np.asfortranarray(np.transpose(win, (1, 2, 0)))
y = series[T:T + n, 0].copy()
nc = bf.destruct(bf.Chain(bf.LSTM(4, 64), bf.Dense(64, 1)))
bnn = bf.BNN(x, y, bf.SeqToOneNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 0.5))
P = bnn.num_total_params
th = np.asfortranarray((0.05 * rng.standard_normal((P, C))).astype(np.float32))
bf.loglikeprior(bnn, th)
torch.cuda.synchronize()
t0 = time.perf_counter()
v = bf.loglikeprior(bnn, th)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3
flop = 2.0 * (4 * 64) * (4 + 64) * T * n * C   # gate contractions only
print(json.dumps({"workload": f"loglikeprior LSTM(4,64)+Dense(64,1) T=50, n={n}, {C} theta columns (forward-only 64-sequence tile)",
                  "ms": ms, "sequences_per_s": n * C / ms * 1e3, "tflops_fp32": flop / (ms * 1e-3) / 1e12,
                  "finite": bool(np.isfinite(v).all())}))
