"""Per-segment gradient error of the tensor-core path vs the float64 oracle (GPU box only)."""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflux_b200 as bf  # noqa: E402
from oracle import bfx_oracle as O  # noqa: E402
from tests.test_gpu_tc import NETS, build  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
like = sys.argv[3] if len(sys.argv) > 3 else "normal"
rng = np.random.default_rng(1)
d = NETS[net][0][0]
n = 400
x = rng.standard_normal((d, n)).astype(np.float32)
y = rng.standard_normal(n).astype(np.float32)
for compute in ("fp32", "tc"):
    bnn, m = build(bf, net, like, "gauss", x, y, compute=compute)
    theta = (0.3 * np.random.default_rng(2).standard_normal((bnn.num_total_params, 2))).astype(np.float32)
    idx = np.stack([np.random.default_rng(3 + c).permutation(n)[:B] for c in range(2)], axis=1)
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=6.0)
    for c in range(2):
        vo, go = O.grad_loglikeprior(m, theta[:, c].astype(np.float64), x[:, idx[:, c]].astype(np.float64),
                                     y[idx[:, c]].astype(np.float64), 6.0)
        print(f"[{compute}] chain {c}: value {v[c]:.6f} oracle {vo:.6f}  total rel {np.linalg.norm(g[:, c]-go)/np.linalg.norm(go):.3e}")
        off = 0
        for li, (i, o, a) in enumerate(NETS[net]):
            for nm, cnt in (("W", i * o), ("b", o)):
                seg = slice(off, off + cnt)
                e = np.linalg.norm(g[seg, c] - go[seg]) / max(np.linalg.norm(go[seg]), 1e-30)
                print(f"    L{li}.{nm}: rel {e:.3e}  |g| {np.linalg.norm(go[seg]):.3e} first got {g[seg, c][:3]} want {go[seg][:3]}")
                off += cnt
        print(f"    like: got {g[-1, c]} want {go[-1]}")
