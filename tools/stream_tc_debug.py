"""Per-layer gradient error of the stream-regime tensor-core path vs the float64 oracle (GPU box only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflux_b200 as bf  # noqa: E402
from oracle import bfx_oracle as O  # noqa: E402
from tests.test_gpu_parity import LIKES, PRIORS, build  # noqa: E402
from tests.test_gpu_stream import TC_NETS, _tc  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "cfg3_small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(1)
layers = TC_NETS[net]
n = 900
x = rng.standard_normal((layers[0].nin, n)).astype(np.float32)
y = rng.standard_normal(n).astype(np.float32)
for mode in ("fp32", "tc"):
    bnn, m = build(bf, layers, LIKES["normal_gamma"], PRIORS["gauss"], x, y)
    if mode == "tc":
        _tc(bf, bnn)
    theta = (0.1 * np.random.default_rng(2).standard_normal((m.n_total, 1))).astype(np.float32)
    idx = np.random.default_rng(3).permutation(n)[:B]
    v, g = bf.grad_loglikeprior(bnn, theta, idx, num_batches=4.0)
    vo, go = O.grad_loglikeprior(m, theta[:, 0].astype(np.float64), x[:, idx].astype(np.float64), y[idx].astype(np.float64), 4.0)
    print(f"[{mode}] value {v[0]:.6f} oracle {vo:.6f} total rel {np.linalg.norm(g[:, 0]-go)/np.linalg.norm(go):.3e}")
    off = 0
    for li, l in enumerate(layers):
        for nm, cnt in (("W", l.nin * l.nout), ("b", l.nout)):
            seg = slice(off, off + cnt)
            e = np.linalg.norm(g[seg, 0] - go[seg]) / max(np.linalg.norm(go[seg]), 1e-30)
            extra = ""
            if nm == "W" and e > 1e-4:
                d = (g[seg, 0] - go[seg]).reshape(l.nout, l.nin, order="F")
                rows = np.linalg.norm(d, axis=1); cols = np.linalg.norm(d, axis=0)
                extra = f" worst rows {np.argsort(-rows)[:6]} worst cols {np.argsort(-cols)[:6]} rowerr[:4] {rows[:4]}"
            print(f"    L{li}.{nm}: rel {e:.3e}{extra}")
            off += cnt
