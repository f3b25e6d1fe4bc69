import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import bayesflux_b200 as bf
rng = np.random.default_rng(1)
x, y = bench.synth_data(1_000_000, rng)
net = bf.Chain(bf.Dense(10, 64, bf.relu), bf.Dense(64, 64, bf.relu), bf.Dense(64, 1))
nc = bf.destruct(net)
bnn = bf.BNN(x, y, bf.FeedforwardNormal(nc, bf.Gamma(2.0, 0.5)), bf.GaussianPrior(nc, 1.0), compute="tc")
P = bnn.num_total_params; C = 1024
th = np.asfortranarray((0.1 * rng.standard_normal((P, C))).astype(np.float32))
s = bf.SGLD()
for U in (1, 128, 128, 1, 128):
    for rep in range(3):
        t0 = time.perf_counter()
        ch = bf.mcmc(bnn, 64, U, s, theta_start=th, keep_every=U, seed=rep)
        t1 = time.perf_counter()
    print(f"U={U}: mcmc call {1e3*(t1-t0):.2f} ms")
# raw copy speeds
d = torch.empty(P * C, dtype=torch.float32, device="cuda")
h = torch.from_numpy(th.ravel(order="F").copy())
hp = h.pin_memory()
for name, src in (("pageable", h), ("pinned", hp)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): d.copy_(src, non_blocking=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"H2D 20MB {name}: {(t1-t0)/5*1e3:.2f} ms")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): src.copy_(d)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"D2H 20MB {name}: {(t1-t0)/5*1e3:.2f} ms")
t0 = time.perf_counter()
for _ in range(5): a = np.empty((P, 1, C), np.float32, order="F"); a[:] = 0
t1 = time.perf_counter()
print(f"np.empty+touch 20MB: {(t1-t0)/5*1e3:.2f} ms")
