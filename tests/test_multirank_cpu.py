"""N > 1 host logic on CPU: world_size-2 gloo process groups.

* chain sharding: the partition of chains over ranks is exact and keyed by global chain ids;
* minibatch sharding of one chain: every rank contributes the likelihood part of the gradient on its
  shard, one all-reduce(sum) of [g ; sum1 ; sum2 ; B_local], prior and sigma-prior added once afterwards
  -- the algebra bfx_stream_grad_local / bfx_stream_apply_update implement -- reproduces the
  single-process gradient of src/model/BNN.jl:61-69.  The compute stand-in here is the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bfx_oracle as O

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    layers = (O.Dense(6, 9, O.TANH), O.Dense(9, 7, O.RELU), O.Dense(7, 1))
    return O.Model(layers, O.Likelihood(O.TDIST, nu=5.0, prior_a=2.0, prior_b=0.5),
                   O.NetPrior(O.NETPRIOR_MIXTURE, sigma1=1.0, sigma2=0.1, pi1=0.5))


def _local_like_part(m, theta, x, y, nb):
    """[num_batches * dloglike/dtheta_net (likelihood only) ; 0 ; sum1 ; sum2 ; B_local] on one shard."""
    tn, _, tl = O.split_params(m, theta)
    sig = float(np.exp(tl[0]))
    yhat, cache = O.net_forward(m.layers, tn, x, want_cache=True)
    z = (y - yhat) / sig
    if m.like.kind == O.NORMAL:
        s1, s2, dl = float(np.sum(z * z)), 0.0, nb * z / sig
    else:
        nu = m.like.nu
        w = (nu + 1.0) * z / (nu + z * z)
        s1, s2, dl = float(np.sum(np.log1p(z * z / nu))), float(np.sum(w * z)), nb * w / sig
    g = O.net_backward(m.layers, tn, x, dl, cache)
    return np.concatenate([g, [0.0, s1, s2, float(y.shape[0])]])


def _finish(m, theta, buf, nb):
    """what bfx_stream_apply_update adds ONCE after the all-reduce: sigma prior + network prior."""
    n_net = m.n_net
    g = buf[:n_net + 1].copy()
    s1, s2, B = buf[n_net + 1], buf[n_net + 2], buf[n_net + 3]
    sl = float(theta[-1])
    dlps = O.dlogpdf_sigma_prior_transformed(m.like, sl)
    dlike = (s1 if m.like.kind == O.NORMAL else s2) - B + dlps
    c0, c2 = O.netprior_coeffs(m.prior)
    g[:n_net] -= c2 * theta[:n_net]
    g[n_net] = nb * dlike
    return g


def _worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bfx_sharded_only", os.path.join(root, "bayesflux.jl_b200", "sharded.py"))
    try:
        rng = np.random.default_rng(6150533)
        n, B = 64, 24
        m = _model()
        x = rng.standard_normal((6, n))
        y = rng.standard_normal(n)
        theta = 0.3 * rng.standard_normal(m.n_total)
        batch = rng.permutation(n)[:B]  # the global minibatch, split evenly over the ranks
        mine = batch[rank * (B // WORLD):(rank + 1) * (B // WORLD)]
        nb = 5.0
        buf = torch.from_numpy(_local_like_part(m, theta, x[:, mine], y[mine], nb))
        dist.all_reduce(buf)  # one collective per grad-eval
        g = _finish(m, theta, buf.numpy(), nb)
        v_ref, g_ref = O.grad_loglikeprior(m, theta, x[:, batch], y[batch], nb)
        err = float(np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref))
        # replicated update with identical noise keeps theta identical on every rank
        noise = np.random.default_rng(99).standard_normal(m.n_total)
        th_new = torch.from_numpy(theta + 1e-3 * g + 1e-2 * noise)
        gathered = [torch.empty_like(th_new) for _ in range(WORLD)]
        dist.all_gather(gathered, th_new)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        # timing contract of bench.py: max over ranks
        tmax = torch.tensor([float(rank + 1)])
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        q.put((rank, err, same, float(tmax.item())))
    finally:
        dist.destroy_process_group()


def test_minibatch_sharded_gradient_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, tmax in res:
        assert err < 1e-12, (rank, err)
        assert same
        assert tmax == float(WORLD)


def _load_sharded():
    """sharded.py's pure helpers without importing the package (which needs libbfx.so + ctypes only -- fine on CPU)."""
    import bayesflux_b200 as bf
    return bf


@pytest.mark.parametrize("nchains,world", [(1024, 1), (1024, 2), (1024, 8), (4096, 8), (10, 4), (3, 8)])
def test_chain_sharding_partition(nchains, world):
    bf = _load_sharded()
    seen = []
    for r in range(world):
        off, cnt = bf.shard_chains(nchains, r, world)
        seen.extend(range(off, off + cnt))
        assert cnt in (nchains // world, nchains // world + 1)
    assert seen == list(range(nchains))  # every global chain id exactly once, in rank order


def test_data_sharding_and_global_num_batches():
    bf = _load_sharded()
    n, world = 2 ** 15 + 3, 8
    cover = np.zeros(n, int)
    for r in range(world):
        cover[bf.shard_data(n, r, world)] += 1
    assert (cover == 1).all()
    assert bf.global_num_batches(2 ** 23, 65536) == 128      # configs[2]
    assert bf.global_num_batches(500, 50) == 10              # configs[0]
    assert bf.global_num_batches(105, 10, partial=False) == 10


# ---- peer-memory exchange of the minibatch-sharded chain (csrc/bfx_stream.cu: k_finish_peer / k_finish_scatter) ----
@pytest.mark.parametrize("P,world", [(559106, 2), (559106, 8), (17, 4), (3, 8), (4096, 3), (1, 2)])
def test_peer_slices_partition_the_parameters(P, world):
    bf = _load_sharded()
    cover = np.zeros(P, int)
    prev_stop = 0
    for r in range(world):
        sl = bf.peer_slice(P, r, world)
        assert sl.start == prev_stop or sl.start == sl.stop  # contiguous, in rank order (late ranks may own nothing)
        assert sl.start == sl.stop or sl.start % 4 == 0        # quad aligned: 16-byte loads / stores on the device
        cover[sl] += 1
        prev_stop = max(prev_stop, sl.stop)
    assert (cover == 1).all()


def _peer_worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import bayesflux_b200 as bf
        P = 1003
        mine = np.random.default_rng(100 + rank).standard_normal(P + 3).astype(np.float32)
        # every rank can read every rank's buffer (peer memory): emulated by an all-gather of the buffers
        got = [torch.empty(P + 3) for _ in range(WORLD)]
        dist.all_gather(got, torch.from_numpy(mine))
        bufs = [g.numpy() for g in got]
        pull = bf.peer_exchange_reference(bufs, rank, "pull")
        scat = bf.peer_exchange_reference(bufs, rank, "scatter")
        # against the collective it replaces (sum order of gloo's all-reduce is not the rank order: tolerance)
        red = torch.from_numpy(mine.copy())
        dist.all_reduce(red)
        err = float(np.max(np.abs(pull - red.numpy())))
        # bit-identical on every rank and between the protocols
        allp = [torch.empty(P + 3) for _ in range(WORLD)]
        dist.all_gather(allp, torch.from_numpy(scat))
        same = all(torch.equal(allp[0], t) for t in allp) and np.array_equal(pull, scat)
        q.put((rank, err, same))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_sums_are_rank_ordered_and_identical_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same in res:
        assert err < 1e-5, (rank, err)
        assert same
