"""Bring-up probe: graph-captured and eager collectives on the library's communicator (2 ranks under torchrun).
usage: torchrun --nproc-per-node 2 tools/nccl_mix_probe.py <variant>"""
import ctypes as C
import faulthandler
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=True)
T0 = time.time()


def log(m):
    sys.stderr.write(f"[rank {os.environ.get('RANK')} +{time.time() - T0:5.1f}s] {m}\n")
    sys.stderr.flush()


def main():
    variant = sys.argv[1]
    import torch
    import torch.distributed as dist
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bayesflux_b200 as bf
    from bayesflux_b200 import _lib as L
    from bayesflux_b200 import api as A
    n_loc, Bl, W = 1 << 16, 2048, 256
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = torch.randn((n_loc, 64), generator=g, device="cuda")
    y = torch.randn((n_loc,), generator=g, device="cuda")
    nc = bf.destruct(bf.Chain(bf.Dense(64, W, bf.relu), bf.Dense(W, W, bf.relu), bf.Dense(W, 1)))
    bnn = bf.BNN(None, None, bf.FeedforwardTDist(nc, bf.Gamma(2.0, 0.5), 5.0), bf.MixtureScalePrior(nc, 1.0, 0.1, 0.5),
                 device=local, compute="tc")
    bnn.set_data_device(x.data_ptr(), (64, n_loc), y.data_ptr())
    P = bnn.num_total_params
    th0 = (0.03 * np.random.default_rng(0).standard_normal((P, 1))).astype(np.float32)
    s = bf.SGNHTS(1e-3, 1.0)
    s.initialise(bnn, th0)
    idt = torch.zeros(L.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(bf.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    bf.comm_init(s, bytes(idt.cpu().numpy().tobytes()), world, rank)
    log("comm up")
    st = torch.cuda.Stream()
    buf = torch.zeros(P + 3, dtype=torch.float32, device="cuda")
    kw = dict(seed=1, chain_offset=rank, stream=st.cuda_stream, ngpus=world, shard_mode="minibatch")

    def eager(k):
        for _ in range(k):
            L.check(L.lib.bfx_stream_allreduce(s._h, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr())))
        torch.cuda.synchronize()
        log(f"{k} eager all-reduces done")

    def graphs(k):
        bf.advance(s, Bl, k, **kw)
        torch.cuda.synchronize()
        log(f"{k} graph updates done (replays {bf.graph_replays(s)})")

    if variant == "eager_then_graph":
        eager(3); graphs(5); graphs(5)
    elif variant == "graph_then_eager":
        graphs(5); eager(1); eager(3); graphs(3)
    elif variant == "graph_then_torch_then_eager":
        graphs(5)
        t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize(); log("torch all_reduce done")
        eager(2)
    elif variant == "threecall":
        run = A._run_desc(Bl, 0, True, True, True, 1, 1, "per_chain", rank, world, "minibatch")
        nb = float(n_loc // Bl)
        graphs(5)
        for it in range(3):
            L.check(L.lib.bfx_stream_grad_local(s._h, C.byref(run), nb, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr()), P + 3))
            torch.cuda.synchronize(); log("grad_local done")
            L.check(L.lib.bfx_stream_allreduce(s._h, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr())))
            torch.cuda.synchronize(); log("allreduce done")
            L.check(L.lib.bfx_stream_apply_update(s._h, C.byref(run), nb, C.c_void_p(st.cuda_stream), C.c_void_p(buf.data_ptr()), None))
            torch.cuda.synchronize(); log("apply done")
    th = s.theta[:, 0]
    chk = torch.tensor([float(np.frombuffer(th.tobytes(), dtype=np.uint32).astype(np.uint64).sum() % (1 << 52))], dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    log(f"replicas identical: {bool((lo == hi).item())}, finite {bool(np.isfinite(th).all())}")
    bf.comm_destroy(s)
    log("comm destroyed")
    dist.barrier()
    dist.destroy_process_group()
    log("variant ok")


if __name__ == "__main__":
    main()
