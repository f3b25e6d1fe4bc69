"""Multi-GPU plumbing of the engine.  One process per GPU.

Two ways the hot path shards:

* **chains** are independent units: rank r runs chains ``[offset, offset + count)`` of ``shard_chains`` with
  ``chain_offset = offset`` keying its Philox / permutation streams -- no collective on the data path, results
  identical for any world size;
* **one large chain, minibatch-sharded** (configs[2]): every rank holds its shard of the data and a replica
  of the sampler; each ``update!`` is ``bfx_stream_grad_local`` -> one all-reduce (sum, fp32, P + 3 floats:
  NCCL over NVLink through ``torch.distributed``) -> ``bfx_stream_apply_update`` with identical Philox
  streams on every rank, so theta / p / xi stay replicated without a broadcast.

torch is used here for device buffers, streams and the process group only.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib as L
from . import api as A


def shard_chains(nchains: int, rank: int, world: int) -> Tuple[int, int]:
    """(offset, count) of the contiguous block of chains rank ``rank`` owns; blocks differ by at most one chain."""
    base, rem = divmod(nchains, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def shard_data(n: int, rank: int, world: int) -> slice:
    """The contiguous block of observations rank ``rank`` holds in the minibatch-sharded mode."""
    off, cnt = shard_chains(n, rank, world)
    return slice(off, off + cnt)


def global_num_batches(n_global: int, batch_global: int, partial: bool = True) -> int:
    """length(DataLoader) of the UNSHARDED problem (src/inference/mcmc/abstract.jl:105)."""
    return -(-n_global // batch_global) if partial else max(1, n_global // batch_global)


def peer_slice(P: int, rank: int, world: int) -> slice:
    """The parameters rank ``rank`` finishes in the SCATTER protocol of the peer-memory exchange
    (csrc/bfx_stream.cu: k_finish_scatter): the quads (groups of four consecutive floats) are split into
    ``world`` contiguous runs of ``ceil(nq / world)`` quads."""
    nq = (P + 3) // 4
    per = -(-nq // world)
    q0 = min(per * rank, nq)
    q1 = min(q0 + per, nq)
    return slice(min(4 * q0, P), min(4 * q1, P))


def peer_exchange_reference(buffers, rank: int, protocol: str = "scatter"):
    """Host restatement of what rank ``rank`` ends up with after the peer-memory exchange of the ranks' buffers
    ``[g ; sums]`` (float32 arrays of equal length): every element is the sum of the ranks' contributions taken IN RANK
    ORDER in float32 -- by this rank for every element (``pull``) or by the slice's owner (``scatter``).  Both orders
    are the same, so every rank holds bit-identical sums whatever the protocol; used by the CPU tests."""
    world = len(buffers)
    P = len(buffers[0])
    out = np.empty(P, np.float32)
    owners = [rank] * world if protocol == "pull" else list(range(world))
    for r in (range(world) if protocol == "scatter" else [rank]):
        sl = peer_slice(P, r, world) if protocol == "scatter" else slice(0, P)
        acc = buffers[0][sl].astype(np.float32).copy()
        for j in range(1, world):
            acc = (acc + buffers[j][sl].astype(np.float32)).astype(np.float32)
        out[sl] = acc
    del owners
    return out


def model_regime(bnn: A.BNN) -> str:
    r = C.c_int32()
    L.check(L.lib.bfx_model_regime(bnn._h, C.byref(r)))
    return "stream" if r.value == L.REGIME_STREAM else "smem"


class ShardedChain:
    """One chain whose minibatch is sharded over the ranks of a process group.

    ``bnn_local`` holds this rank's shard of the data; ``local_batchsize`` observations are drawn from it per
    update; ``num_batches`` is the GLOBAL number of batches (the likelihood scale of src/model/BNN.jl:55).
    ``allreduce`` defaults to ``torch.distributed.all_reduce`` on the default group (a no-op for world size 1);
    tests inject their own to emulate several ranks inside one process."""

    def __init__(self, bnn_local: A.BNN, sampler: A.MCMCState, theta_start, local_batchsize: int, num_batches: int,
                 seed: int = 0, shuffle: bool = True, partial: bool = True, rank: int = 0,
                 allreduce: Optional[Callable] = None):
        import torch
        self.torch = torch
        if model_regime(bnn_local) != "stream":
            raise ValueError("minibatch sharding is implemented for the large-P (stream) regime")
        self.bnn, self.sampler = bnn_local, sampler
        self.P = bnn_local.num_total_params
        sampler.initialise(bnn_local, np.asarray(theta_start, np.float32).reshape(self.P, 1))
        # every rank draws from ITS shard with its own permutation stream (chain_offset = rank), but applies the
        # update with the shared noise key: see bfx_stream_apply_update
        self.run = A._run_desc(local_batchsize, 0, shuffle, partial, True, 1, seed, "per_chain", rank)
        self.nb = float(num_batches)
        dev = torch.device("cuda", bnn_local.device)
        self.buf = torch.empty(self.P + 3, dtype=torch.float32, device=dev)
        self.allreduce = allreduce
        # a dedicated (non-default) stream: the engine enqueues on it and the collective is ordered after it
        # (stream handle 0 would mean "the engine's own stream" to the C ABI)
        self.stream = torch.cuda.Stream(device=dev)

    def _reduce(self):
        with self.torch.cuda.stream(self.stream):
            if self.allreduce is not None:
                self.allreduce(self.buf)
                return
            dist = self.torch.distributed
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self.buf)

    def grad_local(self):
        st = self.stream.cuda_stream
        L.check(L.lib.bfx_stream_grad_local(self.sampler._h, C.byref(self.run), self.nb, C.c_void_p(st),
                                            C.c_void_p(self.buf.data_ptr()), self.P + 3))

    def apply_update(self, sample_out=None):
        st = self.stream.cuda_stream
        ptr = C.c_void_p(sample_out.data_ptr()) if sample_out is not None else None
        L.check(L.lib.bfx_stream_apply_update(self.sampler._h, C.byref(self.run), self.nb, C.c_void_p(st),
                                              C.c_void_p(self.buf.data_ptr()), ptr))

    def step(self, sample_out=None):
        """One update!: local gradient, all-reduce, replicated update."""
        self.grad_local()
        self._reduce()
        self.apply_update(sample_out)

    @property
    def theta(self):
        return self.sampler.theta[:, 0]


def mcmc_sharded(bnn_local: A.BNN, local_batchsize: int, numsamples: int, sampler: A.MCMCState, num_batches: int,
                 theta_start, seed: int = 0, keep_every: int = 1, rank: int = 0, shuffle: bool = True,
                 partial: bool = True, allreduce: Optional[Callable] = None):
    """mcmc(...) for one minibatch-sharded chain: returns the (P, nkept) samples (identical on every rank)."""
    import torch
    ch = ShardedChain(bnn_local, sampler, theta_start, local_batchsize, num_batches, seed, shuffle, partial, rank,
                      allreduce)
    nkept = numsamples // keep_every
    out = torch.empty((max(nkept, 1), ch.P), dtype=torch.float32, device=ch.buf.device)
    for t in range(1, numsamples + 1):
        ch.step(out[t // keep_every - 1] if t % keep_every == 0 else None)
    torch.cuda.synchronize()
    L.check(L.OK)
    st = sampler.get_state(L.STATE_STATUS, np.int32, False)
    if st.any():
        raise L.BfxError(int(st[0]), "NaN in g" if int(st[0]) == L.ERR_NAN_G else "NaN in θ")
    sampler.nsampled = numsamples
    sampler.samples = np.asfortranarray(out[:nkept].cpu().numpy().T)
    return sampler.samples
