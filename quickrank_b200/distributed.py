"""Sharded (multi-GPU) training from a `torch.distributed` process group: one process per GPU, documents
sharded by query.  torch.distributed is only the plumbing that carries the 128-byte
communicator id from rank 0 to the others; everything after that goes through the C ABI
(`qr_ctx_create_sharded`: NCCL for the per-tree scalars, NVLink peer memory for the per-round histograms).

    import torch.distributed as dist
    dist.init_process_group("nccl")          # or "gloo": only a broadcast of 128 bytes is needed
    tr = sharded_trainer(x, labels, qoff, algo="LAMBDAMART", nleaves=64)   # every rank passes the FULL arrays
    for _ in range(1000):
        tree, ndcg = tr.boost_iteration()    # the same tree and metric on every rank
"""
from __future__ import annotations

import numpy as np

from . import api
from .sharding import query_shards


def broadcast_comm_id(dist, device=None) -> bytes:
    """Rank 0 creates the communicator id; every rank returns the same 128 bytes."""
    import torch
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    idt = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().numpy().tobytes())


def local_shard(x, labels, qoff, rank: int, world: int, layout: str = "rowmajor"):
    """This rank's contiguous range of whole queries: (x_local, labels_local, qoff_local, (doc_begin, doc_end))."""
    qoff = np.asarray(qoff)
    q0, q1 = query_shards(qoff, world)[rank]
    d0, d1 = int(qoff[q0]), int(qoff[q1])
    xl = x[d0:d1] if layout == "rowmajor" else x[:, d0:d1]
    return (np.ascontiguousarray(xl, np.float32), np.ascontiguousarray(labels[d0:d1], np.float32),
            (qoff[q0:q1 + 1] - qoff[q0]).astype(np.uint64), (d0, d1))


def sharded_trainer(x, labels, qoff, dist=None, device=None, comm_id: bytes | None = None, trainer_cls=None, **kw):
    """Creates this rank's `api.Trainer` of a sharded run.  `x`, `labels`, `qoff` are the FULL dataset on every
    rank (each keeps its shard); `device` defaults to rank modulo the number of visible GPUs.  With one rank
    (or no process group) this is a plain single-GPU trainer.  The returned trainer carries `.doc_range`, the
    [begin, end) of its documents in the full dataset (for `get_scores()`)."""
    if dist is None:
        import torch.distributed as dist   # noqa: PLC0415
    trainer_cls = trainer_cls or api.Trainer
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    layout = kw.get("layout", "rowmajor")
    if world == 1:
        tr = trainer_cls(x, labels, qoff, device=-1 if device is None else device, **kw)
        tr.doc_range = (0, len(labels))
        return tr
    cid = comm_id if comm_id is not None else broadcast_comm_id(dist)
    xl, ll, ol, rng = local_shard(x, labels, qoff, rank, world, layout)
    if device is None:
        ndev = api.device_count()
        device = rank % ndev if ndev > 0 else rank
    tr = trainer_cls(xl, ll, ol, device=device, comm=(cid, rank, world), **kw)
    tr.doc_range = rng
    return tr
