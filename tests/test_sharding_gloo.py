"""Host-side logic of multi-GPU training on CPU (gloo, world_size 2): query-aligned sharding and
the algebra the NCCL path relies on — per-bin fixed-point histograms of the shards add up exactly
to the histogram of the whole dataset, so every rank scans the same totals and picks the same split."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import qr_testlib as common
from oracle import pyoracle as po
from quickrank_b200.sharding import query_shards


def test_query_shards_are_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lens = rng.integers(1, 200, size=137)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        sh = query_shards(off, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(lens)
        for a, b in zip(sh[:-1], sh[1:]):
            assert a[1] == b[0]
        docs = [int(off[b]) - int(off[a]) for a, b in sh]
        assert min(docs) > 0
        assert max(docs) - min(docs) <= 2 * lens.max()
    with pytest.raises(ValueError):
        query_shards(np.array([0, 5, 9], np.uint64), 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, l, off = common.dataset(n=4000, f=12, q=40, seed=4)
    col = np.ascontiguousarray(x.T)
    ob = po.Binning(col, 0)                      # thresholds of the WHOLE dataset (what the merge step yields)
    bins = ob.bins()
    lam, w = po.lambdas(np.zeros(len(l)), l, off, 10)
    q0, q1 = query_shards(off, world)[rank]
    d0, d1 = int(off[q0]), int(off[q1])
    # fixed-point view with the common scale (max |lambda| over all ranks)
    m = torch.tensor([np.max(np.abs(lam[d0:d1]))], dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    e = int(np.frexp(m.item())[1])
    qexp = (62 - (int(np.ceil(np.log2(len(l)))) + 1)) - e
    lamq = np.rint(np.ldexp(lam, qexp)).astype(np.int64)
    ncell = max(len(ob.thresholds(f)) for f in range(x.shape[1]))
    hist = np.zeros((x.shape[1], ncell), np.int64)
    cnt = np.zeros((x.shape[1], ncell), np.int64)
    for f in range(x.shape[1]):
        np.add.at(hist[f], bins[f, d0:d1], lamq[d0:d1])
        np.add.at(cnt[f], bins[f, d0:d1], 1)
    th, tc = torch.from_numpy(hist), torch.from_numpy(cnt)
    dist.all_reduce(th)
    dist.all_reduce(tc)
    full_h = np.zeros_like(hist)
    full_c = np.zeros_like(cnt)
    for f in range(x.shape[1]):
        np.add.at(full_h[f], bins[f], lamq)
        np.add.at(full_c[f], bins[f], 1)
    ok = np.array_equal(th.numpy(), full_h) and np.array_equal(tc.numpy(), full_c)
    # replicated split decision: best (f, t) from the all-reduced totals == the oracle's root split
    cs = np.cumsum(th.numpy().astype(np.float64) * 2.0 ** -qexp, axis=1)
    cc = np.cumsum(tc.numpy(), axis=1)
    best, bf, bt = -1.0, -1, -1
    for f in range(x.shape[1]):
        nt = len(ob.thresholds(f))
        s, c = cs[f, nt - 1], cc[f, nt - 1]
        for t in range(nt):
            lc, rc = cc[f, t], c - cc[f, t]
            if lc >= 1 and rc >= 1:
                sc = cs[f, t] ** 2 / lc + (s - cs[f, t]) ** 2 / rc
                if sc > best:
                    best, bf, bt = sc, f, t
    tree = ob.fit_tree(lam, w, nleaves=2, minls=1)
    ok = ok and bf == int(tree["feature"][0]) and bt == int(tree["threshold_idx"][0])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_histograms_allreduce_to_the_global_one():
    world = 2
    mgr = tmp.Manager()
    ret = mgr.dict()
    port = _free_port()
    tmp.spawn(_rank_main, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


class _FakeTrainer:
    """Stands in for api.Trainer (which needs a GPU): records what sharded_trainer would create."""

    def __init__(self, x, labels, qoff, device=-1, comm=None, **kw):
        self.x, self.labels, self.qoff, self.device, self.comm, self.kw = x, labels, qoff, device, comm, kw


def _sharded_trainer_main(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quickrank_b200.distributed import sharded_trainer
    x, l, off = common.dataset(n=3000, f=6, q=30, seed=9)
    tr = sharded_trainer(x, l, off, dist=dist, trainer_cls=_FakeTrainer, algo="LAMBDAMART", nleaves=8)
    cid, r, w = tr.comm
    d0, d1 = tr.doc_range
    ret[rank] = dict(cid=cid, rank=r, world=w, d0=d0, d1=d1, device=tr.device, nq=len(tr.qoff) - 1,
                     rows_ok=bool(np.array_equal(tr.x, x[d0:d1]) and np.array_equal(tr.labels, l[d0:d1])),
                     off_ok=bool(tr.qoff[0] == 0 and tr.qoff[-1] == d1 - d0), kw=tr.kw)
    dist.destroy_process_group()


def test_sharded_trainer_helper_hands_every_rank_its_queries_and_the_same_id():
    """quickrank_b200.distributed.sharded_trainer over gloo (world 2): one communicator id for all ranks
    (created by rank 0 through the C ABI), disjoint contiguous document ranges that cover the dataset, local
    query offsets rebased to 0."""
    world = 2
    mgr = tmp.Manager()
    ret = mgr.dict()
    tmp.spawn(_sharded_trainer_main, args=(world, _free_port(), ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    assert a["cid"] == b["cid"] and len(a["cid"]) == 128 and any(a["cid"])
    assert (a["rank"], a["world"], b["rank"], b["world"]) == (0, 2, 1, 2)
    assert a["d0"] == 0 and a["d1"] == b["d0"] and b["d1"] == 3000
    assert a["nq"] + b["nq"] == 30 and a["rows_ok"] and b["rows_ok"] and a["off_ok"] and b["off_ok"]
    assert a["kw"] == dict(algo="LAMBDAMART", nleaves=8)
