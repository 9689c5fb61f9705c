"""Sharded training (one process per GPU, documents sharded by query; 2 ranks, QR_TEST_WORLD for more)
must grow the same trees as one GPU, whether the per-round histograms are exchanged inside the
split-scan kernel (peer-memory loads over NVLink), by the stand-alone peer-memory kernel or by NCCL
all-reduces: histogram sums are fixed-point integers, so the totals do not
depend on the number of ranks or on the order of the additions."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import qr_testlib as common

pytestmark = pytest.mark.gpu

T = 6


MODES = {
    # name: (environment, expected transport)
    "peer-auto": ({}, "peer"),                                   # leaf-wise: feature-sliced scan + winner exchange; oblivious: as below
    "peer-unsliced": ({"QR_PEER_SLICED": "0"}, "peer"),          # every rank scans every feature: fused for small rounds, stand-alone kernel for wide ones
    "peer-fused": ({"QR_PEER_SLICED": "0", "QR_PEER_ONESHOT_MAX": "100000"}, "peer"),   # every round: all-reduce inside the split-scan kernel
    "peer-twoshot": ({"QR_PEER_FUSED": "0"}, "peer"),            # every round: the in-place reduce-scatter + all-gather kernel
    "nccl": ({"QR_PEER_REDUCE": "0"}, "nccl"),
}
WORLD = int(os.environ.get("QR_TEST_WORLD", "2"))


def _worker(rank, world, q, out_q, algo, kw, mode):
    for k in ("QR_PEER_REDUCE", "QR_PEER_FUSED", "QR_PEER_ONESHOT_MAX", "QR_PEER_SLICED"):
        os.environ.pop(k, None)
    os.environ.update(MODES[mode][0])
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from quickrank_b200 import api
    from quickrank_b200.sharding import query_shards
    kw = dict(kw)
    x, l, off = common.dataset(n=20000, f=24, q=200, seed=21, **kw.pop("_data", {}))
    if rank == 0:
        cid = api.comm_unique_id()
        for _ in range(world - 1):
            q.put(cid)
    else:
        cid = q.get()
    q0, q1 = query_shards(off, world)[rank]
    d0, d1 = int(off[q0]), int(off[q1])
    tr = api.Trainer(x[d0:d1], l[d0:d1], (off[q0:q1 + 1] - off[q0]).astype(np.uint64), algo=algo, device=rank,
                     comm=(cid, rank, world), **kw)
    trees, metrics = [], []
    for _ in range(T):
        t, m = tr.boost_iteration()
        trees.append(t)
        metrics.append(m)
    scores = tr.get_scores()
    transport = tr.comm_transport()
    tr.close()
    out_q.put((rank, trees, metrics, d0, d1, scores, transport))


CASES = [("LAMBDAMART", dict(nleaves=16)), ("MART", dict(nleaves=8)), ("OBVLAMBDAMART", dict(treedepth=3)),
         # continuous features: thousands of thresholds per feature (16-bit bins, histograms accumulated in
         # global memory, the split scan's path for wide features) ...
         ("LAMBDAMART", dict(nleaves=12, _data=dict(gridded=False))),
         # ... and the reference's equal-width thresholds over the GLOBAL value range (mart.cc:159-169)
         ("LAMBDAMART", dict(nleaves=12, nthresholds=64, _data=dict(gridded=False)))]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("algo,kw", CASES, ids=["lambdamart", "mart", "oblivious", "continuous", "equal-width"])
def test_sharded_ranks_grow_the_single_gpu_trees(algo, kw, mode):
    from quickrank_b200 import api
    if api.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    tkw = {k: v for k, v in kw.items() if k != "_data"}
    x, l, off = common.dataset(n=20000, f=24, q=200, seed=21, **kw.get("_data", {}))
    with api.Trainer(x, l, off, algo=algo, **tkw) as tr:
        want = [tr.boost_iteration() for _ in range(T)]
        want_scores = tr.get_scores()
    ctx = mp.get_context("spawn")
    q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, WORLD, q, out_q, algo, kw, mode)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = sorted([out_q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got_scores = np.zeros(len(l))
    for rank, trees, metrics, d0, d1, scores, transport in results:
        # the exchange under test must be the one that ran (the peer-memory path falls back to NCCL, with a
        # message on stderr, only where the devices cannot map each other's memory)
        assert transport == MODES[mode][1], transport
        got_scores[d0:d1] = scores
        for m in range(T):
            wt, wm = want[m]
            assert common.same_structure(trees[m], wt), "rank %d tree %d: %s" % (rank, m, common.describe_tree_diff(trees[m], wt))
            assert np.array_equal(trees[m]["count"], wt["count"])
            lv = common.leaves_mask(wt)
            # leaf sums and the NDCG mean are exact integer (fixed-point) sums: bit-identical for any sharding
            assert np.array_equal(trees[m]["value"][lv], wt["value"][lv]), "rank %d tree %d leaf outputs" % (rank, m)
            assert metrics[m] == wm, (rank, m, metrics[m], wm)
    assert np.array_equal(got_scores, want_scores)
    # all ranks hold the identical model
    for r in range(1, WORLD):
        for m in range(T):
            assert common.same_structure(results[0][1][m], results[r][1][m])
            assert np.array_equal(results[0][1][m]["value"], results[r][1][m]["value"])
