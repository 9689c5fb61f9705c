"""Development probe: sharded training (one process per GPU) — ms/tree and, with QR_TRACE=1, the GPU timeline of
the growth rounds on rank 0.  usage: longrun_sharded.py WORLD TREES [N_DOCS]"""
import multiprocessing as mp
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, trees, n_docs, q):
    import numpy as np
    from quickrank_b200 import api, synth
    from quickrank_b200.sharding import query_shards
    if rank != 0:
        os.environ.pop("QR_TRACE_ROUNDS", None)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
    x, l, off = synth.make_dataset(n_docs, 136, n_docs // 100, seed=20260102)
    comm = None
    if world > 1:
        if rank == 0:
            cid = api.comm_unique_id()
            for _ in range(world - 1):
                q.put(cid)
        else:
            cid = q.get()
        q0, q1 = query_shards(off, world)[rank]
        d0, d1 = int(off[q0]), int(off[q1])
        x, l, off = np.ascontiguousarray(x[d0:d1]), l[d0:d1], (off[q0:q1 + 1] - off[q0]).astype(np.uint64)
        comm = (cid, rank, world)
    tr = api.Trainer(x, l, off, algo="LAMBDAMART", nleaves=64, nthresholds=0, cutoff=10, hist_mode=0, device=rank, comm=comm)
    if rank == 0:
        print("exchange:", tr.comm_transport(), flush=True)
    t0 = time.time(); rs = []
    for i in range(trees):
        tr.boost_iteration(want_tree=False, want_metric=True)
        rs.append(tr.last_tree_rounds()[0])
        if i % 50 == 49 and rank == 0:
            dt = (time.time() - t0) / 50
            print("trees %4d-%4d: %.3f ms/tree, rounds/tree %.1f" % (i - 49, i, dt * 1e3, np.mean(rs)), flush=True)
            t0 = time.time(); rs = []
    tr.close()


if __name__ == "__main__":
    world, trees = int(sys.argv[1]), int(sys.argv[2])
    n_docs = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, trees, n_docs, q)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join()
    sys.exit(max(p.exitcode or 0 for p in ps))
