"""Development probe: LambdaMART trained on a per-query document sample — every positive, the
top-scored `rank` share and a random `random` share of each query's negatives, redrawn every `every` trees — GPU sample
contexts against the unmodified reference's LambdaMartSelective::learn on the host cores, same data, same parameters,
same sample sizes.  The draw itself is emulated in numpy here (the C++ host, host/src/sampled_trainers.cc, holds the
reference-exact one; tests/test_sampled_trainers.py pins it) and timed separately.
usage: sampled_probe.py [N_DOCS] [TREES] [EVERY]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import pyref
from quickrank_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
trees = int(sys.argv[2]) if len(sys.argv) > 2 else 12
every = int(sys.argv[3]) if len(sys.argv) > 3 else 3
F, LEAVES, RANK, RANDOM = 136, 64, 0.3, 0.2
x, l, off = synth.make_dataset(n, F, n // 100, seed=5)
N, Q = len(l), len(off) - 1
q_of = np.repeat(np.arange(Q), np.diff(off).astype(np.int64))
qstart = off[:-1].astype(np.int64)[q_of]
pos = l > 0
nneg = np.bincount(q_of, weights=~pos, minlength=Q)
cround = lambda v: np.floor(v + np.float32(0.5)).astype(np.int64)   # std::round on non-negative floats (np.round is half-to-even)
n_top = cround(np.float32(RANK) * nneg.astype(np.float32))
n_rnd = cround(np.float32(RANDOM) * nneg.astype(np.float32))
rng = np.random.default_rng(0)


def draw(scores):
    neg = ~pos
    order = np.lexsort((-scores, ~neg, q_of))                 # per query: negatives first, by decreasing score
    rank = np.empty(N, np.int64)
    rank[order] = np.arange(N) - qstart[order]
    top = neg & (rank < n_top[q_of])
    rest = neg & ~top
    order = np.lexsort((rng.random(N), ~rest, q_of))
    rank[order] = np.arange(N) - qstart[order]
    return np.nonzero(pos | top | (rest & (rank < n_rnd[q_of])))[0]


out = dict(workload="LambdaMART on a document sample: %d docs x %d features x %d queries, %d leaves, %d trees, a new sample "
                    "every %d trees (all positives + %.0f%% top + %.0f%% random negatives per query)"
                    % (N, F, Q, LEAVES, trees, every, 100 * RANK, 100 * RANDOM))
t0 = time.time()
with api.Trainer(x, l, off, nleaves=LEAVES, minleafsupport=1) as full:
    t_create = time.time() - t0
    sm = full.sample_context(x, np.arange(N))
    t_draw = t_ctx = 0.0
    sizes, metric = [], 0.0
    full.evaluate_dataset()
    t1 = time.time()
    for m in range(trees):
        if m > 0 and m % every == 0:
            a = time.time()
            ids = draw(full.get_scores())
            b = time.time()
            sm.redraw(full, ids)
            c = time.time()
            t_draw += b - a
            t_ctx += c - b
            sizes.append(len(ids))
        sm.pull_scores(full)
        sm.compute_pseudoresponses()
        tree = sm.fit_regressor_on_gradient()
        full.apply_tree(tree, full.shrinkage)
        metric = full.evaluate_dataset()
    t_loop = time.time() - t1
    sm.close()
gpu_s = t_loop - t_draw
out["gpu"] = dict(trees_per_s=trees / gpu_s, loop_s=gpu_s, of_which_sample_contexts_s=t_ctx, numpy_draw_s=t_draw,
                  ctx_create_s=t_create, sample_sizes=sizes, ndcg=metric)
print("GPU: %.1f trees/s (%d trees in %.3f s, %.3f s of it loading %d new samples into the sample context; numpy draw %.2f s not counted), "
      "samples %s, NDCG@10 %.4f" % (trees / gpu_s, trees, gpu_s, t_ctx, len(sizes), t_draw, sizes, metric), flush=True)
if pyref.available():
    threads = pyref.set_threads(os.cpu_count())
    sel = dict(sampling_iterations=every, rank_factor=RANK, random_factor=RANDOM)
    with pyref.RefSession("LAMBDAMART-SELECTIVE", x, l, off, ntrees=trees, nleaves=LEAVES, minleafsupport=1, selective=sel) as s:
        t2 = time.time()
        s.learn()
        t3 = time.time()
        log = s.log()
        hist = s.metric_history()
    import re
    ref_sizes = [int(v) for v in re.findall(r"^Reducing training size from \d+ to (\d+)", log, flags=re.M)]
    tt = re.search(r"Training Time: ([0-9.]+)", log)
    train_s = float(tt.group(1)) if tt and float(tt.group(1)) > 0 else t3 - t2
    out["reference"] = dict(trees_per_s=trees / train_s, learn_s=t3 - t2, training_s=train_s, threads=threads,
                            sample_sizes=ref_sizes, ndcg=float(hist[-1]))
    print("reference on %d threads: %.2f trees/s (learn() %.2f s, of which training %.2f s), samples %s, NDCG@10 %.4f"
          % (threads, trees / train_s, t3 - t2, train_s, ref_sizes, hist[-1]), flush=True)
    out["same_sample_sizes"] = ref_sizes == sizes
print(json.dumps(out))
