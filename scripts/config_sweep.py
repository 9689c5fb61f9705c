"""Development probe: the other BASELINE.json configs at reduced size on one GPU (shape checks + rough rates)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quickrank_b200 import api, synth

def train(name, n, f, q, trees, **kw):
    x, l, off = synth.make_dataset(n, f, q, seed=20260100)
    t0 = time.time()
    tr = api.Trainer(x, l, off, nthresholds=0, cutoff=10, hist_mode=api.HIST_FAST, **kw)
    init = time.time() - t0
    for _ in range(3):
        tr.boost_iteration(want_tree=False)
    t0 = time.time()
    m = None
    for _ in range(trees):
        _, m = tr.boost_iteration(want_tree=False)
    dt = (time.time() - t0) / trees
    print("%-46s N=%d F=%d: init %.2fs, %.2f ms/tree = %.1f trees/s, NDCG@10 %.4f, rounds %s" %
          (name, n, f, init, dt * 1e3, 1 / dt, m, tr.last_tree_rounds()[0]), flush=True)
    tr.close()
    return x

train("config 4 shape: OBVLAMBDAMART depth 6", 1000000, 220, 10000, 20, algo="OBVLAMBDAMART", treedepth=6)
x = train("config 5 shape: LAMBDAMART 32 leaves (DART base)", 500000, 500, 5000, 20, algo="LAMBDAMART", nleaves=32)
train("MART 64 leaves", 1000000, 136, 10000, 20, algo="MART", nleaves=64)
# config 3 shape: 5000-tree ensemble x 700 features
import torch
n, f = 1000000, 700
rng = np.random.default_rng(3)
xs = (rng.integers(0, 256, size=(n, f), dtype=np.uint8).astype(np.float32) / 255.0)
trees, w = synth.random_ensemble(5000, 64, f, seed=11)
sc = api.Scorer(trees, w, f)
xd = torch.from_numpy(xs).cuda()
out = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    sc.score_dataset_device(xd.data_ptr(), n, out.data_ptr())
sc.sync()
sc.timer_start()
for _ in range(3):
    sc.score_dataset_device(xd.data_ptr(), n, out.data_ptr())
ms = sc.timer_stop() / 3
print("config 3 shape: 5000 trees x 64 leaves, N=%d F=%d: %.2f ms/pass = %.3g docs/s (%.3g doc-trees/s)" %
      (n, f, ms, n / ms * 1e3, n * 5000 / ms * 1e3), flush=True)
