"""Development probe: ms/tree and growth rounds along a long run (host- vs device-driven rounds)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quickrank_b200 import api, synth
trees = int(sys.argv[1]) if len(sys.argv) > 1 else 300
x, l, off = synth.make_dataset(1000000, 136, 10000, seed=20260102)
tr = api.Trainer(x, l, off, algo="LAMBDAMART", nleaves=64, nthresholds=0, cutoff=10, hist_mode=0)
t0 = time.time(); rs = []
for i in range(trees):
    tr.boost_iteration(want_tree=False, want_metric=True)
    rs.append(tr.last_tree_rounds()[0])
    if i % 50 == 49:
        dt = (time.time() - t0) / 50
        print("trees %4d-%4d: %.3f ms/tree, rounds/tree %.1f, beta %.2f" % (i - 49, i, dt * 1e3, np.mean(rs), tr.last_tree_rounds()[1]), flush=True)
        t0 = time.time(); rs = []
