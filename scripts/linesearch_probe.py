"""Development probe: one line-search iteration over a partial-score-like matrix, GPU vs the unmodified reference on
the host cores.  usage: linesearch_probe.py [N_DOCS] [TREES]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quickrank_b200 import api, synth
from quickrank_b200.linesearch import LineSearch
from oracle import pyref
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
t = int(sys.argv[2]) if len(sys.argv) > 2 else 50
x, l, off = synth.make_dataset(n, t, n // 100, seed=5, gridded=False)
x = (x - 0.5).astype(np.float32) * 0.1
kw = dict(num_points=20, max_iterations=1, window_size=1.0)
t0 = time.time()
with api.LineSearchDevice(x, l, off, cutoff=10) as dev:
    t1 = time.time()
    ls = LineSearch(**kw)
    got = ls.learn(dev)
    t2 = time.time()
    launches = dev.launch_count()
print("GPU: create %.2f s, one iteration over %d columns x %d docs: %.2f s (%d launches), NDCG %.4f"
      % (t1 - t0, t, n, t2 - t1, launches, ls.metric_on_training), flush=True)
if pyref.available():
    pyref.set_threads(os.cpu_count())
    t3 = time.time()
    want = pyref.linesearch(x, l, off, cutoff=10, **kw)
    t4 = time.time()
    print("reference on %d threads: %.2f s; weights equal: %s" % (os.cpu_count(), t4 - t3, np.array_equal(got, want)))
