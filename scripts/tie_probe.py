"""Development probe: fraction of queries whose scores contain ties, per boosting iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quickrank_b200 import api, synth
x, l, off = synth.make_dataset(200000, 136, 2000, seed=20260102)
tr = api.Trainer(x, l, off, nleaves=64)
for it in range(40):
    tr.boost_iteration(want_tree=False, want_metric=False)
    if it in (0, 1, 2, 3, 5, 8, 12, 16, 20, 25, 30, 39):
        s = tr.get_scores()
        ties = 0
        for q in range(len(off) - 1):
            seg = s[int(off[q]):int(off[q + 1])]
            if len(np.unique(seg)) < len(seg):
                ties += 1
        print("iter %2d: %5.1f%% of queries have tied scores" % (it, 100.0 * ties / (len(off) - 1)), flush=True)
