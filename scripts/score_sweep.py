"""Development probe: sweep the scorer's launch shape (threads per document, walks, documents per block)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quickrank_b200 import api, synth
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--f", type=int, default=700)
ap.add_argument("--trees", type=int, default=5000)
ap.add_argument("--leaves", type=int, default=64)
ap.add_argument("--shapes", default="auto,1:1,1:2,2:1,2:2,4:1,4:2,4:4")
a = ap.parse_args()
trees, weights = synth.random_ensemble(a.trees, a.leaves, a.f, seed=7)
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randint(0, 256, (a.n, a.f), device="cuda", generator=g, dtype=torch.int32).float() / 255.0).contiguous()
out = torch.zeros(a.n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
sc = api.Scorer(trees, weights, a.f)
xs = x[:2048].cpu().numpy()
from oracle import pyoracle as po
want = po.score_dataset(trees, weights, xs)
for shape in a.shapes.split(","):
    for k in ("QR_SCORE_TPD", "QR_SCORE_WALKS", "QR_SCORE_DOCS"):
        os.environ.pop(k, None)
    if shape != "auto":
        p = shape.split(":")
        os.environ["QR_SCORE_TPD"] = p[0]
        os.environ["QR_SCORE_WALKS"] = p[1]
        if len(p) > 2:
            os.environ["QR_SCORE_DOCS"] = p[2]
    out.zero_()
    sc.score_dataset_device(x.data_ptr(), a.n, out.data_ptr()); sc.sync()
    sc.timer_start()
    R = 3
    for i in range(R):
        sc.score_dataset_device(x.data_ptr(), a.n, out.data_ptr())
    ms = sc.timer_stop() / R
    ok = np.array_equal(out[:2048].cpu().numpy(), want)
    print("%-10s %d x %d, %d trees: %.2f ms = %.3e docs/s = %.3e doc*trees/s  exact=%s" % (shape, a.n, a.f, a.trees, ms, a.n / ms * 1e3, a.n * a.trees / ms * 1e3, ok), flush=True)
