"""Development probe: ensemble scoring throughput with device-resident documents."""
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
a = ap.parse_args()
t0 = time.time()
trees, weights = synth.random_ensemble(a.trees, a.leaves, a.f, seed=7)
print("model %.1fs" % (time.time() - t0), flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randint(0, 256, (a.n, a.f), device="cuda", generator=g, dtype=torch.int32).float() / 255.0).contiguous()
out = torch.zeros(a.n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
sc = api.Scorer(trees, weights, a.f)
for i in range(2):
    sc.score_dataset_device(x.data_ptr(), a.n, out.data_ptr()); sc.sync()
sc.timer_start()
R = 3
for i in range(R):
    sc.score_dataset_device(x.data_ptr(), a.n, out.data_ptr())
ms = sc.timer_stop() / R
print("%d docs x %d feat, %d trees x %d leaves: %.2f ms/pass = %.3e docs/s = %.3e doc*trees/s" % (a.n, a.f, a.trees, a.leaves, ms, a.n / ms * 1e3, a.n * a.trees / ms * 1e3))
# spot check against a host walk of a few documents
xs = x[:64].cpu().numpy()
from oracle import pyoracle as po
want = po.score_dataset(trees, weights, xs)
print("spot check bit-exact:", np.array_equal(out[:64].cpu().numpy(), want))
