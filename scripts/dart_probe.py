"""Development probe for BASELINE.json configs[4] (DART, rate_drop 0.1, 32 leaves, 2M docs x 500 features, NDCG@10):
the C++ host's `quicklearn --algo DART [--gpus N]` on a synthetic dataset of that shape, next to the unmodified reference's
Dart::learn on the host cores for a bounded sample (the first REF_DOCS documents, a few trees).

The dataset reaches quicklearn through the reader's binary cache (QR_SVML_CACHE=1: `<file>.qrb` = header, labels,
query offsets, row-major floats, keyed to the size and mtime of the text file next to it) — an 8 GB SVMLight text of
this shape would take longer to write than the run takes; the text file here is a one-line placeholder.
usage: dart_probe.py [N_DOCS] [FEATURES] [TREES] [GPUS] [REF_DOCS] [REF_TREES]"""
import json
import os
import re
import struct
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import pyref
from quickrank_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
n, f, trees, gpus, ref_docs, ref_trees = arg(1, 2000000), arg(2, 500), arg(3, 60), arg(4, 1), arg(5, 200000), arg(6, 6)
t0 = time.time()
x, l, off = synth.make_dataset(n, f, n // 100, seed=5)
print("synthetic %d x %d, %d queries: %.0f s" % (n, f, len(off) - 1, time.time() - t0), flush=True)
d = tempfile.mkdtemp()
txt = os.path.join(d, "train.txt")
open(txt, "w").write("0 qid:1 1:0\n")
st = os.stat(txt)
t0 = time.time()
with open(txt + ".qrb", "wb") as fh:
    fh.write(struct.pack("<8sQqqQQQ", b"QRB1", st.st_size, int(st.st_mtime_ns // 10**9), int(st.st_mtime_ns % 10**9),
                         n, f, len(off) - 1))
    fh.write(np.ascontiguousarray(l, np.float32).tobytes())
    fh.write(np.ascontiguousarray(off, np.uint64).tobytes())
    x.tofile(fh)
print("binary dataset written: %.0f s" % (time.time() - t0), flush=True)
out = dict(workload="DART rate_drop 0.1, 32 leaves, %d docs x %d features x %d queries, NDCG@10, %d trees, %d GPU(s)"
                    % (n, f, len(off) - 1, trees, gpus))
cmd = [os.path.join(ROOT, "host", "bin", "quicklearn"), "--algo", "DART", "--train", txt, "--num-trees", str(trees),
       "--num-leaves", "32", "--rate-drop", "0.1", "--end-after-rounds", "0", "--partial", "0", "--gpus", str(gpus)]
t0 = time.time()
r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, QR_SVML_CACHE="1"))
wall = time.time() - t0
assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
m = re.search(r"Training Time: ([0-9.]+)", r.stdout)
init = re.search(r"# Initialization: ([0-9.]+) s", r.stdout)
rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)", r.stdout, flags=re.M)
dropped = [int(v) for v in re.findall(r"(\d+) Dropped Trees", r.stdout)]
train_s = float(m.group(1))
out["gpu"] = dict(trees_per_s=trees / train_s, training_s=train_s, init_s=float(init.group(1)) if init else None,
                  process_wall_s=wall, last_ndcg=float(rows[-1][1]), dropped_trees_total=sum(dropped))
print("quicklearn --algo DART --gpus %d: %.1f trees/s (%d trees in %.2f s; initialisation %s s, whole process %.1f s), "
      "NDCG@10 %.4f, %d tree drops" % (gpus, trees / train_s, trees, train_s, init.group(1) if init else "?", wall,
                                        float(rows[-1][1]), sum(dropped)), flush=True)
if pyref.available() and ref_docs > 0:
    q = int(np.searchsorted(off, ref_docs, side="right")) - 1
    nd = int(off[q])
    threads = pyref.set_threads(os.cpu_count())
    with pyref.RefSession("DART", x[:nd], l[:nd], off[:q + 1], ntrees=ref_trees, nleaves=32, dart=dict(rate_drop=0.1)) as s:
        t0 = time.time()
        s.learn()
        t1 = time.time()
        log = s.log()
    tt = re.search(r"Training Time: ([0-9.]+)", log)
    ref_s = float(tt.group(1)) if tt and float(tt.group(1)) > 0 else t1 - t0
    out["reference"] = dict(docs=nd, trees=ref_trees, threads=threads, training_s=ref_s, learn_s=t1 - t0,
                            trees_per_s=ref_trees / ref_s, trees_per_s_scaled_to_full_size=ref_trees / ref_s * nd / n)
    print("reference Dart::learn on %d threads, first %d docs, %d trees: %.2f trees/s (%.2f s) -> %.3f trees/s if a tree "
          "costs %.1fx more on %d docs" % (threads, nd, ref_trees, ref_trees / ref_s, ref_s, ref_trees / ref_s * nd / n,
                                           n / nd, n), flush=True)
print(json.dumps(out))
for p in (txt, txt + ".qrb"):
    os.remove(p)
