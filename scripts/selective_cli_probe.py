"""Development probe: `quicklearn --algo LAMBDAMART-SELECTIVE` end to end at the config-2 shape —
the C++ host's own draw (host/src/sampled_trainers.cc) included — next to the unmodified reference's
LambdaMartSelective::learn on the same data and parameters.  The dataset reaches quicklearn through the reader's binary
cache, as in scripts/dart_probe.py.  usage: selective_cli_probe.py [N_DOCS] [TREES] [EVERY]"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
trees = int(sys.argv[2]) if len(sys.argv) > 2 else 30
every = int(sys.argv[3]) if len(sys.argv) > 3 else 5
F, LEAVES, RANK, RANDOM = 136, 64, 0.3, 0.2
x, l, off = synth.make_dataset(n, F, n // 100, seed=5)
d = tempfile.mkdtemp()
txt = os.path.join(d, "train.txt")
open(txt, "w").write("0 qid:1 1:0\n")
st = os.stat(txt)
with open(txt + ".qrb", "wb") as fh:
    fh.write(struct.pack("<8sQqqQQQ", b"QRB1", st.st_size, int(st.st_mtime_ns // 10**9), int(st.st_mtime_ns % 10**9),
                         n, F, len(off) - 1))
    fh.write(np.ascontiguousarray(l, np.float32).tobytes())
    fh.write(np.ascontiguousarray(off, np.uint64).tobytes())
    x.tofile(fh)
cmd = [os.path.join(ROOT, "host", "bin", "quicklearn"), "--algo", "LAMBDAMART-SELECTIVE", "--train", txt, "--num-trees", str(trees),
       "--num-leaves", str(LEAVES), "--end-after-rounds", "0", "--partial", "0", "--sampling-iterations", str(every),
       "--rank-sampling-factor", str(RANK), "--random-sampling-factor", str(RANDOM)]
t0 = time.time()
r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, QR_SVML_CACHE="1"))
wall = time.time() - t0
assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
pick = lambda text: [int(v) for v in re.findall(r"^Reducing training size from \d+ to (\d+)", text, flags=re.M)]
train_s = float(re.search(r"Training Time: ([0-9.]+)", r.stdout).group(1))
rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)", r.stdout, flags=re.M)
out = dict(workload="quicklearn --algo LAMBDAMART-SELECTIVE: %d docs x %d features, %d leaves, %d trees, a new sample every %d trees "
                    "(rank %.1f, random %.1f)" % (n, F, LEAVES, trees, every, RANK, RANDOM),
           gpu=dict(trees_per_s=trees / train_s, training_s=train_s, process_wall_s=wall, sample_sizes=pick(r.stdout),
                    ndcg=[float(v[1]) for v in rows]))
print("quicklearn: %.1f trees/s (%d trees in %.2f s, draws included), samples %s, NDCG@10 %.4f"
      % (trees / train_s, trees, train_s, pick(r.stdout), float(rows[-1][1])), flush=True)
if pyref.available():
    threads = pyref.set_threads(os.cpu_count())
    sel = dict(sampling_iterations=every, rank_factor=RANK, random_factor=RANDOM)
    with pyref.RefSession("LAMBDAMART-SELECTIVE", x, l, off, ntrees=trees, nleaves=LEAVES, minleafsupport=1, selective=sel) as s:
        s.learn()
        log = s.log()
        hist = [float(v) for v in s.metric_history()]
    ref_s = float(re.search(r"Training Time: ([0-9.]+)", log).group(1))
    out["reference"] = dict(trees_per_s=trees / ref_s, training_s=ref_s, threads=threads, sample_sizes=pick(log), ndcg=hist)
    out["same_sample_sizes"] = pick(log) == pick(r.stdout)
    out["max_abs_ndcg_difference"] = float(np.max(np.abs(np.array(hist) - np.array(out["gpu"]["ndcg"]))))
    print("reference on %d threads: %.2f trees/s (%.2f s), samples %s; same sample sizes: %s; max |NDCG difference| over the "
          "%d iterations: %.1e (4 decimals are printed)" % (threads, trees / ref_s, ref_s, pick(log), out["same_sample_sizes"],
                                                             trees, out["max_abs_ndcg_difference"]), flush=True)
print(json.dumps(out))
for p in (txt, txt + ".qrb"):
    os.remove(p)
