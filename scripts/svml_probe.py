"""Development probe, CPU only: the host SVMLight reader (all cores, host/bin/svml_check) against
the unmodified reference's io::Svml::read_horizontal (src/io/svml.cc:38-161, through oracle/_ref) on one synthetic
MSLR-shaped text file; same parsed arrays (checksums).  usage: svml_probe.py [N_DOCS] [FEATURES]"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 136
rng = np.random.default_rng(1)
x = np.round(rng.random((n, f)), 6).astype(np.float32)
labels = rng.integers(0, 5, size=n)
qid = np.arange(n) // 100 + 1
path = os.path.join(tempfile.mkdtemp(), "probe.txt")
t0 = time.time()
with open(path, "w") as fh:
    cols = ["%d:" % (j + 1) for j in range(f)]
    for i in range(n):
        fh.write("%d qid:%d %s\n" % (labels[i], qid[i], " ".join(c + ("%.6g" % v) for c, v in zip(cols, x[i]))))
size = os.path.getsize(path)
print("wrote %s: %.0f MB in %.0f s" % (path, size / 1e6, time.time() - t0), flush=True)
out = {"file_mb": size / 1e6, "docs": n, "features": f}
for threads in (os.cpu_count(), 1):
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "host", "bin", "svml_check"), path], capture_output=True, text=True,
                       env=dict(os.environ, QR_SVML_THREADS=str(threads)))
    dt = time.time() - t0
    assert r.returncode == 0, r.stderr
    ours = r.stdout.split()
    out["host_reader_%d_threads" % threads] = {"seconds": dt, "mb_per_s": size / 1e6 / dt}
    print("host reader, %d threads: %.2f s (%.0f MB/s, process start and checksums included)" % (threads, dt, size / 1e6 / dt), flush=True)
shape, sums, sec = pyref.read_svml(path)
out["reference_reader"] = {"seconds": sec, "mb_per_s": size / 1e6 / sec}
out["same_arrays"] = [int(v) for v in ours[:3]] == list(shape) and tuple(int(v, 16) for v in ours[3:6]) == sums
print("reference reader: %.2f s (%.0f MB/s); same arrays: %s" % (sec, size / 1e6 / sec, out["same_arrays"]))
print(json.dumps(out))
os.remove(path)
