"""Multi-GPU plumbing of the C++ host (host/): `quicklearn --gpus N` shards the queries with the same
rule as quickrank_b200/sharding.py, hands the NCCL communicator id to the other processes over TCP and
grows the same model as one GPU."""
import os
import subprocess

import numpy as np
import pytest

import qr_testlib as common
from quickrank_b200.sharding import query_shards
from test_host_cli import QL, write_svml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "host", "bin", "shard_check")


@pytest.mark.parametrize("seed", range(6))
def test_host_query_shards_equal_the_python_rule(seed):
    rng = np.random.default_rng(seed)
    nq = int(rng.integers(8, 200))
    lens = rng.integers(1, 400, size=nq)
    if seed % 2:
        lens[rng.integers(0, nq)] = 20000        # one huge query
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 8):
        out = subprocess.run([SC, "shards", str(world)] + [str(int(o)) for o in off], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = [tuple(int(v) for v in line.split()) for line in out.stdout.split("\n") if line]
        assert got == [tuple(s) for s in query_shards(off, world)]


@pytest.mark.parametrize("addr", ["127.0.0.1", "localhost"])
def test_communicator_id_rendezvous(addr):
    port = 21000 + (os.getpid() + len(addr)) % 20000
    out = subprocess.run([SC, "rendezvous", "4", str(port), addr], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr + out.stdout


def test_rendezvous_ignores_strangers():
    """Connections that do not introduce themselves (a port scanner, another job on the same port) do not use up one
    of the world-1 hand-offs and do not receive the communicator id."""
    import socket
    import threading
    import time
    port = 23000 + os.getpid() % 20000
    got = []

    def stranger():
        deadline = time.time() + 10
        while time.time() < deadline:
            try:
                with socket.create_connection(("127.0.0.1", port), timeout=1) as c:
                    c.sendall(b"GET / HTTP/1.0\r\n\r\n")
                    c.settimeout(2)
                    try:
                        got.append(c.recv(256))
                    except OSError:
                        got.append(b"")
                return
            except OSError:
                time.sleep(0.01)

    threads = [threading.Thread(target=stranger) for _ in range(3)]
    for t in threads:
        t.start()
    out = subprocess.run([SC, "rendezvous", "3", str(port), "127.0.0.1"], capture_output=True, text=True, timeout=120)
    for t in threads:
        t.join()
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr + out.stdout
    assert all(len(g) == 0 for g in got), got


@pytest.mark.gpu
def test_quicklearn_on_two_gpus_grows_the_single_gpu_model(tmp_path):
    from quickrank_b200 import api, modelxml
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, l, off = common.dataset(n=6000, f=12, q=60, seed=8)
    tr = str(tmp_path / "train.txt")
    write_svml(tr, x, l, off)
    models, tables = [], []
    for gpus in (1, 2):
        model = str(tmp_path / ("model%d.xml" % gpus))
        cmd = [QL, "--algo", "LAMBDAMART", "--train", tr, "--num-trees", "6", "--num-leaves", "8", "--model-out", model,
               "--min-leaf-support", "20", "--gpus", str(gpus)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr + out.stdout
        models.append(modelxml.read_model(model))
        tables.append([line for line in out.stdout.split("\n") if line[:8].strip().isdigit()])
    assert "training on 2 GPUs" in out.stdout
    # same stdout table (metrics are printed with 4 decimals) and same trees
    assert tables[0] == tables[1] and len(tables[0]) == 6
    (_i1, t1, w1), (_i2, t2, w2) = models
    assert len(t1) == len(t2) == 6 and np.array_equal(w1, w2)
    for a, b in zip(t1, t2):
        for k in ("feature", "threshold", "left", "right"):
            assert np.array_equal(a[k], b[k]), k
        lv = a["feature"] < 0
        assert np.array_equal(a["value"][lv], b["value"][lv])   # exact integer leaf sums: the same model on any number of GPUs


@pytest.mark.gpu
def test_dart_on_two_gpus_follows_the_single_gpu_run(tmp_path):
    """DART (BASELINE config 5 is 4 GPUs): the host logic (rand() stream, drop sets, normalisation) is replicated on
    every rank, the passes over the documents (Dart::update_modelscores, dart.cc:634-687) run on each rank's shard, the
    trees come from all-reduced integer histograms: same table, same trees, same weights as one GPU."""
    from quickrank_b200 import api, modelxml
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, l, off = common.dataset(n=6000, f=12, q=60, seed=8)
    tr = str(tmp_path / "train.txt")
    write_svml(tr, x, l, off)
    models, tables = [], []
    for gpus in (1, 2):
        model = str(tmp_path / ("dart%d.xml" % gpus))
        cmd = [QL, "--algo", "DART", "--train", tr, "--num-trees", "12", "--num-leaves", "8", "--model-out", model,
               "--min-leaf-support", "20", "--rate-drop", "0.3", "--gpus", str(gpus)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr + out.stdout
        models.append(modelxml.read_model(model))
        tables.append([line for line in out.stdout.split("\n") if line[:8].strip().isdigit()])
    assert tables[0] == tables[1] and len(tables[0]) >= 12
    (_i1, t1, w1), (_i2, t2, w2) = models
    assert len(t1) == len(t2) and np.array_equal(w1, w2)
    for a, b in zip(t1, t2):
        for k in ("feature", "threshold", "left", "right"):
            assert np.array_equal(a[k], b[k]), k


def test_inherited_launcher_variables_alone_do_not_shard(tmp_path):
    """RANK / WORLD_SIZE in the environment (a process started under some launcher) must not switch sharding on:
    it is opt-in through --gpus.  Without a GPU the run still ends at the reference-style error, not in a rendezvous."""
    tr = str(tmp_path / "t.txt")
    open(tr, "w").write("1 qid:1 1:0.5\n0 qid:1 1:0.25\n")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([QL, "--train", tr, "--num-trees", "1"], capture_output=True, text=True, timeout=60, env=env)
    assert "training on" not in out.stdout and "cannot reach rank 0" not in out.stderr
    mismatch = subprocess.run([QL, "--train", tr, "--num-trees", "1", "--gpus", "4"], capture_output=True, text=True,
                              timeout=60, env=env)
    assert mismatch.returncode != 0 and "WORLD_SIZE is 2" in mismatch.stderr
