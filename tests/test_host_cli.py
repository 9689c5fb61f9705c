"""The C++ host layer (host/): quicklearn / quickscore keep the reference's CLI, SVMLight input, stdout
table and XML model format while the hot path runs on the GPU.  Interop both ways with the unmodified
reference (where oracle/_ref exists): a GPU-trained model loads and scores in stock QuickRank, and a
model written by stock QuickRank scores on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import qr_testlib as common
from oracle import pyoracle as po
from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QL = os.path.join(ROOT, "host", "bin", "quicklearn")
QS = os.path.join(ROOT, "host", "bin", "quickscore")


def write_svml(path, x, l, off):
    with open(path, "w") as f:
        for q in range(len(off) - 1):
            for i in range(int(off[q]), int(off[q + 1])):
                feats = " ".join("%d:%.9g" % (j + 1, x[i, j]) for j in range(x.shape[1]))
                f.write("%d qid:%d %s\n" % (int(l[i]), q + 1, feats))


def test_cli_without_gpu_fails_like_the_reference():
    """No CPU fallback: message on stderr and EXIT_FAILURE (the reference's error convention)."""
    from quickrank_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert os.path.exists(QL), "build host/ first"
    out = subprocess.run([QL, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--algo" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("algo,extra", [("LAMBDAMART", ["--num-leaves", "8"]), ("OBVLAMBDAMART", ["--tree-depth", "3"]),
                                        ("MART", ["--num-leaves", "6"])])
def test_quicklearn_trains_saves_and_scores(tmp_path, algo, extra):
    x, l, off = common.dataset(n=3000, f=12, q=30, seed=8)
    xv, lv, offv = common.dataset(n=1500, f=12, q=15, seed=9)
    tr, va = str(tmp_path / "train.txt"), str(tmp_path / "valid.txt")
    write_svml(tr, x, l, off)
    write_svml(va, xv, lv, offv)
    model, scores = str(tmp_path / "model.xml"), str(tmp_path / "scores.txt")
    cmd = [QL, "--algo", algo, "--train", tr, "--valid", va, "--test", va, "--num-trees", "8", "--model-out", model,
           "--scores", scores, "--hist-mode", "reference", "--min-leaf-support", "40"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    # the stdout table of Mart::learn: "iter. training validation" rows
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)\s+([0-9.]+)", out.stdout, flags=re.M)
    assert len(rows) == 8
    # training metric column equals the oracle's trajectory (4 decimals are printed)
    depth = 3 if algo.startswith("OBV") else 0
    leaves = 8 if algo == "LAMBDAMART" else 6
    trees, metric, _ = po.train(algo, x, l, off, 8, nleaves=leaves, depth=depth, minls=40, cutoff=10)
    for (it, mt, _mv), want in zip(rows, metric):
        assert abs(float(mt) - want) <= 6e-5, (it, mt, want)
    got_scores = np.loadtxt(scores)
    # validation picks the best iteration: the saved ensemble is a prefix of the oracle's trees
    xml = open(model).read()
    ntrees = len(re.findall(r"<tree id=", xml))
    assert 1 <= ntrees <= 8
    want_scores = po.score_dataset(trees[:ntrees], [0.1] * ntrees, xv)
    assert np.max(np.abs(got_scores - want_scores)) <= 1e-12 * max(1.0, np.max(np.abs(want_scores)))
    m = re.search(r"NDCG@10 on test data = ([0-9.]+)", out.stdout)
    assert m and abs(float(m.group(1)) - po.ndcg_dataset(lv, want_scores, offv, 10)) <= 6e-5
    if pyref.available():   # the GPU-trained model loads in stock QuickRank and scores identically
        assert np.array_equal(pyref.score_with_model(model, xv), want_scores)


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libqr_ref.so not built")
def test_quickscore_scores_a_stock_quickrank_model(tmp_path):
    x, l, off = common.dataset(n=2000, f=10, q=20, seed=12)
    with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=6, nleaves=8) as s:
        s.learn()
        model = str(tmp_path / "ref_model.xml")
        s.save_model(model)
    data, scores = str(tmp_path / "data.txt"), str(tmp_path / "scores.txt")
    write_svml(data, x, l, off)
    out = subprocess.run([QS, "-d", data, "-m", model, "-r", "2", "-s", scores], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Avg.    Doc. scoring time" in out.stdout
    want = pyref.score_with_model(model, x)
    assert np.max(np.abs(np.loadtxt(scores) - want)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
