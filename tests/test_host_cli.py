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


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libqr_ref.so not built")
@pytest.mark.parametrize("dart,cli,n,minls,full", [
    (dict(rate_drop=0.2), ["--rate-drop", "0.2"], 3000, 20, True),
    (dict(rate_drop=0.2), ["--rate-drop", "0.2"], 20000, 1000, False),
    (dict(rate_drop=0.3, normalize_type=3), ["--rate-drop", "0.3", "--normalize-type", "FOREST"], 3000, 20, False),
    (dict(rate_drop=2.0, skip_drop=0.3, normalize_type=2),
     ["--rate-drop", "2", "--skip-drop", "0.3", "--normalize-type", "WEIGHTED"], 3000, 20, False),
    (dict(rate_drop=0.15, normalize_type=6), ["--rate-drop", "0.15", "--normalize-type", "TREE_BOOST3"], 3000, 20, False),
])
def test_dart_matches_the_reference_learn_loop(tmp_path, dart, cli, n, minls, full):
    """Dart::learn (dart.cc:172-602): dropout selection (std::rand stream), subtraction / re-addition of
    the dropped trees, weight normalisation and best-model bookkeeping on the host; pseudo-responses,
    tree fit, per-tree score updates and NDCG on the GPU, against the reference's own loop.

    What can be compared: subtracting a tree leaves documents that differ only in that tree's leaf TIED in
    exact arithmetic and one ulp apart in floating point, in an order set by the last bits of the leaf
    outputs.  Those bits differ between CUDA's and glibc's exp(), so after the first dropout the two runs
    rank such documents differently, get different lambdas and grow different trees: the reference's
    trajectory is a function of its own rounding noise.  Checked therefore:
      * everything that does not depend on that noise, exactly: number of dropped trees per iteration
        (rand stream), the final ensemble WEIGHTS (normalisation arithmetic, 1e-12), the metric trajectory
        up to the first dropout;
      * the saved model is self-consistent (GPU scores = sum of weight x tree) and of the same quality;
      * `full`: a case where no such near-tie matters — the whole trajectory and the final ensemble."""
    from quickrank_b200 import modelxml
    x, l, off = common.dataset(n=n, f=12, q=n // 100, seed=8)
    ntrees = 25
    with pyref.RefSession("DART", x, l, off, ntrees=ntrees, nleaves=8, minleafsupport=minls, dart=dart) as s:
        s.learn()
        want_metric = s.metric_history()
        ref_model = str(tmp_path / "ref_dart.xml")
        s.save_model(ref_model)
    tr, model, scores = str(tmp_path / "train.txt"), str(tmp_path / "dart.xml"), str(tmp_path / "scores.txt")
    write_svml(tr, x, l, off)
    cmd = [QL, "--algo", "DART", "--train", tr, "--test", tr, "--scores", scores, "--num-trees", str(ntrees),
           "--num-leaves", "8", "--min-leaf-support", str(minls), "--model-out", model, "--hist-mode", "reference",
           "--end-after-rounds", "0", "--partial", "0"] + cli
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)[ *]*\t\[ ([0-9.]+) - ([0-9.]+) - ([0-9.]+) \|.*?(\d+) Dropped Trees", out.stdout, flags=re.M)
    assert len(rows) == len(want_metric), out.stdout
    got_metric = np.array([float(r[1]) for r in rows])
    dropped = np.array([int(r[5]) for r in rows])
    assert dropped.sum() > 0, "no dropout happened: the test would not exercise DART"
    first = int(np.argmax(dropped > 0))          # first iteration with a dropout (0-based)
    upto = len(rows) if full else first
    assert np.max(np.abs(got_metric[:upto] - want_metric[:upto])) <= 6e-5, "\n".join(
        "%s ours %.4f ref %.6f dropped %s" % (r[0], g, w, r[5]) for r, g, w in zip(rows, got_metric, want_metric))
    assert abs(got_metric[-1] - want_metric[-1]) <= 0.03   # same quality at the end of the run
    _gi, got_trees, got_w = modelxml.read_model(model)
    _wi, want_trees, want_w = modelxml.read_model(ref_model)
    assert len(got_trees) == len(want_trees)
    assert np.allclose(got_w, want_w, rtol=1e-12, atol=0)   # rand stream + normalisation arithmetic

    def leaf_of(t, data):
        node = np.zeros(len(data), np.int64)
        while True:
            f = t["feature"][node]
            act = f >= 0
            if not act.any():
                return node
            idx = np.nonzero(act)[0]
            left = data[idx, f[idx]] <= t["threshold"][node[idx]]
            node[idx] = np.where(left, t["left"][node[idx]], t["right"][node[idx]])

    for a, b in list(zip(got_trees, want_trees))[:len(got_trees) if full else first]:
        assert np.array_equal(a["feature"], b["feature"])
        assert np.array_equal(a["left"], b["left"]) and np.array_equal(a["right"], b["right"])
        if not np.array_equal(a["threshold"], b["threshold"]):
            # a different threshold is tolerated only when it cuts the training documents into the very
            # same sets (an empty-bin plateau: an exact tie the reference breaks by rounding noise)
            assert np.array_equal(leaf_of(a, x), leaf_of(b, x))
        lv = a["feature"] < 0
        assert np.allclose(a["value"][lv], b["value"][lv], rtol=1e-5, atol=0)
    # the saved DART model reloads in stock QuickRank (type dispatch) and scores like the GPU did
    assert "<type>DART</type>" in open(model).read()
    ref_scores_of_our_model = pyref.score_with_model(model, x)
    assert np.allclose(np.loadtxt(scores), ref_scores_of_our_model, rtol=1e-9, atol=1e-12)
    if full:
        assert np.allclose(ref_scores_of_our_model, pyref.score_with_model(ref_model, x), rtol=1e-5, atol=1e-12)
