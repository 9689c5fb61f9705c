"""The XML -> C code generators of the host layer (host/src/quickrank_host.cc: GenOpCond, GenOblivious),
i.e. the producers of the `double ranker(float *v)` that quickscore times
(reference: src/io/generate_conditional_operators.cc, src/io/generate_oblivious.cc, driver.cc:199-223).
CPU only: the generated source is compiled with gcc and called through ctypes; where oracle/_ref is
built, the text is also compared byte for byte with what the reference's own generators emit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref
from quickrank_b200 import modelxml, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QL = os.path.join(ROOT, "host", "bin", "quicklearn")

pytestmark = pytest.mark.skipif(not os.path.exists(QL), reason="host/bin/quicklearn not built")


def compile_ranker(code_path, so_path):
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-x", "c", code_path, "-o", so_path])
    lib = C.CDLL(so_path)
    lib.ranker.restype = C.c_double
    lib.ranker.argtypes = [C.POINTER(C.c_float)]
    return lib


def run_ranker(lib, x):
    x = np.ascontiguousarray(x, np.float32)
    return np.array([lib.ranker(x[i].ctypes.data_as(C.POINTER(C.c_float))) for i in range(len(x))])


def symmetric_tree(rng, depth, n_features):
    """An oblivious tree (one (feature, threshold) per level) in the flat pre-order layout."""
    feats = rng.integers(0, n_features, size=depth)
    thrs = (rng.integers(1, 255, size=depth) / 255.0).astype(np.float32)
    feature, threshold, left, right, value = [], [], [], [], []

    def build(level):
        idx = len(feature)
        feature.append(-1); threshold.append(0.0); left.append(-1); right.append(-1); value.append(0.0)
        if level == depth:
            value[idx] = float(rng.normal())
        else:
            feature[idx] = int(feats[level]); threshold[idx] = float(thrs[level])
            left[idx] = build(level + 1)
            right[idx] = build(level + 1)
        return idx

    build(0)
    return dict(feature=np.array(feature, np.int32), threshold=np.array(threshold, np.float32),
                left=np.array(left, np.int32), right=np.array(right, np.int32), value=np.array(value, np.float64))


def test_condop_generator(tmp_path):
    trees, weights = synth.random_ensemble(12, 9, 17, seed=4)
    weights = np.array([0.1, 0.25, 0.0625, 0.1, 1.0, 0.333, 0.1, 0.05, 0.1, 0.2, 0.1, 0.125])
    model, code = str(tmp_path / "m.xml"), str(tmp_path / "ranker.c")
    modelxml.write_model(model, trees, weights)
    out = subprocess.run([QL, "--model-file", model, "--code-file", code, "--generator", "condop"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    src = open(code).read()
    assert src.startswith("double ranker(float* v) {\n\treturn 0.0 ") and " ? " in src
    rng = np.random.default_rng(1)
    x = (rng.integers(0, 256, size=(400, 17)) / 255.0).astype(np.float32)
    got = run_ranker(compile_ranker(code, str(tmp_path / "ranker.so")), x)
    # the generator prints each weight as a float with 3 decimals (generate_conditional_operators.cc:104-105)
    w3 = np.array([float(np.float32(float("%.3f" % np.float32(w)))) for w in weights])
    want = po.score_dataset(trees, w3, x)
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    if pyref.available():
        ref_code = str(tmp_path / "ref_ranker.c")
        pyref.generate_code(model, ref_code, "condop")
        assert open(ref_code).read() == src


def test_oblivious_generator(tmp_path):
    rng = np.random.default_rng(9)
    depths = [3, 2, 3, 1, 3, 2, 3]
    trees = [symmetric_tree(rng, d, 11) for d in depths]
    weights = np.full(len(trees), 0.1)
    model, code = str(tmp_path / "m.xml"), str(tmp_path / "ranker.c")
    modelxml.write_model(model, trees, weights, algo="OBVLAMBDAMART", nleaves=8)
    text = open(model).read().replace("<leaves>8</leaves>\n", "<leaves>8</leaves>\n\t\t<depth>3</depth>\n")
    open(model, "w").write(text)
    out = subprocess.run([QL, "--model-file", model, "--code-file", code, "--generator", "oblivious"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    src = open(code).read()
    assert "#define N 7 // no. of trees" in src and "#define M 3 // max tree depth" in src and "leaf_id(" in src
    x = (rng.integers(0, 256, size=(300, 11)) / 255.0).astype(np.float32)
    got = run_ranker(compile_ranker(code, str(tmp_path / "ranker.so")), x)
    # weights travel as float with max_digits10 digits: exactly float(0.1)
    want = po.score_dataset(trees, np.full(len(trees), float(np.float32(0.1))), x)
    # (the generated code sums the trees in order of depth: same terms, different order)
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    if pyref.available():
        ref_code = str(tmp_path / "ref_ranker.c")
        pyref.generate_code(model, ref_code, "oblivious")
        assert open(ref_code).read() == src


def parse_vpred(text):
    """Independent reader of the VPRED input format: [(depth, {id: record})] per tree."""
    lines = text.split("\n")
    n = int(lines[0])
    trees, i = [], 1
    for _ in range(n):
        depth = int(lines[i]); i += 1
        recs = {}
        while lines[i] != "end":
            parts = lines[i].split()
            recs[int(parts[1])] = parts
            i += 1
        i += 1
        trees.append((depth, recs))
    return trees


def test_vpred_generator(tmp_path):
    """generate_vpred.cc:92-172: breadth-first records, leaf outputs multiplied by the model's shrinkage and
    printed with the stream's default 6 significant digits; unbalanced trees list shallow leaves as 'node'."""
    trees, weights = synth.random_ensemble(9, 10, 13, seed=6)             # unbalanced leaf-wise trees
    rng = np.random.default_rng(3)
    trees += [symmetric_tree(rng, d, 13) for d in (1, 2, 3)]              # complete trees
    weights = np.concatenate([weights, np.full(3, 0.1)])
    model, out_file = str(tmp_path / "m.xml"), str(tmp_path / "vpred.txt")
    modelxml.write_model(model, trees, weights, shrinkage=0.05)
    out = subprocess.run([QL, "--model-file", model, "--code-file", out_file, "--generator", "vpred"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    text = open(out_file).read()
    parsed = parse_vpred(text)
    assert len(parsed) == len(trees)
    for (depth, recs), t in zip(parsed, trees):
        # breadth-first ids over the flat pre-order tree
        order, parent, is_left = [0], {0: 4294967295}, {0: 0}
        for node in order:
            if t["feature"][node] >= 0:
                for child, lf in ((int(t["left"][node]), 1), (int(t["right"][node]), 0)):
                    parent[child] = order.index(node)
                    is_left[child] = lf
                    order.append(child)

        def node_depth(nd):
            return 0 if t["feature"][nd] < 0 else 1 + max(node_depth(int(t["left"][nd])), node_depth(int(t["right"][nd])))
        assert depth == node_depth(0)
        assert sorted(recs) == list(range(len(order)))
        for bfs_id, node in enumerate(order):
            r = recs[bfs_id]
            if t["feature"][node] >= 0:
                if bfs_id == 0:
                    assert r[0] == "root" and int(r[2]) == int(t["feature"][node])
                    assert np.float32(float(r[3])) == t["threshold"][node]
                else:
                    assert r[0] == "node" and int(r[2]) == parent[node] and int(r[3]) == int(t["feature"][node])
                    assert int(r[4]) == is_left[node] and np.float32(float(r[5])) == t["threshold"][node]
            else:
                want = "%g" % (0.05 * float("%.17g" % t["value"][node]))
                if bfs_id >= 2 ** depth - 1:
                    assert r[0] == "leaf" and int(r[2]) == parent[node] and int(r[3]) == is_left[node] and r[4] == want
                else:
                    assert r[0] == "node" and int(r[2]) == parent[node] and int(r[4]) == is_left[node] and r[5] == want
    if pyref.available():
        ref_file = str(tmp_path / "ref_vpred.txt")
        pyref.generate_code(model, ref_file, "vpred")
        assert open(ref_file).read() == text
