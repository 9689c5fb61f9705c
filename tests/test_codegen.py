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
