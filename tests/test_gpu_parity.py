"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs.  Bit-exact for indices (bins, ranks, split feature / threshold index, doc->leaf); floating
point within the tolerance written at each assert (north_star: 1e-5 relative on leaf outputs and
NDCG@k; most checks here are far tighter, and exact where the arithmetic order is replicated)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from quickrank_b200 import api, synth
import qr_testlib as common

pytestmark = pytest.mark.gpu

REL = 1e-5  # north_star tolerance for leaf outputs / NDCG


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0


@pytest.mark.parametrize("gridded,nthr", [(True, 0), (False, 0), (False, 32), (True, 300)])
def test_binning_matches_oracle(gridded, nthr):
    x, l, off = common.dataset(n=2500, f=19, q=25, gridded=gridded)
    ob = po.Binning(np.ascontiguousarray(x.T), nthr)
    with api.Trainer(x, l, off, nthresholds=nthr) as tr:
        obins = ob.bins()
        for f in range(x.shape[1]):
            assert np.array_equal(tr.thresholds(f), ob.thresholds(f)), "thresholds of feature %d" % f
            assert np.array_equal(tr.get_bins(f), obins[f]), "bins of feature %d" % f


def test_binning_colmajor_entry_point():
    x, l, off = common.dataset(n=1200, f=9, q=12)
    ob = po.Binning(np.ascontiguousarray(x.T), 0)
    with api.Trainer(np.ascontiguousarray(x.T), l, off, layout="colmajor") as tr:
        for f in range(x.shape[1]):
            assert np.array_equal(tr.get_bins(f), ob.bins()[f])


@pytest.mark.parametrize("levels", [1, 3, 9, 100000])
def test_ranking_reproduces_std_sort(levels):
    rng = np.random.default_rng(levels)
    x, l, off = common.dataset(n=4000, f=5, q=45, qlen=(1, 200))
    s = rng.integers(0, levels, size=len(l)).astype(np.float64) / 7.0
    with api.Trainer(x, l, off) as tr:
        tr.set_scores(s)
        got = tr.get_ranking()
    for q in range(len(off) - 1):
        a, b = int(off[q]), int(off[q + 1])
        assert np.array_equal(got[a:b], po.sort_desc(s[a:b])), "query %d (n=%d)" % (q, b - a)


@pytest.mark.parametrize("cutoff", [10, 3, 0])
@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_ndcg_matches_oracle(cutoff, mode):
    rng = np.random.default_rng(5)
    x, l, off = common.dataset(n=3000, f=5, q=40, qlen=(1, 150))
    s = np.round(rng.normal(size=len(l)), 1)
    want = po.ndcg_dataset(l, s, off, cutoff)
    with api.Trainer(x, l, off, cutoff=cutoff, hist_mode=mode) as tr:
        tr.set_scores(s)
        got = tr.evaluate_dataset()
    if mode == api.HIST_REFERENCE:
        assert got == want
    else:
        assert abs(got - want) <= 1e-12 * abs(want)


@pytest.mark.parametrize("cutoff", [10, 5, 40, 0])
def test_lambdas_match_oracle(cutoff):
    rng = np.random.default_rng(cutoff + 1)
    x, l, off = common.dataset(n=2500, f=5, q=30, qlen=(1, 120))
    with api.Trainer(x, l, off, cutoff=cutoff) as tr:
        # all-zero scores: exp(0) == 1 exactly, so the result is bit-identical to the reference
        tr.set_scores(np.zeros(len(l)))
        tr.compute_pseudoresponses()
        lam, w = tr.get_pseudoresponses()
        olam, ow = po.lambdas(np.zeros(len(l)), l, off, cutoff)
        assert np.array_equal(lam, olam) and np.array_equal(w, ow)
        # tie-heavy and generic scores: same order of accumulation, exp() may differ in the last ulp
        for s in (common.tie_heavy_scores(len(l), rng), rng.normal(size=len(l))):
            tr.set_scores(s)
            tr.compute_pseudoresponses()
            lam, w = tr.get_pseudoresponses()
            olam, ow = po.lambdas(s, l, off, cutoff)
            scale = np.max(np.abs(olam))
            assert np.max(np.abs(lam - olam)) <= 1e-13 * scale
            assert np.max(np.abs(w - ow)) <= 1e-13 * max(1e-300, np.max(np.abs(ow)))


def test_mart_pseudoresponses():
    rng = np.random.default_rng(2)
    x, l, off = common.dataset(n=1500, f=5, q=15)
    s = rng.normal(size=len(l))
    with api.Trainer(x, l, off, algo="MART") as tr:
        tr.set_scores(s)
        tr.compute_pseudoresponses()
        lam, _ = tr.get_pseudoresponses()
    assert np.array_equal(lam, l.astype(np.float64) - s)


def _gradients(x, l, off, seed, cutoff=10):
    rng = np.random.default_rng(seed)
    s = rng.normal(size=len(l)) * 0.3
    return po.lambdas(s, l, off, cutoff)


@pytest.mark.parametrize("cfg", [
    dict(n=3000, f=20, q=30, gridded=True, nthr=0, leaves=12, minls=1),
    dict(n=3000, f=20, q=30, gridded=False, nthr=0, leaves=8, minls=1),     # u16 bins, one per value
    dict(n=5000, f=33, q=50, gridded=False, nthr=64, leaves=32, minls=5),
    dict(n=800, f=7, q=9, gridded=True, nthr=0, leaves=64, minls=1),        # many tiny nodes: tie rules
])
def test_fit_tree_reference_mode_is_bit_exact(cfg):
    x, l, off = common.dataset(n=cfg["n"], f=cfg["f"], q=cfg["q"], gridded=cfg["gridded"])
    lam, w = _gradients(x, l, off, 3)
    ob = po.Binning(np.ascontiguousarray(x.T), cfg["nthr"])
    want = ob.fit_tree(lam, w, nleaves=cfg["leaves"], minls=cfg["minls"])
    with api.Trainer(x, l, off, nleaves=cfg["leaves"], minleafsupport=cfg["minls"],
                     nthresholds=cfg["nthr"], hist_mode=api.HIST_REFERENCE) as tr:
        tr.set_pseudoresponses(lam, w)
        got = tr.fit_regressor_on_gradient()
        leaf = tr.get_leaf_assignment()
    assert common.same_structure(got, want), common.describe_tree_diff(got, want)
    assert np.array_equal(got["value"], want["value"]), common.describe_tree_diff(got, want)
    assert np.array_equal(got["count"], want["count"])
    assert np.array_equal(got["deviance"][got["feature"] >= 0], want["deviance"][want["feature"] >= 0])
    assert np.array_equal(leaf, want["leaf_of_doc"])


@pytest.mark.parametrize("cfg", [
    dict(n=20000, f=30, q=200, gridded=True, nthr=0, leaves=32, minls=1),     # 8-bit bins
    dict(n=6000, f=20, q=60, gridded=False, nthr=0, leaves=16, minls=3),      # one bin per distinct value: 16-bit bins
    dict(n=9000, f=17, q=90, gridded=True, nthr=0, leaves=64, minls=1, algo="MART"),
])
def test_reference_mode_large_node_path_is_bit_exact(cfg, monkeypatch):
    """REFERENCE mode accumulates large nodes by walking per-feature lists of documents sorted by (bin, document)
    (hist_exact_walk_kernel) and long squares sums by the parallel ordered scheme: with QR_EXACT_WALK_MIN=1 every
    built child above N/16 documents takes that path; the tree must still be the reference's bit for bit."""
    monkeypatch.setenv("QR_EXACT_WALK_MIN", "1")
    x, l, off = common.dataset(n=cfg["n"], f=cfg["f"], q=cfg["q"], gridded=cfg["gridded"])
    algo = cfg.get("algo", "LAMBDAMART")
    lam, w = _gradients(x, l, off, 3)
    if algo == "MART":
        w = None
    ob = po.Binning(np.ascontiguousarray(x.T), cfg["nthr"])
    want = ob.fit_tree(lam, w, nleaves=cfg["leaves"], minls=cfg["minls"])
    with api.Trainer(x, l, off, algo=algo, nleaves=cfg["leaves"], minleafsupport=cfg["minls"],
                     nthresholds=cfg["nthr"], hist_mode=api.HIST_REFERENCE) as tr:
        tr.set_pseudoresponses(lam, w)
        got = tr.fit_regressor_on_gradient()
        leaf = tr.get_leaf_assignment()
    assert common.same_structure(got, want), common.describe_tree_diff(got, want)
    assert np.array_equal(got["value"], want["value"]), common.describe_tree_diff(got, want)
    assert np.array_equal(got["count"], want["count"])
    assert np.array_equal(got["deviance"][got["feature"] >= 0], want["deviance"][want["feature"] >= 0])
    assert np.array_equal(leaf, want["leaf_of_doc"])


@pytest.mark.parametrize("fused", [False, True])
def test_ordered_squares_scheme_equals_the_sequential_chain(fused):
    """squares_sum_ is ONE sequentially rounded chain over a node's documents (rtnode_histogram.cc:65-69, 199-203);
    REFERENCE mode computes long ones in parallel (per-chunk parity -> increment functions, qr_exact_kernels.cuh).
    Bit equality with the plain chain on inputs chosen to hurt: wide exponent ranges, zeros, runs of exact ties,
    magnitudes that shrink or grow along the list, a huge first addend."""
    rng = np.random.default_rng(3)
    n = 300_000
    cases = {
        "pseudo-responses": rng.normal(0, 1e-3, n),
        "wide exponents": rng.normal(0, 1, n) * np.exp2(rng.integers(-20, 20, n)),
        "zeros and eighths": rng.integers(-8, 9, n) / 8.0 * (rng.random(n) > 0.25),
        "shrinking": rng.random(n) * np.exp2(-(np.arange(n) % 1000)),
        "growing": rng.random(n) * np.exp2(np.arange(n) / n * 60 - 30),
        "ties": np.concatenate([[741456.0], rng.integers(1, 16, n - 1) * 2.0 ** -7 * rng.choice([-1, 1], n - 1)]),
        "tiny": rng.normal(0, 1e-160, n),
        "short": rng.normal(0, 1, 4097),
    }
    for name, v in cases.items():
        par, ser, replayed = api.selftest_ordered_squares(v, fused)
        want = 0.0
        if len(v) <= 5000:   # the chain in Python for the short case (no fused multiply-add in numpy: unfused only)
            for xv in v:
                want = want + float(xv) * float(xv)
            if not fused:
                assert ser == want, name
        assert par == ser, (name, par, ser)
        assert np.isfinite(par) and par > 0
        if name in ("pseudo-responses", "ties"):
            assert replayed <= 40, (name, replayed)    # the scheme, not the fallback, did the work


@pytest.mark.parametrize("cfg", [
    dict(n=3000, f=20, q=30, gridded=True, nthr=0, leaves=12, minls=1),
    dict(n=3000, f=20, q=30, gridded=False, nthr=0, leaves=8, minls=1),
    dict(n=6000, f=40, q=60, gridded=True, nthr=0, leaves=24, minls=20),
    dict(n=20000, f=30, q=200, gridded=True, nthr=0, leaves=32, minls=1),
])
def test_fit_tree_fast_mode(cfg):
    """Fixed-point accumulation: same split indices and doc->leaf map, except where the oracle's own
    best and the GPU's choice score within 1e-12 relative of each other on that node (the reference
    broke that tie by rounding; see qr_testlib.audit_tree); leaf outputs within 1e-5 relative."""
    x, l, off = common.dataset(n=cfg["n"], f=cfg["f"], q=cfg["q"], gridded=cfg["gridded"])
    lam, w = _gradients(x, l, off, 4)
    ob = po.Binning(np.ascontiguousarray(x.T), cfg["nthr"])
    bins = ob.bins()
    want = ob.fit_tree(lam, w, nleaves=cfg["leaves"], minls=cfg["minls"])
    with api.Trainer(x, l, off, nleaves=cfg["leaves"], minleafsupport=cfg["minls"],
                     nthresholds=cfg["nthr"], hist_mode=api.HIST_FAST) as tr:
        tr.set_pseudoresponses(lam, w)
        got = tr.fit_regressor_on_gradient()
        leaf = tr.get_leaf_assignment()
        got2 = tr.fit_regressor_on_gradient()   # determinism: same answer twice
    assert common.same_structure(got, got2) and np.array_equal(got["value"], got2["value"])
    equiv, near, clean = common.audit_tree(got, want, ob, bins, lam, cfg["minls"])
    assert near <= 1, "%d rounding-decided near-ties in one tree" % near
    if equiv == 0 and near == 0:
        assert np.array_equal(leaf, want["leaf_of_doc"])
        assert np.array_equal(got["count"], want["count"])
    # the tree as a function of the training documents
    og, ow = common.tree_outputs(got, bins), common.tree_outputs(want, bins)
    assert rel_err(og[clean], ow[clean]) <= REL


@pytest.mark.parametrize("algo,depth", [("OBVLAMBDAMART", 4), ("OBVMART", 3)])
@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_oblivious_tree(algo, depth, mode):
    x, l, off = common.dataset(n=4000, f=15, q=40)
    lam, w = _gradients(x, l, off, 6)
    if algo == "OBVMART":
        w = None
    ob = po.Binning(np.ascontiguousarray(x.T), 0)
    want = ob.fit_tree(lam, w, nleaves=1 << depth, minls=1, depth=depth)
    with api.Trainer(x, l, off, algo=algo, treedepth=depth, hist_mode=mode) as tr:
        tr.set_pseudoresponses(lam, w)
        got = tr.fit_regressor_on_gradient()
        leaf = tr.get_leaf_assignment()
    assert common.same_structure(got, want), common.describe_tree_diff(got, want)
    assert np.array_equal(leaf, want["leaf_of_doc"])
    lv = common.leaves_mask(want)
    if mode == api.HIST_REFERENCE:
        assert np.array_equal(got["value"][lv], want["value"][lv])
    else:
        assert rel_err(got["value"][lv], want["value"][lv]) <= REL


@pytest.mark.parametrize("algo", ["LAMBDAMART", "MART", "OBVLAMBDAMART"])
@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_boosting_loop_stagewise(algo, mode):
    """Mart::learn's loop body (mart.cc:331-347), iteration by iteration from the oracle's state:
    lambdas within 1e-13, split indices identical up to audited rounding ties, leaf outputs, scores
    and NDCG@10 within 1e-5 relative."""
    T = 10
    x, l, off = common.dataset(n=4000, f=24, q=40)
    col = np.ascontiguousarray(x.T)
    depth = 3 if algo.startswith("OBV") else 0
    lam_algo = "LAMBDA" in algo
    ob = po.Binning(col, 0)
    bins = ob.bins()
    scores = np.zeros(len(l))
    near_total = 0
    with api.Trainer(x, l, off, algo=algo, nleaves=10, treedepth=max(depth, 1), hist_mode=mode) as tr:
        for m in range(T):
            if lam_algo:
                lam, w = po.lambdas(scores, l, off, 10)
            else:
                lam, w = l.astype(np.float64) - scores, None
            want = ob.fit_tree(lam, w, nleaves=(1 << depth) if depth else 10, minls=1, depth=depth)
            tr.set_scores(scores)
            tr.compute_pseudoresponses()
            glam, gw = tr.get_pseudoresponses()
            assert np.max(np.abs(glam - lam)) <= 1e-13 * np.max(np.abs(lam))
            got = tr.fit_regressor_on_gradient()
            if depth:
                assert common.same_structure(got, want), common.describe_tree_diff(got, want)
                ties, clean = 0, np.ones(len(l), bool)
            else:
                _equiv, ties, clean = common.audit_tree(got, want, ob, bins, lam, 1)
            near_total += ties
            og, ow = common.tree_outputs(got, bins), common.tree_outputs(want, bins)
            assert rel_err(og[clean], ow[clean]) <= REL, "tree %d" % m
            tr.update_modelscores()
            new_scores = po.update_scores(want, col, 0.1, scores)
            gs = tr.get_scores()
            assert np.max(np.abs(gs[clean] - new_scores[clean])) <= REL * np.max(np.abs(new_scores))
            if ties == 0:
                metric = tr.evaluate_dataset()
                want_metric = po.ndcg_dataset(l, new_scores, off, 10)
                assert abs(metric - want_metric) <= REL * want_metric
            scores = new_scores
    assert near_total <= 2


@pytest.mark.parametrize("algo", ["LAMBDAMART", "MART"])
@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_boosting_loop_free_running(algo, mode):
    """Free-running training with a minimum leaf support that keeps nodes large: every tree cuts the
    training documents exactly as the oracle's does (split indices identical except inside audited
    exact-arithmetic ties), NDCG@10 trajectory and final scores within 1e-5 relative."""
    T = 12
    x, l, off = common.dataset(n=6000, f=24, q=60)
    col = np.ascontiguousarray(x.T)
    ob = po.Binning(col, 0)
    bins = ob.bins()
    want_trees, want_metric, want_scores = po.train(algo, x, l, off, T, nleaves=8, minls=100, cutoff=10)
    scores = np.zeros(len(l))
    with api.Trainer(x, l, off, algo=algo, nleaves=8, minleafsupport=100, hist_mode=mode) as tr:
        for m in range(T):
            lam = po.lambdas(scores, l, off, 10)[0] if algo == "LAMBDAMART" else l.astype(np.float64) - scores
            tree, metric = tr.boost_iteration()
            _equiv, near, _clean = common.audit_tree(tree, want_trees[m], ob, bins, lam, 100)
            assert near == 0, "tree %d" % m
            og, ow = common.tree_outputs(tree, bins), common.tree_outputs(want_trees[m], bins)
            assert rel_err(og, ow) <= REL
            assert abs(metric - want_metric[m]) <= REL * abs(want_metric[m])
            scores = po.update_scores(want_trees[m], col, 0.1, scores)
        assert np.max(np.abs(tr.get_scores() - want_scores)) <= REL * np.max(np.abs(want_scores))


def test_stepwise_hooks_equal_fused_iteration():
    x, l, off = common.dataset(n=2000, f=10, q=20)
    with api.Trainer(x, l, off) as a, api.Trainer(x, l, off) as b:
        for _ in range(3):
            a.compute_pseudoresponses()
            ta = a.fit_regressor_on_gradient()
            a.update_modelscores()
            ma = a.evaluate_dataset()
            tb, mb = b.boost_iteration()
            assert common.same_structure(ta, tb) and np.array_equal(ta["value"], tb["value"])
            assert ma == mb
        assert np.array_equal(a.get_scores(), b.get_scores())


def test_apply_tree_adds_and_subtracts():
    x, l, off = common.dataset(n=2000, f=10, q=20)
    with api.Trainer(x, l, off) as tr:
        tree, _ = tr.boost_iteration()
        s1 = tr.get_scores()
        tr.apply_tree(tree, -0.1)          # DART-style removal (dart.cc:634-650)
        s0 = tr.get_scores()
        assert np.max(np.abs(s0)) <= 1e-15
        tr.apply_tree(tree, 0.1)
        assert np.array_equal(tr.get_scores(), s1)
        col = np.ascontiguousarray(x.T)
        assert np.array_equal(po.update_scores(tree, col, 0.1, np.zeros(len(l))), s1)


def test_apply_trees_equals_one_tree_at_a_time():
    """qr_apply_trees walks a set of trees in one pass over the documents; per document the operations
    are those of Dart::update_modelscores' tree-by-tree loop (dart.cc:634-650), so the scores are
    bit-identical to applying the trees one after the other."""
    x, l, off = common.dataset(n=5000, f=14, q=50)
    with api.Trainer(x, l, off, nleaves=8) as tr:
        trees = [tr.boost_iteration()[0] for _ in range(5)]
        base = tr.get_scores()
        w = [-0.1, 0.07, -0.033, 0.2, 0.011]
        tr.apply_trees(trees, w)
        batched = tr.get_scores()
        tr.set_scores(base)
        for t, wt in zip(trees, w):
            tr.apply_tree(t, wt)
        assert np.array_equal(tr.get_scores(), batched)
        assert not np.array_equal(batched, base)


def test_tree_contributions_match_oracle():
    """qr_tree_contributions = Dart::update_contribution_scores (dart.cc:689-706): mean |tree(doc)| per tree (the
    reference sums in an OpenMP reduction, i.e. in no fixed order: 1e-12 relative)."""
    x, l, off = common.dataset(n=5000, f=12, q=50, seed=6)
    with api.Trainer(x, l, off, nleaves=10) as tr:
        trees = [tr.boost_iteration()[0] for _ in range(4)]
        got = tr.tree_contributions(trees)
        again = tr.tree_contributions(trees[::-1])[::-1]
    assert np.array_equal(got, again)   # fixed reduction shape: independent of how the trees are batched
    for t, tree in enumerate(trees):
        want = np.mean(np.abs(po.score_dataset([tree], [1.0], x)))
        assert abs(got[t] - want) <= 1e-12 * want, (t, got[t], want)


def test_scoring_matches_oracle_bit_for_bit():
    x, l, off = common.dataset(n=3000, f=30, q=30)
    trees, weights = synth.random_ensemble(40, 16, 30, seed=3)
    want = po.score_dataset(trees, weights, x)
    with api.Scorer(trees, weights, 30) as sc:
        got = sc.score_dataset(x)
        one = sc.score_document(x[17])
    assert np.array_equal(got, want)
    assert one == want[17]


def test_partial_scores_match_oracle():
    """qr_score_partial: the per-tree score matrix of Driver::extract_partial_scores (driver.cc:411-446) =
    Ensemble::partial_scores_instance (ensemble.cc:121-131) per document, cast to float."""
    x, l, off = common.dataset(n=2100, f=30, q=21)
    trees, weights = synth.random_ensemble(37, 16, 30, seed=5)   # three chunks of trees, the last one part-filled
    weights = np.linspace(0.05, 1.0, len(trees))
    with api.Scorer(trees, weights, 30) as sc:
        part, full = sc.partial_scores(x, with_scores=True)
    assert part.shape == (len(x), len(trees)) and part.dtype == np.float32
    for t, (tree, w) in enumerate(zip(trees, weights)):
        want = po.score_dataset([tree], [w], x).astype(np.float32)
        assert np.array_equal(part[:, t], want), t
    assert np.array_equal(full, po.score_dataset(trees, weights, x))
    with api.Scorer(trees, np.ones(len(trees)), 30) as sc:   # ignore_weights = true
        raw = sc.partial_scores(x)
    assert np.array_equal(raw[:, 11], po.score_dataset([trees[11]], [1.0], x).astype(np.float32))


def test_condop_weight_mode_equals_the_generated_ranker(tmp_path):
    """QR_SCORER_CONDOP_WEIGHTS: the GPU scores equal, bit for bit, the `double ranker(float *v)` the
    conditional-operator generator emits (weights as 3-decimal floats, generate_conditional_operators.cc:95-105),
    compiled without floating-point contraction."""
    import ctypes as C
    import subprocess
    from quickrank_b200 import modelxml
    ql = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "bin", "quicklearn")
    if not os.path.exists(ql):
        pytest.skip("host/bin/quicklearn not built")
    trees, _ = synth.random_ensemble(20, 9, 17, seed=4)
    weights = np.array([0.1, 0.25, 0.0625, 0.1, 1.0, 0.333, 0.1, 0.05, 0.1, 0.2] * 2)
    model, code, so = str(tmp_path / "m.xml"), str(tmp_path / "ranker.c"), str(tmp_path / "ranker.so")
    modelxml.write_model(model, trees, weights)
    out = subprocess.run([ql, "--model-file", model, "--code-file", code, "--generator", "condop"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c", code, "-o", so])
    rk = C.CDLL(so)
    rk.ranker.restype = C.c_double
    rk.ranker.argtypes = [C.POINTER(C.c_float)]
    rng = np.random.default_rng(1)
    x = (rng.integers(0, 256, size=(500, 17)) / 255.0).astype(np.float32)
    want = np.array([rk.ranker(x[i].ctypes.data_as(C.POINTER(C.c_float))) for i in range(len(x))])
    with api.Scorer(trees, weights, 17, condop_weights=True) as sc:
        got = sc.score_dataset(x)
    assert np.array_equal(got, want)
    with api.Scorer(trees, weights, 17) as sc:   # (the plain scorer uses the double weights: 0.0625 and 0.333 differ)
        assert not np.array_equal(sc.score_dataset(x), want)


def test_scoring_a_trained_model():
    x, l, off = common.dataset(n=2500, f=12, q=25)
    with api.Trainer(x, l, off, hist_mode=api.HIST_REFERENCE) as tr:
        trees = [tr.boost_iteration()[0] for _ in range(5)]
        train_scores = tr.get_scores()
    with api.Scorer(trees, [0.1] * 5, 12) as sc:
        got = sc.score_dataset(x)
    assert np.array_equal(got, po.score_dataset(trees, [0.1] * 5, x))
    assert rel_err(got, train_scores) <= 1e-12


def test_validation_context_uses_training_thresholds():
    """qr_ctx_create_eval: a second dataset binned with the training thresholds; applying a tree to it
    equals walking the float features (the validation branch of Mart::learn, mart.cc:354-359)."""
    x, l, off = common.dataset(n=3000, f=12, q=30, seed=3)
    xv, lv, offv = common.dataset(n=1700, f=12, q=17, seed=4, gridded=False)   # values outside the training grid
    with api.Trainer(x, l, off, nleaves=12) as tr:
        ev = tr.eval_context(xv, lv, offv)
        want = np.zeros(len(lv))
        for _ in range(4):
            tree, _m = tr.boost_iteration()
            ev.apply_tree(tree, 0.1)
            want = po.update_scores(tree, np.ascontiguousarray(xv.T), 0.1, want)
            assert np.array_equal(ev.get_scores(), want)
            assert abs(ev.evaluate_dataset() - po.ndcg_dataset(lv, want, offv, 10)) <= 1e-12
        with pytest.raises(api.QrError):
            ev.boost_iteration()
        ev.close()


def test_errors_are_reported():
    x, l, off = common.dataset(n=500, f=4, q=5)
    bad = x.copy()
    bad[3, 1] = np.nan
    with pytest.raises(api.QrError):
        api.Trainer(bad, l, off)
    with pytest.raises(api.QrError):
        api.Trainer(x, l, off[:-1])


def _edge_cases():
    rng = np.random.default_rng(77)
    cases = {}
    # one query only
    x = (rng.integers(0, 256, size=(300, 5)) / 255.0).astype(np.float32)
    cases["single_query"] = (x, rng.integers(0, 5, size=300).astype(np.float32), np.array([0, 300], np.uint64))
    # ragged: queries of 1, 2 and 17 documents (17 = first length libstdc++'s introsort permutes ties), one of 600
    lens = [1, 2, 17, 1, 600, 3, 16, 40]
    n = sum(lens)
    x = (rng.integers(0, 256, size=(n, 7)) / 255.0).astype(np.float32)
    cases["ragged"] = (x, rng.integers(0, 5, size=n).astype(np.float32), np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64))
    # a query whose labels are all zero (idcg = 0: contributes 0 to NDCG and no lambdas) next to normal ones
    lab = rng.integers(0, 4, size=400).astype(np.float32)
    lab[100:250] = 0
    x = (rng.integers(0, 256, size=(400, 6)) / 255.0).astype(np.float32)
    cases["zero_label_query"] = (x, lab, np.array([0, 100, 250, 400], np.uint64))
    # one feature; and a dataset whose only informative feature is accompanied by constant columns
    x = (rng.integers(0, 256, size=(500, 1)) / 255.0).astype(np.float32)
    cases["one_feature"] = (x, (x[:, 0] > 0.5).astype(np.float32) + (x[:, 0] > 0.8), np.array([0, 120, 260, 500], np.uint64))
    x = np.zeros((500, 4), np.float32)
    x[:, 2] = (rng.integers(0, 256, size=500) / 255.0).astype(np.float32)
    cases["constant_columns"] = (x, (x[:, 2] > 0.6).astype(np.float32) * 2, np.array([0, 250, 500], np.uint64))
    # nothing to learn from: all features constant -> every tree is a single leaf
    cases["all_constant"] = (np.full((200, 3), 0.5, np.float32), rng.integers(0, 3, size=200).astype(np.float32),
                             np.array([0, 90, 200], np.uint64))
    return cases


@pytest.mark.parametrize("case", sorted(_edge_cases()))
@pytest.mark.parametrize("algo,leaves,minls", [("LAMBDAMART", 6, 1), ("MART", 2, 1), ("LAMBDAMART", 8, 1000)])
def test_edge_case_datasets(case, algo, leaves, minls):
    """Degenerate and ragged inputs (single query, 1-document queries, all-zero labels, one feature, constant
    columns, nothing splittable, a minimum leaf support no split can meet): REFERENCE-order accumulation
    reproduces the oracle's trees node for node, and the metric / scores agree within 1e-5."""
    x, l, off = _edge_cases()[case]
    T = 4
    want_trees, want_metric, want_scores = po.train(algo, x, l, off, T, nleaves=leaves, minls=minls, cutoff=10)
    col = np.ascontiguousarray(x.T)
    ob = po.Binning(col, 0)
    bins = ob.bins()
    scores = np.zeros(len(l))
    with api.Trainer(x, l, off, algo=algo, nleaves=leaves, minleafsupport=minls, cutoff=10,
                     hist_mode=api.HIST_REFERENCE) as tr:
        for m in range(T):
            tree, metric = tr.boost_iteration()
            if m == 0:   # pseudo-responses are bit-identical on the first iteration: so is the tree
                assert common.same_structure(tree, want_trees[m]), common.describe_tree_diff(tree, want_trees[m])
            else:        # later: CUDA's and glibc's exp() differ in the last bit; only audited ties may differ
                lam = po.lambdas(scores, l, off, 10)[0] if algo == "LAMBDAMART" else l.astype(np.float64) - scores
                _equiv, near, _clean = common.audit_tree(tree, want_trees[m], ob, bins, lam, minls)
                assert near == 0, "tree %d" % m
            og, ow = common.tree_outputs(tree, bins), common.tree_outputs(want_trees[m], bins)
            assert np.allclose(og, ow, rtol=REL, atol=1e-300)
            assert abs(metric - want_metric[m]) <= REL * max(abs(want_metric[m]), 1e-300)
            scores = po.update_scores(want_trees[m], col, 0.1, scores)
        got = tr.get_scores()
    assert np.max(np.abs(got - want_scores)) <= REL * max(np.max(np.abs(want_scores)), 1e-300)
    # and the default fixed-point mode runs the same inputs to the same quality
    with api.Trainer(x, l, off, algo=algo, nleaves=leaves, minleafsupport=minls, cutoff=10, hist_mode=api.HIST_FAST) as tr:
        for m in range(T):
            _tree, metric = tr.boost_iteration()
        assert abs(metric - want_metric[-1]) <= 1e-3 + REL * abs(want_metric[-1])


def test_more_than_two_million_documents():
    """BASELINE configs 3-5 are 2-10 M documents: the row-major entry point (device transpose) and the
    column-major one must bin and train identically past 2^21 documents (a grid dimension used to
    overflow there), and the histogram counts must account for every document."""
    n, f, q = 2_300_000, 3, 23_000
    rng = np.random.default_rng(5)
    x = (rng.integers(0, 64, size=(n, f)) / 64.0).astype(np.float32)
    x[:, 2] = x[:, 0]
    labels = rng.integers(0, 5, size=n).astype(np.float32)
    off = (np.arange(q + 1, dtype=np.uint64) * (n // q)).astype(np.uint64)
    off[-1] = n
    with api.Trainer(x, labels, off, algo="LAMBDAMART", nleaves=8) as a, \
            api.Trainer(np.ascontiguousarray(x.T), labels, off, algo="LAMBDAMART", nleaves=8, layout="colmajor") as b:
        for fi in range(f):
            assert np.array_equal(a.get_bins(fi), b.get_bins(fi))
        assert np.array_equal(a.get_bins(0), np.round(x[:, 0] * 64).astype(a.get_bins(0).dtype))
        for _ in range(2):
            ta, ma = a.boost_iteration()
            tb, mb = b.boost_iteration()
            assert common.same_structure(ta, tb)
            assert np.array_equal(ta["value"], tb["value"]) and ma == mb
            assert int(ta["count"][0]) == n
            assert int(ta["count"][common.leaves_mask(ta)].sum()) == n
        assert np.array_equal(a.get_scores(), b.get_scores())
