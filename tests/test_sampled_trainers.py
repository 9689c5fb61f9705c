"""LambdaMART on a per-query document sample: LAMBDAMART-SELECTIVE and STOCHASTIC-NEGATIVE.

CPU: the host's sample selection (host/src/sampled_trainers.cc) against the unmodified reference's
LambdaMartSelective::sampling_query_level on the same labels / scores, every strategy.
GPU: the sample as a training context of its own (qr_ctx_create_sample) — pseudo-responses of a masked dataset
bit-equal to the reference's LambdaMart::compute_pseudoresponses(sample_presence), the identity sample growing the
trees of the full context, and whole `quicklearn --algo LAMBDAMART-SELECTIVE` runs against the reference's learn()."""
import os
import re
import subprocess

import numpy as np
import pytest

import qr_testlib as common
from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QL = os.path.join(ROOT, "host", "bin", "quicklearn")
CHECK = os.path.join(ROOT, "host", "bin", "selective_check")

needs_ref = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libqr_ref.so not built")


def _sampling_case(seed, ties, q=40):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 60, size=q)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    n = int(off[-1])
    labels = rng.choice([0, 0, 0, 1, 2, 3], size=n).astype(np.float32)
    scores = rng.normal(size=n)
    if ties:
        scores = np.round(scores, 1)
    return labels, scores, off


@needs_ref
@pytest.mark.parametrize("negative", ["RATIO", "MUL", "POS"])
@pytest.mark.parametrize("adaptive", ["NO", "FIXED", "RATIO", "MIX"])
def test_sample_selection_matches_the_reference(negative, adaptive):
    """sampling_query_level (lambdamartselective.cc:326-493): same sample size and the same permuted id list —
    std::sort under the reference's comparators, float quotas, random_shuffle's rand() stream after srand(0)."""
    assert os.path.exists(CHECK), "build host/ first"
    checked = 0
    for rank, rnd in [(0.3, 0.2), (0.1, 0.0), (0.0, 0.37), (0.55, 0.45), (0.7, 0.6), (0.013, 0.9)]:
        for adapt in (1.0, 0.37, 0.0):
            for ties in (False, True):
                if negative == "RATIO" and adaptive == "RATIO" and rank + rnd > 1:
                    continue   # rank factor = sum of the two > 1: the reference's unsigned arithmetic underflows
                labels, scores, off = _sampling_case(int(rank * 1000 + rnd * 10), ties)
                inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(labels), " ".join(str(int(o)) for o in off),
                                           "\n".join("%d %.17g" % (l, s) for l, s in zip(labels, scores)))
                out = subprocess.run([CHECK, repr(rank), repr(rnd), adaptive, negative, repr(adapt)], input=inp,
                                     capture_output=True, text=True)
                assert out.returncode == 0, out.stderr
                vals = np.array(out.stdout.split(), dtype=np.uint64)
                want_n, want_ids = pyref.selective_sample(labels, scores, off, rank, rnd, adaptive, negative, adapt)
                assert int(vals[0]) == want_n, (rank, rnd, adapt, ties)
                assert np.array_equal(vals[1:], want_ids), (rank, rnd, adapt, ties)
                checked += 1
    assert checked >= 30


@needs_ref
@pytest.mark.parametrize("negative", ["RATIO", "POS"])
def test_sample_selection_on_many_queries(negative):
    """Same as above at a size where the host sorts the queries on several threads (the rand() stream and the moves
    to the front stay in query order)."""
    labels, scores, off = _sampling_case(77, True, q=4000)
    assert len(labels) > 100000
    inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(labels), " ".join(str(int(o)) for o in off),
                               "\n".join("%d %.17g" % (l, s) for l, s in zip(labels, scores)))
    out = subprocess.run([CHECK, "0.3", "0.25", "NO", negative, "1.0"], input=inp, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = np.array(out.stdout.split(), dtype=np.uint64)
    want_n, want_ids = pyref.selective_sample(labels, scores, off, 0.3, 0.25, "NO", negative, 1.0)
    assert int(vals[0]) == want_n and np.array_equal(vals[1:], want_ids)


@needs_ref
@pytest.mark.parametrize("negative", ["RATIO", "MUL", "POS"])
def test_sample_selection_edge_cases(negative):
    """Queries of one document, without positives, without negatives, with tied scores throughout, and a dataset of a
    single query: the same draw as the reference's."""
    cases = []
    rng = np.random.default_rng(31)
    lens = np.array([1, 1, 2, 7, 30, 1, 12, 3])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    labels = np.zeros(int(off[-1]), np.float32)
    labels[0] = 2                       # a one-document query holding a positive
    labels[int(off[3]):int(off[4])] = 1  # a query of positives only
    labels[int(off[4]) + 3] = 3          # one positive among 30
    labels[int(off[6]):int(off[6]) + 6] = rng.integers(1, 4, size=6)
    cases.append((labels, np.zeros(len(labels)), off))                        # every score tied
    cases.append((labels, np.round(rng.normal(size=len(labels)), 1), off))
    one = rng.choice([0, 0, 1, 2], size=40).astype(np.float32)
    cases.append((one, rng.normal(size=40), np.array([0, 40], np.uint64)))     # a single query
    for labels, scores, off in cases:
        inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(labels), " ".join(str(int(o)) for o in off),
                                   "\n".join("%d %.17g" % (a, b) for a, b in zip(labels, scores)))
        for rank, rnd in ((0.5, 0.5), (0.2, 0.0), (0.0, 0.4)):
            out = subprocess.run([CHECK, repr(rank), repr(rnd), "NO", negative, "1.0"], input=inp, capture_output=True, text=True)
            assert out.returncode == 0, out.stderr
            vals = np.array(out.stdout.split(), dtype=np.uint64)
            want_n, want_ids = pyref.selective_sample(labels, scores, off, rank, rnd, "NO", negative, 1.0)
            assert int(vals[0]) == want_n and np.array_equal(vals[1:], want_ids), (rank, rnd)


def test_selective_rejects_what_the_reference_dies_on(tmp_path):
    """--sampling-iterations 0 with a sampling factor set is a division by zero in the reference
    (lambdamartselective.cc:170-171); here it is an error message — checked before any device work."""
    from quickrank_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present (the check below relies on failing before the device is touched)")
    tr = str(tmp_path / "t.txt")
    open(tr, "w").write("1 qid:1 1:0.5\n0 qid:1 1:0.1\n")
    out = subprocess.run([QL, "--algo", "LAMBDAMART-SELECTIVE", "--train", tr, "--num-trees", "2"], capture_output=True, text=True)
    assert out.returncode != 0
    assert "sampling-iterations" in out.stderr or "CUDA" in out.stderr or "GPU" in out.stderr


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("share", [1.0, 0.6, 0.15])
def test_masked_pseudoresponses_match_the_reference(share):
    """compute_pseudoresponses with sample_presence (lambdamart.cc:84-105): queries compacted to their sampled
    documents and ranked by scores_on_training_[position in query] (the offset is dropped, :94), rho from the
    documents' own scores.  Bit-equal lambdas and weights on the sampled documents, tie-heavy scores included."""
    from quickrank_b200 import api
    x, l, off = common.dataset(n=6000, f=10, q=60, seed=21)
    rng = np.random.default_rng(5)
    for scores in (rng.normal(size=len(l)), common.tie_heavy_scores(len(l), rng)):
        mask = rng.random(len(l)) < share if share < 1 else np.ones(len(l), bool)
        with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=1, nleaves=4) as s:
            s.init()
            s.set_scores(scores)
            s.compute_pseudoresponses_masked(mask)
            want_lam, want_w = s.get_gradients()
        ids = np.nonzero(mask)[0]
        with api.Trainer(x, l, off, nleaves=4) as full:
            full.set_scores(scores)
            with full.sample_context(x, ids, gather=bool(share < 0.5)) as sm:
                sm.pull_scores(full)
                assert np.array_equal(sm.get_scores(), scores[ids])
                sm.compute_pseudoresponses()
                lam, w = sm.get_pseudoresponses()
        assert np.array_equal(lam, want_lam[ids])
        assert np.array_equal(w, want_w[ids])
        assert not want_lam[~mask].any()


@pytest.mark.gpu
@pytest.mark.parametrize("gather", [True, False])
def test_identity_sample_grows_the_trees_of_the_full_context(gather):
    """A sample holding every document, ranked by its own scores, is the training set: the same trees, bit for bit,
    as the full context in fixed-point mode, over several iterations of pull -> lambdas -> fit -> apply."""
    from quickrank_b200 import api
    x, l, off = common.dataset(n=8000, f=16, q=80, seed=4)
    with api.Trainer(x, l, off, nleaves=12, minleafsupport=20) as plain, \
            api.Trainer(x, l, off, nleaves=12, minleafsupport=20) as full:
        with full.sample_context(x, np.arange(len(l)), rank_by_position=False, gather=gather) as sm:
            for _ in range(4):
                want, want_metric = plain.boost_iteration()
                sm.pull_scores(full)
                sm.compute_pseudoresponses()
                got = sm.fit_regressor_on_gradient()
                full.apply_tree(got, full.shrinkage)
                for k in ("feature", "threshold_idx", "left", "right", "value", "count"):
                    assert np.array_equal(got[k], want[k]), k
                assert np.array_equal(full.get_scores(), plain.get_scores())
                assert full.evaluate_dataset() == want_metric


@pytest.mark.gpu
def test_redraw_equals_a_fresh_sample_context():
    """qr_sample_redraw refills a sample context in place: after a draw that is smaller, one that is larger and one that
    drops whole queries, pseudo-responses and the fitted tree equal those of a context created afresh for the same draw."""
    from quickrank_b200 import api
    x, l, off = common.dataset(n=8000, f=16, q=80, seed=4)
    rng = np.random.default_rng(9)
    with api.Trainer(x, l, off, nleaves=10, minleafsupport=5) as full:
        for _ in range(3):
            full.boost_iteration()
        with full.sample_context(x, np.arange(len(l))) as sm:
            sm.pull_scores(full)
            sm.compute_pseudoresponses()
            sm.fit_regressor_on_gradient()     # (leaves a tree and its histogram slots behind, as in training)
            q_of = np.searchsorted(off, np.arange(len(l)), side="right") - 1
            for share, drop_queries in ((0.3, False), (0.8, False), (0.5, True)):
                mask = rng.random(len(l)) < share
                if drop_queries:
                    mask &= (q_of % 3) != 0
                ids = np.nonzero(mask)[0]
                sm.redraw(full, ids)
                assert sm.N == len(ids)
                with full.sample_context(x, ids) as fresh:
                    out = []
                    for ctx in (sm, fresh):
                        ctx.pull_scores(full)
                        ctx.compute_pseudoresponses()
                        lam, w = ctx.get_pseudoresponses()
                        tree = ctx.fit_regressor_on_gradient()
                        out.append((lam, w, tree, ctx.get_leaf_assignment()))
                (lam_a, w_a, tree_a, leaf_a), (lam_b, w_b, tree_b, leaf_b) = out
                assert np.array_equal(lam_a, lam_b) and np.array_equal(w_a, w_b)
                for k in ("feature", "threshold_idx", "left", "right", "value", "count", "deviance"):
                    assert np.array_equal(tree_a[k], tree_b[k]), (share, k)
                assert np.array_equal(leaf_a, leaf_b)
            # a sample that does not fit is refused
            with full.sample_context(x, np.arange(100), gather=False) as small:
                with pytest.raises(api.QrError):
                    small.redraw(full, np.arange(200))


def _write_svml(path, x, l, off):
    with open(path, "w") as f:
        for q in range(len(off) - 1):
            for i in range(int(off[q]), int(off[q + 1])):
                f.write("%d qid:%d %s\n" % (int(l[i]), q + 1, " ".join("%d:%.9g" % (j + 1, x[i, j]) for j in range(x.shape[1]))))


def _leaf_of(t, data):
    node = np.zeros(len(data), np.int64)
    while True:
        f = t["feature"][node]
        act = np.nonzero(f >= 0)[0]
        if len(act) == 0:
            return node
        left = data[act, f[act]] <= t["threshold"][node[act]]
        node[act] = np.where(left, t["left"][node[act]], t["right"][node[act]])


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("sel,cli", [
    (dict(sampling_iterations=3, rank_factor=0.3, random_factor=0.2),
     ["--sampling-iterations", "3", "--rank-sampling-factor", "0.3", "--random-sampling-factor", "0.2"]),
    (dict(sampling_iterations=2, rank_factor=0.5, random_factor=0.0, negative="POS"),
     ["--sampling-iterations", "2", "--rank-sampling-factor", "0.5", "--random-sampling-factor", "0.0", "--negative-strategy", "POS"]),
    (dict(sampling_iterations=2, rank_factor=1.5, random_factor=0.5, negative="MUL", adaptive="MIX", normalization_factor=4.0),
     ["--sampling-iterations", "2", "--rank-sampling-factor", "1.5", "--random-sampling-factor", "0.5", "--negative-strategy", "MUL",
      "--adaptive-strategy", "MIX", "--normalization-factor", "4"]),
    # the full context in reference-order mode (scores, NDCG mean in the reference's order); the sample stays fixed-point
    (dict(sampling_iterations=4, rank_factor=0.25, random_factor=0.25),
     ["--sampling-iterations", "4", "--rank-sampling-factor", "0.25", "--random-sampling-factor", "0.25", "--hist-mode", "reference"]),
])
def test_selective_matches_the_reference_learn_loop(tmp_path, sel, cli):
    """Whole `quicklearn --algo LAMBDAMART-SELECTIVE` runs against the unmodified LambdaMartSelective::learn: RATIO with
    random negatives, POS, MUL with the MIX adaptive strategy (see _selective_against_reference)."""
    _selective_against_reference(tmp_path, sel, cli, n=4000, f=12, ntrees=9, leaves=8, minls=10)


@pytest.mark.gpu
@needs_ref
def test_selective_matches_the_reference_on_a_larger_set(tmp_path):
    """The same comparison on 30 000 documents x 20 continuous features, 16 leaves (several threads sort the queries
    in the host's draw; the sample context holds ~20 000 documents)."""
    _selective_against_reference(tmp_path, dict(sampling_iterations=2, rank_factor=0.4, random_factor=0.2),
                                 ["--sampling-iterations", "2", "--rank-sampling-factor", "0.4", "--random-sampling-factor", "0.2"],
                                 n=30000, f=20, ntrees=7, leaves=16, minls=20, gridded=False)


def _selective_against_reference(tmp_path, sel, cli, n, f, ntrees, leaves, minls, gridded=True):
    """LambdaMartSelective::learn (lambdamartselective.cc:46-313) end to end: the sizes of every sample ("Reducing
    training size from N to M"), the sampling factors it prints, the NDCG trajectory (4 decimals printed) and the
    trees — split features and structure equal, thresholds equal or cutting the training documents into the same
    sets, leaf outputs within 1e-9 relative (fixed-point sums on the sample against the reference's FP64 order)."""
    from quickrank_b200 import modelxml
    x, l, off = common.dataset(n=n, f=f, q=n // 100, seed=8, gridded=gridded)
    with pyref.RefSession("LAMBDAMART-SELECTIVE", x, l, off, ntrees=ntrees, nleaves=leaves, minleafsupport=minls, selective=sel) as s:
        s.learn()
        want_metric = s.metric_history()
        want_log = s.log()
        want_trees = [s.tree(t) for t in range(s.num_trees())]
    tr, model = str(tmp_path / "train.txt"), str(tmp_path / "sel.xml")
    _write_svml(tr, x, l, off)
    cmd = [QL, "--algo", "LAMBDAMART-SELECTIVE", "--train", tr, "--num-trees", str(ntrees), "--num-leaves", str(leaves),
           "--min-leaf-support", str(minls), "--model-out", model, "--end-after-rounds", "0", "--partial", "0"] + cli
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    pick = lambda text, pat: re.findall(pat, text, flags=re.M)
    for pat in (r"^Reducing training size from \d+ to \d+", r"^Rank Factor: .*", r"^N\. Positives: .*"):
        assert pick(out.stdout, pat) == pick(want_log, pat), (pat, out.stdout, want_log)
    assert len(pick(want_log, r"^Reducing")) >= 2
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)", out.stdout, flags=re.M)
    assert len(rows) == ntrees == len(want_metric), out.stdout
    got_metric = np.array([float(r[1]) for r in rows])
    assert np.max(np.abs(got_metric - want_metric)) <= 6e-5, (got_metric, want_metric)
    _info, got_trees, _weights = modelxml.read_model(model)
    assert len(got_trees) == ntrees and "<type>LAMBDAMART-SELECTIVE</type>" in open(model).read()
    for a, b in zip(got_trees, want_trees):
        assert np.array_equal(a["feature"], b["feature"])
        assert np.array_equal(a["left"], b["left"]) and np.array_equal(a["right"], b["right"])
        if not np.array_equal(a["threshold"], b["threshold"]):
            assert np.array_equal(_leaf_of(a, x), _leaf_of(b, x))
        lv = a["feature"] < 0
        assert np.allclose(a["value"][lv], b["value"][lv], rtol=1e-9, atol=1e-14)
    # the saved model loads in stock QuickRank (type dispatch, ltr_algorithm.cc:100-102)
    assert np.all(np.isfinite(pyref.score_with_model(model, x)))


@pytest.mark.gpu
def test_stochastic_negative_keeps_positives_and_a_share_of_negatives(tmp_path):
    """StochasticNegative::learn (stochasticnegative.cc:46-283): every iteration after the first trains on all positives
    plus floor(subsample x negatives) documents per query.  The reference seeds its shuffle from the wall clock, so the
    checks are structural: the sample size it reports, reproducibility under --seed, a different draw under another
    seed, and a model that improves on the training set."""
    x, l, off = common.dataset(n=4000, f=12, q=40, seed=8)
    tr = str(tmp_path / "train.txt")
    _write_svml(tr, x, l, off)
    want = 0
    for q in range(len(off) - 1):
        lab = l[int(off[q]):int(off[q + 1])]
        npos = int((lab > 0).sum())
        want += npos + int(np.floor(np.float32(0.4) * np.float32(len(lab) - npos)))

    def run(seed):
        model = str(tmp_path / ("sn%d.xml" % seed))
        cmd = [QL, "--algo", "STOCHASTIC-NEGATIVE", "--train", tr, "--num-trees", "6", "--num-leaves", "8", "--subsample", "0.4",
               "--min-leaf-support", "10", "--model-out", model, "--end-after-rounds", "0", "--partial", "0", "--seed", str(seed)]
        out = subprocess.run(cmd, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
        sizes = [int(v) for v in re.findall(r"^Reducing training size from 4000 to (\d+)", out.stdout, flags=re.M)]
        assert sizes == [want] * 5, (sizes, want)
        rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)", out.stdout, flags=re.M)
        assert len(rows) == 6
        metric = [float(r[1]) for r in rows]
        assert metric[-1] > metric[0]
        return open(model).read()

    a, b, c = run(1), run(1), run(2)
    assert a == b
    assert a != c
    assert "<type>STOCHASTIC-NEGATIVE</type>" in a


@pytest.mark.gpu
def test_selective_with_a_validation_set_rolls_back_to_the_best_model(tmp_path):
    """The validation branch of the sampled loop (lambdamartselective.cc:217-241, 281-287): validation scores follow
    every tree, the starred rows are the new bests, early stop after --end-after-rounds rows without one, and the saved
    ensemble ends at the best row."""
    x, l, off = common.dataset(n=3000, f=12, q=30, seed=8)
    xv, lv, offv = common.dataset(n=1500, f=12, q=15, seed=9)
    tr, va, model = str(tmp_path / "train.txt"), str(tmp_path / "valid.txt"), str(tmp_path / "m.xml")
    _write_svml(tr, x, l, off)
    _write_svml(va, xv, lv, offv)
    cmd = [QL, "--algo", "LAMBDAMART-SELECTIVE", "--train", tr, "--valid", va, "--num-trees", "30", "--num-leaves", "8",
           "--min-leaf-support", "10", "--model-out", model, "--end-after-rounds", "4", "--partial", "0",
           "--sampling-iterations", "3", "--rank-sampling-factor", "0.3", "--random-sampling-factor", "0.3"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)\s+([0-9.]+)( \*)?$", out.stdout, flags=re.M)
    assert rows, out.stdout
    best = max(int(r[0]) for r in rows if r[3])
    valid = [float(r[2]) for r in rows]
    assert valid[best - 1] == max(valid)
    assert len(rows) == min(30, best + 4)   # mart.cc:309-311: stop when m > best_model_ + esr (both 0-based)
    assert len(re.findall(r"<tree id=", open(model).read())) == best
    m = re.search(r"NDCG@10 on validation data = ([0-9.]+)", out.stdout)
    assert m and abs(float(m.group(1)) - valid[best - 1]) <= 5e-5


# ---------------------------------------------------------------------------------------------------------------
# committed golden vectors (tests/golden/sampled.npz, made from the unmodified reference by tests/golden/make_golden.py):
# the same checks where oracle/_ref is not available
# ---------------------------------------------------------------------------------------------------------------
GOLDEN = os.path.join(ROOT, "tests", "golden", "sampled.npz")


def _golden_masked_case():
    from quickrank_b200 import synth
    g = np.load(GOLDEN)
    x, l, off = synth.make_dataset(1500, 6, 18, seed=111)
    return g, x, l, off


def test_oracle_masked_pseudoresponses_against_the_golden_vectors():
    """oracle/qr_oracle.c qro_lambdas_masked (the sample_presence branch of lambdamart.cc:62-152) reproduces the
    reference's lambdas and weights bit for bit."""
    from oracle import pyoracle as po
    g, _x, l, off = _golden_masked_case()
    lam, w = po.lambdas_masked(g["masked_scores"], l, off, 10, g["masked_presence"])
    assert np.array_equal(lam, g["masked_lambda"]) and np.array_equal(w, g["masked_weight"])
    assert not lam[g["masked_presence"] == 0].any()


@needs_ref
def test_oracle_masked_pseudoresponses_against_the_reference():
    from oracle import pyoracle as po
    x, l, off = common.dataset(n=5000, f=6, q=50, seed=23)
    rng = np.random.default_rng(6)
    for scores in (rng.normal(size=len(l)), common.tie_heavy_scores(len(l), rng)):
        for share in (1.0, 0.5, 0.08):
            mask = (rng.random(len(l)) < share).astype(np.uint8)
            with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=1, nleaves=4) as s:
                s.init()
                s.set_scores(scores)
                s.compute_pseudoresponses_masked(mask)
                want_lam, want_w = s.get_gradients()
            lam, w = po.lambdas_masked(scores, l, off, 10, mask)
            assert np.array_equal(lam, want_lam) and np.array_equal(w, want_w)


def test_host_draw_against_the_golden_vectors():
    g = np.load(GOLDEN)
    labels, scores, off = g["draw_labels"], g["draw_scores"], g["draw_offsets"]
    inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(labels), " ".join(str(int(o)) for o in off),
                               "\n".join("%d %.17g" % (a, b) for a, b in zip(labels, scores)))
    for i in range(4):
        rank, rnd, adaptive, negative, adapt = eval(str(g["draw%d_params" % i]))
        out = subprocess.run([CHECK, repr(rank), repr(rnd), adaptive, negative, repr(adapt)], input=inp, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        vals = np.array(out.stdout.split(), dtype=np.uint64)
        assert int(vals[0]) == int(g["draw%d_n" % i]) and np.array_equal(vals[1:], g["draw%d_ids" % i]), i


@pytest.mark.gpu
def test_sample_context_against_the_golden_vectors(tmp_path):
    """The GPU path against the committed reference outputs: masked pseudo-responses bit-equal; a whole
    LAMBDAMART-SELECTIVE run with the reference's sample sizes, NDCG trajectory and trees."""
    from quickrank_b200 import api, modelxml, synth
    g, x, l, off = _golden_masked_case()
    ids = np.nonzero(g["masked_presence"])[0]
    with api.Trainer(x, l, off, nleaves=4) as full:
        full.set_scores(g["masked_scores"])
        with full.sample_context(x, ids) as sm:
            sm.pull_scores(full)
            sm.compute_pseudoresponses()
            lam, w = sm.get_pseudoresponses()
    assert np.array_equal(lam, g["masked_lambda"][ids]) and np.array_equal(w, g["masked_weight"][ids])
    c = eval(str(g["learn_case"]))
    x, l, off = synth.make_dataset(c["n"], c["f"], c["q"], seed=c["seed"])
    tr, model = str(tmp_path / "train.txt"), str(tmp_path / "sel.xml")
    _write_svml(tr, x, l, off)
    sel = c["selective"]
    cmd = [QL, "--algo", "LAMBDAMART-SELECTIVE", "--train", tr, "--num-trees", str(c["trees"]), "--num-leaves", str(c["leaves"]),
           "--min-leaf-support", str(c["minls"]), "--model-out", model, "--end-after-rounds", "0", "--partial", "0",
           "--sampling-iterations", str(sel["sampling_iterations"]), "--rank-sampling-factor", str(sel["rank_factor"]),
           "--random-sampling-factor", str(sel["random_factor"])]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    want_log = str(g["learn_log"])
    pick = lambda text, pat: re.findall(pat, text, flags=re.M)
    for pat in (r"^Reducing training size from \d+ to \d+", r"^N\. Positives: .*"):
        assert pick(out.stdout, pat) == pick(want_log, pat) and pick(want_log, pat)
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)", out.stdout, flags=re.M)
    assert np.max(np.abs(np.array([float(r[1]) for r in rows]) - g["learn_metric"])) <= 6e-5
    _info, trees, _w = modelxml.read_model(model)
    assert len(trees) == c["trees"]
    for t, a in enumerate(trees):
        assert np.array_equal(a["feature"], g["learn_tree%d_feature" % t])
        assert np.array_equal(a["left"], g["learn_tree%d_left" % t]) and np.array_equal(a["right"], g["learn_tree%d_right" % t])
        b = dict(feature=g["learn_tree%d_feature" % t], threshold=g["learn_tree%d_threshold" % t],
                 left=g["learn_tree%d_left" % t], right=g["learn_tree%d_right" % t])
        if not np.array_equal(a["threshold"], b["threshold"]):
            assert np.array_equal(_leaf_of(a, x), _leaf_of(b, x))
        lv = a["feature"] < 0
        assert np.allclose(a["value"][lv], g["learn_tree%d_value" % t][lv], rtol=1e-9, atol=1e-14)


@pytest.mark.gpu
def test_sample_context_edge_cases_against_the_oracle():
    """Samples that leave one document per query, keep a single query, or consist of one-document queries: pseudo-
    responses equal the oracle's masked restatement (itself bit-identical to the reference), and a tree can be fitted."""
    from oracle import pyoracle as po
    from quickrank_b200 import api
    x, l, off = common.dataset(n=2400, f=8, q=24, seed=14, qlen=(1, 200))
    rng = np.random.default_rng(2)
    scores = common.tie_heavy_scores(len(l), rng) + np.where(rng.random(len(l)) < 0.3, rng.normal(size=len(l)), 0.0)
    q_of = np.searchsorted(off, np.arange(len(l)), side="right") - 1
    masks = {
        "one document per query": np.isin(np.arange(len(l)), off[:-1].astype(np.int64) + (np.diff(off).astype(np.int64) // 2)),
        "a single query": q_of == 5,
        "two documents of every third query": (q_of % 3 == 0) & (np.arange(len(l)) - off[q_of].astype(np.int64) < 2),
        "everything but the first document of each query": ~np.isin(np.arange(len(l)), off[:-1].astype(np.int64)),
    }
    with api.Trainer(x, l, off, nleaves=4, minleafsupport=1) as full:
        full.set_scores(scores)
        sm = None
        for name, mask in masks.items():
            ids = np.nonzero(mask)[0]
            want_lam, want_w = po.lambdas_masked(scores, l, off, 10, mask.astype(np.uint8))
            if sm is None:
                sm = full.sample_context(x, ids)
            else:
                sm.redraw(full, ids)
            sm.pull_scores(full)
            sm.compute_pseudoresponses()
            lam, w = sm.get_pseudoresponses()
            assert np.array_equal(lam, want_lam[ids]) and np.array_equal(w, want_w[ids]), name
            tree = sm.fit_regressor_on_gradient()
            assert int(tree["count"][0]) == len(ids), name
        sm.close()


def test_cpu_replay_of_a_selective_run_is_bit_identical_to_the_reference():
    """CPU only: the oracle's pieces (masked pseudo-responses, tree fit over `sampleids` in the reference's order, score
    update, NDCG) driven by the HOST's draw (host/bin/selective_check, the product's sampling code, continuing one
    rand() stream across draws) replay the golden LambdaMartSelective::learn run of tests/golden/sampled.npz bit for
    bit: every tree (structure, thresholds, leaf outputs), every iteration's NDCG, every sample size."""
    from oracle import pyoracle as po
    from quickrank_b200 import synth
    g = np.load(GOLDEN)
    c = eval(str(g["learn_case"]))
    sel = c["selective"]
    x, l, off = synth.make_dataset(c["n"], c["f"], c["q"], seed=c["seed"])
    state = {"rand_calls": 0}

    def draw(scores):
        inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(l), " ".join(str(int(o)) for o in off),
                                   "\n".join("%d %.17g" % (a, b) for a, b in zip(l, scores)))
        out = subprocess.run([CHECK, repr(sel["rank_factor"]), repr(sel["random_factor"]), "NO", "RATIO", "1.0",
                              str(state["rand_calls"])], input=inp, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        state["rand_calls"] += int(re.search(r"rand calls: (\d+)", out.stderr).group(1))
        vals = np.array(out.stdout.split(), dtype=np.uint64)
        return int(vals[0]), vals[1:]

    trees, metric, _scores, sizes = po.train_sampled(x, l, off, c["trees"], draw, lambda m: m % sel["sampling_iterations"] == 0,
                                                     nleaves=c["leaves"], minls=c["minls"], cutoff=c["cutoff"])
    want_sizes = [int(v) for v in re.findall(r"^Reducing training size from \d+ to (\d+)", str(g["learn_log"]), flags=re.M)]
    assert sizes == want_sizes and len(sizes) >= 2 and state["rand_calls"] > 0
    assert np.array_equal(metric, g["learn_metric"])
    for t, tree in enumerate(trees):
        for k in ("feature", "threshold_idx", "threshold", "left", "right", "value", "count"):
            assert np.array_equal(tree[k], g["learn_tree%d_%s" % (t, k)]), (t, k)


@needs_ref
@pytest.mark.parametrize("negative,rank,rnd,every", [("POS", 0.5, 0.3, 2), ("MUL", 1.5, 0.5, 2), ("RATIO", 0.2, 0.0, 3)])
def test_cpu_replay_against_the_live_reference(negative, rank, rnd, every):
    """The same CPU replay (oracle pieces + the host's draw) against LambdaMartSelective::learn run here, for the other
    negative strategies: trees and NDCG trajectory bit for bit."""
    from oracle import pyoracle as po
    x, l, off = common.dataset(n=2000, f=8, q=20, seed=17)
    ntrees = 7
    sel = dict(sampling_iterations=every, rank_factor=rank, random_factor=rnd, negative=negative)
    with pyref.RefSession("LAMBDAMART-SELECTIVE", x, l, off, ntrees=ntrees, nleaves=6, minleafsupport=3, selective=sel) as s:
        s.learn()
        want_metric = s.metric_history()
        want_trees = [s.tree(t) for t in range(s.num_trees())]
    state = {"rand_calls": 0}

    def draw(scores):
        inp = "%d %d\n%s\n%s\n" % (len(off) - 1, len(l), " ".join(str(int(o)) for o in off),
                                   "\n".join("%d %.17g" % (a, b) for a, b in zip(l, scores)))
        out = subprocess.run([CHECK, repr(rank), repr(rnd), "NO", negative, "1.0", str(state["rand_calls"])], input=inp,
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        state["rand_calls"] += int(re.search(r"rand calls: (\d+)", out.stderr).group(1))
        vals = np.array(out.stdout.split(), dtype=np.uint64)
        return int(vals[0]), vals[1:]

    trees, metric, _scores, _sizes = po.train_sampled(x, l, off, ntrees, draw, lambda m: m % every == 0, nleaves=6, minls=3)
    assert np.array_equal(metric, want_metric)
    for t, tree in enumerate(trees):
        for k in ("feature", "threshold_idx", "threshold", "left", "right", "value", "count"):
            assert np.array_equal(tree[k], want_trees[t][k]), (t, k)
