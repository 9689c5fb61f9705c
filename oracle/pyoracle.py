"""TEST INFRASTRUCTURE: ctypes view of oracle/libqr_oracle.so (the C restatement of the
reference's hot path, oracle/qr_oracle.c).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this; the product never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqr_oracle.so")


class Tree(C.Structure):
    _fields_ = [
        ("nnodes", C.c_uint32), ("nleaves", C.c_uint32),
        ("feature", C.POINTER(C.c_int32)), ("threshold_idx", C.POINTER(C.c_uint32)),
        ("threshold", C.POINTER(C.c_float)), ("left", C.POINTER(C.c_int32)),
        ("right", C.POINTER(C.c_int32)), ("value", C.POINTER(C.c_double)),
        ("deviance", C.POINTER(C.c_double)), ("count", C.POINTER(C.c_uint64)),
    ]


class Bins(C.Structure):
    _fields_ = [
        ("N", C.c_size_t), ("F", C.c_size_t), ("thr", C.POINTER(C.POINTER(C.c_float))),
        ("thr_size", C.POINTER(C.c_size_t)), ("bins", C.POINTER(C.c_uint32)),
        ("colmajor", C.POINTER(C.c_float)),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        sz, dp, fp = C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_float)
        u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.qro_sort_desc.argtypes = [dp, sz, u32p]
        for name in ("qro_dcg_labels", "qro_idcg"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [fp, sz, sz]
        for name in ("qro_dcg_query", "qro_ndcg_query"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [fp, dp, sz, sz]
        L.qro_ndcg_dataset.restype = C.c_double
        L.qro_ndcg_dataset.argtypes = [fp, dp, u64p, sz, sz]
        L.qro_delta_ndcg.restype = C.c_double
        L.qro_delta_ndcg.argtypes = [fp, sz, sz, C.c_double, sz, sz]
        L.qro_lambdas.argtypes = [dp, fp, u64p, sz, sz, dp, dp]
        L.qro_lambdas_masked.argtypes = [dp, fp, u64p, sz, sz, C.POINTER(C.c_uint8), dp, dp]
        L.qro_mart_pseudo.argtypes = [dp, fp, sz, dp]
        L.qro_radix_argsort.argtypes = [fp, sz, u64p]
        L.qro_binning.restype = C.POINTER(Bins)
        L.qro_binning.argtypes = [fp, sz, sz, sz]
        L.qro_bins_free.argtypes = [C.POINTER(Bins)]
        L.qro_fit_tree.restype = C.POINTER(Tree)
        L.qro_fit_tree.argtypes = [C.POINTER(Bins), dp, dp, sz, sz, sz, u32p]
        L.qro_fit_tree_sampled.restype = C.POINTER(Tree)
        L.qro_fit_tree_sampled.argtypes = [C.POINTER(Bins), dp, dp, u64p, sz, sz, sz, sz, u32p]
        L.qro_tree_free.argtypes = [C.POINTER(Tree)]
        L.qro_split_scores.argtypes = [C.POINTER(Bins), dp, u64p, sz, sz, u32p, u32p, sz, dp]
        L.qro_update_scores.argtypes = [C.POINTER(Tree), fp, sz, C.c_double, dp]
        L.qro_score_dataset.argtypes = [C.POINTER(C.POINTER(Tree)), dp, sz, fp, sz, sz, dp]
        L.qro_train.argtypes = [C.c_int, fp, fp, u64p, sz, sz, sz, sz, C.c_double, sz, sz, sz, sz,
                                sz, C.POINTER(C.POINTER(Tree)), dp, dp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


ALGOS = {"MART": 0, "LAMBDAMART": 1, "OBVMART": 2, "OBVLAMBDAMART": 3}
TREE_FIELDS = (("feature", np.int32), ("threshold_idx", np.uint32), ("threshold", np.float32),
               ("left", np.int32), ("right", np.int32), ("value", np.float64),
               ("deviance", np.float64), ("count", np.uint64))


def tree_to_dict(tp):
    t = tp.contents
    n = t.nnodes
    out = {}
    for name, dt in TREE_FIELDS:
        out[name] = np.ctypeslib.as_array(getattr(t, name), shape=(n,)).astype(dt, copy=True)
    out["nleaves"] = int(t.nleaves)
    return out


class CTree:
    """A flat tree (dict of numpy arrays) presented to the oracle as a qro_tree."""

    def __init__(self, d):
        self.arrs = {}
        n = len(d["feature"])
        self.t = Tree()
        self.t.nnodes = n
        self.t.nleaves = int(np.sum(np.asarray(d["feature"]) < 0))
        for name, dt in TREE_FIELDS:
            a = np.ascontiguousarray(d[name], dtype=dt) if name in d else np.zeros(n, dt)
            self.arrs[name] = a
            setattr(self.t, name, a.ctypes.data_as(dict(Tree._fields_)[name]))


def sort_desc(scores):
    scores = np.ascontiguousarray(scores, np.float64)
    idx = np.empty(len(scores), np.uint32)
    lib().qro_sort_desc(_p(scores, C.c_double), len(scores), _p(idx, C.c_uint32))
    return idx


def dcg_query(labels, scores, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    return lib().qro_dcg_query(_p(labels, C.c_float), _p(scores, C.c_double), len(labels), cutoff)


def ndcg_query(labels, scores, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    return lib().qro_ndcg_query(_p(labels, C.c_float), _p(scores, C.c_double), len(labels), cutoff)


def idcg(labels, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    return lib().qro_idcg(_p(labels, C.c_float), len(labels), cutoff)


def delta_ndcg(sorted_labels, cutoff, idcg_v, i, j):
    sl = np.ascontiguousarray(sorted_labels, np.float32)
    return lib().qro_delta_ndcg(_p(sl, C.c_float), len(sl), cutoff, idcg_v, i, j)


def ndcg_dataset(labels, scores, qoff, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    return lib().qro_ndcg_dataset(_p(labels, C.c_float), _p(scores, C.c_double),
                                  _p(qoff, C.c_uint64), len(qoff) - 1, cutoff)


def lambdas(scores, labels, qoff, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    lam = np.zeros(len(labels), np.float64)
    w = np.zeros(len(labels), np.float64)
    lib().qro_lambdas(_p(scores, C.c_double), _p(labels, C.c_float), _p(qoff, C.c_uint64),
                      len(qoff) - 1, cutoff, _p(lam, C.c_double), _p(w, C.c_double))
    return lam, w


def lambdas_masked(scores, labels, qoff, cutoff, presence):
    """compute_pseudoresponses with sample_presence (lambdamart.cc:84-105): the document-sampling trainers' call"""
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    presence = np.ascontiguousarray(presence, np.uint8)
    lam = np.zeros(len(labels), np.float64)
    w = np.zeros(len(labels), np.float64)
    lib().qro_lambdas_masked(_p(scores, C.c_double), _p(labels, C.c_float), _p(qoff, C.c_uint64), len(qoff) - 1, cutoff,
                             _p(presence, C.c_uint8), _p(lam, C.c_double), _p(w, C.c_double))
    return lam, w


def radix_argsort(v):
    v = np.ascontiguousarray(v, np.float32)
    out = np.empty(len(v), np.uint64)
    lib().qro_radix_argsort(_p(v, C.c_float), len(v), _p(out, C.c_uint64))
    return out


class Binning:
    """thresholds + bin map for a column-major feature matrix [F, N]."""

    def __init__(self, colmajor, nthresholds=0):
        self.col = np.ascontiguousarray(colmajor, np.float32)
        self.F, self.N = self.col.shape
        self.h = lib().qro_binning(_p(self.col, C.c_float), self.N, self.F, nthresholds)

    def thresholds(self, f):
        b = self.h.contents
        n = b.thr_size[f]
        return np.ctypeslib.as_array(b.thr[f], shape=(n,)).copy()

    def bins(self):
        b = self.h.contents
        return np.ctypeslib.as_array(b.bins, shape=(self.F, self.N)).copy()

    def fit_tree(self, lam, w=None, nleaves=10, minls=1, depth=0, sampleids=None):
        """sampleids: fit on these documents only, in this order (the document-sampling trainers)"""
        lam = np.ascontiguousarray(lam, np.float64)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
            wp = _p(w, C.c_double)
        leaf = np.zeros(self.N, np.uint32)
        if sampleids is not None:
            ids = np.ascontiguousarray(sampleids, np.uint64)
            tp = lib().qro_fit_tree_sampled(self.h, _p(lam, C.c_double), wp, _p(ids, C.c_uint64), len(ids), nleaves, minls,
                                            depth, _p(leaf, C.c_uint32))
        else:
            tp = lib().qro_fit_tree(self.h, _p(lam, C.c_double), wp, nleaves, minls, depth,
                                    _p(leaf, C.c_uint32))
        d = tree_to_dict(tp)
        lib().qro_tree_free(tp)
        d["leaf_of_doc"] = leaf
        return d

    def split_scores(self, lam, ids, minls, cands):
        """Reference split score of each (feature, threshold_idx) candidate on the node `ids`."""
        lam = np.ascontiguousarray(lam, np.float64)
        ids = np.ascontiguousarray(ids, np.uint64)
        cf = np.ascontiguousarray([c[0] for c in cands], np.uint32)
        ct = np.ascontiguousarray([c[1] for c in cands], np.uint32)
        out = np.zeros(len(cands), np.float64)
        lib().qro_split_scores(self.h, _p(lam, C.c_double), _p(ids, C.c_uint64), len(ids), minls,
                               _p(cf, C.c_uint32), _p(ct, C.c_uint32), len(cands), _p(out, C.c_double))
        return out

    def close(self):
        if self.h:
            lib().qro_bins_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def update_scores(tree, colmajor, shrinkage, scores):
    col = np.ascontiguousarray(colmajor, np.float32)
    scores = np.ascontiguousarray(scores, np.float64).copy()
    ct = CTree(tree)
    lib().qro_update_scores(C.byref(ct.t), _p(col, C.c_float), col.shape[1], shrinkage,
                            _p(scores, C.c_double))
    return scores


def score_dataset(trees, weights, rowmajor):
    x = np.ascontiguousarray(rowmajor, np.float32)
    cts = [CTree(t) for t in trees]
    arr = (C.POINTER(Tree) * len(cts))(*[C.pointer(c.t) for c in cts])
    w = np.ascontiguousarray(weights, np.float64)
    out = np.zeros(x.shape[0], np.float64)
    lib().qro_score_dataset(arr, _p(w, C.c_double), len(cts), _p(x, C.c_float), x.shape[0],
                            x.shape[1], _p(out, C.c_double))
    return out


def train_sampled(rowmajor, labels, qoff, ntrees, draw, resample_due, shrinkage=0.1, nthresholds=0, nleaves=10, minls=1,
                  cutoff=10):
    """LambdaMART on a redrawn document sample — the loop of LambdaMartSelective::learn / StochasticNegative::learn
    (lambdamartselective.cc:163-215) restated over the oracle's pieces, in the reference's arithmetic and ORDER (root
    histogram and leaf sums run over `sampleids` as drawn).  draw(scores) -> (n, ids): a permutation of the documents
    with the sample in front (the reference's sampling_query_level; the tests pass the host's draw); resample_due(m):
    is a new sample drawn before iteration m.  Returns (trees, metric per iteration, scores, sample sizes)."""
    x = np.ascontiguousarray(rowmajor, np.float32)
    col = np.ascontiguousarray(x.T)
    labels = np.ascontiguousarray(labels, np.float32)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    N = x.shape[0]
    bins = Binning(col, nthresholds)
    scores = np.zeros(N, np.float64)
    ids, n = np.arange(N, dtype=np.uint64), N
    presence = np.ones(N, np.uint8)
    trees, metric, sizes = [], [], []
    for m in range(ntrees):
        if m > 0 and resample_due(m):
            n, ids = draw(scores)
            sizes.append(n)
            if n < N:                                   # lambdamartselective.cc:188-192
                presence[:] = 0
                presence[ids[:n]] = 1
        lam, w = lambdas_masked(scores, labels, qoff, cutoff, presence)
        tree = bins.fit_tree(lam, w, nleaves=nleaves, minls=minls, sampleids=ids[:n])
        scores = update_scores(tree, col, shrinkage, scores)
        trees.append(tree)
        metric.append(ndcg_dataset(labels, scores, qoff, cutoff))
    bins.close()
    return trees, np.array(metric), scores, sizes


def train(algo, rowmajor, labels, qoff, ntrees, shrinkage=0.1, nthresholds=0, nleaves=10, depth=0,
          minls=1, cutoff=10, keep_trees=True):
    x = np.ascontiguousarray(rowmajor, np.float32)
    col = np.ascontiguousarray(x.T)
    labels = np.ascontiguousarray(labels, np.float32)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    N, F = x.shape
    trees = (C.POINTER(Tree) * ntrees)()
    metric = np.zeros(ntrees, np.float64)
    scores = np.zeros(N, np.float64)
    lib().qro_train(ALGOS[algo], _p(col, C.c_float), _p(labels, C.c_float), _p(qoff, C.c_uint64), N,
                    F, len(qoff) - 1, ntrees, shrinkage, nthresholds, nleaves, depth, minls, cutoff,
                    trees if keep_trees else None, _p(metric, C.c_double), _p(scores, C.c_double))
    out = []
    if keep_trees:
        for i in range(ntrees):
            out.append(tree_to_dict(trees[i]))
            lib().qro_tree_free(trees[i])
    return out, metric, scores
