"""TEST INFRASTRUCTURE: ctypes view of oracle/_ref/libqr_ref.so (the unmodified reference
sources + oracle/ref_harness.cc).  Only tests/, __graft_entry__.smoke() and bench.py's
reference arm may import this module; the product never does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libqr_ref.so")

ALGOS = {"MART": 0, "LAMBDAMART": 1, "OBVMART": 2, "OBVLAMBDAMART": 3, "DART": 4, "LAMBDAMART-SELECTIVE": 5}
SEL_ADAPTIVE = ["NO", "FIXED", "RATIO", "MIX"]
SEL_NEGATIVE = ["RATIO", "MUL", "POS"]


class Params(C.Structure):
    _fields_ = [
        ("algo", C.c_int32),
        ("ntrees", C.c_uint64),
        ("shrinkage", C.c_double),
        ("nthresholds", C.c_uint64),
        ("nleaves", C.c_uint64),
        ("treedepth", C.c_uint64),
        ("minleafsupport", C.c_uint64),
        ("cutoff", C.c_uint64),
        ("dart_sample_type", C.c_int32),
        ("dart_normalize_type", C.c_int32),
        ("dart_adaptive_type", C.c_int32),
        ("dart_rate_drop", C.c_double),
        ("dart_skip_drop", C.c_double),
        ("dart_keep_drop", C.c_int32),
        ("dart_best_on_train", C.c_int32),
        ("dart_random_keep", C.c_double),
        ("dart_drop_on_best", C.c_double),
        ("sel_sampling_iterations", C.c_int32),
        ("sel_adaptive", C.c_int32),
        ("sel_negative", C.c_int32),
        ("sel_pad", C.c_int32),
        ("sel_rank_factor", C.c_double),
        ("sel_random_factor", C.c_double),
        ("sel_normalization_factor", C.c_double),
    ]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u64, dp, fp = C.c_void_p, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.qref_open.restype = vp
        L.qref_open.argtypes = [C.POINTER(Params), fp, fp, C.POINTER(u64), u64, u64, u64]
        L.qref_close.argtypes = [vp]
        L.qref_learn.argtypes = [vp, C.c_int, C.c_int]
        L.qref_init.argtypes = [vp]
        L.qref_set_scores.argtypes = [vp, dp]
        L.qref_get_scores.argtypes = [vp, dp]
        L.qref_compute_pseudoresponses.argtypes = [vp]
        L.qref_compute_pseudoresponses_masked.argtypes = [vp, C.POINTER(C.c_uint8)]
        L.qref_get_gradients.argtypes = [vp, dp, dp]
        L.qref_set_gradients.argtypes = [vp, dp, dp]
        L.qref_fit_tree.argtypes = [vp, C.c_int]
        L.qref_evaluate.restype = C.c_double
        L.qref_evaluate.argtypes = [vp]
        L.qref_num_trees.restype = u64
        L.qref_num_trees.argtypes = [vp]
        L.qref_tree_nodes.restype = u64
        L.qref_tree_nodes.argtypes = [vp, u64]
        L.qref_tree_get.argtypes = [vp, u64, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), fp,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp, dp,
                                    C.POINTER(u64), dp]
        L.qref_thresholds_size.restype = u64
        L.qref_thresholds_size.argtypes = [vp, u64]
        L.qref_thresholds_get.argtypes = [vp, u64, fp]
        L.qref_num_metric.restype = u64
        L.qref_num_metric.argtypes = [vp]
        L.qref_metric_get.argtypes = [vp, dp]
        L.qref_num_recorded.restype = u64
        L.qref_num_recorded.argtypes = [vp, C.c_int]
        L.qref_recorded_get.argtypes = [vp, C.c_int, u64, dp]
        L.qref_save_model.argtypes = [vp, C.c_char_p]
        L.qref_generate_code.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.qref_score_with_model.argtypes = [C.c_char_p, fp, u64, u64, dp]
        L.qref_dcg.restype = C.c_double
        L.qref_dcg.argtypes = [fp, dp, u64, u64]
        L.qref_ndcg.restype = C.c_double
        L.qref_ndcg.argtypes = [fp, dp, u64, u64]
        L.qref_ndcg_jacobian.argtypes = [fp, dp, u64, u64, dp]
        L.qref_sort_indices.argtypes = [dp, u64, C.POINTER(u64)]
        L.qref_radix_argsort.argtypes = [fp, u64, C.POINTER(u64)]
        L.qref_linesearch.argtypes = [fp, u64, u64, fp, C.POINTER(u64), u64, u64, C.c_uint32, C.c_double, C.c_double,
                                      C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, dp, dp]
        L.qref_linesearch_valid.argtypes = [fp, u64, u64, fp, C.POINTER(u64), u64, fp, u64, fp, C.POINTER(u64), u64, u64,
                                            C.c_uint32, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                                            dp, dp]
        L.qref_cleaver.argtypes = [C.c_int, fp, u64, u64, fp, C.POINTER(u64), u64, u64, C.c_double, dp, C.c_uint32,
                                   C.c_double, C.c_double, C.c_uint32, dp]
        L.qref_log.restype = u64
        L.qref_log.argtypes = [vp, C.c_char_p, u64]
        L.qref_selective_sample.restype = u64
        L.qref_selective_sample.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, fp, dp, C.POINTER(u64),
                                            u64, u64, C.POINTER(u64)]
        L.qref_read_svml.argtypes = [C.c_char_p, C.POINTER(u64), C.POINTER(u64), dp]
        L.qref_set_threads.argtypes = [C.c_int]
        L.qref_max_threads.restype = C.c_int
        _lib = L
    return _lib


def set_threads(n: int) -> int:
    """Sets the OpenMP team size of the reference's loops; returns what the runtime will use."""
    lib().qref_set_threads(int(n))
    return int(lib().qref_max_threads())


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefSession:
    """One reference algorithm object bound to one dataset."""

    def __init__(self, algo, x, labels, qoff, ntrees=10, shrinkage=0.1, nthresholds=0, nleaves=10,
                 treedepth=3, minleafsupport=1, cutoff=10, dart=None, selective=None):
        L = lib()
        self.x = np.ascontiguousarray(x, dtype=np.float32)
        self.labels = np.ascontiguousarray(labels, dtype=np.float32)
        self.qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        self.N, self.F = self.x.shape
        self.Q = len(self.qoff) - 1
        p = Params()
        p.algo = ALGOS[algo]
        p.ntrees, p.shrinkage, p.nthresholds = ntrees, shrinkage, nthresholds
        p.nleaves, p.treedepth, p.minleafsupport, p.cutoff = nleaves, treedepth, minleafsupport, cutoff
        d = dict(sample_type=0, normalize_type=0, adaptive_type=0, rate_drop=0.1, skip_drop=0.0,
                 keep_drop=0, best_on_train=0, random_keep=0.0, drop_on_best=0.0)
        d.update(dart or {})
        p.dart_sample_type, p.dart_normalize_type = d["sample_type"], d["normalize_type"]
        p.dart_adaptive_type = d["adaptive_type"]
        p.dart_rate_drop, p.dart_skip_drop = d["rate_drop"], d["skip_drop"]
        p.dart_keep_drop, p.dart_best_on_train = d["keep_drop"], d["best_on_train"]
        p.dart_random_keep, p.dart_drop_on_best = d["random_keep"], d["drop_on_best"]
        sel = dict(sampling_iterations=0, rank_factor=1.0, random_factor=0.0, normalization_factor=100.0,
                   adaptive="NO", negative="RATIO")
        sel.update(selective or {})
        p.sel_sampling_iterations = sel["sampling_iterations"]
        p.sel_adaptive, p.sel_negative = SEL_ADAPTIVE.index(sel["adaptive"]), SEL_NEGATIVE.index(sel["negative"])
        p.sel_rank_factor, p.sel_random_factor = sel["rank_factor"], sel["random_factor"]
        p.sel_normalization_factor = sel["normalization_factor"]
        self.h = L.qref_open(C.byref(p), _p(self.x, C.c_float), _p(self.labels, C.c_float),
                             _p(self.qoff, C.c_uint64), self.N, self.F, self.Q)
        if not self.h:
            raise RuntimeError("qref_open failed")

    def close(self):
        if self.h:
            lib().qref_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # full reference learn() loop
    def learn(self, keep_gradients=False, quiet=True):
        lib().qref_learn(self.h, int(keep_gradients), int(quiet))

    # step-wise protocol
    def log(self):
        """stdout of the last quiet learn()"""
        n = int(lib().qref_log(self.h, None, 0))
        buf = C.create_string_buffer(n + 1)
        lib().qref_log(self.h, buf, n + 1)
        return buf.value.decode()

    def init(self):
        lib().qref_init(self.h)

    def set_scores(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        lib().qref_set_scores(self.h, _p(s, C.c_double))

    def get_scores(self):
        s = np.empty(self.N, dtype=np.float64)
        lib().qref_get_scores(self.h, _p(s, C.c_double))
        return s

    def compute_pseudoresponses(self):
        lib().qref_compute_pseudoresponses(self.h)

    def compute_pseudoresponses_masked(self, presence):
        """LambdaMart::compute_pseudoresponses with sample_presence (lambdamart.cc:84-105)"""
        m = np.ascontiguousarray(presence, dtype=np.uint8)
        assert len(m) == self.N
        lib().qref_compute_pseudoresponses_masked(self.h, _p(m, C.c_uint8))

    def get_gradients(self):
        lam = np.zeros(self.N, dtype=np.float64)
        w = np.zeros(self.N, dtype=np.float64)
        lib().qref_get_gradients(self.h, _p(lam, C.c_double), _p(w, C.c_double))
        return lam, w

    def set_gradients(self, lam, w=None):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64)
            wp = _p(w, C.c_double)
        lib().qref_set_gradients(self.h, _p(lam, C.c_double), wp)

    def fit_tree(self, update_scores=True):
        return lib().qref_fit_tree(self.h, int(update_scores))

    def evaluate(self):
        return lib().qref_evaluate(self.h)

    # recordings
    def num_trees(self):
        return lib().qref_num_trees(self.h)

    def tree(self, t):
        L = lib()
        n = L.qref_tree_nodes(self.h, t)
        out = dict(feature=np.empty(n, np.int32), threshold_idx=np.empty(n, np.uint32),
                   threshold=np.empty(n, np.float32), left=np.empty(n, np.int32),
                   right=np.empty(n, np.int32), value=np.empty(n, np.float64),
                   deviance=np.empty(n, np.float64), count=np.empty(n, np.uint64))
        w = C.c_double()
        L.qref_tree_get(self.h, t, _p(out["feature"], C.c_int32), _p(out["threshold_idx"], C.c_uint32),
                        _p(out["threshold"], C.c_float), _p(out["left"], C.c_int32),
                        _p(out["right"], C.c_int32), _p(out["value"], C.c_double),
                        _p(out["deviance"], C.c_double), _p(out["count"], C.c_uint64), C.byref(w))
        out["weight"] = w.value
        return out

    def thresholds(self, f):
        L = lib()
        n = L.qref_thresholds_size(self.h, f)
        out = np.empty(n, np.float32)
        L.qref_thresholds_get(self.h, f, _p(out, C.c_float))
        return out

    def metric_history(self):
        L = lib()
        n = L.qref_num_metric(self.h)
        out = np.empty(n, np.float64)
        L.qref_metric_get(self.h, _p(out, C.c_double))
        return out

    def recorded(self, kind, it):
        k = {"lambdas": 0, "weights": 1, "scores": 2}[kind]
        out = np.empty(self.N, np.float64)
        lib().qref_recorded_get(self.h, k, it, _p(out, C.c_double))
        return out

    def num_recorded(self, kind):
        k = {"lambdas": 0, "weights": 1, "scores": 2}[kind]
        return lib().qref_num_recorded(self.h, k)

    def save_model(self, path):
        lib().qref_save_model(self.h, path.encode())


def generate_code(xml_path, code_path, kind):
    """The reference's own XML -> C generators (kind: "condop" | "oblivious" | "vpred")."""
    lib().qref_generate_code(xml_path.encode(), code_path.encode(), {"condop": 0, "oblivious": 1, "vpred": 2}[kind])


def score_with_model(xml_path, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape[0], np.float64)
    rc = lib().qref_score_with_model(xml_path.encode(), _p(x, C.c_float), x.shape[0], x.shape[1],
                                     _p(out, C.c_double))
    if rc:
        raise RuntimeError("reference could not load %s" % xml_path)
    return out


def dcg(labels, scores, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    return lib().qref_dcg(_p(labels, C.c_float), _p(scores, C.c_double), len(labels), cutoff)


def ndcg(labels, scores, cutoff):
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    return lib().qref_ndcg(_p(labels, C.c_float), _p(scores, C.c_double), len(labels), cutoff)


def ndcg_jacobian(labels, scores, cutoff):
    """Dense symmetric n x n matrix of Ndcg::jacobian for one result list."""
    labels = np.ascontiguousarray(labels, np.float32)
    scores = np.ascontiguousarray(scores, np.float64)
    n = len(labels)
    packed = np.zeros(n * (n + 1) // 2, np.float64)
    lib().qref_ndcg_jacobian(_p(labels, C.c_float), _p(scores, C.c_double), n, cutoff,
                             _p(packed, C.c_double))
    m = np.zeros((n, n))
    k = 0
    for i in range(n):
        for j in range(i, n):
            m[i, j] = m[j, i] = packed[k]
            k += 1
    return m


def sort_indices(scores):
    scores = np.ascontiguousarray(scores, np.float64)
    out = np.empty(len(scores), np.uint64)
    lib().qref_sort_indices(_p(scores, C.c_double), len(scores), _p(out, C.c_uint64))
    return out


def radix_argsort(v):
    v = np.ascontiguousarray(v, np.float32)
    out = np.empty(len(v), np.uint64)
    lib().qref_radix_argsort(_p(v, C.c_float), len(v), _p(out, C.c_uint64))
    return out


def linesearch(x, labels, qoff, cutoff=10, num_points=20, window_size=1.0, reduction_factor=0.95, max_iterations=5,
               max_failed_vali=20, adaptive=False, last_only=0, init_weights=None, valid=None):
    """The reference's LineSearch::learn (line_search.cc:153-416) on a row-major matrix, NDCG@cutoff; valid = (x, labels,
    query offsets) of a validation set (not larger than the training set) or None; returns the learned weights."""
    x = np.ascontiguousarray(x, np.float32)
    labels = np.ascontiguousarray(labels, np.float32)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    out = np.zeros(x.shape[1], np.float64)
    iw = None
    if init_weights is not None:
        init_weights = np.ascontiguousarray(init_weights, np.float64)
        iw = _p(init_weights, C.c_double)
    if valid is not None:
        xv = np.ascontiguousarray(valid[0], np.float32)
        lv = np.ascontiguousarray(valid[1], np.float32)
        ov = np.ascontiguousarray(valid[2], np.uint64)
        rc = lib().qref_linesearch_valid(_p(x, C.c_float), x.shape[0], x.shape[1], _p(labels, C.c_float), _p(qoff, C.c_uint64),
                                         len(qoff) - 1, _p(xv, C.c_float), xv.shape[0], _p(lv, C.c_float), _p(ov, C.c_uint64),
                                         len(ov) - 1, cutoff, num_points, window_size, reduction_factor, max_iterations,
                                         max_failed_vali, int(adaptive), last_only, iw, _p(out, C.c_double))
        if rc:
            raise RuntimeError("reference line search (with validation) failed: %d" % rc)
        return out
    rc = lib().qref_linesearch(_p(x, C.c_float), x.shape[0], x.shape[1], _p(labels, C.c_float), _p(qoff, C.c_uint64),
                               len(qoff) - 1, cutoff, num_points, window_size, reduction_factor, max_iterations,
                               max_failed_vali, int(adaptive), last_only, iw, _p(out, C.c_double))
    if rc:
        raise RuntimeError("reference line search failed")
    return out


CLEAVER_METHODS = {"LAST": 0, "SKIP": 1, "LOW_WEIGHTS": 2, "QUALITY_LOSS": 3, "QUALITY_LOSS_ADV": 4, "SCORE_LOSS": 5}


def cleaver(method, x, labels, qoff, weights, pruning_rate, cutoff=10, num_points=0, window_size=1.0,
            reduction_factor=0.95, max_iterations=5):
    """The reference's Cleaver::optimize (cleaver.cc:166-412) on a partial-score matrix; returns the new weights
    (0 for pruned trees).  num_points == 0: no line search."""
    x = np.ascontiguousarray(x, np.float32)
    labels = np.ascontiguousarray(labels, np.float32)
    qoff = np.ascontiguousarray(qoff, np.uint64)
    w = np.ascontiguousarray(weights, np.float64)
    out = np.zeros(x.shape[1], np.float64)
    rc = lib().qref_cleaver(CLEAVER_METHODS[method], _p(x, C.c_float), x.shape[0], x.shape[1], _p(labels, C.c_float),
                            _p(qoff, C.c_uint64), len(qoff) - 1, cutoff, pruning_rate, _p(w, C.c_double), num_points,
                            window_size, reduction_factor, max_iterations, _p(out, C.c_double))
    if rc:
        raise RuntimeError("reference cleaver failed (%d)" % rc)
    return out


def selective_sample(labels, scores, qoff, rank_factor, random_factor, adaptive="NO", negative="RATIO", adapt_factor=1.0):
    """One draw of the reference's LambdaMartSelective::sampling_query_level after srand(0): (sample size, id list)."""
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
    ids = np.zeros(len(labels), dtype=np.uint64)
    n = lib().qref_selective_sample(float(rank_factor), float(random_factor), SEL_ADAPTIVE.index(adaptive),
                                    SEL_NEGATIVE.index(negative), float(adapt_factor), _p(labels, C.c_float),
                                    _p(scores, C.c_double), _p(qoff, C.c_uint64), len(labels), len(qoff) - 1,
                                    _p(ids, C.c_uint64))
    return int(n), ids


def read_svml(path):
    """The reference's io::Svml::read_horizontal on a file: ((N, F, Q), (fnv labels, fnv offsets, fnv matrix), seconds)."""
    shape = (C.c_uint64 * 3)()
    sums = (C.c_uint64 * 3)()
    sec = C.c_double()
    rc = lib().qref_read_svml(path.encode(), shape, sums, C.byref(sec))
    if rc:
        raise RuntimeError("reference SVMLight reader failed")
    return tuple(int(v) for v in shape), tuple(int(v) for v in sums), sec.value
