"""ctypes binding of the C ABI (include/quickrank_b200.h) — the host-side mirror used by the
tests and the benchmark.  The product path fails loudly when the CUDA library is missing or no
GPU is present: there is no CPU fallback here (the CPU restatement lives in oracle/ and is test
infrastructure only).

Names follow the reference's hooks (include/learning/forests/mart.h:118-147 of hpclab/quickrank):
``compute_pseudoresponses``, ``fit_regressor_on_gradient``, ``update_modelscores``,
``evaluate_dataset``, ``score_dataset``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libquickrank_b200.so")

QR_OK = 0
ALGOS = {"MART": 0, "LAMBDAMART": 1, "OBVMART": 2, "OBVLAMBDAMART": 3}
HIST_FAST, HIST_REFERENCE = 0, 1
COMM_ID_BYTES = 128
PHASES = ("pseudo", "hist", "scan", "partition", "leaf", "rank")


class QrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("quickrank_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("algo", C.c_uint32), ("nleaves", C.c_uint32), ("treedepth", C.c_uint32),
        ("minleafsupport", C.c_uint32), ("nthresholds", C.c_uint64), ("ndcg_cutoff", C.c_uint64),
        ("shrinkage", C.c_double), ("hist_mode", C.c_uint32), ("device", C.c_int32),
    ]


class FlatTree(C.Structure):
    _fields_ = [
        ("capacity", C.c_uint32), ("nnodes", C.c_uint32), ("nleaves", C.c_uint32),
        ("feature", C.POINTER(C.c_int32)), ("threshold_idx", C.POINTER(C.c_uint32)),
        ("threshold", C.POINTER(C.c_float)), ("left", C.POINTER(C.c_int32)),
        ("right", C.POINTER(C.c_int32)), ("value", C.POINTER(C.c_double)),
        ("deviance", C.POINTER(C.c_double)), ("count", C.POINTER(C.c_uint64)),
    ]


TREE_FIELDS = (("feature", np.int32), ("threshold_idx", np.uint32), ("threshold", np.float32),
               ("left", np.int32), ("right", np.int32), ("value", np.float64),
               ("deviance", np.float64), ("count", np.uint64))

_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QrError(-1, "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (nvcc, sm_100a)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, sz, dp, fp = C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_float)
        u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.qr_last_error.restype = C.c_char_p
        L.qr_device_count.restype = C.c_int
        for name in ("qr_ctx_create", "qr_ctx_create_rowmajor"):
            getattr(L, name).argtypes = [fp, sz, sz, fp, u64p, sz, C.POINTER(Params), C.POINTER(vp)]
        L.qr_ctx_create_eval.argtypes = [vp, fp, sz, sz, fp, u64p, sz, C.POINTER(vp)]
        L.qr_ctx_create_sample.argtypes = [vp, fp, sz, sz, fp, u64p, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(vp)]
        L.qr_sample_redraw.argtypes = [vp, vp, sz, fp, u64p, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.qr_sample_pull_scores.argtypes = [vp, vp]
        L.qr_ctx_destroy.argtypes = [vp]
        L.qr_get_thresholds.argtypes = [vp, sz, C.POINTER(fp), C.POINTER(sz)]
        L.qr_compute_pseudoresponses.argtypes = [vp]
        L.qr_fit_tree.argtypes = [vp, C.POINTER(FlatTree)]
        L.qr_update_modelscores.argtypes = [vp, C.c_double]
        L.qr_apply_tree.argtypes = [vp, C.POINTER(FlatTree), C.c_double]
        L.qr_apply_trees.argtypes = [vp, C.POINTER(FlatTree), dp, C.c_size_t]
        L.qr_tree_contributions.argtypes = [vp, C.POINTER(FlatTree), C.c_size_t, dp]
        L.qr_evaluate.argtypes = [vp, dp]
        L.qr_ls_create.argtypes = [fp, sz, sz, fp, u64p, sz, C.c_uint32, C.c_int, C.POINTER(vp)]
        L.qr_ls_destroy.argtypes = [vp]
        L.qr_ls_evaluate.argtypes = [vp, dp, dp]
        L.qr_ls_feature_points.argtypes = [vp, dp, C.c_uint32, dp, C.c_uint32, dp]
        L.qr_ls_line_points.argtypes = [vp, dp, dp, C.c_uint32, dp]
        L.qr_ls_drop_points.argtypes = [vp, dp, C.POINTER(C.c_uint32), C.c_uint32, dp]
        L.qr_ls_drop_column.argtypes = [vp, dp, C.c_uint32]
        L.qr_ls_score_loss.argtypes = [vp, dp, dp]
        L.qr_ls_launch_count.argtypes = [vp]
        L.qr_ls_launch_count.restype = C.c_uint64
        L.qr_selftest_ordered_squares.argtypes = [dp, sz, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_uint64)]
        L.qr_boost_iteration.argtypes = [vp, C.POINTER(FlatTree), dp]
        L.qr_get_scores.argtypes = [vp, dp]
        L.qr_set_scores.argtypes = [vp, dp]
        L.qr_get_pseudoresponses.argtypes = [vp, dp, dp]
        L.qr_set_pseudoresponses.argtypes = [vp, dp, dp]
        L.qr_get_leaf_assignment.argtypes = [vp, u32p]
        L.qr_get_bins.argtypes = [vp, sz, u32p]
        L.qr_get_ranking.argtypes = [vp, u32p]
        L.qr_last_tree_stats.argtypes = [vp, dp, dp, u32p]
        L.qr_last_tree_rounds.argtypes = [vp, u32p, dp]
        L.qr_launch_count.restype = C.c_uint64
        L.qr_launch_count.argtypes = [vp]
        L.qr_phase_times.argtypes = [vp, dp, u64p, C.c_int]
        L.qr_set_profiling.argtypes = [vp, C.c_int]
        L.qr_hist_kernel_time.argtypes = [vp, dp, u64p, dp, C.c_int]
        L.qr_timer_start.argtypes = [vp]
        L.qr_timer_stop.argtypes = [vp, dp]
        L.qr_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
        L.qr_ctx_create_sharded.argtypes = [fp, C.c_int, sz, sz, fp, u64p, sz, C.POINTER(Params),
                                            C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.POINTER(vp)]
        L.qr_ctx_comm_transport.argtypes = [vp]
        L.qr_scorer_create.argtypes = [C.POINTER(FlatTree), dp, sz, sz, C.c_int, C.POINTER(vp)]
        L.qr_scorer_create_ex.argtypes = [C.POINTER(FlatTree), dp, sz, sz, C.c_int, C.c_uint, C.POINTER(vp)]
        L.qr_scorer_destroy.argtypes = [vp]
        L.qr_score_partial.argtypes = [vp, fp, sz, sz, fp, dp]
        L.qr_score_partial_device.argtypes = [vp, vp, sz, sz, vp, vp]
        L.qr_score_dataset.argtypes = [vp, fp, sz, sz, dp]
        L.qr_score_dataset_device.argtypes = [vp, vp, sz, sz, vp]
        L.qr_scorer_sync.argtypes = [vp]
        L.qr_scorer_launch_count.restype = C.c_uint64
        L.qr_scorer_launch_count.argtypes = [vp]
        L.qr_scorer_timer.argtypes = [vp, C.c_int, dp]
        L.qr_score_document.argtypes = [vp, fp, sz, dp]
        _lib = L
    return _lib


def _check(rc):
    if rc != QR_OK:
        raise QrError(rc, lib().qr_last_error().decode(errors="replace"))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def device_count() -> int:
    return lib().qr_device_count()


class TreeBuffer:
    """Caller-owned arrays behind a qr_flat_tree."""

    def __init__(self, capacity):
        self.arrs = {name: np.zeros(capacity, dt) for name, dt in TREE_FIELDS}
        self.t = FlatTree()
        self.t.capacity = capacity
        for name, _ in TREE_FIELDS:
            setattr(self.t, name, self.arrs[name].ctypes.data_as(dict(FlatTree._fields_)[name]))

    @classmethod
    def from_dict(cls, d):
        n = len(d["feature"])
        tb = cls(n)
        for name, dt in TREE_FIELDS:
            if name in d:
                tb.arrs[name][:] = np.asarray(d[name], dt)
        tb.t.nnodes = n
        tb.t.nleaves = int(np.sum(np.asarray(d["feature"]) < 0))
        return tb

    def to_dict(self):
        n = self.t.nnodes
        out = {name: self.arrs[name][:n].copy() for name, _ in TREE_FIELDS}
        out["nleaves"] = int(self.t.nleaves)
        return out


class Trainer:
    """One training run on one GPU: the device-resident counterpart of Mart/LambdaMart between
    ``init`` and ``clear`` (mart.cc:117-206)."""

    def __init__(self, x, labels, qoff, algo="LAMBDAMART", nleaves=10, treedepth=3, minleafsupport=1,
                 nthresholds=0, cutoff=10, shrinkage=0.1, hist_mode=HIST_FAST, device=-1,
                 layout="rowmajor", comm=None):
        """comm = (id_bytes, rank, world) makes this the context of one rank of a sharded run
        (x, labels, qoff are then the rank's own whole queries)."""
        L = lib()
        self.labels = np.ascontiguousarray(labels, np.float32)
        self.qoff = np.ascontiguousarray(qoff, np.uint64)
        x = np.ascontiguousarray(x, np.float32)
        if layout == "rowmajor":
            self.N, self.F = x.shape
        else:
            self.F, self.N = x.shape
        self.Q = len(self.qoff) - 1
        p = Params()
        p.algo = ALGOS[algo]
        p.nleaves, p.treedepth, p.minleafsupport = nleaves, treedepth, minleafsupport
        p.nthresholds, p.ndcg_cutoff, p.shrinkage = nthresholds, cutoff, shrinkage
        p.hist_mode, p.device = hist_mode, device
        self.params = p
        self.shrinkage = shrinkage
        self.max_nodes = 2 * ((1 << treedepth) if algo.startswith("OBV") else max(1, nleaves)) + 1
        self.h = C.c_void_p()
        if comm is not None:
            cid, rank, world = comm
            buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(cid)
            _check(L.qr_ctx_create_sharded(_p(x, C.c_float), int(layout == "rowmajor"), self.N, self.F,
                                           _p(self.labels, C.c_float), _p(self.qoff, C.c_uint64), self.Q,
                                           C.byref(p), buf, rank, world, C.byref(self.h)))
            return
        fn = L.qr_ctx_create_rowmajor if layout == "rowmajor" else L.qr_ctx_create
        _check(fn(_p(x, C.c_float), self.N, self.F, _p(self.labels, C.c_float),
                  _p(self.qoff, C.c_uint64), self.Q, C.byref(p), C.byref(self.h)))

    def eval_context(self, x, labels, qoff):
        """Validation / test set binned with this trainer's thresholds (qr_ctx_create_eval)."""
        ev = Trainer.__new__(Trainer)
        ev.labels = np.ascontiguousarray(labels, np.float32)
        ev.qoff = np.ascontiguousarray(qoff, np.uint64)
        x = np.ascontiguousarray(x, np.float32)
        ev.N, ev.F = x.shape
        ev.Q = len(ev.qoff) - 1
        ev.params, ev.shrinkage, ev.max_nodes = self.params, self.shrinkage, self.max_nodes
        ev.h = C.c_void_p()
        _check(lib().qr_ctx_create_eval(self.h, _p(x, C.c_float), ev.N, ev.F, _p(ev.labels, C.c_float),
                                        _p(ev.qoff, C.c_uint64), ev.Q, C.byref(ev.h)))
        return ev

    def _sample_layout(self, doc_ids, rank_by_position):
        doc_ids = np.ascontiguousarray(doc_ids, np.uint32)
        assert np.all(np.diff(doc_ids.astype(np.int64)) > 0), "sampled documents must be ascending"
        q_of = np.searchsorted(self.qoff, doc_ids, side="right") - 1
        counts = np.bincount(q_of, minlength=self.Q)
        labels = np.ascontiguousarray(self.labels[doc_ids], np.float32)
        qoff = np.concatenate([[0], np.cumsum(counts[counts > 0])]).astype(np.uint64)
        key = np.ascontiguousarray(doc_ids - self.qoff[q_of].astype(np.uint32), np.uint32) if rank_by_position else None
        return doc_ids, labels, qoff, key

    def sample_context(self, x, doc_ids, rank_by_position=True, gather=True):
        """The documents `doc_ids` (ascending; whole set `x` row-major) as a training context of their own, binned with
        this trainer's thresholds (qr_ctx_create_sample): what LambdaMartSelective / StochasticNegative fit a tree on.
        rank_by_position: the reference's ranking key (lambdamart.cc:94: the score of the document whose index is this
        one's position inside its query); False: the document's own score.  gather: take the bins from this context on
        the device (x is not read) instead of uploading and binning the sampled rows."""
        doc_ids, labels, qoff, key = self._sample_layout(doc_ids, rank_by_position)
        sm = Trainer.__new__(Trainer)
        sm.labels, sm.qoff = labels, qoff
        rows = None if gather else np.ascontiguousarray(np.asarray(x, np.float32)[doc_ids])
        sm.N, sm.F = len(doc_ids), self.F
        sm.Q = len(sm.qoff) - 1
        sm.params, sm.shrinkage, sm.max_nodes = self.params, self.shrinkage, self.max_nodes
        sm.h = C.c_void_p()
        _check(lib().qr_ctx_create_sample(self.h, _p(rows, C.c_float) if rows is not None else None, sm.N, sm.F,
                                          _p(sm.labels, C.c_float), _p(sm.qoff, C.c_uint64), sm.Q, _p(doc_ids, C.c_uint32),
                                          _p(key, C.c_uint32) if key is not None else None, C.byref(sm.h)))
        return sm

    def redraw(self, full, doc_ids, rank_by_position=True):
        """(sample context created with gather=True) a new draw in place: qr_sample_redraw"""
        doc_ids, labels, qoff, key = full._sample_layout(doc_ids, rank_by_position)
        _check(lib().qr_sample_redraw(self.h, full.h, len(doc_ids), _p(labels, C.c_float), _p(qoff, C.c_uint64), len(qoff) - 1,
                                      _p(doc_ids, C.c_uint32), _p(key, C.c_uint32) if key is not None else None))
        self.labels, self.qoff, self.N, self.Q = labels, qoff, len(doc_ids), len(qoff) - 1

    def pull_scores(self, full):
        """(sample context) copies the current scores of `full` (qr_sample_pull_scores)"""
        _check(lib().qr_sample_pull_scores(self.h, full.h))

    def close(self):
        if self.h:
            lib().qr_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def thresholds(self, f):
        ptr, n = C.POINTER(C.c_float)(), C.c_size_t()
        _check(lib().qr_get_thresholds(self.h, f, C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def compute_pseudoresponses(self):
        _check(lib().qr_compute_pseudoresponses(self.h))

    def fit_regressor_on_gradient(self, want_tree=True):
        if not want_tree:
            _check(lib().qr_fit_tree(self.h, None))
            return None
        tb = TreeBuffer(self.max_nodes)
        _check(lib().qr_fit_tree(self.h, C.byref(tb.t)))
        return tb.to_dict()

    def update_modelscores(self, weight=None):
        _check(lib().qr_update_modelscores(self.h, self.shrinkage if weight is None else weight))

    def apply_tree(self, tree, weight):
        tb = TreeBuffer.from_dict(tree)
        _check(lib().qr_apply_tree(self.h, C.byref(tb.t), weight))

    def apply_trees(self, trees, weights):
        bufs = [TreeBuffer.from_dict(t) for t in trees]
        arr = (FlatTree * len(bufs))(*[b.t for b in bufs])
        w = np.ascontiguousarray(weights, np.float64)
        _check(lib().qr_apply_trees(self.h, arr, _p(w, C.c_double), len(bufs)))

    def tree_contributions(self, trees):
        """mean |tree(doc)| over the dataset for each tree (Dart::update_contribution_scores, dart.cc:689-706)"""
        bufs = [TreeBuffer.from_dict(t) for t in trees]
        arr = (FlatTree * len(bufs))(*[b.t for b in bufs])
        out = np.zeros(len(bufs), np.float64)
        _check(lib().qr_tree_contributions(self.h, arr, len(bufs), _p(out, C.c_double)))
        return out

    def evaluate_dataset(self):
        m = C.c_double()
        _check(lib().qr_evaluate(self.h, C.byref(m)))
        return m.value

    def boost_iteration(self, want_tree=True, want_metric=True):
        tb = TreeBuffer(self.max_nodes) if want_tree else None
        m = C.c_double()
        _check(lib().qr_boost_iteration(self.h, C.byref(tb.t) if tb else None,
                                        C.byref(m) if want_metric else None))
        return (tb.to_dict() if tb else None), (m.value if want_metric else None)

    # parity taps
    def get_scores(self):
        s = np.empty(self.N, np.float64)
        _check(lib().qr_get_scores(self.h, _p(s, C.c_double)))
        return s

    def set_scores(self, s):
        s = np.ascontiguousarray(s, np.float64)
        _check(lib().qr_set_scores(self.h, _p(s, C.c_double)))

    def get_pseudoresponses(self):
        lam, w = np.empty(self.N, np.float64), np.empty(self.N, np.float64)
        _check(lib().qr_get_pseudoresponses(self.h, _p(lam, C.c_double), _p(w, C.c_double)))
        return lam, w

    def set_pseudoresponses(self, lam, w=None):
        lam = np.ascontiguousarray(lam, np.float64)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
            wp = _p(w, C.c_double)
        _check(lib().qr_set_pseudoresponses(self.h, _p(lam, C.c_double), wp))

    def get_leaf_assignment(self):
        a = np.empty(self.N, np.uint32)
        _check(lib().qr_get_leaf_assignment(self.h, _p(a, C.c_uint32)))
        return a

    def get_bins(self, f):
        a = np.empty(self.N, np.uint32)
        _check(lib().qr_get_bins(self.h, f, _p(a, C.c_uint32)))
        return a

    def get_ranking(self):
        a = np.empty(self.N, np.uint32)
        _check(lib().qr_get_ranking(self.h, _p(a, C.c_uint32)))
        return a

    def last_tree_stats(self):
        rho, sigma, ns = C.c_double(), C.c_double(), C.c_uint32()
        _check(lib().qr_last_tree_stats(self.h, C.byref(rho), C.byref(sigma), C.byref(ns)))
        return rho.value, sigma.value, ns.value

    def last_tree_rounds(self):
        r, b = C.c_uint32(), C.c_double()
        _check(lib().qr_last_tree_rounds(self.h, C.byref(r), C.byref(b)))
        return r.value, b.value

    def launch_count(self):
        return int(lib().qr_launch_count(self.h))

    def comm_transport(self) -> str:
        """How the per-round histogram exchange travels: 'none' (one GPU), 'nccl', 'peer' (one
        peer-memory kernel per round over NVLink)."""
        return ("none", "nccl", "peer")[int(lib().qr_ctx_comm_transport(self.h))]

    def set_profiling(self, on):
        _check(lib().qr_set_profiling(self.h, int(on)))

    def timer_start(self):
        _check(lib().qr_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _check(lib().qr_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def hist_kernel_time(self, reset=False):
        """(ms, launches, documents) of the histogram kernel since the last reset (profiling mode only)."""
        ms, ln, docs = C.c_double(), C.c_uint64(), C.c_double()
        _check(lib().qr_hist_kernel_time(self.h, C.byref(ms), C.byref(ln), C.byref(docs), int(reset)))
        return ms.value, int(ln.value), docs.value

    def phase_times(self, reset=False):
        ms = (C.c_double * 6)()
        ln = (C.c_uint64 * 6)()
        _check(lib().qr_phase_times(self.h, ms, ln, int(reset)))
        return dict(zip(PHASES, list(ms))), dict(zip(PHASES, [int(v) for v in ln]))


class LineSearchDevice:
    """The device side of a line search over a row-major score matrix [N][T] (qr_ls_*): weighted sums per document
    and NDCG@cutoff of the rankings they induce, in the reference's arithmetic (line_search.cc)."""

    def __init__(self, x, labels, qoffsets, cutoff=10, device=-1):
        x = np.ascontiguousarray(x, np.float32)
        labels = np.ascontiguousarray(labels, np.float32)
        qoffsets = np.ascontiguousarray(qoffsets, np.uint64)
        self.N, self.T = x.shape
        self.h = C.c_void_p()
        _check(lib().qr_ls_create(_p(x, C.c_float), self.N, self.T, _p(labels, C.c_float), _p(qoffsets, C.c_uint64),
                                  len(qoffsets) - 1, cutoff, device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().qr_ls_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def evaluate(self, weights):
        w = np.ascontiguousarray(weights, np.float64)
        m = C.c_double()
        _check(lib().qr_ls_evaluate(self.h, _p(w, C.c_double), C.byref(m)))
        return m.value

    def feature_points(self, weights, f, points):
        w = np.ascontiguousarray(weights, np.float64)
        pts = np.ascontiguousarray(points, np.float64)
        out = np.zeros(len(pts), np.float64)
        _check(lib().qr_ls_feature_points(self.h, _p(w, C.c_double), int(f), _p(pts, C.c_double), len(pts), _p(out, C.c_double)))
        return out

    def line_points(self, weights, step, npoints):
        w = np.ascontiguousarray(weights, np.float64)
        st = np.ascontiguousarray(step, np.float64)
        out = np.zeros(npoints, np.float64)
        _check(lib().qr_ls_line_points(self.h, _p(w, C.c_double), _p(st, C.c_double), npoints, _p(out, C.c_double)))
        return out

    def drop_points(self, weights, cols):
        """metric of the ensemble without column c, for every c in cols (qr_ls_drop_points)"""
        w = np.ascontiguousarray(weights, np.float64)
        cols = np.ascontiguousarray(cols, np.uint32)
        out = np.zeros(len(cols), np.float64)
        _check(lib().qr_ls_drop_points(self.h, _p(w, C.c_double), _p(cols, C.c_uint32), len(cols), _p(out, C.c_double)))
        return out

    def drop_column(self, weights, f):
        """the device's running sums lose column f in place (qr_ls_drop_column); continue with weights[f] = 0"""
        w = np.ascontiguousarray(weights, np.float64)
        _check(lib().qr_ls_drop_column(self.h, _p(w, C.c_double), int(f)))

    def score_loss(self, weights):
        """per column: sum over documents of the column's share of the score (qr_ls_score_loss)"""
        w = np.ascontiguousarray(weights, np.float64)
        out = np.zeros(self.T, np.float64)
        _check(lib().qr_ls_score_loss(self.h, _p(w, C.c_double), _p(out, C.c_double)))
        return out

    def launch_count(self):
        return int(lib().qr_ls_launch_count(self.h))


def selftest_ordered_squares(values, fused, device=-1):
    """(parallel, serial, replayed_chunks): sum of values**2 in index order by REFERENCE mode's parallel scheme and
    by the plain chain (qr_selftest_ordered_squares)."""
    v = np.ascontiguousarray(values, np.float64)
    par, ser, rep = C.c_double(), C.c_double(), C.c_uint64()
    _check(lib().qr_selftest_ordered_squares(_p(v, C.c_double), len(v), 1 if fused else 0, device, C.byref(par),
                                             C.byref(ser), C.byref(rep)))
    return par.value, ser.value, rep.value


def comm_unique_id() -> bytes:
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(lib().qr_comm_unique_id(buf))
    return bytes(buf)


class Scorer:
    """An ensemble resident on one GPU (the quickscore path)."""

    def __init__(self, trees, weights, n_features, device=-1, condop_weights=False):
        """condop_weights: use the tree weights of the ranker() the reference's conditional-operator generator
        emits (floats printed with three decimals, generate_conditional_operators.cc:95-105)."""
        L = lib()
        self.bufs = [TreeBuffer.from_dict(t) for t in trees]
        arr = (FlatTree * len(self.bufs))(*[b.t for b in self.bufs])
        w = np.ascontiguousarray(weights, np.float64)
        self.F = n_features
        self.ntrees = len(self.bufs)
        self.h = C.c_void_p()
        _check(L.qr_scorer_create_ex(arr, _p(w, C.c_double), len(self.bufs), n_features, device,
                                     1 if condop_weights else 0, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().qr_scorer_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def score_dataset(self, x, out=None):
        """Row-major host documents -> scores (LTR_Algorithm::score_dataset).  `out` may be a
        caller-owned float64 array (e.g. page-locked) to receive the scores."""
        x = np.ascontiguousarray(x, np.float32)
        if out is None:
            out = np.empty(x.shape[0], np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape[0] == x.shape[0]
        _check(lib().qr_score_dataset(self.h, _p(x, C.c_float), x.shape[0], x.shape[1],
                                      _p(out, C.c_double)))
        return out

    def partial_scores(self, x, with_scores=False):
        """Row-major host documents -> the per-tree score matrix [N][ntrees] (float32) of
        Driver::extract_partial_scores (driver.cc:411-446); with_scores: also the ensemble scores."""
        x = np.ascontiguousarray(x, np.float32)
        part = np.empty((x.shape[0], self.ntrees), np.float32)
        sc = np.empty(x.shape[0], np.float64) if with_scores else None
        _check(lib().qr_score_partial(self.h, _p(x, C.c_float), x.shape[0], x.shape[1], _p(part, C.c_float),
                                      _p(sc, C.c_double) if with_scores else None))
        return (part, sc) if with_scores else part

    def score_dataset_device(self, docs_ptr, n, scores_ptr):
        _check(lib().qr_score_dataset_device(self.h, C.c_void_p(docs_ptr), n, self.F,
                                             C.c_void_p(scores_ptr)))

    def sync(self):
        _check(lib().qr_scorer_sync(self.h))

    def launch_count(self):
        return int(lib().qr_scorer_launch_count(self.h))

    def timer_start(self):
        _check(lib().qr_scorer_timer(self.h, 0, None))

    def timer_stop(self):
        ms = C.c_double()
        _check(lib().qr_scorer_timer(self.h, 1, C.byref(ms)))
        return ms.value

    def score_document(self, d):
        d = np.ascontiguousarray(d, np.float32)
        s = C.c_double()
        _check(lib().qr_score_document(self.h, _p(d, C.c_float), len(d), C.byref(s)))
        return s.value
