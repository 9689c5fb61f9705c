"""Line search and CLEAVER over a score matrix: the host side.

``LineSearch.learn`` follows ``LineSearch::learn`` of the reference (src/learning/linear/line_search.cc:153-416)
decision for decision — window, candidate points, first-maximum acceptance, the two steps of an iteration — while
every pass over the documents (weighted sums, rankings, NDCG@k of each candidate) runs on the GPU through
``api.LineSearchDevice`` (quickrank_b200/csrc/qr_linesearch.cu), in the reference's own arithmetic: given the same
matrix the learned weights equal the reference's bit for bit (tests/test_linesearch.py, against oracle/_ref).

``Cleaver.optimize`` is ``Cleaver::optimize`` (src/optimization/post_learning/cleaver/cleaver.cc:166-412) with the
pruning strategies LAST, SKIP, LOW_WEIGHTS (last_pruning.cc, skip_pruning.cc, low_weights_pruning.cc), QUALITY_LOSS
(quality_loss_pruning.cc: NDCG of the ensemble without each tree, 32 trees per ranking launch), QUALITY_LOSS_ADV
(quality_loss_adv_pruning.cc: the same greedily, one tree at a time, on the reference's running score vector),
SCORE_LOSS (score_loss_pruning.cc: each tree's summed share of the document scores, one ordered FP64 chain per tree on
the device) — all pinned bit for bit against the unmodified reference — and RANDOM (random_pruning.cc; the reference
seeds rand() from the wall clock, here from ``seed``).  RANDOM_ADV is not provided (its OpenMP loop shares one rand()
stream and one never-reset candidate set between threads: the result is not a function of its inputs).
The matrix is what ``api.Scorer.partial_scores`` returns for an ensemble (driver.cc:411-446).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import api

_libm = C.CDLL("libm.so.6")
_libm.fma.restype = C.c_double
_libm.fma.argtypes = [C.c_double, C.c_double, C.c_double]


def _fma(a, b, c):
    return _libm.fma(float(a), float(b), float(c))


class LineSearch:
    """Parameters as in the reference's constructor (include/learning/linear/line_search.h:41-44)."""

    def __init__(self, num_points=20, window_size=1.0, reduction_factor=0.95, max_iterations=5, max_failed_vali=20,
                 adaptive=False, last_only=0):
        self.num_points = int(num_points)
        self.window_size = float(window_size)
        self.reduction_factor = float(reduction_factor)
        self.max_iterations = int(max_iterations)
        self.max_failed_vali = int(max_failed_vali)
        self.adaptive = bool(adaptive)
        self.last_only = int(last_only)
        self.weights = None          # best_weights_
        self.history = []            # (iteration, training metric, validation metric or None, gain, window)

    def learn(self, train: api.LineSearchDevice, valid: api.LineSearchDevice | None = None):
        """train / valid: device-resident score matrices.  Returns the learned weights (also kept in self.weights)."""
        T = train.T
        num_points = self.num_points - 1 if self.num_points % 2 else self.num_points   # line_search.cc:163-165
        if self.weights is None:
            self.weights = np.ones(T, np.float64)
        elif len(self.weights) != T:
            raise ValueError("initial line search weights do not correspond to the dataset's size")
        weights = np.array(self.weights, np.float64)
        weights_prev = weights.copy()
        best_train = train.evaluate(weights)
        best_valid = valid.evaluate(weights) if valid is not None else 0.0
        self.history = [(0, best_train, best_valid if valid is not None else None, 0.0, 0.0)]
        starting_window = _mean_seq(self.weights)                                   # line_search.cc:234-238
        window = starting_window * self.window_size
        first = T - self.last_only if self.last_only else 0
        failed = 0
        for it in range(self.max_iterations):
            step1 = 2 * window / num_points
            for f in range(first, T):                       # step 1: every weight on its own
                points = []
                point = weights_prev[f] - window
                while point <= weights_prev[f] + window:
                    if point >= 0:
                        points.append(point)
                    point += step1
                if not points:
                    continue
                metrics = train.feature_points(weights_prev, f, points)
                p = int(np.argmax(metrics))                 # std::max_element: the first maximum
                if metrics[p] > best_train:
                    weights[f] = points[p]
            step2 = (weights - weights_prev) / num_points   # step 2: along the line from weights_prev to weights
            gain = 0.0
            if np.any(step2 != 0):
                metrics = train.line_points(weights_prev, step2, num_points + 1)
                p = int(np.argmax(metrics))
                if metrics[p] > best_train:
                    weights = np.array([_fma(step2[f], float(p), weights_prev[f]) for f in range(T)], np.float64)
                    gain = metrics[p] - best_train
                    best_train = metrics[p]
                    weights_prev = weights.copy()
            factor = self.reduction_factor
            if self.adaptive:                               # line_search.cc:349-358
                max_gain = 0.005
                relative_gain = min((gain - max_gain) / max_gain, 1.0)
                factor = 1 + max(relative_gain, -0.5)
            stop = False
            vmetric = None
            if valid is not None:
                vmetric = valid.evaluate(weights)
                if vmetric > best_valid:
                    failed = 0
                    best_valid = vmetric
                    self.weights = weights.copy()
                else:
                    failed += 1
                    stop = failed >= self.max_failed_vali
            self.history.append((it + 1, best_train, vmetric, gain, window))
            if stop:
                break
            window *= factor
            if self.adaptive and window < starting_window / 10:
                break
        if valid is None:
            self.weights = weights.copy()
        self.metric_on_training = best_train
        return self.weights


def _mean_seq(w):
    acc = 0.0
    for v in w:                                             # std::accumulate, left to right
        acc += float(v)
    return acc / len(w)


def std_sort(a, less):
    """libstdc++'s std::sort (introsort: median-of-3 quicksort to depth 2*floor(log2 n), heapsort below it, one final
    insertion sort; bits/stl_algo.h) on a Python list, in place: the order of elements that compare equal is the one
    the reference's `std::sort(idx...)` calls produce (low_weights_pruning.cc:47-51, quality_loss_pruning.cc:79-84)."""
    n = len(a)
    if n < 2:
        return a

    def swap(i, j):
        a[i], a[j] = a[j], a[i]

    def unguarded_linear_insert(last):
        val = a[last]
        nxt = last - 1
        while less(val, a[nxt]):
            a[last] = a[nxt]
            last = nxt
            nxt -= 1
        a[last] = val

    def insertion_sort(first, last):
        for i in range(first + 1, last):
            if less(a[i], a[first]):
                val = a[i]
                a[first + 1:i + 1] = a[first:i]
                a[first] = val
            else:
                unguarded_linear_insert(i)

    def adjust_heap(first, hole, length, value):
        top = hole
        child = hole
        while child < (length - 1) // 2:
            child = 2 * (child + 1)
            if less(a[first + child], a[first + child - 1]):
                child -= 1
            a[first + hole] = a[first + child]
            hole = child
        if (length & 1) == 0 and child == (length - 2) // 2:
            child = 2 * (child + 1)
            a[first + hole] = a[first + child - 1]
            hole = child - 1
        parent = (hole - 1) // 2                       # __push_heap
        while hole > top and less(a[first + parent], value):
            a[first + hole] = a[first + parent]
            hole = parent
            parent = (hole - 1) // 2
        a[first + hole] = value

    def heapsort(first, last):                         # std::__partial_sort(first, last, last)
        length = last - first
        if length >= 2:                                # __make_heap
            parent = (length - 2) // 2
            while True:
                adjust_heap(first, parent, length, a[first + parent])
                if parent == 0:
                    break
                parent -= 1
        while last - first > 1:                        # __sort_heap
            last -= 1
            value = a[last]
            a[last] = a[first]
            adjust_heap(first, 0, last - first, value)

    def move_median_to_first(result, x, y, z):
        if less(a[x], a[y]):
            if less(a[y], a[z]):
                swap(result, y)
            elif less(a[x], a[z]):
                swap(result, z)
            else:
                swap(result, x)
        elif less(a[x], a[z]):
            swap(result, x)
        elif less(a[y], a[z]):
            swap(result, z)
        else:
            swap(result, y)

    def introsort_loop(first, last, depth):
        while last - first > 16:
            if depth == 0:
                heapsort(first, last)
                return
            depth -= 1
            mid = first + (last - first) // 2
            move_median_to_first(first, first + 1, mid, last - 1)
            lo, hi, pivot = first + 1, last, first     # __unguarded_partition
            while True:
                while less(a[lo], a[pivot]):
                    lo += 1
                hi -= 1
                while less(a[pivot], a[hi]):
                    hi -= 1
                if not lo < hi:
                    break
                swap(lo, hi)
                lo += 1
            introsort_loop(lo, last, depth)
            last = lo

    introsort_loop(0, n, 2 * (n.bit_length() - 1))
    if n > 16:
        insertion_sort(0, 16)
        for i in range(16, n):
            unguarded_linear_insert(i)
    else:
        insertion_sort(0, n)
    return a


PRUNING_METHODS = ("LAST", "SKIP", "LOW_WEIGHTS", "QUALITY_LOSS", "QUALITY_LOSS_ADV", "SCORE_LOSS", "RANDOM")
_PRE_PRUNING_LS = {"LAST": False, "SKIP": False, "LOW_WEIGHTS": True, "QUALITY_LOSS": True, "QUALITY_LOSS_ADV": True,
                   "SCORE_LOSS": True, "RANDOM": False}


class Cleaver:
    """Ensemble pruning + re-weighting (cleaver.cc).  pruning_rate < 1: a fraction of the trees, else a count."""

    def __init__(self, pruning_rate, method="QUALITY_LOSS", line_search: LineSearch | None = None, last_only=0, seed=0):
        if method not in PRUNING_METHODS:
            raise ValueError("pruning method %s is not supported (supported: %s)" % (method, ", ".join(PRUNING_METHODS)))
        self.pruning_rate = float(pruning_rate)
        self.method = method
        self.line_search = line_search
        self.last_only = int(last_only)
        self.seed = int(seed)
        self.weights = None
        self.pruned = set()

    def _prune(self, dev: api.LineSearchDevice, weights, last, to_prune):
        T = dev.T
        start_last = T - last
        if self.method == "LAST":
            return {T - i for i in range(1, to_prune + 1)}
        if self.method == "SKIP":
            to_select = last - to_prune
            step = last / to_select
            selected = {int(math.ceil(step * i + start_last)) for i in range(to_select)}
            return {f for f in range(start_last, T) if f not in selected}
        if self.method == "LOW_WEIGHTS":
            idx = std_sort(list(range(start_last, T)), lambda a, b: weights[a] < weights[b])
            return set(idx[:to_prune])
        if self.method == "RANDOM":   # random_pruning.cc:47-55 (srand(time(NULL)) there)
            libc = C.CDLL("libc.so.6")
            libc.srand(C.c_uint(self.seed & 0xffffffff))
            pruned = set()
            while len(pruned) < to_prune:
                pruned.add(libc.rand() % last + start_last)
            return pruned
        if self.method == "SCORE_LOSS":   # score_loss_pruning.cc:58-77: the trees with the smallest summed share go
            loss = dev.score_loss(weights)
            idx = std_sort(list(range(start_last, T)), lambda a, b: loss[a] < loss[b])
            return set(idx[:to_prune])
        if self.method == "QUALITY_LOSS_ADV":
            # quality_loss_adv_pruning.cc:58-93: one tree per step — the one whose removal leaves the best metric
            # (std::max_element: the first maximum) — on running scores that lose the pruned column in place
            w, pruned = np.array(weights, np.float64), set()
            for _ in range(to_prune):
                cand = [f for f in range(start_last, T) if f not in pruned]
                metrics = dev.drop_points(w, cand)
                f_prune = cand[int(np.argmax(metrics))]   # (np.argmax: first maximum; pruned trees hold `lowest` there)
                pruned.add(f_prune)
                dev.drop_column(w, f_prune)
                w[f_prune] = 0.0
            return pruned
        # QUALITY_LOSS: the metric of the ensemble without tree f, for every f; the trees whose removal hurts least go
        metrics = dev.drop_points(weights, list(range(start_last, T)))
        idx = std_sort(list(range(start_last, T)), lambda a, b: metrics[a - start_last] > metrics[b - start_last])
        return set(idx[:to_prune])

    def optimize(self, x, labels, qoffsets, weights, cutoff=10, device=-1):
        """x: the per-tree partial-score matrix [N][T] (unit weights: Driver::extract_partial_scores with
        ignore_weights), weights: the ensemble's weights.  Returns (new weights with 0 for pruned trees, pruned set)."""
        x = np.ascontiguousarray(x, np.float32)
        T = x.shape[1]
        last, opt_last_only = (self.last_only, True) if self.last_only else (T, False)
        to_prune = int(round(self.pruning_rate * last)) if self.pruning_rate < 1 else int(self.pruning_rate)
        if to_prune >= last:
            raise ValueError("incorrect pruning rate value (too high)")
        w = np.array(weights, np.float64)
        start = w.copy()
        ls = self.line_search
        with api.LineSearchDevice(x, labels, qoffsets, cutoff, device) as dev:
            self.metric_before = dev.evaluate(w)
            if _PRE_PRUNING_LS[self.method] and to_prune > 0 and ls is not None:
                ls.last_only = last if opt_last_only else 0
                ls.weights = w.copy()
                w = ls.learn(dev).copy()
            pruned = self._prune(dev, w, last, to_prune) if to_prune > 0 else set()
            w = start.copy()
            for f in pruned:
                w[f] = 0.0
        if ls is not None:
            keep = [f for f in range(T) if f not in pruned]
            ls.weights = w[keep].copy()
            ls.last_only = last - to_prune if opt_last_only else 0
            with api.LineSearchDevice(np.ascontiguousarray(x[:, keep]), labels, qoffsets, cutoff, device) as dev:
                lw = ls.learn(dev)
            w[keep] = lw
        with api.LineSearchDevice(x, labels, qoffsets, cutoff, device) as dev:
            self.metric_after = dev.evaluate(w)
        self.weights, self.pruned = w, pruned
        return w, pruned
