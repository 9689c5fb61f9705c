"""TEST INFRASTRUCTURE (CPU suite only): a numpy stand-in for api.LineSearchDevice, so that the HOST logic of
quickrank_b200/linesearch.py (LineSearch.learn, Cleaver.optimize: windows, points, first-maximum acceptance, pruning
strategies) is exercised against the unmodified reference without a GPU.  The passes over documents are done here in
the reference's arithmetic (what quickrank_b200/csrc/qr_linesearch.cu does on the device: separate multiply and add in
feature order for the totals, multiply-then-subtract for the dropped column, fma where the reference's build fuses)
and NDCG comes from the oracle.  Never imported by the product; tests monkeypatch it in."""
import ctypes as C

import numpy as np

from oracle import pyoracle as po

_libm = C.CDLL("libm.so.6")
_libm.fma.restype = C.c_double
_libm.fma.argtypes = [C.c_double, C.c_double, C.c_double]
_fma = np.vectorize(lambda a, b, c: _libm.fma(a, b, c), otypes=[np.float64])


class CpuLineSearchDevice:
    def __init__(self, x, labels, qoffsets, cutoff=10, device=-1):
        self.x = np.ascontiguousarray(x, np.float32).astype(np.float64)     # (double) x[s][f], exact
        self.labels = np.ascontiguousarray(labels, np.float32)
        self.qoff = np.ascontiguousarray(qoffsets, np.uint64)
        self.cutoff = cutoff
        self.N, self.T = self.x.shape
        self.calls = 0
        self._total_w, self._total = None, None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass

    def launch_count(self):
        return self.calls

    def _metric(self, scores):
        self.calls += 1
        return po.ndcg_dataset(self.labels, scores, self.qoff, self.cutoff)

    def _totals(self, w):
        w = np.asarray(w, np.float64)
        if self._total_w is not None and np.array_equal(self._total_w, w):
            return self._total
        acc = np.zeros(self.N)
        for f in range(self.T):                      # line_search.cc:447-482: acc += w[f] * x[s][f], from 0
            acc = acc + w[f] * self.x[:, f]
        self._total_w, self._total = w.copy(), acc
        return acc

    def evaluate(self, weights):
        return self._metric(self._totals(weights))

    def feature_points(self, weights, f, points):
        pre = self._totals(weights) - float(weights[f]) * self.x[:, f]
        return np.array([self._metric(_fma(float(p), self.x[:, f], pre)) for p in points])

    def line_points(self, weights, step, npoints):
        out = []
        for p in range(npoints):
            acc = np.zeros(self.N)
            for f in range(self.T):
                acc = _fma(_libm.fma(float(step[f]), float(p), float(weights[f])), self.x[:, f], acc)
            out.append(self._metric(acc))
        return np.array(out)

    def drop_points(self, weights, cols):
        tot = self._totals(weights)
        return np.array([self._metric(tot - float(weights[c]) * self.x[:, c]) for c in cols])

    def drop_column(self, weights, f):
        tot = self._totals(weights)
        self._total = tot - float(weights[f]) * self.x[:, f]
        self._total_w = np.asarray(weights, np.float64).copy()
        self._total_w[f] = 0.0

    def score_loss(self, weights):
        tot = self._totals(weights)
        out = np.zeros(self.T)
        for f in range(self.T):
            acc = 0.0
            for v in (float(weights[f]) * self.x[:, f]) / tot:       # one ordered chain per column
                acc += v
            out[f] = acc
        return out
