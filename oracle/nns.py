"""Oracle for cosine nearest-neighbour matching (TEST INFRASTRUCTURE ONLY).

Restates cslam/nns_matching.py:6-76 (class NearestNeighborsMatching) and the
part of scipy.spatial.distance.cosine -> correlation(centered=False) it calls
(scipy >= 1.8, `dist = 1.0 - uv / math.sqrt(uu * vv); np.clip(dist, 0, 2)`).

Two scorers are provided:
  * `search_loop`  — the reference's per-row Python loop, literally
    (nns_matching.py:55-61); used for small cases and to pin `search_vec`.
  * `search_vec`   — the same arithmetic vectorised (float64 matrix-vector
    product; x.x accumulated per row by np.dot on the float32 row exactly as
    correlation() does).  Used for large pools and as the timed CPU baseline.
"""
import math

import numpy as np


class NNSOracle(object):
    def __init__(self, dim=None):
        # nns_matching.py:10-21
        self.n = 0
        self.dim = dim
        self.items = dict()
        self.data = []
        if dim is not None:
            self.data = np.zeros((1000, dim), dtype='float32')
        self._vv = None  # cached float32 squared norms for search_vec

    def add_item(self, vector, item):
        # nns_matching.py:23-40 (float32 store, capacity 1000 -> x2)
        vector = np.asarray(vector)
        assert vector.ndim == 1
        if self.n >= len(self.data):
            if self.dim is None:
                self.dim = len(vector)
                self.data = np.zeros((1000, self.dim), dtype='float32')
            else:
                new = np.zeros((2 * len(self.data), self.dim), dtype='float32')
                new[:self.n] = self.data[:self.n]
                self.data = new
        self.items[self.n] = item
        self.data[self.n] = vector
        self.n += 1
        self._vv = None

    def add_items(self, vectors, items):
        for v, it in zip(vectors, items):
            self.add_item(v, it)

    @staticmethod
    def _cosine_distance(u, v):
        # scipy.spatial.distance.correlation(u, v, centered=False)
        uv = np.dot(u, v)
        uu = np.dot(u, u)
        vv = np.dot(v, v)
        dist = 1.0 - uv / math.sqrt(uu * vv)
        return np.clip(dist, 0.0, 2.0)

    def similarities_loop(self, query):
        # nns_matching.py:55-58
        query = np.asarray(query)
        sims = np.zeros(self.n)
        for i in range(self.n):
            sims[i] = 1 - self._cosine_distance(query, self.data[i, :].squeeze())
        return sims

    def similarities_vec(self, query):
        query = np.asarray(query)
        d = self.data[:self.n]
        if self._vv is None:
            # np.dot(v, v) on each float32 row, as correlation() computes vv
            self._vv = np.array([np.dot(r, r) for r in d], dtype=np.float32)
        if query.dtype == np.float32:
            # float32 query: uv, uu, uu*vv stay float32 (numpy promotion rules)
            uv = d @ query
            uu = np.dot(query, query)
            den = np.sqrt((uu * self._vv).astype(np.float64))
            dist = (1.0 - (uv / den.astype(np.float64)).astype(np.float32)).astype(np.float64)
        else:
            q = query.astype(np.float64)
            uv = d.astype(np.float64) @ q
            uu = np.dot(q, q)
            dist = 1.0 - uv / np.sqrt(uu * self._vv.astype(np.float64))
        return 1 - np.clip(dist, 0.0, 2.0)

    def _rank(self, sims, k):
        # nns_matching.py:60-61
        ns = np.argsort(sims)[::-1][:k]
        return [self.items[n] for n in ns], sims[ns], ns

    def search_loop(self, query, k):
        if len(self.data) == 0:
            return [], []
        items, sims, _ = self._rank(self.similarities_loop(query), k)
        return items, sims

    def search_vec(self, query, k):
        if len(self.data) == 0:
            return [], []
        items, sims, _ = self._rank(self.similarities_vec(query), k)
        return items, sims

    search = search_vec

    def search_best(self, query):
        # nns_matching.py:63-76
        if len(self.data) == 0:
            return None, None
        items, sims = self.search(query, 1)
        return items[0], sims[0]


def lists_match_modulo_ties(idx_a, idx_b, sims_full, tol=1e-6):
    """True if the two ranked id lists are identical except for swaps among
    entries whose oracle similarities differ by < tol (the tie tolerance of
    the reference's own test, tests/test_sparse_matching.py:71-80)."""
    if len(idx_a) != len(idx_b):
        return False
    sims_full = np.asarray(sims_full)
    # a NaN / inf similarity compares False with everything and would let any list through
    if not np.isfinite(sims_full).all():
        return False
    for a, b in zip(idx_a, idx_b):
        if a != b and not (abs(sims_full[a] - sims_full[b]) < tol):
            return False
    return True
