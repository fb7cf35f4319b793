"""NearestNeighborsMatching — same class API as the reference
(cslam/nns_matching.py:6-76), arithmetic on the GPU through libcslam_b200.

Differences a caller can observe (all documented in DESIGN.md):
  * the pool lives in HBM; `.data` is a property that downloads it (float32,
    reference capacity semantics 1000 -> x2, nns_matching.py:21,31-37);
  * `search_batch` / `add_items` are batched extensions (the reference API is
    one vector per call);
  * exact ties are returned by descending row id.
"""
import ctypes

import numpy as np

from . import _lib


class NearestNeighborsMatching(object):
    """Nearest Neighbor matching of description vectors (GPU resident)."""

    def __init__(self, dim=None, device=None):
        """
        Args:
            dim (int, optional): Global descriptor size. Defaults to None
                (taken from the first added vector, nns_matching.py:31-34).
            device (int, optional): CUDA ordinal; default = current torch device or 0.
        """
        self.n = 0
        self.dim = dim
        self.items = dict()
        self._device = device
        self._h = None
        self._capacity = 0
        if dim is not None:
            self._create(dim)
            self._capacity = 1000

    # -- handle management -------------------------------------------------
    def _resolve_device(self):
        if self._device is not None:
            return int(self._device)
        try:
            import torch
            if torch.cuda.is_available():
                return int(torch.cuda.current_device())
        except Exception:
            pass
        return 0

    def _create(self, dim):
        lib = _lib.load()
        _lib.require_device()
        h = ctypes.c_void_p()
        _lib.check(lib.cslam_nns_create(int(dim), self._resolve_device(), ctypes.byref(h)))
        self._h = h
        self.dim = int(dim)

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().cslam_nns_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- reference attributes ------------------------------------------------
    @property
    def data(self):
        """float32 [capacity, dim] copy of the pool (rows >= n are zero), like
        the reference's `self.data`; `[]` before the first add when dim is None."""
        if self._h is None:
            return []
        out = np.zeros((self._capacity, self.dim), dtype=np.float32)
        if self.n > 0:
            _lib.check(_lib.load().cslam_nns_read_rows(self._h, 0, self.n, _lib.ptr(out)))
        return out

    def read_rows(self, start, count):
        """float32 [count, dim] copy of pool rows [start, start + count)."""
        out = np.empty((int(count), self.dim), dtype=np.float32)
        if count > 0:
            _lib.check(_lib.load().cslam_nns_read_rows(self._h, int(start), int(count), _lib.ptr(out)))
        return out

    # -- reference methods ---------------------------------------------------
    def add_item(self, vector, item):
        """Add item to the matching list (nns_matching.py:23-40)."""
        vector = np.asarray(vector)
        assert vector.ndim == 1
        if self._h is None:
            self._create(len(vector))
        if self.n >= self._capacity:
            self._capacity = 1000 if self._capacity == 0 else 2 * self._capacity
        if len(vector) != self.dim:
            raise ValueError(
                f"could not broadcast input array from shape ({len(vector)},) into shape ({self.dim},)")
        self.items[self.n] = item
        self._add_rows(vector.reshape(1, -1))
        self.n += 1

    def add_items(self, vectors, items):
        """Batched extension: append rows of a [m, dim] array."""
        vectors = np.asarray(vectors)
        assert vectors.ndim == 2 and len(items) == vectors.shape[0]
        if self._h is None:
            self._create(vectors.shape[1])
        for j, it in enumerate(items):
            self.items[self.n + j] = it
        self._add_rows(vectors)
        self.n += vectors.shape[0]
        while self.n > self._capacity:
            self._capacity = 1000 if self._capacity == 0 else 2 * self._capacity

    def _add_rows(self, rows):
        if rows.dtype == np.float32:
            dt = _lib.DTYPE_F32
        else:
            rows = rows.astype(np.float64, copy=False)
            dt = _lib.DTYPE_F64
        rows = np.ascontiguousarray(rows)
        _lib.check(_lib.load().cslam_nns_add_host(self._h, _lib.ptr(rows), dt, rows.shape[0]))

    def search_batch(self, queries, k):
        """Batched search: queries [nq, dim] -> (idx int32 [nq, k'], sims float64 [nq, k'])
        with k' = min(k, n); idx are pool row ids (use `.items` to map)."""
        queries = np.asarray(queries)
        assert queries.ndim == 2
        nq = queries.shape[0]
        kk = min(int(k), self.n)
        if self._h is None or kk <= 0 or nq == 0:
            return np.zeros((nq, 0), dtype=np.int32), np.zeros((nq, 0), dtype=np.float64)
        if queries.dtype == np.float32:
            dt = _lib.DTYPE_F32
        else:
            queries = queries.astype(np.float64, copy=False)
            dt = _lib.DTYPE_F64
        queries = np.ascontiguousarray(queries)
        idx = np.empty((nq, kk), dtype=np.int32)
        sims = np.empty((nq, kk), dtype=np.float64)
        info = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.load().cslam_nns_search_host(self._h, _lib.ptr(queries), dt, nq, kk,
                                                     _lib.ptr(idx), _lib.ptr(sims),
                                                     _lib.ptr(info)))
        self.last_info = info
        return idx, sims

    def search(self, query, k):
        """Search for nearest neighbors (nns_matching.py:42-61).

        Returns:
            list, np.array: item ids of the best matches and their similarities
        """
        if self._capacity == 0:  # reference: `if len(self.data) == 0`
            return [], []
        query = np.asarray(query)
        if self.n == 0 or k <= 0:
            return [], np.zeros(0)
        idx, sims = self.search_batch(query.reshape(1, -1), k)
        return [self.items[int(n)] for n in idx[0]], sims[0]

    def search_best(self, query):
        """Search for the nearest neighbor (nns_matching.py:63-76)."""
        if self._capacity == 0:
            return None, None
        items, similarities = self.search(query, 1)
        return items[0], similarities[0]

    # -- device-resident extensions (torch CUDA tensors; no host copies) ---------
    def add_items_device(self, rows, items=None):
        """Append a float32 CUDA tensor [m, dim]; items default to the row ids."""
        import torch
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2
        rows = rows.contiguous()
        if self._h is None:
            self._device = rows.device.index
            self._create(rows.shape[1])
        m = rows.shape[0]
        if items is None:
            self.items.update((self.n + j, self.n + j) for j in range(m))
        else:
            self.items.update((self.n + j, it) for j, it in enumerate(items))
        stream = torch.cuda.current_stream(rows.device).cuda_stream
        _lib.check(_lib.load().cslam_nns_add_device(self._h, _lib.ptr(rows), m,
                                                    ctypes.c_void_p(stream)))
        self.n += m
        while self.n > self._capacity:
            self._capacity = 1000 if self._capacity == 0 else 2 * self._capacity

    def search_batch_device(self, queries, k, out=None):
        """queries: CUDA tensor [nq, dim] float32/float64 -> (idx int32, sims float64)
        CUDA tensors [nq, min(k, n)], computed on torch's current stream."""
        import torch
        assert queries.is_cuda and queries.dim() == 2
        queries = queries.contiguous()
        dt = _lib.DTYPE_F32 if queries.dtype == torch.float32 else _lib.DTYPE_F64
        if dt == _lib.DTYPE_F64 and queries.dtype != torch.float64:
            queries = queries.double()
        nq = queries.shape[0]
        kk = min(int(k), self.n)
        if out is None:
            idx = torch.empty((nq, kk), dtype=torch.int32, device=queries.device)
            sims = torch.empty((nq, kk), dtype=torch.float64, device=queries.device)
        else:
            idx, sims = out
        info = np.zeros(4, dtype=np.int64)
        stream = torch.cuda.current_stream(queries.device).cuda_stream
        _lib.check(_lib.load().cslam_nns_search_device(self._h, _lib.ptr(queries), dt, nq, kk,
                                                       _lib.ptr(idx), _lib.ptr(sims),
                                                       ctypes.c_void_p(stream), _lib.ptr(info)))
        self.last_info = info
        return idx, sims

    def last_timing(self):
        """(coarse_ms, coarse_launches, total_ms) of the last search (CUDA events)."""
        c, n, t = ctypes.c_float(), ctypes.c_int(), ctypes.c_float()
        _lib.check(_lib.load().cslam_nns_last_timing(self._h, ctypes.byref(c), ctypes.byref(n),
                                                     ctypes.byref(t)))
        return c.value, n.value, t.value

    # -- tuning hooks ----------------------------------------------------------
    def set_mode(self, mode):
        """0 = tensor-core coarse pass + exact re-rank (default); 1 = exact fp64 scan."""
        _lib.check(_lib.load().cslam_nns_set_mode(self._h, int(mode)))

    def set_sample_rows(self, sample_rows=32768):
        _lib.check(_lib.load().cslam_nns_set_sample_rows(self._h, int(sample_rows)))
