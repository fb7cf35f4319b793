"""ScanContextMatching — lidar place recognition by Scan Context, on the GPU.

Same class, constructor, attributes and return conventions as the reference
(cslam/lidar_pr/scancontext_matching.py:6-104), so that
`LoopClosureSparseMatching` can use it wherever the reference does
(cslam/loop_closure_sparse_matching.py:21-22,28-29).  The pool (scan contexts stored column by
column, column norms, ring keys) lives in HBM behind `cslam_sc_*` (csrc/scancontext.cu):
ring-key nearest neighbours by an exhaustive fp64 scan, column-shift distance of every
candidate in one CTA each.  There is no CPU path.

Reference conventions kept on purpose:
  * `search` returns ONE match however large `k` is (:87-88), as two lists;
  * an empty pool answers `([None], [None])` / `(None, None)` (:55-56, :99-100);
  * when no candidate is closer than 1 the answer is item 0 with similarity 0.0 (:81-84);
  * `threshold` is stored and unused (:16).
Extensions: `add_items` / `search_batch` for whole batches.
"""
import ctypes

import numpy as np

from .. import _lib


class ScanContextMatching(object):
    """Pool of Scan Context descriptors with ring-key candidate search and column-shift
    scoring (reference class of the same name)."""

    def __init__(self, shape=[20, 60], num_candidates=10, threshold=0.15, device=None):
        """shape = [rings, sectors] of a descriptor, num_candidates = ring-key neighbours scored
        per query, threshold kept for signature compatibility (unused, as in the reference);
        defaults as in the reference (:10)."""
        lib = _lib.load()
        _lib.require_device()
        self.shape = shape
        self.num_candidates = num_candidates
        self.threshold = threshold
        self.items = dict()
        self.nb_items = 0
        if device is None:
            try:
                import torch
                device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            except Exception:
                device = 0
        self._device = int(device)
        h = ctypes.c_void_p()
        _lib.check(lib.cslam_sc_create(int(shape[0]), int(shape[1]), int(num_candidates),
                                       self._device, ctypes.byref(h)))
        self._h = h
        self.last_yaw_diff_deg = None     # of the last search() (`nn_yawdiff_deg`, :86)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().cslam_sc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- the reference's array attributes, read back from the device ----------------
    def _read(self, want_sc):
        lib = _lib.load()
        cap = int(lib.cslam_sc_capacity(self._h))
        R, S = int(self.shape[0]), int(self.shape[1])
        out = np.zeros((cap, R, S) if want_sc else (cap, R))
        if self.nb_items:
            live = np.empty((self.nb_items, R, S) if want_sc else (self.nb_items, R))
            _lib.check(lib.cslam_sc_read(self._h, 0, self.nb_items, _lib.ptr(live) if want_sc else None,
                                         None if want_sc else _lib.ptr(live)))
            out[:self.nb_items] = live
        return out

    @property
    def scancontexts(self):
        """[capacity, rings, sectors], zero beyond nb_items (:18)."""
        return self._read(True)

    @property
    def ringkeys(self):
        """[capacity, rings] (:19)."""
        return self._read(False)

    # ---- pool ---------------------------------------------------------------------------
    @staticmethod
    def _as_rows(descriptors, cells):
        a = np.asarray(descriptors)
        if a.dtype != np.float32:
            a = a.astype(np.float64, copy=False)
        a = np.ascontiguousarray(a.reshape(-1, cells))
        return a, (_lib.DTYPE_F32 if a.dtype == np.float32 else _lib.DTYPE_F64)

    def add_items(self, descriptors, items):
        """`add_item` for a batch: descriptors [B, rings*sectors] (or [B, rings, sectors])."""
        rows, dtype = self._as_rows(descriptors, int(self.shape[0]) * int(self.shape[1]))
        items = list(items)
        assert len(items) == len(rows)
        _lib.check(_lib.load().cslam_sc_add_host(self._h, _lib.ptr(rows), dtype, len(rows)))
        for item in items:
            self.items[self.nb_items] = item
            self.nb_items += 1

    def add_item(self, descriptor, item):
        """Append one descriptor (flat or [rings, sectors]) under the caller's id `item` (:24-46)."""
        self.add_items(np.asarray(descriptor).reshape(1, -1), [item])

    # ---- search -------------------------------------------------------------------------
    def search_batch(self, queries, details=False):
        """All queries [B, rings*sectors] in one call.

        Returns:
            rows int32 [B] (pool row of the match, -1 = none closer than 1), similarities
            float64 [B], yaw shifts int32 [B]; with `details` also the candidate rows
            [B, num_candidates] and their column-shift distances.
        """
        q, dtype = self._as_rows(queries, int(self.shape[0]) * int(self.shape[1]))
        B, C = len(q), int(self.num_candidates)
        rows = np.empty(B, dtype=np.int32)
        sims = np.empty(B, dtype=np.float64)
        yaw = np.empty(B, dtype=np.int32)
        cand = np.empty((B, C), dtype=np.int32) if details else None
        cdist = np.empty((B, C), dtype=np.float64) if details else None
        if self.nb_items < 1:
            raise ValueError("search_batch on an empty pool")
        _lib.check(_lib.load().cslam_sc_search_host(self._h, _lib.ptr(q), dtype, B, _lib.ptr(rows),
                                                    _lib.ptr(sims), _lib.ptr(yaw), _lib.ptr(cand),
                                                    _lib.ptr(cdist)))
        return (rows, sims, yaw, cand, cdist) if details else (rows, sims, yaw)

    def search(self, query, k):
        """Best match of `query` as two one-element lists `[item], [similarity]`; `k` is accepted
        and ignored like in the reference, which returns a single match (:48-89)."""
        if self.nb_items < 1:
            return [None], [None]
        rows, sims, yaw = self.search_batch(np.asarray(query).reshape(1, -1))
        row = int(rows[0])
        if row < 0:
            self.last_yaw_diff_deg = 0
            return [self.items[0]], [0.0]
        self.last_yaw_diff_deg = int(yaw[0]) * (360 / self.shape[1])
        return [self.items[row]], [sims[0]]

    def search_best(self, query):
        """`(item, similarity)` of the best match, `(None, None)` for an empty pool (:91-104)."""
        if self.nb_items < 1:
            return None, None
        idxs, sims = self.search(query, 1)
        return idxs[0], sims[0]

    def last_timing(self):
        """CUDA-event milliseconds of the last search: (ring-key kNN, distance + pick)."""
        a, b = ctypes.c_float(), ctypes.c_float()
        _lib.check(_lib.load().cslam_sc_last_timing(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value
