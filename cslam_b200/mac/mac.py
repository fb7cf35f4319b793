"""MAC — maximising algebraic connectivity by Frank-Wolfe, same class API as the
reference (cslam/mac/mac.py:19-233); the Laplacian, the Fiedler pair, the edge gradient,
the top-k direction and the FW update live on the GPU (libcslam_b200, csrc/mac.cu).
"""
import ctypes
from collections import namedtuple

import numpy as np

from .. import _lib
from .utils import (edges_to_arrays, weight_graph_lap_from_edge_list,
                    weight_graph_lap_from_edges)

MACResult = namedtuple('MACResult', ['w', 'F_unrounded', 'objective_values', 'duality_gaps'])


def _resolve_device(device):
    if device is not None:
        return int(device)
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class MAC:

    def __init__(self, fixed_measurements, candidate_measurements, num_poses, device=None):
        """MAC(fixed, candidates, num_poses) (mac.py:21-33).  Edge lists are lists of
        `Edge(i, j, weight)`; (i, j, w) array triples are accepted as well."""
        lib = _lib.load()
        _lib.require_device()
        self.num_poses = int(num_poses)
        fi, fj, fw = self._as_arrays(fixed_measurements)
        ci, cj, cw = self._as_arrays(candidate_measurements)
        self._fixed = (fi, fj, fw)
        self.weights = cw
        self._cand_ij = (ci, cj)
        self._edge_list = None      # [m, 2] view of the reference (mac.py:27), built on first use
        self._device = _resolve_device(device)
        h = ctypes.c_void_p()
        _lib.check(lib.cslam_mac_create(self.num_poses, len(fi), _lib.ptr(fi), _lib.ptr(fj),
                                        _lib.ptr(fw), len(ci), _lib.ptr(ci), _lib.ptr(cj),
                                        _lib.ptr(cw), self._device, ctypes.byref(h)))
        self._h = h
        self._L_odom = None
        self.last_fw_iters = 0
        self.last_trace = None

    @staticmethod
    def _as_arrays(meas):
        if isinstance(meas, tuple) and len(meas) == 3 and isinstance(meas[0], np.ndarray):
            i, j, w = meas
        else:
            i, j, w = edges_to_arrays(list(meas))
        return (np.ascontiguousarray(i, dtype=np.int32), np.ascontiguousarray(j, dtype=np.int32),
                np.ascontiguousarray(w, dtype=np.float64))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().cslam_mac_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_options(self, tol=1e-10, block_size=2, max_lobpcg_iters=20000):
        _lib.check(_lib.load().cslam_mac_set_options(self._h, float(tol), int(block_size),
                                                     int(max_lobpcg_iters)))

    def stats(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        _lib.check(_lib.load().cslam_mac_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"lobpcg_iters": a.value, "spmv_columns": b.value, "jacobi_fallback": bool(c.value)}

    def solver_timing(self):
        """Cumulative CUDA-event time / launches / iterations / algorithmic bytes of the
        persistent eigen-solver kernel (for the bench roofline)."""
        ms, a, b, c = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.load().cslam_mac_solver_timing(self._h, ctypes.byref(ms), ctypes.byref(a),
                                                       ctypes.byref(b), ctypes.byref(c)))
        return {"kernel_ms": ms.value, "launches": a.value, "iterations": b.value,
                "algorithmic_bytes": c.value}

    # ---- reference API -----------------------------------------------------
    @property
    def edge_list(self):
        """[m, 2] int array of the candidate endpoints (mac.py:27); assembled on demand — the
        solver works on the separate index arrays it was created from."""
        if self._edge_list is None:
            ci, cj = self._cand_ij
            self._edge_list = np.stack([ci, cj], axis=1) if len(ci) else np.zeros((0, 2), dtype=np.int32)
        return self._edge_list

    @property
    def L_odom(self):
        """scipy CSR Laplacian of the fixed edges (mac.py:22-23), built on demand."""
        if self._L_odom is None:
            from .utils import _laplacian
            self._L_odom = _laplacian(*self._fixed, self.num_poses)
        return self._L_odom

    def find_fiedler_pair(self, L, method='tracemin_lu', tol=1e-8):
        """(lambda_2, v_2) of a caller-assembled Laplacian (mac.py:35-59).  `method` is
        accepted for compatibility; the solver is the GPU LOBPCG of csrc/mac.cu, run to
        min(tol, 1e-10)."""
        assert method != 'lobpcg'
        from scipy.sparse import csr_matrix
        L = csr_matrix(L)
        L.sum_duplicates()
        n = L.shape[0]
        indptr = np.ascontiguousarray(L.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(L.indices, dtype=np.int32)
        data = np.ascontiguousarray(L.data, dtype=np.float64)
        lam = ctypes.c_double()
        vec = np.empty(n, dtype=np.float64)
        iters = ctypes.c_int()
        _lib.check(_lib.load().cslam_fiedler_csr(n, _lib.ptr(indptr), _lib.ptr(indices),
                                                 _lib.ptr(data), min(float(tol), 1e-10), 1,
                                                 self._device, ctypes.byref(lam), _lib.ptr(vec),
                                                 ctypes.byref(iters)))
        return lam.value, vec

    def combined_laplacian(self, w, tol=1e-10):
        """scipy CSR of L(w) (mac.py:61-77); host-side convenience, not on the hot path."""
        w = np.asarray(w, dtype=np.float64)
        idx = np.where(w > tol)
        prod = w[idx] * self.weights[idx]
        return self.L_odom + weight_graph_lap_from_edges(self.edge_list[idx], prod, self.num_poses)

    def evaluate_fiedler_pair(self, w, method='tracemin_lu', tol=1e-8):
        """(lambda_2(L(w)), v_2(L(w))) (mac.py:79-97), assembled and solved on the GPU."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert len(w) == len(self.weights)
        lam = ctypes.c_double()
        vec = np.empty(self.num_poses, dtype=np.float64)
        iters = ctypes.c_int()
        _lib.check(_lib.load().cslam_mac_fiedler(self._h, _lib.ptr(w), ctypes.byref(lam),
                                                 _lib.ptr(vec), ctypes.byref(iters)))
        self.last_lobpcg_iters = iters.value
        return lam.value, vec

    def evaluate_objective(self, w):
        """F(w) = lambda_2(L(w)) (mac.py:99-110)."""
        return self.evaluate_fiedler_pair(w)[0]

    def grad_from_fiedler(self, fiedler_vec):
        """grad_e = weight_e (v_i - v_j)^2 (mac.py:112-130)."""
        v = np.ascontiguousarray(fiedler_vec, dtype=np.float64)
        assert len(v) == self.num_poses
        g = np.empty(len(self.weights), dtype=np.float64)
        _lib.check(_lib.load().cslam_mac_grad(self._h, _lib.ptr(v), _lib.ptr(g)))
        return g

    def round_solution(self, w, k):
        """0/1 indicator of the k largest entries (mac.py:132-147).  Stand-alone helper;
        inside fw_subset the same selection runs on the GPU (k_topk_*)."""
        w = np.asarray(w)
        rounded = np.zeros(len(w))
        if k > 0:
            rounded[np.argpartition(w, -k)[-k:]] = 1.0
        return rounded

    def simple_random_round(self, w, k):
        """Randomised rounding with E[#selected] = k (mac.py:149-166)."""
        w = np.asarray(w)
        return (w > np.random.rand(len(w))).astype(float)

    def round_solution_tiebreaker(self, w, k, decimal_tol=10):
        """Top-k by (round(w, decimal_tol), weight) (mac.py:168-189)."""
        w = np.asarray(w)
        zipped = np.zeros(len(w), dtype=[('w', 'float'), ('weight', 'float')])
        zipped['w'] = w.round(decimals=decimal_tol)
        zipped['weight'] = self.weights
        rounded = np.zeros(len(w))
        if k > 0:
            rounded[np.argpartition(zipped, -k, order=['w', 'weight'])[-k:]] = 1.0
        return rounded

    def fw_subset_sparse(self, init_idx, init_val, k, max_iters=5, duality_gap_tol=1e-8, trace=False,
                         want_support=True):
        """`fw_subset` without dense vectors: the start vector by its non-zero entries
        (`init_idx`, `init_val`); returns (sel_idx, (sup_idx, sup_val), upper_bound) - the ascending
        ids of the rounded selection and the non-zero entries of the unrounded iterate
        (`want_support=False`: None instead - select_candidates only uses the selection)."""
        init_idx = np.ascontiguousarray(init_idx, dtype=np.int32)
        init_val = np.ascontiguousarray(init_val, dtype=np.float64)
        assert init_idx.shape == init_val.shape and init_idx.ndim == 1
        m, k = len(self.weights), int(k)
        cap = max(1, min(m, len(init_idx) + k * max(int(max_iters), 1)))
        sel = np.empty(max(k, 1), dtype=np.int32)
        sup_idx = np.empty(cap, dtype=np.int32) if want_support else None
        sup_val = np.empty(cap, dtype=np.float64) if want_support else None
        n_sup, u, iters = ctypes.c_int64(), ctypes.c_double(), ctypes.c_int()
        tsel = np.full((max(max_iters, 1), max(k, 1)), -1, dtype=np.int32) if trace else None
        tf = np.full(max(max_iters, 1), np.nan) if trace else None
        _lib.check(_lib.load().cslam_mac_fw_subset_sparse(
            self._h, len(init_idx), _lib.ptr(init_idx), _lib.ptr(init_val), k, int(max_iters),
            float(duality_gap_tol), _lib.ptr(sel), cap, _lib.ptr(sup_idx), _lib.ptr(sup_val),
            ctypes.byref(n_sup), ctypes.byref(u), ctypes.byref(iters), _lib.ptr(tsel), _lib.ptr(tf)))
        self.last_fw_iters = iters.value
        self.last_trace = (tsel, tf) if trace else None
        sup = (sup_idx[:n_sup.value], sup_val[:n_sup.value]) if want_support else None
        return sel[:k], sup, u.value

    def fw_subset(self, w_init, k, max_iters=5, duality_gap_tol=1e-8, trace=False):
        """Frank-Wolfe subset selection (mac.py:191-233), entirely on the GPU.

        returns (solution, unrounded, upper_bound) like the reference.
        """
        w_init = np.asarray(w_init, dtype=np.float64)
        m = len(self.weights)
        assert len(w_init) == m
        idx = np.flatnonzero(w_init > 0.0)
        sel, (sup_idx, sup_val), u = self.fw_subset_sparse(idx, w_init[idx], k, max_iters,
                                                           duality_gap_tol, trace)
        rounded = np.zeros(m)
        rounded[sel] = 1.0
        w = np.zeros(m)
        w[sup_idx] = sup_val
        return rounded, w, u
