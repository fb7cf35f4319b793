"""Oracle for the MAC sparsification solver (TEST INFRASTRUCTURE ONLY).

Restates, in numpy/scipy:
  * cslam/mac/utils.py:47-126  — Laplacian assembly from edge lists
  * cslam/mac/mac.py:19-233    — class MAC (combined_laplacian, find_fiedler_pair,
                                 grad_from_fiedler, round_solution[_tiebreaker], fw_subset)
  * networkx (un-vendored dependency; reference pins networkx==2.7.1 in
    cslam/mac/requirements.txt:2, 2.8.4 in spec-file.txt:84; the build container
    has 3.6.1) linalg/algebraicconnectivity.py `_tracemin_fiedler` with the
    `tracemin_lu` solver, reached from cslam/mac/mac.py:52-58 through the private
    `_get_fiedler_func`.  Published algorithm: Manguoglu, Cox, Saied, Sameh,
    "TRACEMIN-Fiedler: A Parallel Algorithm for Computing the Fiedler Vector"
    (VECPAR 2010): block trace minimisation with q = min(4, n-1) vectors,
    X0 = RandomState(7).normal(size=(q, n)).T, exact solves with the Laplacian
    made non-singular by setting the diagonal entry of the densest row to +inf.

Pinned by tests/test_oracle_mac.py against golden (lambda_2, v_2, selections)
produced by running the reference's own MAC / networkx code (oracle/make_golden.py).
"""
from collections import namedtuple

import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg

Edge = namedtuple('Edge', ['i', 'j', 'weight'])  # cslam/mac/utils.py:13


def laplacian_from_arrays(i, j, w, n):
    """cslam/mac/utils.py:47-126: 4 triplets per edge, COO -> CSR (duplicates summed)."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    # same triplet order as the reference loops: (u,u) (v,v) (u,v) (v,u) per edge
    rows = np.stack([i, j, i, j], axis=1).ravel()
    cols = np.stack([i, j, j, i], axis=1).ravel()
    data = np.stack([w, w, -w, -w], axis=1).ravel()
    return sp.csr_matrix(sp.coo_matrix((data, (rows, cols)), shape=[n, n]))


def tracemin_fiedler_lu(L, tol=1e-8, seed=7):
    """networkx _get_fiedler_func('tracemin_lu') + _tracemin_fiedler (normalized=False)."""
    n = L.shape[0]
    q = min(4, n - 1)
    X = np.asarray(np.random.RandomState(seed).normal(size=(q, n))).T

    def project(X):
        for j in range(X.shape[1]):
            X[:, j] -= X[:, j].sum() / n

    A = sp.csc_array(L, dtype=float, copy=True)
    i = (A.indptr[1:] - A.indptr[:-1]).argmax()
    A[i, i] = np.inf
    lu = scipy.sparse.linalg.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                                  options={"Equil": True, "SymmetricMode": True})

    Lnorm = abs(L).sum(axis=1).flatten().max()
    project(X)
    W = np.ndarray(X.shape, order="F")
    while True:
        X = np.linalg.qr(X)[0]
        W[:, :] = L @ X
        H = X.T @ W
        sigma, Y = scipy.linalg.eigh(H, overwrite_a=True)
        X = X @ Y
        res = scipy.linalg.blas.dasum(W @ Y[:, 0] - sigma[0] * X[:, 0]) / Lnorm
        if res < tol:
            break
        for c in range(X.shape[1]):
            W[:, c] = lu.solve(np.asarray(X[:, c]))
        X = (scipy.linalg.inv(W.T @ X) @ W.T).T
        project(X)
    return sigma[0], np.asarray(X[:, 0])


class MACOracle:
    """cslam/mac/mac.py:19-233."""

    def __init__(self, fixed_measurements, candidate_measurements, num_poses):
        fi = [e.i for e in fixed_measurements]
        fj = [e.j for e in fixed_measurements]
        fw = [e.weight for e in fixed_measurements]
        self.L_odom = laplacian_from_arrays(fi, fj, fw, num_poses)
        self.num_poses = num_poses
        self.weights = np.array([e.weight for e in candidate_measurements])
        self.edge_list = np.array([(e.i, e.j) for e in candidate_measurements])

    @classmethod
    def from_arrays(cls, fixed, cand, num_poses):
        """Same object from (i, j, weight) array triples (large synthetic graphs)."""
        self = cls.__new__(cls)
        self.L_odom = laplacian_from_arrays(*fixed, num_poses)
        self.num_poses = num_poses
        self.weights = np.asarray(cand[2], dtype=np.float64)
        self.edge_list = np.stack([cand[0], cand[1]], axis=1)
        return self

    def find_fiedler_pair(self, L, method='tracemin_lu', tol=1e-8):
        return tracemin_fiedler_lu(L, tol=tol, seed=7)  # mac.py:52-58

    def combined_laplacian(self, w, tol=1e-10):
        idx = np.where(w > tol)  # mac.py:72
        prod = w[idx] * self.weights[idx]
        e = self.edge_list[idx]
        C1 = laplacian_from_arrays(e[:, 0] if len(e) else [], e[:, 1] if len(e) else [], prod,
                                   self.num_poses)
        return self.L_odom + C1

    def evaluate_fiedler_pair(self, w, method='tracemin_lu', tol=1e-8):
        return self.find_fiedler_pair(self.combined_laplacian(w), method, tol)

    def evaluate_objective(self, w):
        return self.find_fiedler_pair(self.combined_laplacian(w))[0]

    def grad_from_fiedler(self, fiedler_vec):
        # mac.py:112-130, vectorised: grad_k = (weight_k * (v_i - v_j)) * (v_i - v_j)
        d = fiedler_vec[self.edge_list[:, 0]] - fiedler_vec[self.edge_list[:, 1]]
        return (self.weights * d) * d

    def round_solution(self, w, k):
        # mac.py:132-147
        idx = np.argpartition(w, -k)[-k:]
        rounded = np.zeros(len(w))
        if k > 0:
            rounded[idx] = 1.0
        return rounded

    def round_solution_tiebreaker(self, w, k, decimal_tol=10):
        # mac.py:168-189
        truncated_w = w.round(decimals=decimal_tol)
        zipped = np.zeros(len(w), dtype=[('w', 'float'), ('weight', 'float')])
        zipped['w'] = truncated_w
        zipped['weight'] = self.weights
        idx = np.argpartition(zipped, -k, order=['w', 'weight'])[-k:]
        rounded = np.zeros(len(w))
        if k > 0:
            rounded[idx] = 1.0
        return rounded

    def fw_subset(self, w_init, k, max_iters=5, duality_gap_tol=1e-8, trace=None):
        # mac.py:191-233
        u_i = float("inf")
        w_i = w_init
        for it in range(max_iters):
            f_i, vec_i = self.evaluate_fiedler_pair(w_i)
            grad_i = self.grad_from_fiedler(vec_i)
            s_i = self.round_solution(grad_i, k)
            u_i = min(u_i, f_i + grad_i @ (s_i - w_i))
            if trace is not None:
                trace.append(dict(f=f_i, vec=vec_i, grad=grad_i, s=s_i, u=u_i))
            if u_i - f_i < duality_gap_tol:
                return self.round_solution_tiebreaker(w_i, k), w_i, u_i
            alpha = 2.0 / (it + 2.0)
            w_i = w_i + alpha * (s_i - w_i)
        return self.round_solution_tiebreaker(w_i, k), w_i, u_i


def topk_boundary_gap(values, k):
    """Gap between the k-th and (k+1)-th largest value (inf if k >= len)."""
    if k <= 0 or k >= len(values):
        return float("inf")
    part = np.partition(values, -k - 1)
    return float(part[-k:].min() - part[-k - 1])
