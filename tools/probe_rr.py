"""Cycles per call of the small Rayleigh-Ritz solvers (one warp), per section, on problems
shaped like those of the eigen-solver: X block orthonormal with diagonal Rayleigh quotients,
W and P blocks small and coupled."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import eigh
from cslam_b200 import _lib

lib = _lib.load()
rng = np.random.default_rng(0)
n = 400
M = rng.normal(size=(n, n)); M = M @ M.T / n
w, U = np.linalg.eigh(M)
for label, eps in (("early", 3e-1), ("late", 1e-5)):
    X = U[:, :2] + eps * rng.normal(size=(n, 2))
    X, _ = np.linalg.qr(X)
    th = np.diag(X.T @ M @ X)
    W = M @ X - X * th
    P = eps * rng.normal(size=(n, 2))
    S = np.concatenate([X, W, P], axis=1)
    GA = np.ascontiguousarray(S.T @ M @ S); GB = np.ascontiguousarray(S.T @ S)
    ref = eigh(GA, GB, eigvals_only=True)[:2]
    # impl 2 = two-stage solve: sweeps > 0 takes the RQI stage 1 (Jacobi only as its fallback), sweeps < 0 the
    # Jacobi stage 1 with |sweeps| sweeps; it solves a restricted problem, so its theta is compared with
    # the Jacobi form of itself (last column), not with the 6 x 6 reference
    t_jac = None
    for sweeps in (3, -3):
        for impl in ((0, 1, 2) if sweeps > 0 else (2,)):
            C = np.zeros((6, 2)); t = np.zeros(2); ok = ctypes.c_int(); cyc = np.zeros(6, dtype=np.int64)
            _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(GA), _lib.ptr(GB), 6, 2, impl, sweeps, 200, 0,
                                                     _lib.ptr(C), _lib.ptr(t), ctypes.byref(ok), _lib.ptr(cyc)))
            print(f"{label} sweeps<={sweeps} impl {impl}: ok {ok.value} total {cyc[0]} | setup {cyc[1]} chol {cyc[2]} tri {cyc[3]} jacobi {cyc[4]} back {cyc[5]} | "
                  f"theta err {np.abs(t - ref).max() / np.abs(ref).max():.1e}")
            if impl == 2 and sweeps > 0:
                t_rqi, C_rqi = t.copy(), C.copy()
            if impl == 2 and sweeps < 0:
                sgn = np.sign((C * C_rqi).sum(axis=0))
                print(f"   two-stage RQI vs Jacobi stage 1: theta rel.diff {np.abs(t - t_rqi).max() / np.abs(t).max():.1e}, "
                      f"C max diff {np.abs(C - C_rqi * sgn).max() / np.abs(C).max():.1e}")
