"""CPU study (numpy/scipy, no GPU) of the eigen-solver's ITERATION COUNT on the C5 graph: the
GPU solver's algorithm (LOBPCG on 1-perp, block m, preconditioner = tridiagonal part of L, residual
||L x - theta x||_1 / ||L||_inf < 1e-10) restated in numpy, and variants of it.  What an iteration
costs is measured on the GPU (profiles/r1_rr_variants.txt); how many are needed is a property of
the algorithm and can be explored here.

    python tools/lobpcg_study.py [--scale 1.0]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.linalg import eigh, solve_banded

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.inputs import mac_scale_graph  # noqa: E402


def laplacian(n, i, j, w):
    A = sp.coo_matrix((np.r_[w, w], (np.r_[i, j], np.r_[j, i])), shape=(n, n)).tocsr()
    d = np.asarray(A.sum(axis=1)).ravel()
    return (sp.diags(d) - A).tocsr()


class TriPrec:
    """x -> tridiag(L)^-1 x (banded solve)."""
    def __init__(self, L):
        n = L.shape[0]
        self.ab = np.zeros((3, n))
        self.ab[1] = L.diagonal()
        off = L.diagonal(1)
        self.ab[0, 1:] = off
        self.ab[2, :-1] = off

    def __call__(self, R):
        return solve_banded((1, 1), self.ab, R)


def chain_hats(n, robots, poses, h):
    """Prolongation Z [n, nc]: piecewise-linear hat functions along every robot's pose chain, one
    coarse node every h poses (plus the chain end)."""
    rows, cols, vals, col = [], [], [], 0
    for r in range(robots):
        base = r * poses
        nodes = np.arange(0, poses, h)
        if nodes[-1] != poses - 1:
            nodes = np.r_[nodes, poses - 1]
        for a in range(len(nodes) - 1):
            x0, x1 = nodes[a], nodes[a + 1]
            idx = np.arange(x0, x1 + (1 if a == len(nodes) - 2 else 0))
            t = (idx - x0) / (x1 - x0)
            rows += list(base + idx) * 2
            cols += [col + a] * len(idx) + [col + a + 1] * len(idx)
            vals += list(1 - t) + list(t)
        col += len(nodes)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, col))


class ShiftInvert:
    """x -> (L + sigma I)^-1 x by sparse LU: what an ideal preconditioner at shift 0 would give."""
    def __init__(self, L, sigma):
        import scipy.sparse.linalg as spl
        self.lu = spl.splu((L + sigma * sp.identity(L.shape[0])).tocsc())

    def __call__(self, R):
        return self.lu.solve(R)


class TwoLevel:
    """Multiplicative two-level preconditioner: exact coarse solve on span(Z), then one
    tridiagonal smoothing step on the remaining residual."""
    def __init__(self, L, Z):
        import scipy.sparse.linalg as spl
        self.L, self.Z, self.tri = L, Z, TriPrec(L)
        Lc = (Z.T @ L @ Z).tocsc()
        self.lu = spl.splu((Lc + 1e-9 * sp.identity(Lc.shape[0])).tocsc())

    def __call__(self, R):
        y = self.Z @ self.lu.solve(self.Z.T @ R)
        return y + self.tri(R - self.L @ y)


RQI_STATS = {"calls": 0, "fallback": 0, "steps": 0}


def rqi3_lowest(GA, GB, max_steps=6):
    """Lowest pair of the (unit-diagonal-scaled) 3 x 3 pencil by Rayleigh-quotient iteration from e_0 with
    adjugate solves, checked for lowestness through the leading minors of A - (theta - delta) B.  When the
    iteration lands on another eigenpair (theta_f, y_f) that pair is deflated: the 2 x 2 pencil on the
    B-orthogonal complement of y_f is solved in closed form and its lowest vector is the start of a second
    iteration.  None = use the Jacobi path (the device routine geig3_lowest_rqi follows this)."""
    A, B = 0.5 * (GA + GA.T), 0.5 * (GB + GB.T)
    RQI_STATS["calls"] += 1
    y = np.array([1.0, 0.0, 0.0])
    th = A[0, 0] / B[0, 0]
    scale = abs(A[0, 0]) + abs(A[1, 1]) + abs(A[2, 2])
    for attempt in range(2):
        ok = False
        for step in range(max_steps):
            M = A - th * B
            r = B @ y
            c00 = M[1, 1] * M[2, 2] - M[1, 2] ** 2
            c01 = M[0, 2] * M[1, 2] - M[0, 1] * M[2, 2]
            c02 = M[0, 1] * M[1, 2] - M[0, 2] * M[1, 1]
            c11 = M[0, 0] * M[2, 2] - M[0, 2] ** 2
            c12 = M[0, 1] * M[0, 2] - M[0, 0] * M[1, 2]
            c22 = M[0, 0] * M[1, 1] - M[0, 1] ** 2
            z = np.array([c00 * r[0] + c01 * r[1] + c02 * r[2], c01 * r[0] + c11 * r[1] + c12 * r[2],
                          c02 * r[0] + c12 * r[1] + c22 * r[2]])
            zm = np.abs(z).max()
            if not (zm > 0 and np.isfinite(zm)):
                break
            z /= zm
            zb = z @ B @ z
            if not zb > 0:
                break
            thn = (z @ A @ z) / zb
            y = z / np.sqrt(zb)
            RQI_STATS["steps"] += 1
            done = abs(thn - th) <= 1e-10 * scale
            th = thn
            if done:
                ok = True
                break
        if not ok:
            break
        # A - (th - delta) B positive definite (leading minors, no divisions) <=> nothing below th - delta
        M = A - (th - 1e-9 * scale) * B
        m2 = M[0, 0] * M[1, 1] - M[0, 1] ** 2
        det = (M[0, 0] * (M[1, 1] * M[2, 2] - M[1, 2] ** 2) + M[0, 1] * (M[0, 2] * M[1, 2] - M[0, 1] * M[2, 2])
               + M[0, 2] * (M[0, 1] * M[1, 2] - M[0, 2] * M[1, 1]))
        if M[0, 0] > 0 and m2 > 0 and det > 0:
            return th, y
        ok = False
        if attempt == 1:
            break
        # deflate (th, y): q_i = e_i - beta_i y spans the B-orthogonal complement for the two indices i, j
        # with the smallest |beta|; A2 = A_ij - th beta_i beta_j, B2 = B_ij - beta_i beta_j
        beta = B @ y
        drop = int(np.argmax(np.abs(beta * y)))
        i, j = [k for k in range(3) if k != drop]
        a11, a12, a22 = A[i, i] - th * beta[i] ** 2, A[i, j] - th * beta[i] * beta[j], A[j, j] - th * beta[j] ** 2
        b11, b12, b22 = B[i, i] - beta[i] ** 2, B[i, j] - beta[i] * beta[j], B[j, j] - beta[j] ** 2
        qa = b11 * b22 - b12 * b12
        qb = a11 * b22 + a22 * b11 - 2.0 * a12 * b12      # = -(linear coefficient)
        qc = a11 * a22 - a12 * a12
        disc = qb * qb - 4.0 * qa * qc
        if not (qa > 0 and qb > 0 and disc >= 0):
            break
        lam = 2.0 * qc / (qb + np.sqrt(disc))
        r1 = np.array([a12 - lam * b12, -(a11 - lam * b11)])   # null vector of row 1
        r2 = np.array([a22 - lam * b22, -(a12 - lam * b12)])   # null vector of row 2
        z2 = r1 if (r1 @ r1) >= (r2 @ r2) else r2
        ynew = np.zeros(3)
        ynew[i], ynew[j] = z2[0], z2[1]
        ynew -= (z2[0] * beta[i] + z2[1] * beta[j]) * y
        nb2 = ynew @ B @ ynew
        if not nb2 > 0:
            break
        y = ynew / np.sqrt(nb2)
        th = y @ A @ y
        RQI_STATS["deflated"] = RQI_STATS.get("deflated", 0) + 1
    RQI_STATS["fallback"] += 1
    return None


def two_stage_rr(S, AS, m, rqi=False):
    """The restricted Rayleigh-Ritz step of rr_two_stage (csrc/mac.cu): per column c the lowest
    pair on span{x_c, w_c, p_c} (3 x 3), then the m x m problem on the results.  Returns
    (theta [m], Y [3m, m]) like the full step."""
    nb = S.shape[1] // m
    Y6 = np.zeros((S.shape[1], m))
    for c in range(m):
        idx = [b * m + c for b in range(nb)]
        Sc, ASc = S[:, idx], AS[:, idx]
        d = 1.0 / np.sqrt((Sc * Sc).sum(axis=0))
        GA, GB = (Sc.T @ ASc) * np.outer(d, d), (Sc.T @ Sc) * np.outer(d, d)
        got = rqi3_lowest(GA, GB) if (rqi and nb == 3) else None
        if got is not None:
            Y6[idx, c] = got[1] * d
        else:
            lam, Y = eigh(0.5 * (GA + GA.T), 0.5 * (GB + GB.T))
            Y6[idx, c] = Y[:, 0] * d
    U, AU = S @ Y6, AS @ Y6
    GA, GB = U.T @ AU, U.T @ U
    lam, Y2 = eigh(0.5 * (GA + GA.T), 0.5 * (GB + GB.T))
    return lam[:m], Y6 @ Y2


def lobpcg(L, m, prec, tol=1e-10, max_iters=3000, inner=0, inner_omega=None, X0=None, seed=7, want_x=False,
           rr="full"):
    """Returns (theta0, iterations, SpMM count).  inner > 0: W = result of `inner` extra steps of
    preconditioned Richardson/Chebyshev on (L - theta I) w = r starting from prec(r)."""
    n = L.shape[0]
    lnorm = abs(L).sum(axis=1).max()
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, m)) if X0 is None else X0.copy()
    X -= X.mean(axis=0)
    AX = L @ X
    th, C = eigh(X.T @ AX, X.T @ X)
    X, AX = X @ C, AX @ C
    P = AP = None
    spmm = 1
    for it in range(max_iters):
        R = AX - X * th
        if np.abs(R[:, 0]).sum() / lnorm < tol:
            return (th[0], it, spmm, X) if want_x else (th[0], it, spmm)
        W = prec(R)
        for s in range(inner):            # polynomial smoothing of the correction
            Rw = R - (L @ W - W * th)      # residual of (L - theta) W = R
            spmm += 1
            W = W + (inner_omega if inner_omega else 1.0) * prec(Rw)
        W -= W.mean(axis=0)
        AW = L @ W
        spmm += 1
        S = [X, W] + ([P] if P is not None else [])
        AS = [AX, AW] + ([AP] if P is not None else [])
        S, AS = np.hstack(S), np.hstack(AS)
        if rr in ("two-stage", "two-stage-rqi") and P is not None:
            th, Y = two_stage_rr(S, AS, m, rqi=rr == "two-stage-rqi")
            Pn = S[:, m:] @ Y[m:]
            APn = AS[:, m:] @ Y[m:]
            X = X @ Y[:m] + Pn
            AX = AX @ Y[:m] + APn
            P, AP = Pn, APn
            continue
        # column scaling like the GPU solve
        d = 1.0 / np.sqrt((S * S).sum(axis=0))
        GA, GB = (S.T @ AS) * np.outer(d, d), (S.T @ S) * np.outer(d, d)
        try:
            lam, Y = eigh(0.5 * (GA + GA.T), 0.5 * (GB + GB.T))
        except np.linalg.LinAlgError:
            S, AS = S[:, :2 * m], AS[:, :2 * m]
            d = d[:2 * m]
            GA, GB = (S.T @ AS) * np.outer(d, d), (S.T @ S) * np.outer(d, d)
            lam, Y = eigh(0.5 * (GA + GA.T), 0.5 * (GB + GB.T))
        Y = Y[:, :m] * d[:S.shape[1], None]
        th = lam[:m]
        Pn = S[:, m:] @ Y[m:]
        APn = AS[:, m:] @ Y[m:]
        X = X @ Y[:m] + Pn
        AX = AX @ Y[:m] + APn
        P, AP = Pn, APn
    return (th[0], max_iters, spmm, X) if want_x else (th[0], max_iters, spmm)


def frank_wolfe(n, fixed, cand, k, iters, make_prec=TriPrec, **kw):
    """The Frank-Wolfe loop of MAC.fw_subset (mac.py:191-233) around warm-started solves; returns the
    LOBPCG iterations of every solve, the SpMM total and the final selection."""
    ci, cj, cw = cand
    w = np.zeros(len(cw))
    w[np.argpartition(cw, -k)[-k:]] = 1.0
    X = None
    counts, spmm_total = [], 0
    for it in range(iters):
        act = np.flatnonzero(w > 1e-10)
        L = laplacian(n, np.r_[fixed[0], ci[act]], np.r_[fixed[1], cj[act]], np.r_[fixed[2], w[act] * cw[act]])
        th, its, spmm, X = lobpcg(L, prec=make_prec(L), X0=X, want_x=True, **kw)
        counts.append(its)
        spmm_total += spmm
        v = X[:, 0] / np.linalg.norm(X[:, 0])
        g = cw * (v[ci] - v[cj]) ** 2
        s = np.zeros(len(cw))
        s[np.argpartition(g, -k)[-k:]] = 1.0
        w = w + 2.0 / (it + 2.0) * (s - w)
    return counts, spmm_total, np.sort(np.argpartition(w, -k)[-k:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the C5 graph (poses, candidates)")
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--fw", type=int, default=0, help="run this many Frank-Wolfe iterations per variant instead")
    ap.add_argument("--rqi", action="store_true", help="with --fw: only the two-stage variants, Jacobi against RQI stage 1")
    a = ap.parse_args()
    R, Pn, mc = 8, int(12500 * a.scale), int(1_000_000 * a.scale)
    fixed, cand, n = mac_scale_graph(R, Pn, mc)
    k = max(10, int(a.k * a.scale))
    sel = np.argpartition(cand[2], -k)[-k:]
    i = np.r_[fixed[0], cand[0][sel]]
    j = np.r_[fixed[1], cand[1][sel]]
    w = np.r_[fixed[2], cand[2][sel]]
    L = laplacian(n, i, j, w)
    prec = TriPrec(L)
    if a.fw:
        print(f"n = {n}, {mc} candidates, budget {k}: {a.fw} Frank-Wolfe iterations, warm-started solves")
        base = None
        hats = {h: chain_hats(n, R, Pn, h) for h in (128, 64, 32)}
        variants = (("block 2 (the GPU solver)", dict(m=2)), ("block 2, two-stage Rayleigh-Ritz", dict(m=2, rr="two-stage")),
                          ("block 1", dict(m=1)),
                          ("block 2 + 1 smoothing step", dict(m=2, inner=1)),
                          ("block 1 + 1 smoothing step", dict(m=1, inner=1)),
                          ("block 2, two-level h=128", dict(m=2, make_prec=lambda L: TwoLevel(L, hats[128]))),
                          ("block 2, two-level h=64", dict(m=2, make_prec=lambda L: TwoLevel(L, hats[64]))),
                          ("block 2, two-level h=32", dict(m=2, make_prec=lambda L: TwoLevel(L, hats[32]))))
        if a.rqi:
            variants = (("block 2, two-stage Rayleigh-Ritz", dict(m=2, rr="two-stage")),
                        ("block 2, two-stage, RQI stage 1", dict(m=2, rr="two-stage-rqi")))
        for label, kw in variants:
            t0 = time.time()
            counts, spmm, selection = frank_wolfe(n, fixed, cand, k, a.fw, **kw)
            base = selection if base is None else base
            print(f"  {label:30s} LOBPCG iterations {sum(counts):5d} (first {counts[0]}, then mean {np.mean(counts[1:]):.0f})  "
                  f"SpMM {spmm:5d}  selection differs in {len(set(base) ^ set(selection))} edges  {time.time() - t0:.0f} s", flush=True)
            if kw.get("rr") == "two-stage-rqi":
                print(f"    RQI stage 1: {RQI_STATS['calls']} solves, {RQI_STATS['fallback']} fell back to the Jacobi path, "
                      f"{RQI_STATS['steps'] / max(1, RQI_STATS['calls']):.2f} steps per solve, {RQI_STATS.get('deflated', 0)} second attempts after deflation")
                print("    per solve:", counts)
        return
    print(f"n = {n}, nnz = {L.nnz}, first Frank-Wolfe Laplacian (greedy start, {k} active candidates)")
    ref = None
    for label, kw in (("block 2, tridiagonal (the GPU solver)", dict(m=2)),
                      ("block 1", dict(m=1)),
                      ("block 3", dict(m=3)),
                      ("block 4", dict(m=4)),
                      ("block 2 + 1 smoothing step", dict(m=2, inner=1)),
                      ("block 2 + 2 smoothing steps", dict(m=2, inner=2)),
                      ("block 2 + 4 smoothing steps", dict(m=2, inner=4)),
                      ("block 2, two-level h=128 (coarse+smooth)", dict(m=2, prec=TwoLevel(L, chain_hats(n, R, Pn, 128)))),
                      ("block 2, two-level h=64", dict(m=2, prec=TwoLevel(L, chain_hats(n, R, Pn, 64)))),
                      ("block 2, two-level h=32", dict(m=2, prec=TwoLevel(L, chain_hats(n, R, Pn, 32)))),
                      ("block 2, exact (L + 1e-5 I)^-1", dict(m=2, prec=ShiftInvert(L, 1e-5)))):
        t0 = time.time()
        kw = dict(kw)
        kw.setdefault("prec", prec)
        th, it, spmm = lobpcg(L, **kw)
        ref = th if ref is None else ref
        print(f"  {label:40s} iterations {it:5d}  SpMM {spmm:5d}  lambda2 {th:.12e}  (d {abs(th - ref):.1e})  {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
