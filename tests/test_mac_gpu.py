"""GPU parity of the MAC solver (csrc/mac.cu through the C ABI) against (a) golden outputs
of the reference's own MAC / networkx code (tests/golden/mac.npz, oracle/make_golden.py)
and (b) the numpy oracle (oracle/mac.py)."""
import os

import numpy as np
import pytest

from oracle.inputs import MAC_CASES
from oracle.mac import Edge as OEdge
from oracle.mac import MACOracle, topk_boundary_gap

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mac.npz"))


def _case(tag):
    rf = GOLD[f"{tag}_rekey_fixed"]
    rc = GOLD[f"{tag}_rekey_cand"]
    R, P, m, k, seed = MAC_CASES[tag]
    n = R * P
    return rf, rc, n, k


def _mac(tag):
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    rf, rc, n, k = _case(tag)
    fixed = [Edge(int(a), int(b), float(w)) for a, b, w in rf]
    cand = [Edge(int(a), int(b), float(w)) for a, b, w in rc]
    return MAC(fixed, cand, n), k


def _align(v, ref):
    return v if np.dot(v, ref) >= 0 else -v


@pytest.mark.parametrize("tag", list(MAC_CASES))
def test_fiedler_pair_matches_reference(tag):
    mac, k = _mac(tag)
    w0 = GOLD[f"{tag}_w0"]
    lam, vec = mac.evaluate_fiedler_pair(w0)
    lam_ref, vec_ref = float(GOLD[f"{tag}_lambda2"]), GOLD[f"{tag}_fiedler"]
    # north_star tolerance: Fiedler values within 1e-3; the reference itself stops at a
    # 1e-8 relative residual, we assert far tighter than 1e-3
    assert abs(lam - lam_ref) <= 1e-6 * max(1.0, abs(lam_ref)) + 1e-9
    assert abs(np.linalg.norm(vec) - 1.0) < 1e-9
    assert np.abs(_align(vec, vec_ref) - vec_ref).max() < 1e-3
    # our own residual is tighter than the reference's tolerance
    L = mac.combined_laplacian(w0)
    res = np.abs(L @ vec - lam * vec).sum() / abs(L).sum(axis=1).max()
    assert res < 1e-9


@pytest.mark.parametrize("tag", list(MAC_CASES))
def test_gradient_matches_reference(tag):
    mac, k = _mac(tag)
    vec_ref = GOLD[f"{tag}_fiedler"]
    g = mac.grad_from_fiedler(vec_ref)
    # same inputs, same operation order (mac.py:123-129): bit-identical
    assert np.array_equal(g, GOLD[f"{tag}_grad"])


@pytest.mark.parametrize("tag", list(MAC_CASES))
def test_fw_subset_matches_reference(tag):
    mac, k = _mac(tag)
    rf, rc, n, _ = _case(tag)
    w0 = GOLD[f"{tag}_w0"]
    rounded, w, u = mac.fw_subset(w0.copy(), k, max_iters=20, trace=True)
    ref_rounded, ref_w, ref_u = GOLD[f"{tag}_fw_rounded"], GOLD[f"{tag}_fw_w"], float(GOLD[f"{tag}_fw_u"])
    assert rounded.sum() == k and set(np.unique(rounded)) <= {0.0, 1.0}
    # Replay the reference algorithm with the oracle to learn how well separated every
    # top-k decision was; the selections must be identical whenever they were decidable
    # (gap well above the eigen-solver tolerance), else compare objective values.
    orc = MACOracle([OEdge(int(a), int(b), float(c)) for a, b, c in rf],
                    [OEdge(int(a), int(b), float(c)) for a, b, c in rc], n)
    trace = []
    orc.fw_subset(w0.copy(), k, max_iters=20, trace=trace)
    decidable = all(topk_boundary_gap(t["grad"], k) > 1e-6 * max(t["grad"].max(), 1e-300)
                    for t in trace)
    if decidable:
        tsel, tf = mac.last_trace
        for it, t in enumerate(trace):
            assert set(tsel[it]) == set(np.nonzero(t["s"])[0]), f"direction s_{it} differs"
            assert abs(tf[it] - t["f"]) <= 1e-6 * max(1.0, abs(t["f"]))
        assert np.array_equal(rounded, ref_rounded)
        np.testing.assert_allclose(w, ref_w, atol=1e-12)
        assert abs(u - ref_u) <= 1e-6 * max(1.0, abs(ref_u))
    else:
        f_ours = orc.evaluate_objective(rounded)
        f_ref = orc.evaluate_objective(ref_rounded)
        assert f_ours >= f_ref * (1 - 1e-3)


def test_find_fiedler_pair_on_scipy_matrix_and_disconnected_graph():
    from cslam_b200._lib import SingularLaplacianError
    mac, k = _mac("g0")
    L = mac.combined_laplacian(GOLD["g0_w0"])
    lam, vec = mac.find_fiedler_pair(L)
    assert abs(lam - float(GOLD["g0_lambda2"])) < 1e-8
    # no candidate active and no bridge between the chains -> singular, like SuperLU raising
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    fixed = [Edge(i, i + 1, 1.0) for i in range(4)] + [Edge(i, i + 1, 1.0) for i in range(5, 9)]
    cand = [Edge(0, 7, 0.5), Edge(2, 9, 0.25)]
    m2 = MAC(fixed, cand, 10)
    with pytest.raises(SingularLaplacianError):
        m2.evaluate_fiedler_pair(np.zeros(2))
    lam2, _ = m2.evaluate_fiedler_pair(np.array([1.0, 0.0]))
    assert lam2 > 0


def test_block_size_two_and_medium_graph_vs_oracle():
    # 6 robots x 400 poses, 3000 candidates: compare with the numpy TraceMIN oracle
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    rng = np.random.default_rng(11)
    R, P, m, k = 6, 400, 3000, 60
    n = R * P
    fi = [r * P + i for r in range(R) for i in range(P - 1)] + [(r + 1) * P - 1 for r in range(R - 1)]
    fj = [r * P + i + 1 for r in range(R) for i in range(P - 1)] + [(r + 2) * P - 1 for r in range(R - 1)]
    r0 = rng.integers(0, R, m)
    r1 = (r0 + rng.integers(1, R, m)) % R
    ci = r0 * P + rng.integers(0, P, m)
    cj = r1 * P + rng.integers(0, P, m)
    cw = rng.random(m)
    fixed = [Edge(a, b, 1.0) for a, b in zip(fi, fj)]
    cand = [Edge(int(a), int(b), float(c)) for a, b, c in zip(ci, cj, cw)]
    mac = MAC(fixed, cand, n)
    orc = MACOracle([OEdge(*e) for e in fixed], [OEdge(*e) for e in cand], n)
    w0 = np.zeros(m)
    w0[np.argpartition(cw, -k)[-k:]] = 1.0
    lam_ref, vec_ref = orc.evaluate_fiedler_pair(w0)
    for bs in (1, 2):
        mac.set_options(block_size=bs)
        lam, vec = mac.evaluate_fiedler_pair(w0)
        assert abs(lam - lam_ref) < 1e-7 * max(1.0, lam_ref)
        assert np.abs(_align(vec, vec_ref) - vec_ref).max() < 1e-3
    assert not mac.stats()["jacobi_fallback"]


@pytest.mark.parametrize("impl", [1, 0])
def test_small_rayleigh_ritz_solvers_against_scipy(impl):
    """The 6 x 6 generalised eigenproblem solved once per LOBPCG iteration (register-resident
    and shared-memory warp solvers) against scipy.linalg.eigh: Ritz values to 1e-10 relative,
    vectors GB-orthonormal with a small residual, for full and padded bases and for the badly
    scaled bases of a nearly converged iteration."""
    import ctypes
    from scipy.linalg import eigh
    from cslam_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for s in (2, 4, 6):
        for trial in range(12):
            S = rng.normal(size=(40, s))
            if trial % 3 == 1:
                S[:, 2:] *= 1e-5          # W, P tiny next to X
            if trial % 3 == 2:
                S[:, s // 2:] = S[:, :s - s // 2] + 1e-3 * S[:, s // 2:]   # nearly dependent
            M = rng.normal(size=(40, 40))
            M = M @ M.T
            GA = np.zeros((6, 6))
            GB = np.zeros((6, 6))
            GA[:s, :s] = S.T @ M @ S
            GB[:s, :s] = S.T @ S
            m = 2
            C = np.zeros((6, 2))
            th = np.zeros(2)
            ok = ctypes.c_int()
            _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(GA), _lib.ptr(GB), s, m, impl, 8, 1, 0,
                                                     _lib.ptr(C), _lib.ptr(th), ctypes.byref(ok), None))
            assert ok.value == 1
            w = eigh(GA[:s, :s], GB[:s, :s], eigvals_only=True)
            cond = np.linalg.cond(GB[:s, :s] / np.sqrt(np.outer(np.diag(GB)[:s], np.diag(GB)[:s])))
            np.testing.assert_allclose(th, w[:2], rtol=1e-13 * max(cond, 1e3), atol=0)
            assert not C[s:].any()
            G = C[:s].T @ GB[:s, :s] @ C[:s]
            np.testing.assert_allclose(G, np.eye(2), atol=1e-13 * max(cond, 1e3))
            res = GA[:s, :s] @ C[:s] - GB[:s, :s] @ C[:s] * th
            scale = np.abs(GA[:s, :s] @ C[:s]).max()
            assert np.abs(res).max() <= 1e-6 * scale     # single-precision angles, 8 sweeps
    bad = np.eye(6)
    bad[1, 1] = -1.0
    _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(np.eye(6)), _lib.ptr(bad), 4, 2, impl, 3, 1, 0,
                                             _lib.ptr(C), _lib.ptr(th), ctypes.byref(ok), None))
    assert ok.value == 0


def _two_stage_reference(GA, GB, s, m):
    """numpy restatement of rr_two_stage (csrc/mac.cu): per column the lowest pair on
    span{x_c, w_c, p_c}, then the m x m problem on the results."""
    from scipy.linalg import eigh
    nb = s // m
    Y6 = np.zeros((6, m))
    for c in range(m):
        idx = [b * m + c for b in range(nb)]
        lam, Y = eigh(GA[np.ix_(idx, idx)], GB[np.ix_(idx, idx)])
        Y6[idx, c] = Y[:, 0]
    lam2, Y2 = eigh(Y6.T @ GA @ Y6, Y6.T @ GB @ Y6)
    return lam2, Y6 @ Y2


def test_two_stage_rayleigh_ritz_against_numpy():
    """rr_impl = 2: the restricted (3 x 3 per column, then 2 x 2) Rayleigh-Ritz step.  Ritz values
    equal the numpy restatement's, the returned vectors are GB-orthonormal, their Rayleigh
    quotients are the returned thetas, and theta_0 is never below the full problem's lowest
    value (it is a Rayleigh-Ritz value on a subspace of the trial space)."""
    import ctypes
    from scipy.linalg import eigh
    from cslam_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    for s in (2, 4, 6):
        for trial in range(12):
            S = rng.normal(size=(40, s))
            if trial % 3 == 1:
                S[:, 2:] *= 1e-5
            if trial % 3 == 2 and s > 2:
                S[:, 2:] = S[:, :s - 2] @ rng.normal(size=(s - 2, s - 2)) * 1e-2 + 1e-3 * S[:, 2:]
            M = rng.normal(size=(40, 40))
            M = M @ M.T
            GA, GB = np.zeros((6, 6)), np.zeros((6, 6))
            GA[:s, :s] = S.T @ M @ S
            GB[:s, :s] = S.T @ S
            C, th, ok = np.zeros((6, 2)), np.zeros(2), ctypes.c_int()
            _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(GA), _lib.ptr(GB), s, 2, 2, 8, 1, 0,
                                                     _lib.ptr(C), _lib.ptr(th), ctypes.byref(ok), None))
            assert ok.value == 1
            lam_ref, C_ref = _two_stage_reference(GA, GB, s, 2)
            cond = np.linalg.cond(GB[:s, :s] / np.sqrt(np.outer(np.diag(GB)[:s], np.diag(GB)[:s])))
            np.testing.assert_allclose(th, lam_ref, rtol=1e-12 * max(cond, 1e3))
            assert not C[s:].any()
            np.testing.assert_allclose(C[:s].T @ GB[:s, :s] @ C[:s], np.eye(2), atol=1e-12 * max(cond, 1e3))
            rq = np.diag(C[:s].T @ GA[:s, :s] @ C[:s])
            np.testing.assert_allclose(rq, th, rtol=1e-10 * max(cond, 1e3))
            assert th[0] >= eigh(GA[:s, :s], GB[:s, :s], eigvals_only=True)[0] * (1 - 1e-12)
    bad = np.eye(6)
    bad[1, 1] = -1.0
    _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(np.eye(6)), _lib.ptr(bad), 4, 2, 2, 3, 1, 0,
                                             _lib.ptr(C), _lib.ptr(th), ctypes.byref(ok), None))
    assert ok.value == 0


def test_two_stage_rqi_lands_on_the_lowest_pair():
    """The 3 x 3 problems of the two-stage step are solved by Rayleigh-quotient iteration from e_0; when
    e_0 sits next to the SECOND eigenvector of its 3 x 3 pencil the iteration converges there, the
    lowestness check (leading minors) fails and the pair is deflated.  Either way the result must be
    the lowest pair: compared with the Jacobi-only form (negative sweep count) and with scipy."""
    import ctypes
    from scipy.linalg import eigh
    from cslam_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(21)
    for trial in range(40):
        GA, GB = np.zeros((6, 6)), np.zeros((6, 6))
        for c in range(2):
            idx = [c, 2 + c, 4 + c]
            Rm = rng.normal(size=(3, 3))
            B3 = np.eye(3) + 0.2 * (Rm + Rm.T) / 2            # SPD, moderately coupled
            lam = np.sort(rng.uniform(1e-6, 1.0, 3)) * (10.0 ** rng.integers(-6, 1))
            if trial % 4 == 3:
                lam[1] = lam[0] * (1 + 1e-4)                  # a close lowest pair
            V = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            # B-orthonormal eigenvectors Y with Y[:, k] ~ e_0 for k = trial % 3 (k = 1, 2: RQI lands elsewhere)
            Lc = np.linalg.cholesky(B3)
            k = trial % 3
            V[:, k] = Lc.T @ np.array([1.0, 0.0, 0.0]) + 0.05 * rng.normal(size=3)
            V = np.linalg.qr(V[:, [k] + [q for q in range(3) if q != k]])[0]
            order = [k] + [q for q in range(3) if q != k]
            Y = np.linalg.solve(Lc.T, V)                      # Y^T B Y = I
            lam_c = np.empty(3)
            lam_c[0], lam_c[1:] = lam[k], lam[[q for q in range(3) if q != k]]
            Yi = np.linalg.inv(Y)
            A3 = Yi.T @ np.diag(lam_c) @ Yi
            GA[np.ix_(idx, idx)] = 0.5 * (A3 + A3.T)
            GB[np.ix_(idx, idx)] = B3
        out = {}
        for sweeps in (8, -8):
            C, th, ok = np.zeros((6, 2)), np.zeros(2), ctypes.c_int()
            _lib.check(lib.cslam_debug_rayleigh_ritz(_lib.ptr(GA), _lib.ptr(GB), 6, 2, 2, sweeps, 1, 0,
                                                     _lib.ptr(C), _lib.ptr(th), ctypes.byref(ok), None))
            assert ok.value == 1
            out[sweeps] = (th.copy(), C.copy())
        lam_ref, _ = _two_stage_reference(GA, GB, 6, 2)
        # (a lowest pair closer than the check's delta = 1e-9 x scale is one value for the purposes of the step)
        rtol = 2e-4 if trial % 4 == 3 else 1e-9
        for sweeps in (8, -8):
            np.testing.assert_allclose(out[sweeps][0], lam_ref, rtol=rtol, atol=1e-30)
        np.testing.assert_allclose(out[8][0], out[-8][0], rtol=max(rtol, 1e-10))


def test_grid_barrier_hook_runs():
    """The barrier micro-benchmark (profiles/r2k_barrier_ubench.txt) returns a plausible cycle count
    for every recipe; recipe 0 is the barrier of the persistent kernels."""
    import ctypes
    from cslam_b200 import _lib
    lib = _lib.load()
    for variant in range(7):
        c = ctypes.c_int64()
        _lib.check(lib.cslam_debug_grid_barrier(16, 256, 200, 4, variant, 0, ctypes.byref(c)))
        assert 100 < c.value < 200000


def test_two_stage_solver_gives_the_same_selection(monkeypatch):
    """The whole Frank-Wolfe selection with the two-stage small eigen-solve: same Fiedler values
    and per-iteration sets as the reference goldens (the eigen-solver converges to the same pair;
    only its path differs)."""
    monkeypatch.setenv("CSLAM_RR_IMPL", "2")
    for tag in ("g1", "g2"):
        mac, k = _mac(tag)
        w0 = GOLD[f"{tag}_w0"]
        rounded, w, u = mac.fw_subset(w0.copy(), k, max_iters=20)
        assert np.array_equal(rounded, GOLD[f"{tag}_fw_rounded"])
        np.testing.assert_allclose(w, GOLD[f"{tag}_fw_w"], atol=1e-12)


def test_fused_tail_edge_cases():
    """k_fw_select / k_fw_prepare on the corner cases of mac.py:191-233: all gradients tied (equal
    weights: the reference's own tests use them), k = m, an immediate duality-gap exit (w must come
    back untouched, mac.py:223-225), and agreement of the fused kernels with the multi-kernel
    sequences they replace."""
    import os
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    rng = np.random.default_rng(4)
    n = 60
    fixed = [Edge(i, i + 1, 1.0) for i in range(n - 1)]
    pairs = set()
    while len(pairs) < 200:
        i, j = (int(x) for x in rng.integers(0, n, 2))
        if abs(i - j) > 1:
            pairs.add((min(i, j), max(i, j)))
    pairs = sorted(pairs)
    for weights in (np.ones(len(pairs)), rng.random(len(pairs))):
        cand = [Edge(i, j, float(w)) for (i, j), w in zip(pairs, weights)]
        k = 12
        w0 = np.zeros(len(cand))
        w0[np.argpartition(weights, -k)[-k:]] = 1.0
        results = {}
        for mode in ("fused", "multi"):
            os.environ["CSLAM_FW_FUSED"] = "1" if mode == "fused" else "0"
            os.environ["CSLAM_FW_FUSED_PREPARE"] = "1" if mode == "fused" else "0"
            try:
                mac = MAC(fixed, cand, n)
                results[mode] = mac.fw_subset(w0.copy(), k, max_iters=12, trace=True) + (mac.last_trace,)
            finally:
                os.environ.pop("CSLAM_FW_FUSED", None)
                os.environ.pop("CSLAM_FW_FUSED_PREPARE", None)
        (ra, wa, ua, ta), (rb, wb, ub, tb) = results["fused"], results["multi"]
        assert np.array_equal(ta[0], tb[0]), "per-iteration selections differ between the two forms"
        assert np.array_equal(ra, rb) and np.allclose(wa, wb, atol=1e-15) and abs(ua - ub) <= 1e-12 * abs(ub)
        assert ra.sum() == k and abs(wa.sum() - k) < 1e-9
        for it in range(12):
            assert len(set(ta[0][it].tolist())) == k and list(ta[0][it]) == sorted(ta[0][it])
    # k = m: everything is selected in every iteration
    mac = MAC(fixed, cand, n)
    r, w, u = mac.fw_subset(np.ones(len(cand)), len(cand), max_iters=3)
    assert r.sum() == len(cand) and np.allclose(w, 1.0)
    # immediate exit: with a huge gap tolerance the first iteration stops before the update
    mac = MAC(fixed, cand, n)
    r, w, u = mac.fw_subset(w0.copy(), k, max_iters=5, duality_gap_tol=1e9)
    assert mac.last_fw_iters == 1 and np.array_equal(w, w0)
    assert r.sum() == k


def test_fw_subset_with_an_empty_budget_and_an_empty_start():
    """k = 0 (mac.py:143-146: nothing is rounded up) and an all-zero start vector on a connected
    fixed graph: no active edge, the solver runs on the odometry Laplacian alone."""
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    n = 50
    fixed = [Edge(i, i + 1, 1.0) for i in range(n - 1)]
    cand = [Edge(0, 30, 0.5), Edge(5, 45, 0.7), Edge(10, 20, 0.2)]
    mac = MAC(fixed, cand, n)
    r, w, u = mac.fw_subset(np.zeros(3), 0, max_iters=3)
    assert r.sum() == 0 and not w.any()
    lam_path = 2.0 * (1.0 - np.cos(np.pi / n))            # Fiedler value of a path graph
    assert abs(mac.evaluate_objective(np.zeros(3)) - lam_path) < 1e-9
    r, w, u = mac.fw_subset(np.zeros(3), 2, max_iters=4)
    assert r.sum() == 2 and abs(w.sum() - 2) < 1e-12
