"""GPU parity of ScanContextMatching (csrc/scancontext.cu through the C ABI) against the
reference's own outputs (tests/golden/scancontext.npz) and the numpy oracle.

Bars: ring keys bit-exact; candidate rows and matched items exact; column-shift distances and
similarities within 1e-12 (float64 throughout; only the summation order inside a 20-element dot
product differs from BLAS); yaw shift exact."""
import os

import numpy as np
import pytest

from oracle.inputs import SC_CASES, sc_case
from oracle.scancontext import ScanContextMatchingOracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "scancontext.npz"))


def _gpu(pool, items, one_by_one=False, **kw):
    from cslam_b200.lidar_pr.scancontext_matching import ScanContextMatching
    m = ScanContextMatching(**kw)
    if one_by_one:
        for row, item in zip(pool, items):
            m.add_item(row, item)
    else:
        m.add_items(pool, items)
    return m


@pytest.mark.parametrize("tag", list(SC_CASES))
def test_matches_reference_golden(tag):
    pool, items, queries = sc_case(tag)
    m = _gpu(pool, items, one_by_one=(tag in ("p4", "p1")))
    assert m.nb_items == len(pool) and m.items[0] == items[0]
    assert m.ringkeys.shape == (int(GOLD[tag + "_capacity"]), 20)
    assert np.array_equal(m.ringkeys[:len(pool)], GOLD[tag + "_ringkeys"])
    assert np.array_equal(m.scancontexts[:len(pool)].reshape(len(pool), -1), pool)
    assert not m.scancontexts[len(pool):].any()
    rows, sims, yaw, cand, cdist = m.search_batch(queries, details=True)
    ref_cand = GOLD[tag + "_cand"]
    assert np.array_equal(cand, ref_cand)
    live = ref_cand >= 0
    np.testing.assert_allclose(cdist[live], GOLD[tag + "_cand_dist"][live], rtol=0, atol=1e-12)
    for t, q in enumerate(queries):               # the reference API, one query at a time
        a, s = m.search(q, 1)
        b, s2 = m.search_best(q)
        assert a[0] == b == GOLD[tag + "_items"][t]
        assert abs(s[0] - GOLD[tag + "_sims"][t]) <= 1e-12 and s[0] == s2
        if rows[t] >= 0:
            c = list(cand[t]).index(rows[t])
            assert yaw[t] == GOLD[tag + "_cand_yaw"][t][c]
            assert m.last_yaw_diff_deg == yaw[t] * 6.0
    assert rows[-1] == -1 and sims[-1] == 0.0      # the all-zero scan: item 0, similarity 0 (:81-84)
    assert m.search(queries[-1], 1) == ([items[0]], [0.0])


def test_empty_pool_and_float32_input():
    from cslam_b200.lidar_pr.scancontext_matching import ScanContextMatching
    m = ScanContextMatching()
    assert m.search(np.zeros(1200), 5) == ([None], [None])
    assert m.search_best(np.zeros(1200)) == (None, None)
    pool, items, queries = sc_case("p4")
    m.add_items(pool.astype(np.float32), items)    # lidar heights arrive as float32
    o = ScanContextMatchingOracle()
    for row, item in zip(pool.astype(np.float32).astype(np.float64), items):
        o.add_item(row, item)
    assert np.array_equal(m.ringkeys[:4], o.ringkeys[:4])
    for q in queries:
        (a,), (s,) = m.search(q.astype(np.float32), 1)
        (b,), (s2,) = o.search(q.astype(np.float32).astype(np.float64), 1)
        assert a == b and abs(s - s2) <= 1e-12


@pytest.mark.parametrize("shape,ncand", [((20, 60), 3), ((12, 40), 10), ((8, 136), 16)])
def test_other_shapes_and_candidate_counts_against_oracle(shape, ncand):
    rng = np.random.default_rng(shape[1] + ncand)
    n = 500
    pool = rng.random((n, shape[0] * shape[1])) * (rng.random((n, shape[0] * shape[1])) > 0.25)
    queries = np.stack([np.roll(pool[i].reshape(shape), i % shape[1], axis=1).ravel() for i in range(0, 60, 10)])
    queries = queries + (queries > 0) * rng.normal(0, 0.05, queries.shape)
    m = _gpu(pool, list(range(n)), shape=list(shape), num_candidates=ncand)
    o = ScanContextMatchingOracle(shape=list(shape), num_candidates=ncand)
    for i in range(n):
        o.add_item(pool[i], i)
    assert np.array_equal(m.ringkeys[:n], o.ringkeys[:n])
    rows, sims, yaw, cand, cdist = m.search_batch(queries, details=True)
    for t, q in enumerate(queries):
        row, sim, y, oc, od = o.search_details(q)
        assert list(cand[t]) == list(oc)
        np.testing.assert_allclose(cdist[t], od, rtol=0, atol=1e-12)
        assert rows[t] == row and yaw[t] == y and abs(sims[t] - sim) <= 1e-12


def test_large_pool_properties():
    """100k entries (no oracle at this size): a rotated copy of a pool entry finds that entry
    with similarity 1 and the rotation as yaw shift; the nearest ring key of an exact copy is
    the entry itself; answers do not depend on how the queries are batched."""
    rng = np.random.default_rng(3)
    n, R, S = 100_000, 20, 60
    pool = (rng.random((n, R * S)) * 5).astype(np.float32)
    pool *= rng.random((n, R * S)) > 0.3
    m = _gpu(pool, list(range(n)))
    picks = rng.integers(0, n, 48)
    shifts = rng.integers(1, S, 48)
    queries = np.stack([np.roll(pool[p].reshape(R, S), s, axis=1).ravel() for p, s in zip(picks, shifts)])
    rows, sims, yaw, cand, cdist = m.search_batch(queries, details=True)
    assert np.array_equal(rows, picks)
    assert np.array_equal(cand[:, 0], picks)       # rolling columns leaves the ring key unchanged
    np.testing.assert_allclose(sims, 1.0, rtol=0, atol=1e-12)
    assert np.array_equal(yaw, shifts)               # candidate rolled by `shift` columns == query
    one = [m.search_batch(q[None])[0][0] for q in queries[:6]]
    assert one == list(picks[:6])
    assert m.nb_items == n and m.ringkeys.shape[0] == 128_000
