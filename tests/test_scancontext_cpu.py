"""The Scan Context oracle (oracle/scancontext.py) against the reference's own outputs
(tests/golden/scancontext.npz, made by oracle/make_golden_sc.py from
cslam/lidar_pr/scancontext_matching.py).  CPU only; sized to run in a few seconds."""
import os

import numpy as np
import pytest

from oracle.inputs import SC_CASES, sc_case
from oracle.scancontext import ScanContextMatchingOracle, distance_sc, sc2rk

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "scancontext.npz"))


@pytest.mark.parametrize("tag", ["p300", "p4", "p1"])
def test_oracle_reproduces_reference(tag):
    pool, items, queries = sc_case(tag)
    m = ScanContextMatchingOracle()
    for row, item in zip(pool, items):
        m.add_item(row, item)
    assert np.array_equal(m.ringkeys[:m.nb_items], GOLD[tag + "_ringkeys"])     # bit-exact
    assert len(m.ringkeys) == int(GOLD[tag + "_capacity"])
    nq = min(len(queries), 8)                      # the pure-Python distance is slow
    for t in list(range(nq - 1)) + [len(queries) - 1]:
        row, sim, yaw, cand, dist = m.search_details(queries[t])
        ref_cand = GOLD[tag + "_cand"][t]
        ref_cand = ref_cand[ref_cand >= 0]
        assert list(cand) == list(ref_cand), (tag, t)
        np.testing.assert_allclose(dist, GOLD[tag + "_cand_dist"][t][:len(cand)], rtol=0, atol=1e-12)
        a, s = m.search(queries[t], 1)
        assert a[0] == GOLD[tag + "_items"][t]
        assert abs(s[0] - GOLD[tag + "_sims"][t]) <= 1e-12
        if row >= 0:
            assert yaw == GOLD[tag + "_cand_yaw"][t][list(cand).index(row)]


def test_oracle_growth_and_empty_pool():
    m = ScanContextMatchingOracle()
    assert m.search(np.zeros(1200), 1) == ([None], [None])
    assert m.search_best(np.zeros(1200)) == (None, None)
    pool, items, _ = sc_case("p1100")
    for row, item in zip(pool, items):
        m.add_item(row, item)
    assert len(m.ringkeys) == int(GOLD["p1100_capacity"]) == 2000
    assert np.array_equal(m.ringkeys[:1100], GOLD["p1100_ringkeys"])


def test_distance_is_rotation_invariant_and_zero_safe():
    rng = np.random.default_rng(0)
    sc = rng.random((20, 60))
    sc[:, ::7] = 0.0
    for shift in (1, 17, 59):
        d, yaw = distance_sc(np.roll(sc, -shift, axis=1), sc)       # candidate rolled back by `shift`
        assert abs(d) < 1e-12 and yaw == shift
    assert distance_sc(np.zeros((20, 60)), sc) == (1.0, 1)
    assert np.allclose(sc2rk(sc), sc.mean(axis=1))
