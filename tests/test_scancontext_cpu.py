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


def test_sparse_matching_facade_routes_lidar_descriptors(monkeypatch):
    """`frontend.sensor_type: lidar` makes LoopClosureSparseMatching build Scan Context matchers
    (reference cslam/loop_closure_sparse_matching.py:21-31); the routing (:36-72) and the batched
    forms give the same candidate edges.  The GPU matcher is replaced by the oracle here (host
    logic only)."""
    import types
    from collections import namedtuple
    import cslam_b200.lidar_pr.scancontext_matching as mod
    from cslam_b200.loop_closure_sparse_matching import LoopClosureSparseMatching
    monkeypatch.setattr(mod, "ScanContextMatching", ScanContextMatchingOracle)
    Msg = namedtuple("Msg", ["robot_id", "keyframe_id", "descriptor"])
    params = {"frontend.sensor_type": "lidar", "robot_id": 0, "max_nb_robots": 3,
              "frontend.similarity_threshold": 0.6, "frontend.nb_best_matches": 5,
              "frontend.intra_loop_min_inbetween_keyframes": 2,
              "frontend.enable_sparsification": True, "evaluation.enable_sparsification_comparison": False}
    pool, _, queries = sc_case("p4")
    one, many = LoopClosureSparseMatching(dict(params)), LoopClosureSparseMatching(dict(params))
    assert isinstance(one.local_nnsm, ScanContextMatchingOracle)
    assert all(isinstance(m, ScanContextMatchingOracle) for m in one.other_robots_nnsm.values())
    remote = [Msg(1 + t % 2, 10 + t, pool[t]) for t in range(4)]
    local = [np.roll(pool[t].reshape(20, 60), 7 + t, axis=1).ravel() for t in range(4)]
    a = [one.add_other_robot_global_descriptor(m) for m in remote]
    b = many.add_other_robot_global_descriptors(remote)
    assert a == b == [None] * 4                       # nothing local yet
    a = [m for t, d in enumerate(local) for m in one.add_local_global_descriptor(d, t)]
    b = many.add_local_global_descriptors(np.stack(local), range(4))
    assert [tuple(e) for e in a] == [tuple(e) for e in b] and len(a) == 4
    assert [(e.robot1_id, e.robot1_keyframe_id) for e in a] == [(1 + t % 2, 10 + t) for t in range(4)]
    assert all(e.weight > 0.999 for e in a)           # rotated revisits
    more = [Msg(2, 20 + t, np.roll(local[t].reshape(20, 60), 3, axis=1).ravel()) for t in range(4)]
    a2 = [one.add_other_robot_global_descriptor(m) for m in more]
    b2 = many.add_other_robot_global_descriptors(more)
    assert [tuple(e) for e in a2] == [tuple(e) for e in b2] and [e.robot0_keyframe_id for e in a2] == [0, 1, 2, 3]
    assert dict(one.candidate_selector.candidate_edges.items()) == dict(many.candidate_selector.candidate_edges.items())
    assert one.match_local_loop_closures(local[0], 9)[0] == 0
    with pytest.raises(NotImplementedError):
        many.match_local_loop_closures_batch(np.stack(local), [4, 5, 6, 7], 4)
