"""Device hash index of the candidate graph (csrc/keymap.cu,
candidate_table.py device mode): the key -> slot map lives in an open-addressing table in HBM.
Same contracts as the host-mode tests of tests/test_candidate_table_cpu.py - the table is
indistinguishable from the reference's dict (cslam/algebraic_connectivity_maximization.py:58,
132-203,559-572), bulk `add_matches` equals sequential `add_match` (pinned to the reference's own
outputs in tests/golden/candidates.npz) - plus the raw key map against a Python dict."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _snapshot(table):
    return [(k, tuple(v)) for k, v in table.items()]


def test_keymap_against_a_python_dict():
    from cslam_b200.candidate_table import DeviceKeyMap
    rng = np.random.default_rng(0)
    km, ref = DeviceKeyMap(0, 16), {}
    universe = rng.integers(0, 1 << 62, 50000).astype(np.uint64)
    for rnd in range(30):
        keys = np.unique(rng.choice(universe, size=int(rng.integers(1, 8000))))
        op = rnd % 3
        if op in (0, 1):
            vals = rng.integers(0, 1 << 30, len(keys)).astype(np.int32)
            km.insert(keys, vals)
            ref.update(zip(keys.tolist(), vals.tolist()))
        else:
            erased = km.erase(keys)
            expect = np.array([ref.pop(k, -1) for k in keys.tolist()], dtype=np.int32)
            assert np.array_equal(erased, expect)
        probe = rng.choice(universe, size=5000)
        got = km.lookup(probe)
        assert np.array_equal(got, np.array([ref.get(k, -1) for k in probe.tolist()], dtype=np.int32))
        assert len(km) == len(ref)
    assert len(ref) > 1000
    with pytest.raises(Exception):
        km.insert(np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64), np.array([1], dtype=np.int32))


def test_table_in_device_mode_behaves_like_a_dict_under_random_edits():
    from cslam_b200.algebraic_connectivity_maximization import EdgeInterRobot
    from cslam_b200.candidate_table import CandidateTable
    rng = random.Random(1)
    table, plain = CandidateTable(EdgeInterRobot, capacity=4), {}
    assert table.use_device_index(0) and table.device_mode
    for step in range(3000):
        key = (rng.randint(0, 2), rng.randint(0, 12), rng.randint(3, 4), rng.randint(0, 12))
        op = rng.random()
        if op < 0.55:
            e = EdgeInterRobot(*key, rng.random())
            table[key] = e
            plain[key] = e
        elif op < 0.8:
            assert table.pop(key, None) == plain.pop(key, None)
        elif op < 0.9 and key in plain:
            del table[key]
            del plain[key]
        else:
            assert (key in table) == (key in plain)
            assert table.get(key) == plain.get(key)
        if step % 500 == 0:
            assert _snapshot(table) == _snapshot(plain)       # same content AND same order
            ends, w = table.columns()
            assert [tuple(r) for r in ends.tolist()] == [tuple(e[:4]) for e in plain.values()]
    assert len(table) == len(plain) and list(table) == list(plain) and table.device_mode
    with pytest.raises(KeyError):
        table[(9, 9, 9, 9)]
    # a key that does not fit the 64-bit packing moves the table back to the Python dict, intact
    big = (0, 1 << 30, 1, 5)
    table[big] = EdgeInterRobot(*big, 0.5)
    plain[big] = EdgeInterRobot(*big, 0.5)
    assert not table.device_mode and _snapshot(table) == _snapshot(plain)


@pytest.mark.parametrize("seed", range(4))
def test_bulk_add_matches_with_the_device_index_equals_sequential_add_match(seed):
    from cslam_b200.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization as ACM, EdgeInterRobot)
    rng = np.random.default_rng(seed)
    R = 4
    one = ACM(0, R, extra_params={"frontend.enable_sparsification": True, "frontend.candidate_index": "host",
                                  "evaluation.enable_sparsification_comparison": False})
    bulk = ACM(0, R)
    for rnd in range(4):
        n = 400
        r0 = rng.integers(0, R, n)
        r1 = (r0 + rng.integers(1, R, n)) % R
        m = (r0, rng.integers(0, 6, n), r1, rng.integers(0, 6, n), np.round(rng.random(n), 1))
        for t in range(n):
            one.add_match(EdgeInterRobot(*(int(c[t]) for c in m[:4]), float(m[4][t])))
        bulk.add_matches(*m)
        assert bulk.candidate_edges.device_mode and not one.candidate_edges.device_mode
        assert _snapshot(one.candidate_edges) == _snapshot(bulk.candidate_edges)
        assert one.nb_poses == bulk.nb_poses
        gone = list(one.candidate_edges.values())[:5 + rnd]
        one.remove_candidate_edges(list(gone))
        bulk.remove_candidate_edges(list(gone))
        assert one.already_considered_matches == bulk.already_considered_matches
        assert _snapshot(one.candidate_edges) == _snapshot(bulk.candidate_edges)
    assert len(one.candidate_edges) > 0


def test_device_index_reproduces_the_reference_golden_sequences():
    """tests/golden/candidates.npz: the REFERENCE class's candidate dictionary, blacklist and
    nb_poses after seeded rounds of add_match / remove_candidate_edges / candidate_edges_to_fixed
    (oracle/make_golden_candidates.py).  Replayed through the bulk path with the device index."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "candidate_table_cpu_tests", os.path.join(os.path.dirname(__file__), "test_candidate_table_cpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    host_test = mod.test_bulk_path_reproduces_the_reference_golden_sequences
    import cslam_b200.algebraic_connectivity_maximization as acm
    seen = []
    orig = acm.CandidateTable.use_device_index

    def spy(self, device=0):
        seen.append(1)
        return orig(self, device)
    acm.CandidateTable.use_device_index = spy
    try:
        for seed in range(3):
            host_test(seed)
    finally:
        acm.CandidateTable.use_device_index = orig
    assert seen, "the bulk path never switched to the device index"


def test_million_matches_bulk_insert_and_removal():
    """C5 scale: 1M distinct inter-robot matches inserted in one call, 1000 removed (what a
    selection does), a second overlapping batch merged - sizes, weights and order as the dict rules say."""
    import time
    from cslam_b200.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization as ACM, EdgeInterRobot)
    rng = np.random.default_rng(0)
    R, P, m = 8, 12500, 1000000
    r0 = rng.integers(0, R, m)
    r1 = (r0 + rng.integers(1, R, m)) % R
    k0, k1, w = rng.integers(0, P, m), rng.integers(0, P, m), rng.random(m)
    acm = ACM(0, R)
    t0 = time.time()
    acm.add_matches(r0, k0, r1, k1, w)
    t_insert = time.time() - t0
    table = acm.candidate_edges
    assert table.device_mode
    # distinct normalised pairs
    lo_first = r0 < r1
    a = np.where(lo_first, r0, r1) * P + np.where(lo_first, k0, k1)
    b = np.where(lo_first, r1, r0) * P + np.where(lo_first, k1, k0)
    assert len(table) == len(np.unique(a * (R * P) + b))
    ends, wt = table.columns()
    assert len(wt) == len(table) and wt.min() >= 0
    gone = [EdgeInterRobot(*(int(x) for x in ends[i]), float(wt[i])) for i in range(0, 100000, 100)]
    acm.remove_candidate_edges(gone)
    assert len(table) == len(wt) - len(gone)
    for e in gone[:5]:
        assert acm.edge_key(e) not in table and acm.edge_key(e) in acm.already_considered_matches
    # the removed pairs are blacklisted: offering them again changes nothing
    n_before = len(table)
    acm.add_matches([e.robot0_id for e in gone], [e.robot0_keyframe_id for e in gone],
                    [e.robot1_id for e in gone], [e.robot1_keyframe_id for e in gone], [1.0] * len(gone))
    assert len(table) == n_before
    print(f"bulk insert of {m} matches: {t_insert:.3f} s")
    assert t_insert < 1.0
