"""Candidate table: the columnar store must be indistinguishable from
the reference's plain dict (cslam/algebraic_connectivity_maximization.py:58,132-203,559-572),
`add_matches` from a sequence of `add_match` calls, and the vectorised set-up of
`select_candidates` from the edge-by-edge one.  Host logic only: the solver is replaced by a
deterministic stand-in, so this runs without a GPU."""
import random

import numpy as np
import pytest

from cslam_b200.algebraic_connectivity_maximization import (
    AlgebraicConnectivityMaximization as ACM, EdgeInterRobot, _count, _weights_of)
from cslam_b200.candidate_table import CandidateTable


def _snapshot(table):
    return [(k, tuple(v)) for k, v in table.items()]


def test_table_behaves_like_a_dict_under_random_edits():
    rng = random.Random(1)
    table, plain = CandidateTable(EdgeInterRobot, capacity=4), {}
    for step in range(6000):
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
            assert w.tolist() == [e.weight for e in plain.values()]
            raw_ends, raw_w, alive = table.raw()
            live = np.ones(len(raw_w), bool) if alive is None else alive
            assert raw_ends[live].tolist() == ends.tolist() and raw_w[live].tolist() == w.tolist()
            assert table.robots_present() == sorted({e[0] for e in plain.values()} | {e[2] for e in plain.values()})
    assert len(table) == len(plain) and list(table) == list(plain)
    assert table == plain
    with pytest.raises(KeyError):
        table[(9, 9, 9, 9)]


def _random_matches(rng, n, R, frames):
    r0 = rng.integers(0, R, n)
    r1 = (r0 + rng.integers(1, R, n)) % R                      # another robot
    return r0, rng.integers(0, frames, n), r1, rng.integers(0, frames, n), \
        np.round(rng.random(n), 1)                            # coarse weights: many exact ties


@pytest.mark.parametrize("seed", range(6))
def test_add_matches_equals_sequential_add_match(seed):
    rng = np.random.default_rng(seed)
    R = 4
    one, bulk = ACM(0, R), ACM(0, R)
    for rnd in range(4):
        m = _random_matches(rng, 400, R, 6)                   # 4*6 vertices: heavy duplication
        for t in range(len(m[0])):
            one.add_match(EdgeInterRobot(*(int(c[t]) for c in m[:4]), float(m[4][t])))
        bulk.add_matches(*m)
        assert _snapshot(one.candidate_edges) == _snapshot(bulk.candidate_edges)
        assert one.nb_poses == bulk.nb_poses
        # blacklist a few pairs (what a selection does) and go on
        gone = list(one.candidate_edges.values())[:5 + rnd]
        one.remove_candidate_edges(list(gone))
        bulk.remove_candidate_edges(list(gone))
        assert one.already_considered_matches == bulk.already_considered_matches
    assert len(one.candidate_edges) > 0


def test_add_matches_falls_back_for_intra_robot_pairs():
    one, bulk = ACM(0, 3), ACM(0, 3)
    m = (np.array([0, 1, 1]), np.array([1, 2, 5]), np.array([1, 1, 0]), np.array([4, 7, 2]),
         np.array([0.5, 0.25, 0.75]))
    for t in range(3):
        one.add_match(EdgeInterRobot(*(int(c[t]) for c in m[:4]), float(m[4][t])))
    bulk.add_matches(*m)
    assert _snapshot(one.candidate_edges) == _snapshot(bulk.candidate_edges)
    assert one.nb_poses == bulk.nb_poses


def _stub_solver(self, fixed, candidates, w_init, budget):
    """Deterministic stand-in for the GPU solver: remembers what it was given and picks the
    `budget` candidates with the largest weight * |i - j|."""
    def cols(edges):
        if isinstance(edges, tuple):
            return tuple(np.asarray(c).tolist() for c in edges)
        return ([e.i for e in edges], [e.j for e in edges], [e.weight for e in edges])
    self.seen = (cols(fixed), cols(candidates), np.asarray(w_init).tolist(), budget, self.total_nb_poses)
    i, j, w = (np.asarray(c, dtype=np.float64) for c in cols(candidates))
    out = np.zeros(len(w))
    out[np.argsort(-(w * np.abs(i - j)), kind="stable")[:budget]] = 1.0
    return out


@pytest.mark.parametrize("greedy", [True, False])
@pytest.mark.parametrize("seed", range(5))
def test_columnar_select_candidates_equals_edge_by_edge(monkeypatch, seed, greedy):
    monkeypatch.setattr(ACM, "run_mac_solver", _stub_solver)
    rng = np.random.default_rng(100 + seed)
    R = 5
    cfg = {"frontend.enable_sparsification": True, "evaluation.enable_sparsification_comparison": True}
    fast, slow = ACM(0, R, extra_params=cfg), ACM(0, R, extra_params=cfg)
    slow.candidate_edges = {}                                  # plain dict: the edge-by-edge path
    nrob = 4 if seed % 2 else R                                # odd seeds: robot 4 never appears
    for acm in (fast, slow):
        for r in range(nrob - 1):
            acm.add_fixed_edge(EdgeInterRobot(r, 3 + r, r + 1, 2, 1.0))
    m = _random_matches(rng, 300, nrob, 40)
    fast.add_matches(*m)
    for t in range(300):
        slow.add_match(EdgeInterRobot(*(int(c[t]) for c in m[:4]), float(m[4][t])))
    assert _snapshot(fast.candidate_edges) == _snapshot(slow.candidate_edges)
    in_range = {r: True for r in range(R)}
    if seed == 2:
        in_range[3] = False                                    # a robot out of range: edges filtered
    for rnd in range(3):
        np.random.seed(seed)
        a = fast.select_candidates(7, in_range, greedy_initialization=greedy)
        np.random.seed(seed)
        b = slow.select_candidates(7, in_range, greedy_initialization=greedy)
        assert [tuple(e) for e in a] == [tuple(e) for e in b] and len(a) == 7
        assert fast.seen == slow.seen
        assert fast.offsets == slow.offsets and fast.total_nb_poses == slow.total_nb_poses
        assert [tuple(e) for e in fast.log_mac_edges] == [tuple(e) for e in slow.log_mac_edges]
        assert [tuple(e) for e in fast.log_greedy_edges] == [tuple(e) for e in slow.log_greedy_edges]
        assert _snapshot(fast.candidate_edges) == _snapshot(slow.candidate_edges)
        assert fast.already_considered_matches == slow.already_considered_matches
        # verified edges come back as measurements (…detection.py:449-484)
        fast.candidate_edges_to_fixed(list(a[:2]))
        slow.candidate_edges_to_fixed(list(b[:2]))
        assert [tuple(e) for e in fast.fixed_edges] == [tuple(e) for e in slow.fixed_edges]


def test_columnar_connection_biased_branch(monkeypatch):
    """No measured inter-robot edge yet: the greedy branch (:519-524) runs on the same edges."""
    rng = np.random.default_rng(5)
    fast, slow = ACM(0, 3), ACM(0, 3)
    slow.candidate_edges = {}
    m = _random_matches(rng, 60, 3, 15)
    fast.add_matches(*m)
    for t in range(60):
        slow.add_match(EdgeInterRobot(*(int(c[t]) for c in m[:4]), float(m[4][t])))
    a = fast.select_candidates(5, {0: True, 1: True, 2: True})
    b = slow.select_candidates(5, {0: True, 1: True, 2: True})
    assert [tuple(e) for e in a] == [tuple(e) for e in b] and len(a) == 5


def test_helpers_accept_both_forms():
    cols = (np.array([0, 1], np.int32), np.array([5, 9], np.int32), np.array([0.25, 0.5]))
    assert _count(cols) == 2 and list(_weights_of(cols)) == [0.25, 0.5]
    assert ACM().greedy_initialization(1, cols).tolist() == [0.0, 1.0]


@pytest.mark.parametrize("seed", range(8))
def test_bulk_path_reproduces_the_reference_golden_sequences(seed):
    """tests/golden/candidates.npz holds the REFERENCE class's candidate dictionary after seeded
    rounds of add_match / remove_candidate_edges / candidate_edges_to_fixed
    (oracle/make_golden_candidates.py); `add_matches` + the candidate table must land on the same
    dictionary (keys, stored spelling, weights, ORDER), nb_poses, blacklist and fixed edges."""
    import os
    from oracle.make_golden_candidates import scenario
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "candidates.npz"))
    R, rounds = scenario(seed)
    acm = ACM(robot_id=0, max_nb_robots=R)

    def snap():
        return np.array([list(k) + list(v) for k, v in acm.candidate_edges.items()],
                        dtype=np.float64).reshape(-1, 9)

    for rnd, (m, n_remove, n_fix) in enumerate(rounds):
        acm.add_matches(m[:, 0].astype(int), m[:, 1].astype(int), m[:, 2].astype(int), m[:, 3].astype(int), m[:, 4])
        assert np.array_equal(snap(), gold[f"s{seed}_r{rnd}_after_add"])
        assert [acm.nb_poses[r] for r in range(R)] == gold[f"s{seed}_r{rnd}_nb_poses"].tolist()
        oldest = list(acm.candidate_edges.values())
        acm.remove_candidate_edges(oldest[:n_remove])
        acm.candidate_edges_to_fixed(list(oldest[n_remove:n_remove + n_fix]))
        assert np.array_equal(snap(), gold[f"s{seed}_r{rnd}_after_edit"])
        assert sorted(acm.already_considered_matches) == [tuple(r) for r in gold[f"s{seed}_r{rnd}_considered"].tolist()]
        assert np.array_equal(np.array([list(e) for e in acm.fixed_edges], dtype=np.float64).reshape(-1, 5),
                              gold[f"s{seed}_r{rnd}_fixed"])
