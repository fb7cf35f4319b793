"""Host-side bookkeeping of AlgebraicConnectivityMaximization against the reference class itself
on random graphs (runs where the reference checkout is mounted, i.e. in the build container;
skipped on the GPU box).  Only methods that need no solver are exercised here; the solver
path is compared through tests/golden/mac.npz and frontend.npz on the GPU."""
import os
import random
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cslam")),
                                reason="reference checkout not mounted")


def _tuples(edges):
    return [tuple(e) for e in edges]


def test_bookkeeping_matches_reference_on_random_graphs():
    sys.path.insert(0, REF)
    try:
        from cslam.algebraic_connectivity_maximization import (
            AlgebraicConnectivityMaximization as RefACM, EdgeInterRobot as RefEdge)
    finally:
        sys.path.remove(REF)
    from cslam_b200.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization as ACM, EdgeInterRobot as Edge)
    rng = random.Random(0)
    for trial in range(150):
        R = rng.randint(2, 5)
        ref, me = RefACM(robot_id=0, max_nb_robots=R), ACM(robot_id=0, max_nb_robots=R)
        for _ in range(rng.randint(1, 30)):
            r0, r1 = rng.sample(range(R), 2)
            e = (r0, rng.randint(0, 20), r1, rng.randint(0, 20), rng.random())
            if rng.random() < 0.2:
                ref.add_fixed_edge(RefEdge(*e))
                me.add_fixed_edge(Edge(*e))
            else:   # includes the un-normalised key lookup quirk of add_match
                ref.add_match(RefEdge(*e))
                me.add_match(Edge(*e))
        assert {k: tuple(v) for k, v in ref.candidate_edges.items()} == \
            {k: tuple(v) for k, v in me.candidate_edges.items()}
        assert ref.nb_poses == me.nb_poses
        assert ref.initial_fixed_edge_exists == me.initial_fixed_edge_exists
        in_range = {r: (r == 0 or rng.random() < 0.7) for r in range(R)}
        inc_ref, inc = ref.check_graph_disconnections(in_range), me.check_graph_disconnections(in_range)
        assert inc_ref == inc
        assert ref.check_initial_fixed_measurements_exists(inc_ref) == \
            me.check_initial_fixed_measurements_exists(inc)
        ref.compute_offsets(inc_ref)
        me.compute_offsets(inc)
        assert ref.offsets == me.offsets
        rk_ref = ref.rekey_edges(ref.candidate_edges.values(), inc_ref)
        rk = me.rekey_edges(me.candidate_edges.values(), inc)
        assert _tuples(rk_ref) == _tuples(rk)
        assert _tuples(ref.fill_odometry()) == _tuples(me.fill_odometry())
        assert _tuples(ref.recover_inter_robot_edges(rk_ref, inc_ref)) == \
            _tuples(me.recover_inter_robot_edges(rk, inc))
        e_ref = ref.get_included_edges(ref.candidate_edges.values(), inc_ref)
        e_me = me.get_included_edges(me.candidate_edges.values(), inc)
        if e_ref:
            k = rng.randint(1, len(e_ref))
            assert np.array_equal(ref.connection_biased_greedy_selection(k, list(e_ref), inc_ref),
                                  me.connection_biased_greedy_selection(k, list(e_me), inc))
            assert np.array_equal(ref.greedy_initialization(k, rk_ref), me.greedy_initialization(k, rk))
            np.random.seed(trial)
            a = ref.pseudo_greedy_initialization(k, min(k, 2), rk_ref)
            np.random.seed(trial)
            b = me.pseudo_greedy_initialization(k, min(k, 2), rk)
            assert np.array_equal(a, b)
            np.random.seed(trial)
            a = ref.random_initialization(k, list(rk_ref))
            np.random.seed(trial)
            b = me.random_initialization(k, list(rk))
            assert np.array_equal(a, b)
        if me.candidate_edges:
            key = sorted(me.candidate_edges)[0]
            ref.candidate_edges_to_fixed([ref.candidate_edges[key]])
            me.candidate_edges_to_fixed([me.candidate_edges[key]])
            assert _tuples(ref.fixed_edges) == _tuples(me.fixed_edges)
            assert set(ref.candidate_edges) == set(me.candidate_edges)
            assert ref.already_considered_matches == me.already_considered_matches
            # a blacklisted pair cannot come back
            me.add_match(Edge(*key, 0.99))
            ref.add_match(RefEdge(*key, 0.99))
            assert set(ref.candidate_edges) == set(me.candidate_edges)


def test_bulk_add_matches_equals_the_reference_class_fed_one_by_one():
    """`add_matches` (candidate table, bulk) against the REFERENCE class receiving the same matches
    through its own `add_match` one at a time: same dictionary (keys, stored spelling, weights,
    order), same nb_poses — including reversed-key overwrites, equal weights and blacklisted pairs."""
    sys.path.insert(0, REF)
    try:
        from cslam.algebraic_connectivity_maximization import (
            AlgebraicConnectivityMaximization as RefACM, EdgeInterRobot as RefEdge)
    finally:
        sys.path.remove(REF)
    from cslam_b200.algebraic_connectivity_maximization import AlgebraicConnectivityMaximization as ACM
    rng = np.random.default_rng(3)
    for trial in range(40):
        R = int(rng.integers(2, 6))
        ref, me = RefACM(robot_id=0, max_nb_robots=R), ACM(robot_id=0, max_nb_robots=R)
        for rnd in range(3):
            n = int(rng.integers(1, 120))
            r0 = rng.integers(0, R, n)
            r1 = (r0 + rng.integers(1, R, n)) % R
            k0, k1 = rng.integers(0, 5, n), rng.integers(0, 5, n)
            w = np.round(rng.random(n), 1)
            for t in range(n):
                ref.add_match(RefEdge(int(r0[t]), int(k0[t]), int(r1[t]), int(k1[t]), float(w[t])))
            me.add_matches(r0, k0, r1, k1, w)
            assert [(k, tuple(v)) for k, v in ref.candidate_edges.items()] == \
                [(k, tuple(v)) for k, v in me.candidate_edges.items()], (trial, rnd)
            assert ref.nb_poses == me.nb_poses
            gone = list(ref.candidate_edges.values())[:3]
            ref.remove_candidate_edges([RefEdge(*e) for e in gone])
            me.remove_candidate_edges(list(me.candidate_edges.values())[:3])
            assert set(ref.candidate_edges) == set(me.candidate_edges)
            assert ref.already_considered_matches == me.already_considered_matches


def test_columnar_select_candidates_equals_the_reference_end_to_end(monkeypatch):
    """Whole `select_candidates` of this package (candidate table, vectorised set-up) against the
    REFERENCE `select_candidates` running its own MAC (scipy/networkx) on small multi-robot graphs.
    The GPU solver is replaced by the oracle's restatement of MAC here (no GPU in the build
    container); what is compared is everything around it: inclusion, offsets, rekeying, the greedy
    start, recovery of the (robot, keyframe) edges and the removal of the selection, over
    successive rounds."""
    sys.path.insert(0, REF)
    try:
        from cslam.algebraic_connectivity_maximization import (
            AlgebraicConnectivityMaximization as RefACM, EdgeInterRobot as RefEdge)
    finally:
        sys.path.remove(REF)
    from cslam_b200.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization as ACM, EdgeInterRobot as Edge)
    from oracle.inputs import multi_robot_graph
    from oracle.mac import MACOracle

    def oracle_solver(self, fixed, candidates, w_init, budget):
        assert isinstance(fixed, tuple) and isinstance(candidates, tuple)      # the columnar path
        mac = MACOracle.from_arrays(fixed, candidates, self.total_nb_poses)
        return mac.fw_subset(w_init, budget, max_iters=self.max_iters)[0]

    monkeypatch.setattr(ACM, "run_mac_solver", oracle_solver)
    for R, P, m, k, seed in ((3, 15, 40, 5, 0), (4, 12, 60, 6, 1)):
        fixed, cand = multi_robot_graph(R, P, m, seed)
        ref, me = RefACM(robot_id=0, max_nb_robots=R), ACM(robot_id=0, max_nb_robots=R)
        ref.set_graph([RefEdge(*e) for e in fixed], [RefEdge(*e) for e in cand])
        me.set_graph([Edge(*e) for e in fixed], [Edge(*e) for e in cand])
        in_range = {r: True for r in range(R)}
        for rnd in range(2):
            a = ref.select_candidates(k, in_range, greedy_initialization=True)
            b = me.select_candidates(k, in_range, greedy_initialization=True)
            assert sorted(tuple(e)[:4] for e in a) == sorted(tuple(e)[:4] for e in b) and len(b) == k
            assert set(ref.candidate_edges) == set(me.candidate_edges)
            assert ref.already_considered_matches == me.already_considered_matches
            assert ref.offsets == me.offsets and ref.total_nb_poses == me.total_nb_poses
