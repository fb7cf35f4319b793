"""CPU tests of the host-side logic around the GPU path: front-end oracle vs the golden
reference outputs, Broker (modelled on the reference's tests/test_broker.py), message
chunking, neighbour bookkeeping and the in-process node/bus."""
import math
import os
import random

import numpy as np
import pytest

from cslam_b200.algebraic_connectivity_maximization import EdgeInterRobot
from cslam_b200.broker import Broker
from cslam_b200.local_node import LocalBus, LocalNode
from cslam_b200.msgs import UInt32
from cslam_b200.neighbors_manager import NeighborManager
from cslam_b200.utils.misc import dict_to_list_chunks, list_chunks, list_range

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_frontend_oracle_matches_reference_golden():
    from oracle.frontend import FrontendOracle
    from oracle.inputs import FRONTEND_PARAMS, frontend_scenario
    g = np.load(os.path.join(GOLD, "frontend.npz"))
    orc = FrontendOracle(dict(FRONTEND_PARAMS))
    intra, matches = [], []
    for ev in frontend_scenario():
        if ev[0] == 'local':
            for kf, d in zip(ev[1], ev[2]):
                i, new = orc.local_keyframe(d, kf)
                intra.append((kf, -1 if i is None else i))
                matches.extend(new)
        else:
            for kf, d in zip(ev[2], ev[3]):
                m = orc.remote_keyframe(ev[1], np.asarray(d.tolist()), kf)   # float32[] on the wire
                if m is not None:
                    matches.append(m)
    assert np.array_equal(np.array(intra), g["intra"])
    got = np.array(matches)
    assert np.array_equal(got[:, :4].astype(int), g["matches"][:, :4].astype(int))
    np.testing.assert_allclose(got[:, 4], g["matches"][:, 4], atol=1e-12)
    keys = sorted(orc.candidates)
    assert np.array_equal(np.array(keys), g["cand_keys"])
    np.testing.assert_allclose([orc.candidates[k][4] for k in keys], g["cand_weights"], atol=1e-12)


# ---- Broker ---------------------------------------------------------------------------
def _random_edges(nb_robots, nb_poses, nb_edges, rng):
    edges, seen = [], set()
    while len(edges) < nb_edges:
        r0, r1 = rng.sample(range(nb_robots), 2)
        e = EdgeInterRobot(r0, rng.randrange(nb_poses), r1, rng.randrange(nb_poses), rng.random())
        key = (min((e[0], e[1]), (e[2], e[3])), max((e[0], e[1]), (e[2], e[3])))
        if key not in seen:
            seen.add(key)
            edges.append(e)
    return edges


def _check_cover(edges, robots, components):
    vertices = [v for c in components for v in c]
    assert len(vertices) == len(set(vertices))           # no duplicates
    cover = set(vertices)
    for e in edges:
        if e.robot0_id in robots and e.robot1_id in robots:
            assert (e.robot0_id, e.robot0_keyframe_id) in cover or \
                (e.robot1_id, e.robot1_keyframe_id) in cover
    return cover


@pytest.mark.parametrize("nb_robots,nb_edges", [(2, 10), (2, 60), (3, 40), (5, 100)])
@pytest.mark.parametrize("use_vertex_cover", [True, False])
def test_broker_cover_properties(nb_robots, nb_edges, use_vertex_cover):
    rng = random.Random(nb_robots * 1000 + nb_edges)
    np.random.seed(0)
    edges = _random_edges(nb_robots, 30, nb_edges, rng)
    robots = list(range(nb_robots))
    broker = Broker(edges, robots)
    comps = broker.brokerage(use_vertex_cover)
    cover = _check_cover(edges, robots, comps)
    assert len(cover) <= nb_edges
    if nb_robots == 2 and use_vertex_cover:
        nb_vertices = len({(e[0], e[1]) for e in edges} | {(e[2], e[3]) for e in edges})
        assert len(cover) <= math.ceil(nb_vertices / 2)      # reference test_broker.py:41-102


def test_broker_bipartite_cover_is_minimum():
    nx = pytest.importorskip("networkx")
    rng = random.Random(5)
    for trial in range(20):
        edges = _random_edges(2, 12, rng.randrange(5, 40), rng)
        comps = Broker(edges, [0, 1]).brokerage(True)
        cover = _check_cover(edges, [0, 1], comps)
        g = nx.Graph()
        g.add_edges_from(((e[0], e[1]), (e[2], e[3])) for e in edges)
        expected = 0
        for c in nx.connected_components(g):
            sub = g.subgraph(c)
            expected += len(nx.bipartite.maximum_matching(sub)) // 2   # König: |cover| = |matching|
        assert len(cover) == expected
        assert len(comps) == nx.number_connected_components(g)


def test_broker_manual_star_and_subset_of_robots():
    # reference test_broker.py:213-265: two stars -> two components, two vertices
    edges = [EdgeInterRobot(0, 1, 1, i, 1.0) for i in range(2, 6)] + \
            [EdgeInterRobot(1, 20, 0, i, 1.0) for i in range(10, 14)]
    comps = Broker(edges, [0, 1]).brokerage(True)
    assert len(comps) == 2 and sorted(len(c) for c in comps) == [1, 1]
    assert {v for c in comps for v in c} == {(0, 1), (1, 20)}
    # robot 2 not involved: its edges are ignored; a single robot left -> nothing to broker
    edges = [EdgeInterRobot(0, 1, 2, 3, 1.0), EdgeInterRobot(0, 2, 2, 5, 1.0)]
    assert Broker(edges, [0, 1]).brokerage(True) == []
    assert Broker([], [0, 1]).brokerage(True) == []


# ---- misc / neighbours / bus ----------------------------------------------------------------
def test_chunk_helpers():
    d = {k: f"v{k}" for k in range(3, 14)}
    assert dict_to_list_chunks(d, 5, 4) == [["v5", "v6", "v7", "v8"], ["v9", "v10", "v11", "v12"], ["v13"]]
    assert dict_to_list_chunks(d, 100, 4) == []
    assert list_chunks(list(range(7)), 2, 3) == [[2, 3, 4], [5, 6]]
    assert list_range([1, 2, 3, 4], 1) == [2, 3]


def _params(robot_id, n=3, monitoring=True):
    return {'robot_id': robot_id, 'max_nb_robots': n,
            'neighbor_management.enable_neighbor_monitoring': monitoring,
            'neighbor_management.init_delay_sec': 0.0,
            'neighbor_management.max_heartbeat_delay_sec': 5.0}


def test_neighbor_manager_liveness_broker_and_send_windows():
    bus = LocalBus()
    nm = NeighborManager(LocalNode(bus, "/r1"), _params(1))
    flags, in_range = nm.check_neighbors_in_range()
    assert in_range == [1] and nm.local_robot_is_broker()
    bus.publish('/r2/cslam/heartbeat', UInt32(data=2))
    flags, in_range = nm.check_neighbors_in_range()
    assert in_range == [1, 2] and flags == {0: False, 1: True, 2: True}
    assert nm.local_robot_is_broker()                      # 1 < 2
    bus.publish('/r0/cslam/heartbeat', UInt32(data=0))
    assert not nm.local_robot_is_broker()                  # robot 0 is alive and lower
    # first broadcast starts from keyframe 0, the next one from where it stopped
    assert nm.select_from_which_kf_to_send(9) == 0
    assert nm.select_from_which_kf_to_send(15) == 10
    assert nm.useless_descriptors(15) == 15
    assert nm.select_from_which_match_to_send(4) == 0 and nm.useless_matches(4) == 4

    class D:  # descriptor stub
        def __init__(self, r, k):
            self.robot_id, self.keyframe_id = r, k
    assert nm.get_unknown_range([D(2, 0), D(2, 1), D(2, 2)]) == [0, 1, 2]
    assert nm.get_unknown_range([D(2, 1), D(2, 2), D(2, 3)]) == [2]
    # monitoring disabled: the reference's is_alive() returns None -> nobody is in range
    nm2 = NeighborManager(LocalNode(LocalBus(), "/r0"), _params(0, monitoring=False))
    assert nm2.check_neighbors_in_range()[1] == [0]


def test_local_node_topics_timers_parameters():
    bus = LocalBus()
    a, b = LocalNode(bus, "/r0", {"x": 3}), LocalNode(bus, "/r1")
    got = []
    b.create_subscription(None, "/cslam/global", got.append)
    b.create_subscription(None, "cslam/private", got.append)
    a.create_publisher(None, "/cslam/global").publish("g")
    a.create_publisher(None, "cslam/private").publish("not for r1")   # resolves to /r0/...
    assert got == ["g"] and a.get_parameter("x").value == 3
    fired = []
    a.create_timer(1000.0, lambda: fired.append(1))
    a.spin_once()
    assert fired == []
    a.spin_once(force=True)
    assert fired == [1]
