"""AlgebraicConnectivityMaximization — inter-robot loop-closure candidate selection.

Same class, method and attribute names as the reference
(cslam/algebraic_connectivity_maximization.py:9-572); the graph bookkeeping stays on the
host (dicts of `EdgeInterRobot`, as the reference tests read them), the numerical work
(Laplacian, Fiedler pair, gradient, Frank-Wolfe) is delegated to `cslam_b200.mac.mac.MAC`,
i.e. to the GPU.

Quirks of the reference that are reproduced on purpose:
  * `add_match` looks the un-normalised key up in a dict keyed by the normalised key, so a
    match with robot0_id > robot1_id always overwrites the stored candidate (:559-572);
  * `total_nb_poses` sums the poses of ALL robots, included or not (:501-503), and
    `fill_odometry` chains every robot's poses on its (possibly zero) offset (:348-362);
    with excluded robots the Laplacian is singular, the solver raises and, after
    `nb_candidates_to_choose` re-initialisations, the greedy guess is returned (:448-466).
Differences: `remove_candidate_edges` is O(candidates + edges) instead of the reference's
O(candidates * edges) list scan (:178-190) — same result.
"""
from typing import NamedTuple

import numpy as np

from ._lib import CslamError
from .mac.mac import MAC
from .mac.utils import Edge


class EdgeInterRobot(NamedTuple):
    """ Inter-robot loop closure edge (reference :9-31)."""
    robot0_id: int
    robot0_keyframe_id: int
    robot1_id: int
    robot1_keyframe_id: int
    weight: float

    def __eq__(self, other):
        """Equality ignores the weight and the direction of the edge."""
        a = (self.robot0_id, self.robot0_keyframe_id)
        b = (self.robot1_id, self.robot1_keyframe_id)
        c = (other.robot0_id, other.robot0_keyframe_id)
        d = (other.robot1_id, other.robot1_keyframe_id)
        return (a == c and b == d) or (a == d and b == c)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__


class AlgebraicConnectivityMaximization(object):

    def __init__(self,
                 robot_id=0,
                 max_nb_robots=1,
                 max_iters=20,
                 fixed_weight=1.0,
                 extra_params={
                     "frontend.enable_sparsification": True,
                     "evaluation.enable_sparsification_comparison": False,
                 }):
        """
        Args:
            robot_id (int, optional): ID of the robot
            max_nb_robots (int, optional): number of robots. Defaults to 1.
            max_iters (int, optional): Frank-Wolfe iterations. Defaults to 20.
            fixed_weight (float, optional): weight of fixed measurements. Defaults to 1.0.
        """
        self.fixed_weight = fixed_weight
        self.params = extra_params

        self.fixed_edges = []
        self.candidate_edges = {}
        self.already_considered_matches = set()

        self.max_iters = max_iters
        self.max_nb_robots = max_nb_robots
        self.robot_id = robot_id
        self.total_nb_poses = 0

        self.nb_poses = {i: 0 for i in range(max_nb_robots)}
        self.initial_fixed_edge_exists = {i: False for i in range(max_nb_robots)}

        self.log_greedy_edges = []
        self.log_mac_edges = []
        self.last_mac = None          # MAC instance of the last run (stats / traces)
        self.last_mac_trials = 0      # re-initialisations needed by the last run

    # ---- keys / small helpers ------------------------------------------------
    def edge_key(self, edge):
        """Direction-independent key, lower robot id first (reference :76-90)."""
        if edge.robot0_id < edge.robot1_id:
            return (edge.robot0_id, edge.robot0_keyframe_id, edge.robot1_id,
                    edge.robot1_keyframe_id)
        return (edge.robot1_id, edge.robot1_keyframe_id, edge.robot0_id,
                edge.robot0_keyframe_id)

    def replace_weight(self, edge, weight):
        """Copy of `edge` with another weight (reference :92-108)."""
        if type(edge) is EdgeInterRobot:
            return EdgeInterRobot(edge.robot0_id, edge.robot0_keyframe_id, edge.robot1_id,
                                  edge.robot1_keyframe_id, weight)
        elif type(edge) is Edge:
            return Edge(edge.i, edge.j, weight)

    def update_nb_poses(self, edge):
        """nb_poses[r] = 1 + largest keyframe id seen for robot r (reference :110-119)."""
        for r, kf in ((edge.robot0_id, edge.robot0_keyframe_id),
                      (edge.robot1_id, edge.robot1_keyframe_id)):
            if kf + 1 > self.nb_poses[r]:
                self.nb_poses[r] = kf + 1

    def update_initial_fixed_edge_exists(self, fixed_edge):
        """Remember which robots already share a fixed inter-robot edge (reference :121-130)."""
        if fixed_edge.robot0_id != fixed_edge.robot1_id:
            self.initial_fixed_edge_exists[fixed_edge.robot0_id] = True
            self.initial_fixed_edge_exists[fixed_edge.robot1_id] = True

    # ---- graph editing ---------------------------------------------------------
    def set_graph(self, fixed_edges, candidate_edges):
        """Fill the graph (reference :132-152)."""
        self.fixed_edges = fixed_edges
        for e in self.fixed_edges:
            self.update_nb_poses(e)
            self.update_initial_fixed_edge_exists(e)
        for e in candidate_edges:
            self.update_nb_poses(e)
        for e in candidate_edges:
            self.candidate_edges[self.edge_key(e)] = e

    def add_fixed_edge(self, edge):
        """Add an already computed edge (reference :154-163)."""
        self.fixed_edges.append(edge)
        self.update_nb_poses(edge)
        self.update_initial_fixed_edge_exists(edge)

    def add_candidate_edge(self, edge):
        """Add a candidate unless it was already tried or fixed (reference :165-178)."""
        key = self.edge_key(edge)
        if key in self.already_considered_matches:
            return
        self.candidate_edges[key] = edge
        self.update_nb_poses(edge)

    def remove_candidate_edges(self, edges, failed=False):
        """Drop candidates and blacklist them (reference :178-190)."""
        for edge in edges:
            key = self.edge_key(edge)
            self.candidate_edges.pop(key, None)
            self.already_considered_matches.add(key)

    def candidate_edges_to_fixed(self, edges):
        """Candidates that became measurements: fixed weight, fixed list (reference :192-203)."""
        for i in range(len(edges)):
            edges[i] = self.replace_weight(edges[i], weight=self.fixed_weight)
            self.update_initial_fixed_edge_exists(edges[i])
        self.fixed_edges.extend(edges)
        self.remove_candidate_edges(edges)

    # ---- initial guesses -------------------------------------------------------
    def greedy_initialization(self, nb_candidates_to_choose, edges):
        """1.0 on the `nb_candidates_to_choose` largest weights (reference :205-218)."""
        weights = [e.weight for e in edges]
        w_init = np.zeros(len(weights))
        indices = np.argpartition(weights, -nb_candidates_to_choose)[-nb_candidates_to_choose:]
        w_init[indices] = 1.0
        return w_init

    def pseudo_greedy_initialization(self, nb_candidates_to_choose, nb_random, edges):
        """Greedy for all but `nb_random` picks, those at random (reference :220-245)."""
        w_init = self.greedy_initialization(nb_candidates_to_choose - nb_random, edges)
        nb_edges = len(edges)
        picked, trial, max_trials = 0, 0, 2 * nb_random
        while picked < nb_random and trial < max_trials:
            j = int(np.random.rand() * nb_edges)
            if w_init[j] < 0.5:
                w_init[j] = 1.0
                picked += 1
            trial += 1
        if trial >= max_trials:
            w_init = self.greedy_initialization(nb_candidates_to_choose, edges)
        return w_init

    def random_initialization(self, nb_candidates_to_choose, edges):
        """Random weights, then greedy (reference :247-255; mutates `edges`)."""
        for e in range(len(edges)):
            edges[e] = self.replace_weight(edges[e], np.random.rand())
        return self.greedy_initialization(nb_candidates_to_choose, edges)

    def connection_biased_greedy_selection(self, nb_candidates_to_choose, edges,
                                           is_robot_included):
        """Greedy selection that first links robots without a fixed edge (reference :257-289)."""
        edges_copy = edges.copy()
        forced = []
        for rid in [r for r in is_robot_included.keys() if is_robot_included[r]]:
            if self.initial_fixed_edge_exists[rid]:
                continue
            best, best_w = None, -1
            for i, e in enumerate(edges_copy):
                if (e.robot0_id == rid or e.robot1_id == rid) and e.weight > best_w:
                    best, best_w = i, e.weight
            if best is not None:
                forced.append(best)
                edges_copy[best] = self.replace_weight(edges_copy[best], weight=0.0)
        w_init = np.zeros(len(edges))
        if nb_candidates_to_choose - len(forced) > 0:
            w_init = self.greedy_initialization(nb_candidates_to_choose - len(forced),
                                                self.rekey_edges(edges_copy, is_robot_included))
        for i in forced:
            w_init[i] = 1.0
        return w_init

    # ---- rekeying ----------------------------------------------------------------
    def compute_offsets(self, is_robot_included):
        """Node-id offset of every included robot: running sum of nb_poses (reference :291-310)."""
        self.offsets = {i: 0 for i in range(self.max_nb_robots)}
        running = 0
        for rid in range(self.max_nb_robots):
            if is_robot_included[rid]:
                self.offsets[rid] = running
                running += self.nb_poses[rid]

    def rekey_edges(self, edges, is_robot_included):
        """(robot, keyframe) pairs -> global node ids; edges touching an excluded robot are
        dropped (reference :312-335)."""
        out = []
        for e in edges:
            if is_robot_included[e.robot0_id] and is_robot_included[e.robot1_id]:
                out.append(Edge(self.offsets[e.robot0_id] + e.robot0_keyframe_id,
                                self.offsets[e.robot1_id] + e.robot1_keyframe_id, e.weight))
        return out

    def get_included_edges(self, edges, is_robot_included):
        """Edges whose two robots are included (reference :337-346)."""
        return [e for e in edges
                if is_robot_included[e.robot0_id] and is_robot_included[e.robot1_id]]

    def fill_odometry(self):
        """Implicit odometry chains, weight `fixed_weight` (reference :348-362)."""
        odom = []
        for i in range(len(self.nb_poses)):
            base = self.offsets[i]
            for k in range(self.nb_poses[i] - 1):
                odom.append(Edge(base + k, base + k + 1, self.fixed_weight))
        return odom

    def recover_inter_robot_edges(self, edges, is_robot_included):
        """Inverse of `rekey_edges` (reference :364-389)."""
        out = []
        for e in edges:
            r0 = r1 = 0
            for o in self.offsets:
                if o != 0 and is_robot_included[o]:
                    if e.i >= self.offsets[o]:
                        r0 = o
                    if e.j >= self.offsets[o]:
                        r1 = o
            out.append(EdgeInterRobot(r0, e.i - self.offsets[r0], r1, e.j - self.offsets[r1],
                                      e.weight))
        return out

    # ---- inclusion logic --------------------------------------------------------------
    def check_graph_disconnections(self, is_other_robot_considered):
        """A robot is included if it is the local one or appears in any fixed/candidate edge
        while being in range (reference :391-417)."""
        connected = {i: (i == self.robot_id) for i in range(self.max_nb_robots)}
        for edge in list(self.fixed_edges) + list(self.candidate_edges.values()):
            if is_other_robot_considered[edge.robot0_id]:
                connected[edge.robot0_id] = True
            if is_other_robot_considered[edge.robot1_id]:
                connected[edge.robot1_id] = True
        return connected

    def check_initial_fixed_measurements_exists(self, is_robot_included):
        """True when every included robot already has a fixed inter-robot edge (reference :419-434)."""
        return all(self.initial_fixed_edge_exists[rid]
                   for rid in is_robot_included if is_robot_included[rid])

    # ---- solver ---------------------------------------------------------------------------
    def run_mac_solver(self, fixed_edges, candidate_edges, w_init, nb_candidates_to_choose):
        """Frank-Wolfe on the GPU with the reference's retry policy (reference :436-466): if the
        Laplacian is singular (graph disconnected) the initial guess is re-drawn with
        increasing randomness, at most `nb_candidates_to_choose` times, else the initial
        guess is returned."""
        mac = MAC(fixed_edges, candidate_edges, self.total_nb_poses)
        self.last_mac = mac
        result = w_init.copy()
        trial = 0
        while trial < nb_candidates_to_choose:
            try:
                result, _, _ = mac.fw_subset(w_init, nb_candidates_to_choose,
                                             max_iters=self.max_iters)
                break
            except CslamError:
                # reference: bare `except` around a SuperLU "Factor is exactly singular"
                trial += 1
                w_init = self.pseudo_greedy_initialization(nb_candidates_to_choose, trial,
                                                           candidate_edges)
        self.last_mac_trials = trial
        return result

    def select_candidates(self, nb_candidates_to_choose, is_other_robot_considered,
                          greedy_initialization=True):
        """Solve algebraic connectivity maximisation (reference :468-543).

        Args:
            nb_candidates_to_choose (int): budget
            is_other_robot_considered: dict(int, bool): robots in communication range
            greedy_initialization: initialise from the similarity weights

        Returns:
            list(EdgeInterRobot): selected edges
        """
        is_robot_included = self.check_graph_disconnections(is_other_robot_considered)

        self.compute_offsets(is_robot_included)
        rekeyed_fixed_edges = self.rekey_edges(self.fixed_edges, is_robot_included)
        rekeyed_fixed_edges.extend(self.fill_odometry())
        rekeyed_candidate_edges = self.rekey_edges(self.candidate_edges.values(),
                                                   is_robot_included)

        if nb_candidates_to_choose > len(rekeyed_candidate_edges):
            nb_candidates_to_choose = len(rekeyed_candidate_edges)
        if len(rekeyed_candidate_edges) == 0:
            return []

        self.total_nb_poses = sum(self.nb_poses[n] for n in range(len(self.nb_poses)))

        if greedy_initialization:
            w_init = self.greedy_initialization(nb_candidates_to_choose, rekeyed_candidate_edges)
        else:
            w_init = self.random_initialization(nb_candidates_to_choose, rekeyed_candidate_edges)

        if self.params["frontend.enable_sparsification"] and \
                self.check_initial_fixed_measurements_exists(is_robot_included):
            result = self.run_mac_solver(rekeyed_fixed_edges, rekeyed_candidate_edges, w_init,
                                         nb_candidates_to_choose)
        else:
            result = self.connection_biased_greedy_selection(
                nb_candidates_to_choose,
                self.get_included_edges(self.candidate_edges.values(), is_robot_included),
                is_robot_included)

        if self.params["evaluation.enable_sparsification_comparison"]:
            self.sparsification_comparison_logs(rekeyed_candidate_edges, is_robot_included,
                                                w_init, result)

        selected_edges = [rekeyed_candidate_edges[i] for i in np.nonzero(result.astype(int))[0]]
        inter_robot_edges = self.recover_inter_robot_edges(selected_edges, is_robot_included)
        self.remove_candidate_edges(inter_robot_edges)
        return inter_robot_edges

    def sparsification_comparison_logs(self, rekeyed_candidate_edges, is_robot_included,
                                       greedy_result, mac_result):
        """Keep the greedy and the MAC selections for evaluation logs (reference :545-557)."""
        self.log_greedy_edges = self.recover_inter_robot_edges(
            [rekeyed_candidate_edges[i] for i in np.nonzero(greedy_result.astype(int))[0]],
            is_robot_included)
        self.log_mac_edges = self.recover_inter_robot_edges(
            [rekeyed_candidate_edges[i] for i in np.nonzero(mac_result.astype(int))[0]],
            is_robot_included)

    def add_match(self, match):
        """Add a potential match, keeping the better weight (reference :559-572, including its
        un-normalised key lookup)."""
        key = (match.robot0_id, match.robot0_keyframe_id, match.robot1_id,
               match.robot1_keyframe_id)
        if key in self.candidate_edges:
            if match.weight > self.candidate_edges[key].weight:
                self.add_candidate_edge(match)
        else:
            self.add_candidate_edge(match)
