"""AlgebraicConnectivityMaximization — inter-robot loop-closure candidate selection.

Class, method and attribute names are those of the reference
(cslam/algebraic_connectivity_maximization.py:9-572) because its callers and tests address
them (`candidate_edges` is a dict keyed by a 4-tuple, `fixed_edges` a list, `nb_poses`,
`offsets`, ...).  The graph bookkeeping is host code; all numerical work (Laplacian,
Fiedler pair, gradient, Frank-Wolfe) runs on the GPU behind `cslam_b200.mac.mac.MAC`.

Behaviour of the reference that is reproduced on purpose:
  * `add_match` probes the candidate dict with the key exactly as the match spells it, while
    the dict is keyed lower-robot-first: a match written with robot0_id > robot1_id never
    finds the stored entry and always overwrites it (:559-572);
  * `total_nb_poses` counts the poses of ALL robots, included or not (:501-503), and
    `fill_odometry` chains every robot from its (possibly zero) offset (:348-362); with an
    excluded robot the Laplacian is singular, the solver raises, and after
    `nb_candidates_to_choose` re-initialisations the greedy guess is returned (:448-466).
Difference: removing candidates is a dict pop per edge, not the reference's scan of the whole
candidate list per edge (:178-190) — same result.
"""
import bisect
from typing import NamedTuple

import numpy as np

from ._lib import CslamError
from .mac.mac import MAC
from .mac.utils import Edge


class EdgeInterRobot(NamedTuple):
    """Loop closure between keyframes of two robots (reference :9-31)."""
    robot0_id: int
    robot0_keyframe_id: int
    robot1_id: int
    robot1_keyframe_id: int
    weight: float

    def _ends(self):
        return frozenset(((self.robot0_id, self.robot0_keyframe_id),
                          (self.robot1_id, self.robot1_keyframe_id)))

    def __eq__(self, other):
        # same two vertices, in either direction; the weight does not take part
        return self._ends() == EdgeInterRobot._ends(other)

    def __ne__(self, other):
        return not self == other

    __hash__ = tuple.__hash__


def _largest(values, count):
    """0/1 vector marking the `count` largest entries (np.argpartition, like the reference)."""
    mask = np.zeros(len(values))
    mask[np.argpartition(values, -count)[-count:]] = 1.0
    return mask


class AlgebraicConnectivityMaximization(object):

    def __init__(self, robot_id=0, max_nb_robots=1, max_iters=20, fixed_weight=1.0,
                 extra_params={"frontend.enable_sparsification": True,
                               "evaluation.enable_sparsification_comparison": False}):
        """
        Args:
            robot_id (int): id of the local robot
            max_nb_robots (int): number of robots in the swarm
            max_iters (int): Frank-Wolfe iterations of the solver
            fixed_weight (float): weight given to measured (fixed) edges
            extra_params (dict): the node's parameter dict
        """
        self.robot_id = robot_id
        self.max_nb_robots = max_nb_robots
        self.max_iters = max_iters
        self.fixed_weight = fixed_weight
        self.params = extra_params
        robots = range(max_nb_robots)
        self.nb_poses = dict.fromkeys(robots, 0)
        self.initial_fixed_edge_exists = dict.fromkeys(robots, False)
        self.fixed_edges = []
        self.candidate_edges = {}
        self.already_considered_matches = set()
        self.total_nb_poses = 0
        self.log_greedy_edges = []
        self.log_mac_edges = []
        self.last_mac = None        # solver object of the last run (stats / traces)
        self.last_mac_trials = 0    # re-initialisations the last run needed

    # ------------------------------------------------------------------ small helpers
    def edge_key(self, edge):
        """Direction-free dictionary key: the lower robot id comes first (reference :76-90)."""
        a = (edge.robot0_id, edge.robot0_keyframe_id)
        b = (edge.robot1_id, edge.robot1_keyframe_id)
        return a + b if edge.robot0_id < edge.robot1_id else b + a

    def replace_weight(self, edge, weight):
        """Same edge, other weight (reference :92-108); None for foreign types, as there."""
        if type(edge) in (EdgeInterRobot, Edge):
            return edge._replace(weight=weight)
        return None

    def update_nb_poses(self, edge):
        """A robot has at least (largest keyframe id seen) + 1 poses (reference :110-119)."""
        for robot, keyframe in ((edge.robot0_id, edge.robot0_keyframe_id),
                                (edge.robot1_id, edge.robot1_keyframe_id)):
            self.nb_poses[robot] = max(self.nb_poses[robot], keyframe + 1)

    def update_initial_fixed_edge_exists(self, fixed_edge):
        """Robots joined by a measured inter-robot edge (reference :121-130)."""
        if fixed_edge.robot0_id == fixed_edge.robot1_id:
            return
        self.initial_fixed_edge_exists[fixed_edge.robot0_id] = True
        self.initial_fixed_edge_exists[fixed_edge.robot1_id] = True

    # ------------------------------------------------------------------ graph editing
    def set_graph(self, fixed_edges, candidate_edges):
        """Install a whole graph (reference :132-152)."""
        self.fixed_edges = fixed_edges
        for edge in fixed_edges:
            self.update_nb_poses(edge)
            self.update_initial_fixed_edge_exists(edge)
        for edge in candidate_edges:
            self.update_nb_poses(edge)
            self.candidate_edges[self.edge_key(edge)] = edge

    def add_fixed_edge(self, edge):
        """A measured edge (reference :154-163)."""
        self.fixed_edges.append(edge)
        self.update_nb_poses(edge)
        self.update_initial_fixed_edge_exists(edge)

    def add_candidate_edge(self, edge):
        """A candidate, unless that pair was already selected or measured (reference :165-178)."""
        key = self.edge_key(edge)
        if key not in self.already_considered_matches:
            self.candidate_edges[key] = edge
            self.update_nb_poses(edge)

    def remove_candidate_edges(self, edges, failed=False):
        """Forget candidates for good (reference :178-190)."""
        for key in map(self.edge_key, edges):
            self.candidate_edges.pop(key, None)
            self.already_considered_matches.add(key)

    def candidate_edges_to_fixed(self, edges):
        """Verified candidates become measurements with the fixed weight (reference :192-203;
        like there, the caller's list is rewritten in place)."""
        edges[:] = [self.replace_weight(e, weight=self.fixed_weight) for e in edges]
        for edge in edges:
            self.update_initial_fixed_edge_exists(edge)
        self.fixed_edges.extend(edges)
        self.remove_candidate_edges(edges)

    def add_match(self, match):
        """Candidate from a descriptor match; a stored candidate is only replaced by a heavier
        one — when the lookup finds it (see the module docstring; reference :559-572)."""
        stored = self.candidate_edges.get((match.robot0_id, match.robot0_keyframe_id,
                                           match.robot1_id, match.robot1_keyframe_id))
        if stored is None or match.weight > stored.weight:
            self.add_candidate_edge(match)

    # ------------------------------------------------------------------ initial guesses
    def greedy_initialization(self, nb_candidates_to_choose, edges):
        """The heaviest candidates (reference :205-218)."""
        return _largest([e.weight for e in edges], nb_candidates_to_choose)

    def pseudo_greedy_initialization(self, nb_candidates_to_choose, nb_random, edges):
        """Greedy except for `nb_random` picks drawn with np.random.rand (at most 2*nb_random
        draws, else plain greedy) (reference :220-245)."""
        w_init = self.greedy_initialization(nb_candidates_to_choose - nb_random, edges)
        draws_left, missing = 2 * nb_random, nb_random
        while missing > 0 and draws_left > 0:
            j = int(np.random.rand() * len(edges))
            draws_left -= 1
            if w_init[j] < 0.5:
                w_init[j] = 1.0
                missing -= 1
        if draws_left <= 0:
            return self.greedy_initialization(nb_candidates_to_choose, edges)
        return w_init

    def random_initialization(self, nb_candidates_to_choose, edges):
        """Greedy over weights redrawn from U(0,1); rewrites `edges` like the reference (:247-255)."""
        edges[:] = [self.replace_weight(e, np.random.rand()) for e in edges]
        return self.greedy_initialization(nb_candidates_to_choose, edges)

    def connection_biased_greedy_selection(self, nb_candidates_to_choose, edges, is_robot_included):
        """Greedy selection that first gives every included robot without a measured
        inter-robot edge its heaviest candidate (reference :257-289)."""
        pool = list(edges)
        forced = []
        for robot in (r for r, inc in is_robot_included.items() if inc):
            if self.initial_fixed_edge_exists[robot]:
                continue
            touching = [(e.weight, -i) for i, e in enumerate(pool)
                        if robot in (e.robot0_id, e.robot1_id) and e.weight > -1]
            if touching:
                best = -max(touching)[1]          # heaviest, first one among equals
                forced.append(best)
                pool[best] = self.replace_weight(pool[best], weight=0.0)
        w_init = np.zeros(len(edges))
        free = nb_candidates_to_choose - len(forced)
        if free > 0:
            w_init = self.greedy_initialization(free, self.rekey_edges(pool, is_robot_included))
        w_init[forced] = 1.0
        return w_init

    # ------------------------------------------------------------------ (robot, keyframe) <-> node id
    def compute_offsets(self, is_robot_included):
        """First node id of every included robot: running sum of nb_poses; excluded robots
        keep offset 0 (reference :291-310)."""
        self.offsets = dict.fromkeys(range(self.max_nb_robots), 0)
        start = 0
        for robot in range(self.max_nb_robots):
            if is_robot_included[robot]:
                self.offsets[robot] = start
                start += self.nb_poses[robot]

    def get_included_edges(self, edges, is_robot_included):
        """Edges between two included robots (reference :337-346)."""
        return [e for e in edges
                if is_robot_included[e.robot0_id] and is_robot_included[e.robot1_id]]

    def rekey_edges(self, edges, is_robot_included):
        """Inter-robot edges -> `Edge(i, j, weight)` on global node ids (reference :312-335)."""
        off = self.offsets
        return [Edge(off[e.robot0_id] + e.robot0_keyframe_id,
                     off[e.robot1_id] + e.robot1_keyframe_id, e.weight)
                for e in self.get_included_edges(edges, is_robot_included)]

    def fill_odometry(self):
        """Consecutive poses of each robot, weight `fixed_weight` (reference :348-362)."""
        chains = []
        for robot in range(len(self.nb_poses)):
            first = self.offsets[robot]
            chains.extend(Edge(i, i + 1, self.fixed_weight)
                          for i in range(first, first + self.nb_poses[robot] - 1))
        return chains

    def recover_inter_robot_edges(self, edges, is_robot_included):
        """Node ids back to (robot, keyframe): the owner of a node is the last included robot
        (other than robot 0, the default) whose offset does not exceed it (reference :364-389)."""
        owners = [r for r in self.offsets if r != 0 and is_robot_included[r]]
        starts = [self.offsets[r] for r in owners]

        def owner(node):
            pos = bisect.bisect_right(starts, node)
            return owners[pos - 1] if pos > 0 else 0

        out = []
        for e in edges:
            r0, r1 = owner(e.i), owner(e.j)
            out.append(EdgeInterRobot(r0, e.i - self.offsets[r0], r1, e.j - self.offsets[r1], e.weight))
        return out

    # ------------------------------------------------------------------ which robots take part
    def check_graph_disconnections(self, is_other_robot_considered):
        """The local robot, plus every robot in range that appears in some edge (reference :391-417)."""
        seen = {self.robot_id}
        for edge in list(self.fixed_edges) + list(self.candidate_edges.values()):
            seen.update(r for r in (edge.robot0_id, edge.robot1_id) if is_other_robot_considered[r])
        return {r: r in seen for r in range(self.max_nb_robots)}

    def check_initial_fixed_measurements_exists(self, is_robot_included):
        """Does every included robot already share a measured edge with another one? (reference :419-434)"""
        return all(self.initial_fixed_edge_exists[r] for r, inc in is_robot_included.items() if inc)

    # ------------------------------------------------------------------ solver
    def run_mac_solver(self, fixed_edges, candidate_edges, w_init, nb_candidates_to_choose):
        """GPU Frank-Wolfe with the reference's retry policy (:436-466): a singular Laplacian
        (disconnected graph) re-draws the start with one more random pick each time, at most
        `nb_candidates_to_choose` times, then the start vector itself is the answer."""
        mac = self.last_mac = MAC(fixed_edges, candidate_edges, self.total_nb_poses)
        answer = w_init.copy()
        for trial in range(nb_candidates_to_choose):
            try:
                answer = mac.fw_subset(w_init, nb_candidates_to_choose, max_iters=self.max_iters)[0]
                self.last_mac_trials = trial
                return answer
            except CslamError:    # the reference swallows SuperLU's "Factor is exactly singular"
                w_init = self.pseudo_greedy_initialization(nb_candidates_to_choose, trial + 1,
                                                           candidate_edges)
        self.last_mac_trials = nb_candidates_to_choose
        return answer

    def select_candidates(self, nb_candidates_to_choose, is_other_robot_considered,
                          greedy_initialization=True):
        """Choose the candidates that maximise the algebraic connectivity (reference :468-543).

        Args:
            nb_candidates_to_choose (int): budget
            is_other_robot_considered (dict(int, bool)): robots in communication range
            greedy_initialization (bool): start from the heaviest candidates (else random)
        Returns:
            list(EdgeInterRobot): the selection; it is removed from the candidates
        """
        included = self.check_graph_disconnections(is_other_robot_considered)
        self.compute_offsets(included)
        fixed = self.rekey_edges(self.fixed_edges, included) + self.fill_odometry()
        candidates = self.rekey_edges(self.candidate_edges.values(), included)
        if not candidates:
            return []
        budget = min(nb_candidates_to_choose, len(candidates))
        self.total_nb_poses = sum(self.nb_poses.values())

        start = (self.greedy_initialization if greedy_initialization
                 else self.random_initialization)(budget, candidates)
        if self.params["frontend.enable_sparsification"] and \
                self.check_initial_fixed_measurements_exists(included):
            chosen = self.run_mac_solver(fixed, candidates, start, budget)
        else:
            chosen = self.connection_biased_greedy_selection(
                budget, self.get_included_edges(self.candidate_edges.values(), included), included)

        if self.params["evaluation.enable_sparsification_comparison"]:
            self.sparsification_comparison_logs(candidates, included, start, chosen)
        picked = [candidates[i] for i in np.flatnonzero(np.asarray(chosen).astype(int))]
        selection = self.recover_inter_robot_edges(picked, included)
        self.remove_candidate_edges(selection)
        return selection

    def sparsification_comparison_logs(self, rekeyed_candidate_edges, is_robot_included,
                                       greedy_result, mac_result):
        """Keep both selections for the evaluation topics (reference :545-557)."""
        def as_edges(mask):
            return self.recover_inter_robot_edges(
                [rekeyed_candidate_edges[i] for i in np.flatnonzero(np.asarray(mask).astype(int))],
                is_robot_included)
        self.log_greedy_edges = as_edges(greedy_result)
        self.log_mac_edges = as_edges(mac_result)
