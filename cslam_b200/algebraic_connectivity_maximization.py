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

Scale: `candidate_edges` is a `CandidateTable` — the reference's dict
plus numpy columns kept in step with it — and `select_candidates` sets the problem up from
those columns (inclusion, offsets, rekeying, greedy start, recovery of the picked edges) without
walking the candidates in Python; `add_matches` is the bulk form of `add_match`.  The
edge-by-edge methods of the reference API remain and are used whenever `candidate_edges` was
replaced by a plain dict.
"""
import bisect
import logging
from typing import NamedTuple

import numpy as np

from ._lib import CslamError, SingularLaplacianError
from .candidate_table import CandidateTable, keys_packable, pack_keys
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

    def __hash__(self):
        # hashes what __eq__ compares (the reference defines __eq__ only, which makes its class
        # unhashable; here edges can live in sets/dicts without duplicating equal edges)
        return hash(self._ends())


def _largest(values, count):
    """0/1 vector marking the `count` largest entries (np.argpartition, like the reference)."""
    mask = np.zeros(len(values))
    mask[np.argpartition(values, -count)[-count:]] = 1.0
    return mask


def _is_columns(edges):
    """True for the (i, j, weight) array triple that stands in for a list of `Edge`."""
    return isinstance(edges, tuple) and len(edges) == 3 and isinstance(edges[0], np.ndarray)


def _weights_of(edges):
    return edges[2] if _is_columns(edges) else [e.weight for e in edges]


def _count(edges):
    return len(edges[2]) if _is_columns(edges) else len(edges)


def _picked(edges, mask):
    """The edges whose entry of the 0/1 vector `mask` is set, as a list of `Edge`."""
    idx = np.flatnonzero(np.asarray(mask).astype(int))
    if _is_columns(edges):
        i, j, w = edges
        return [Edge(int(i[t]), int(j[t]), float(w[t])) for t in idx]
    return [edges[t] for t in idx]


_KF_BITS = 40   # keyframe ids below 2^40, robot ids below 2^23: one int64 per (robot, keyframe)


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
        self.candidate_edges = CandidateTable(EdgeInterRobot)
        self.already_considered_matches = set()
        self.total_nb_poses = 0
        self.log_greedy_edges = []
        self.log_mac_edges = []
        self.last_mac = None        # solver object of the last run (stats / traces)
        self.last_mac_trials = 0    # re-initialisations the last run needed
        self.last_mac_error = None  # non-singular solver failure of the last run, if any

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
        keys = [self.edge_key(e) for e in edges]
        if isinstance(self.candidate_edges, CandidateTable):
            self.candidate_edges.remove_keys(keys)
        else:
            for key in keys:
                self.candidate_edges.pop(key, None)
        self.already_considered_matches.update(keys)

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

    def add_matches(self, robot0_id, robot0_keyframe_id, robot1_id, robot1_keyframe_id, weight):
        """`add_match` for a whole batch given as five equally long arrays: the candidate
        table ends up exactly as if the matches had been passed to `add_match` one after the
        other (including the reversed-key overwrite of :565-569 and the blacklist of
        :170-171), but the work per match is a few numpy passes; Python touches each DISTINCT
        vertex pair once."""
        r0 = np.asarray(robot0_id, dtype=np.int64).ravel()
        k0 = np.asarray(robot0_keyframe_id, dtype=np.int64).ravel()
        r1 = np.asarray(robot1_id, dtype=np.int64).ravel()
        k1 = np.asarray(robot1_keyframe_id, dtype=np.int64).ravel()
        w = np.asarray(weight, dtype=np.float64).ravel()
        n = len(w)
        if n == 0:
            return
        table = self.candidate_edges
        if (not isinstance(table, CandidateTable) or np.any(r0 == r1) or np.any(np.isnan(w))
                or max(k0.max(), k1.max()) >= (1 << _KF_BITS) or min(k0.min(), k1.min()) < 0):
            for t in range(n):
                self.add_match(EdgeInterRobot(int(r0[t]), int(k0[t]), int(r1[t]), int(k1[t]), float(w[t])))
            return
        direct = r0 < r1            # spelled the way the table is keyed: the lookup of :568 can hit
        lo = np.where(direct, (r0 << _KF_BITS) | k0, (r1 << _KF_BITS) | k1)
        hi = np.where(direct, (r1 << _KF_BITS) | k1, (r0 << _KF_BITS) | k0)
        order = np.lexsort((hi, lo))                      # stable: batch order inside a pair
        slo, shi = lo[order], hi[order]
        first = np.flatnonzero(np.r_[True, (slo[1:] != slo[:-1]) | (shi[1:] != shi[:-1])])
        group = np.cumsum(np.r_[True, (slo[1:] != slo[:-1]) | (shi[1:] != shi[:-1])]) - 1
        pos = np.arange(n)
        sw, srev = w[order], ~direct[order]
        # the last reversed spelling of a pair overwrites whatever was stored; after it only a
        # strictly heavier direct spelling replaces the entry: the survivor is the first maximum
        # among the positions from that overwrite on (from the start when there is none)
        last_rev = np.maximum.reduceat(np.where(srev, pos, -1), first)
        eligible = pos >= last_rev[group]
        best_w = np.maximum.reduceat(np.where(eligible, sw, -np.inf), first)
        winner = np.minimum.reduceat(np.where(eligible & (sw == best_w[group]), pos, n), first)
        mask = (1 << _KF_BITS) - 1
        glo, ghi = slo[first], shi[first]
        g_r0, g_k0, g_r1, g_k1 = glo >> _KF_BITS, glo & mask, ghi >> _KF_BITS, ghi & mask
        # every accepted match raises nb_poses (:110-119); a refused (lighter) one names the same two
        # vertices as the stored edge, so only blacklisted pairs must be left out (`blocked` below)
        def raise_nb_poses(blocked):
            live = ~blocked[group]
            if np.any(live):
                src = order[live]
                for robots, frames in ((r0[src], k0[src]), (r1[src], k1[src])):
                    top = np.zeros(self.max_nb_robots, dtype=np.int64)
                    np.maximum.at(top, robots, frames + 1)
                    for robot in np.flatnonzero(top).tolist():
                        self.nb_poses[robot] = max(self.nb_poses[robot], int(top[robot]))

        if not table.device_mode and self._device_index_wanted() and keys_packable(g_r0, g_k0, g_r1, g_k1):
            table.use_device_index(self._index_device)
        if table.device_mode and keys_packable(g_r0, g_k0, g_r1, g_k1):
            # ---- device index: packed keys, one lookup kernel, one insert kernel, no tuples
            key64 = pack_keys(g_r0, g_k0, g_r1, g_k1)
            considered = self._considered_packed()
            blocked = np.isin(key64, considered) if len(considered) else np.zeros(len(key64), dtype=bool)
            keep = ~blocked
            slots = table.lookup_packed(key64)
            if len(table):
                # a stored entry survives a batch of direct spellings that are not heavier
                cand = np.flatnonzero(keep & (last_rev < 0) & (slots >= 0))
                if len(cand):
                    keep[cand[~(best_w[cand] > table._w[slots[cand]])]] = False
            raise_nb_poses(blocked)
            chosen = np.flatnonzero(keep)
            if len(chosen) == 0:
                return
            chosen = chosen[np.argsort(order[first[chosen]], kind="stable")]   # new pairs in batch order
            src = order[winner[chosen]]
            table.put_rows_packed(key64[chosen], np.stack([r0[src], k0[src], r1[src], k1[src]], axis=1),
                                  w[src], slots=slots[chosen])
            return
        keys = list(zip(g_r0.tolist(), g_k0.tolist(), g_r1.tolist(), g_k1.tolist()))
        considered = self.already_considered_matches
        blocked = np.zeros(len(keys), dtype=bool)
        if considered:
            blocked = np.fromiter((key in considered for key in keys), dtype=bool, count=len(keys))
        keep = ~blocked
        if len(table):
            # a stored entry survives a batch of direct spellings that are not heavier
            stored = table.weight_of
            for g in np.flatnonzero(keep & (last_rev < 0)).tolist():
                sw0 = stored(keys[g])
                if sw0 is not None and not best_w[g] > sw0:
                    keep[g] = False
        raise_nb_poses(blocked)
        chosen = np.flatnonzero(keep)
        if len(chosen) == 0:
            return
        chosen = chosen[np.argsort(order[first[chosen]], kind="stable")]   # new pairs in batch order
        src = order[winner[chosen]]
        table.put_rows([keys[g] for g in chosen.tolist()],
                       np.stack([r0[src], k0[src], r1[src], k1[src]], axis=1), w[src])

    # ---- device hash index of the candidate keys (candidate_table.py, csrc/keymap.cu)
    def _device_index_wanted(self):
        """Bulk inserts move the key -> slot map to the GPU when one is present (the per-edge
        reference API works in either mode).  `frontend.candidate_index: host` keeps the dict."""
        want = getattr(self, "_index_wanted", None)
        if want is None:
            want = False
            if str(self.params.get("frontend.candidate_index", "auto")).lower() != "host":
                try:
                    from . import _lib
                    want = _lib.device_count() > 0
                except Exception:
                    want = False
            self._index_device = 0
            if want:
                try:
                    import torch
                    self._index_device = int(torch.cuda.current_device())
                except Exception:
                    pass
            self._index_wanted = want
        return want

    def _considered_packed(self):
        """`already_considered_matches` (a set of 4-tuples, like the reference's) as a sorted array
        of packed keys; rebuilt when the set has grown."""
        c = self.already_considered_matches
        cache = getattr(self, "_considered_cache", None)
        if cache is None or cache[0] != len(c):
            if c:
                arr = np.array(list(c), dtype=np.int64).reshape(-1, 4)
                ok = (arr[:, 0] >= 0) & (arr[:, 0] < 256) & (arr[:, 2] >= 0) & (arr[:, 2] < 256) & \
                     (arr[:, 1] >= 0) & (arr[:, 1] < (1 << 24)) & (arr[:, 3] >= 0) & (arr[:, 3] < (1 << 24))
                arr = arr[ok]
                packed = np.unique(pack_keys(arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]))
            else:
                packed = np.zeros(0, dtype=np.uint64)
            cache = self._considered_cache = (len(c), packed)
        return cache[1]

    # ------------------------------------------------------------------ initial guesses
    def greedy_initialization(self, nb_candidates_to_choose, edges):
        """The heaviest candidates (reference :205-218)."""
        return _largest(_weights_of(edges), nb_candidates_to_choose)

    def pseudo_greedy_initialization(self, nb_candidates_to_choose, nb_random, edges):
        """Greedy except for `nb_random` picks drawn with np.random.rand (at most 2*nb_random
        draws, else plain greedy) (reference :220-245)."""
        w_init = self.greedy_initialization(nb_candidates_to_choose - nb_random, edges)
        draws_left, missing = 2 * nb_random, nb_random
        while missing > 0 and draws_left > 0:
            j = int(np.random.rand() * _count(edges))
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

    def _candidate_columns(self, is_robot_included):
        """`rekey_edges(self.candidate_edges.values(), ...)` as an (i, j, weight) array triple
        straight from the candidate table's columns."""
        ends, w, alive = self.candidate_edges.raw()
        robots = range(self.max_nb_robots)
        inc = np.array([bool(is_robot_included[r]) for r in robots])
        off = np.array([self.offsets[r] for r in robots], dtype=np.int64)
        r0, r1 = ends[:, 0], ends[:, 2]
        # computed over every slot, live or not, and filtered once (a removed slot keeps its old,
        # valid, contents): one 16 MB gather instead of copying the 40 MB of live columns first
        i = (off[r0] + ends[:, 1]).astype(np.int32)
        j = (off[r1] + ends[:, 3]).astype(np.int32)
        keep = inc[r0] & inc[r1]
        if alive is not None:
            keep &= alive
        if not keep.all():
            return i[keep], j[keep], w[keep]
        return i, j, np.array(w, dtype=np.float64)

    def _fixed_columns(self, is_robot_included):
        """`rekey_edges(self.fixed_edges, ...) + fill_odometry()` as an array triple."""
        fixed = self.rekey_edges(self.fixed_edges, is_robot_included)
        first = np.array([self.offsets[r] for r in range(len(self.nb_poses))], dtype=np.int64)
        links = np.array([max(self.nb_poses[r] - 1, 0) for r in range(len(self.nb_poses))], dtype=np.int64)
        # chain r = first[r], first[r]+1, ... : position inside the chain by a segmented arange
        base = np.repeat(first - np.r_[0, np.cumsum(links)[:-1]], links)
        tail = base + np.arange(int(links.sum()), dtype=np.int64)
        i = np.r_[np.array([e.i for e in fixed], dtype=np.int64), tail]
        j = np.r_[np.array([e.j for e in fixed], dtype=np.int64), tail + 1]
        w = np.r_[np.array([e.weight for e in fixed], dtype=np.float64),
                  np.full(len(tail), float(self.fixed_weight))]
        return i.astype(np.int32), j.astype(np.int32), w

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
        edges = list(self.fixed_edges)
        if isinstance(self.candidate_edges, CandidateTable):
            seen.update(r for r in self.candidate_edges.robots_present() if is_other_robot_considered[r])
        else:
            edges += list(self.candidate_edges.values())
        for edge in edges:
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
        self.last_mac_error = None
        answer = w_init.copy()
        for trial in range(nb_candidates_to_choose):
            try:
                start = np.flatnonzero(w_init > 0.0)
                chosen, _, _ = mac.fw_subset_sparse(start, np.asarray(w_init, dtype=np.float64)[start],
                                                    nb_candidates_to_choose, max_iters=self.max_iters,
                                                    want_support=False)
                answer = np.zeros(len(w_init))
                answer[chosen] = 1.0
                self.last_mac_trials = trial
                return answer
            except SingularLaplacianError:
                # the reference swallows SuperLU's "Factor is exactly singular" (disconnected graph)
                w_init = self.pseudo_greedy_initialization(nb_candidates_to_choose, trial + 1,
                                                           candidate_edges)
            except CslamError as err:
                # anything else (CUDA failure, out of memory, eigen-solver not converging) is not
                # a property of the start vector: no retries, keep the greedy start, say why
                self.last_mac_error = str(err)
                logging.getLogger("cslam_b200.acm").warning(
                    "MAC solver failed (%s); falling back to the greedy selection", err)
                self.last_mac_trials = trial
                return answer
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
        if isinstance(self.candidate_edges, CandidateTable):
            fixed = self._fixed_columns(included)
            candidates = self._candidate_columns(included)
        else:
            fixed = self.rekey_edges(self.fixed_edges, included) + self.fill_odometry()
            candidates = self.rekey_edges(self.candidate_edges.values(), included)
        if _count(candidates) == 0:
            return []
        budget = min(nb_candidates_to_choose, _count(candidates))
        self.total_nb_poses = sum(self.nb_poses.values())

        if greedy_initialization:
            start = self.greedy_initialization(budget, candidates)
        elif _is_columns(candidates):
            # random_initialization redraws the weights of the rekeyed edges themselves (:247-255),
            # one np.random.rand() per edge in order = one vectorised draw from the same stream
            candidates = candidates[:2] + (np.random.rand(_count(candidates)),)
            start = self.greedy_initialization(budget, candidates)
        else:
            start = self.random_initialization(budget, candidates)
        if self.params["frontend.enable_sparsification"] and \
                self.check_initial_fixed_measurements_exists(included):
            chosen = self.run_mac_solver(fixed, candidates, start, budget)
        else:
            chosen = self.connection_biased_greedy_selection(
                budget, self.get_included_edges(self.candidate_edges.values(), included), included)

        if self.params["evaluation.enable_sparsification_comparison"]:
            self.sparsification_comparison_logs(candidates, included, start, chosen)
        selection = self.recover_inter_robot_edges(_picked(candidates, chosen), included)
        self.remove_candidate_edges(selection)
        return selection

    def sparsification_comparison_logs(self, rekeyed_candidate_edges, is_robot_included,
                                       greedy_result, mac_result):
        """Keep both selections for the evaluation topics (reference :545-557)."""
        def as_edges(mask):
            return self.recover_inter_robot_edges(_picked(rekeyed_candidate_edges, mask),
                                                  is_robot_included)
        self.log_greedy_edges = as_edges(greedy_result)
        self.log_mac_edges = as_edges(mac_result)
