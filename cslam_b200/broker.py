"""Broker — decides which keyframe vertices of the selected inter-robot edges are shipped
between robots.  Same class API as the reference (cslam/broker.py:8-129):
`Broker(edges, robots_involved).brokerage(use_vertex_cover) -> list(set((robot, keyframe)))`.

This is the step right after `select_candidates` (reference
cslam/global_descriptor_loop_closure_detection.py:328-339).  The graphs are tiny (at most
`inter_robot_loop_closure_budget` edges), so it is host code; unlike the reference it does
not need networkx:

  * two robots  -> every component is bipartite: maximum matching by augmenting paths
    (Hopcroft-Karp style layered search) and the minimum vertex cover from König's theorem,
    which is what `nx.bipartite.maximum_matching` + `to_vertex_cover` compute
    (broker.py:104-107);
  * more robots -> the local-ratio 2-approximation of Bar-Yehuda & Even with unit weights,
    the algorithm behind `min_weighted_vertex_cover` (broker.py:108-111);
  * `use_vertex_cover=False` -> the "simple dialog": one random endpoint per uncovered edge
    (broker.py:114-129).
"""
from collections import deque

import numpy as np


class Broker(object):
    """The broker decides which vertices in the matching graph are shared between robots."""

    def __init__(self, edges, robots_involved):
        """
        Args:
            edges (list(EdgeInterRobot)): selected inter-robot edges
            robots_involved (list(int)): ids of the robots taking part in the exchange
        """
        self.edges = edges
        involved = set(robots_involved)
        with_edges = []
        for e in edges:
            for rid in (e.robot0_id, e.robot1_id):
                if rid in involved and rid not in with_edges:
                    with_edges.append(rid)
        with_edges.sort()
        self.robots_involved_with_edges = with_edges
        self.is_multi_robot_graph = len(with_edges) >= 2
        self.is_bipartite = len(with_edges) == 2
        # adjacency in insertion order (dict keeps it): vertex -> list of neighbours
        self.adjacency = {}
        if not self.is_multi_robot_graph:
            return
        for e in edges:
            u = (e.robot0_id, e.robot0_keyframe_id)
            v = (e.robot1_id, e.robot1_keyframe_id)
            for x in (u, v):
                if x[0] in with_edges and x not in self.adjacency:
                    self.adjacency[x] = []
            if u[0] in with_edges and v[0] in with_edges and u != v:
                if v not in self.adjacency[u]:
                    self.adjacency[u].append(v)
                    self.adjacency[v].append(u)

    # ------------------------------------------------------------------
    def brokerage(self, use_vertex_cover):
        """Vertices to transmit: vertex cover per connected component, or simple dialog."""
        if not self.is_multi_robot_graph:
            return []
        return self.vertex_cover() if use_vertex_cover else self.simple_dialog()

    def connected_components(self):
        seen, comps = set(), []
        for root in self.adjacency:
            if root in seen:
                continue
            comp, queue = [], deque([root])
            seen.add(root)
            while queue:
                x = queue.popleft()
                comp.append(x)
                for y in self.adjacency[x]:
                    if y not in seen:
                        seen.add(y)
                        queue.append(y)
            comps.append(comp)
        return comps

    def vertex_cover(self):
        """Minimum (two robots) or 2-approximate (more robots) vertex cover of every
        connected component of the matching graph."""
        covers = []
        for comp in self.connected_components():
            if self.is_bipartite:
                covers.append(self._konig_cover(comp))
            else:
                covers.append(self._local_ratio_cover(comp))
        return covers

    # -- bipartite: maximum matching + König ---------------------------------
    def _konig_cover(self, comp):
        left_robot = self.robots_involved_with_edges[0]
        left = [x for x in comp if x[0] == left_robot]
        adj = self.adjacency
        match = {}  # vertex -> partner, both directions

        def layers():
            """BFS from the free left vertices; returns distance labels or None when no
            augmenting path exists."""
            dist, queue, found = {}, deque(), False
            for u in left:
                if u not in match:
                    dist[u] = 0
                    queue.append(u)
            while queue:
                u = queue.popleft()
                for v in adj[u]:
                    w = match.get(v)
                    if w is None:
                        found = True
                    elif w not in dist:
                        dist[w] = dist[u] + 1
                        queue.append(w)
            return dist if found else None

        def augment(u, dist):
            for v in adj[u]:
                w = match.get(v)
                if w is None or (dist.get(w) == dist[u] + 1 and augment(w, dist)):
                    match[u], match[v] = v, u
                    return True
            dist[u] = None  # dead end for this phase
            return False

        while True:
            dist = layers()
            if dist is None:
                break
            progressed = False
            for u in left:
                if u not in match and augment(u, dist):
                    progressed = True
            if not progressed:
                break

        # König: Z = vertices reachable from free left vertices along alternating paths
        # (non-matching edge left->right, matching edge right->left);
        # cover = (left \ Z) | (right & Z)
        z, queue = set(), deque()
        for u in left:
            if u not in match:
                z.add(u)
                queue.append(u)
        while queue:
            u = queue.popleft()
            for v in adj[u]:
                if v in z or match.get(u) == v:
                    continue
                z.add(v)
                w = match.get(v)
                if w is not None and w not in z:
                    z.add(w)
                    queue.append(w)
        left_set = set(left)
        return {x for x in comp if (x in left_set and x not in z) or (x not in left_set and x in z)}

    # -- general graphs: local-ratio 2-approximation --------------------------
    def _local_ratio_cover(self, comp):
        cost = {x: 1.0 for x in comp}
        cover = set()
        for u in comp:
            for v in self.adjacency[u]:
                if u in cover or v in cover:
                    continue
                if cost[u] <= cost[v]:
                    cover.add(u)
                    cost[v] -= cost[u]
                else:
                    cover.add(v)
                    cost[u] -= cost[v]
        return cover

    # -- no vertex cover ---------------------------------------------------------
    def simple_dialog(self):
        """For each edge send one endpoint picked at random unless one is already sent."""
        chosen = set()
        for e in self.edges:
            ends = ((e.robot0_id, e.robot0_keyframe_id), (e.robot1_id, e.robot1_keyframe_id))
            if ends[0] not in chosen and ends[1] not in chosen:
                chosen.add(ends[np.random.randint(2)])
        return [chosen]
