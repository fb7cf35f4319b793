"""Oracle for the descriptor routing of the loop-closure front end (TEST INFRASTRUCTURE ONLY).

Restates cslam/loop_closure_sparse_matching.py:36-92 (which pool a descriptor is added to,
which pools are searched, the similarity gate, the intra-robot filter) and the candidate
dictionary update of cslam/algebraic_connectivity_maximization.py:559-572 (`add_match`,
including its un-normalised key lookup) on top of `oracle.nns.NNSOracle`.
Pinned by tests/golden/frontend.npz (reference outputs, oracle/make_golden.py:gen_frontend).
"""
from oracle.nns import NNSOracle


class FrontendOracle(object):
    def __init__(self, params):
        self.p = params
        self.local = NNSOracle()
        self.others = {i: NNSOracle() for i in range(params['max_nb_robots'])
                       if i != params['robot_id']}
        self.candidates = {}   # normalised key -> (r0, k0, r1, k1, weight)

    @staticmethod
    def _key(e):
        return (e[0], e[1], e[2], e[3]) if e[0] < e[2] else (e[2], e[3], e[0], e[1])

    def add_match(self, e):
        # acm.py:559-572: lookup with the raw key, store under the normalised key
        raw = (e[0], e[1], e[2], e[3])
        if raw in self.candidates and not e[4] > self.candidates[raw][4]:
            return
        self.candidates[self._key(e)] = e

    @staticmethod
    def _best(nn, d):
        # nns_matching.py:63-76
        if len(nn.data) == 0:
            return None, None
        ids, sims = nn.search_loop(d, 1)
        return ids[0], sims[0]

    def local_keyframe(self, d, kf):
        """-> (intra match or None, [new candidate tuples])   lcsm.py:36-54, :74-92"""
        intra = None
        if self.p['frontend.enable_intra_robot_loop_closures']:
            ids, sims = ([], []) if len(self.local.data) == 0 else \
                self.local.search_loop(d, self.p['frontend.nb_best_matches'])
            ids, sims = list(ids), list(sims)
            if ids and ids[0] == kf:
                ids, sims = ids[1:], sims[1:]
            for i, s in zip(ids, sims):
                if abs(i - kf) < self.p['frontend.intra_loop_min_inbetween_keyframes']:
                    continue
                if s < self.p['frontend.similarity_threshold']:
                    continue
                intra = i
                break
        self.local.add_item(d, kf)
        out = []
        for r in sorted(self.others):
            i, s = self._best(self.others[r], d)
            if i is not None and s >= self.p['frontend.similarity_threshold']:
                e = (self.p['robot_id'], kf, r, i, float(s))
                self.add_match(e)
                out.append(e)
        return intra, out

    def remote_keyframe(self, robot, d, kf):
        """lcsm.py:56-72"""
        self.others[robot].add_item(d, kf)
        i, s = self._best(self.local, d)
        if i is not None and s >= self.p['frontend.similarity_threshold']:
            e = (self.p['robot_id'], i, robot, kf, float(s))
            self.add_match(e)
            return e
        return None
