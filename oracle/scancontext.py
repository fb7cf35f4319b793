"""Oracle: Scan Context matching (numpy).  TEST INFRASTRUCTURE ONLY.

Restates cslam/lidar_pr/scancontext_matching.py:6-104 and the helpers of
cslam/lidar_pr/scancontext_utils.py it calls (sc2rk :78-79, distance_sc :81-113).  The
reference rebuilds a scipy KDTree over the ring keys on every search (:61-66); a KD-tree is
only an index — the candidates are the `num_candidates` nearest ring keys in Euclidean
distance — so this restatement scans all ring keys.  Equal ring-key distances: ascending pool
row (scipy leaves the order of exact ties unspecified).

Pinned by tests/golden/scancontext.npz: outputs of the reference class itself on the seeded
pools of oracle/inputs.py (oracle/make_golden_sc.py, run in the build container).
"""
import numpy as np


def sc2rk(sc):
    """Ring key: mean of every ring (scancontext_utils.py:78-79)."""
    return np.mean(sc, axis=1)


def distance_sc(sc1, sc2):
    """Column-shift distance (scancontext_utils.py:81-113): for each shift s = 1..sectors of
    sc1, the mean cosine similarity over the columns that are non-zero in both; returns
    (1 - best similarity, best shift).  Same arithmetic as the reference's double loop
    (np.dot per column pair, division by the product of the two norms, running sum in column
    order), one shift at a time."""
    sectors = sc1.shape[1]
    sims = np.zeros(sectors)
    live2 = np.any(sc2, axis=0)
    norm2 = np.array([np.linalg.norm(sc2[:, j]) for j in range(sectors)])
    for i in range(sectors):
        sc1 = np.roll(sc1, 1, axis=1)
        total, engaged = 0, 0
        for j in range(sectors):
            a = sc1[:, j]
            if not np.any(a) or not live2[j]:
                continue
            total = total + np.dot(a, sc2[:, j]) / (np.linalg.norm(a) * norm2[j])
            engaged += 1
        sims[i] = total / engaged if engaged else 0.0
    return 1 - np.max(sims), int(np.argmax(sims)) + 1


class ScanContextMatchingOracle(object):
    """ScanContextMatching (scancontext_matching.py:6-104) with the same growth, the same
    answers for an empty pool and the same fallback when no candidate is closer than 1."""

    def __init__(self, shape=[20, 60], num_candidates=10, threshold=0.15):
        self.shape = shape
        self.num_candidates = num_candidates
        self.threshold = threshold
        self.scancontexts = np.zeros((1000, shape[0], shape[1]))
        self.ringkeys = np.zeros((1000, shape[0]))
        self.items = dict()
        self.nb_items = 0

    def add_item(self, descriptor, item):                      # :24-46
        sc = np.asarray(descriptor).reshape(self.shape)
        if self.nb_items >= len(self.ringkeys):
            grown = np.zeros((2 * len(self.scancontexts),) + tuple(self.shape))
            grown[:self.nb_items] = self.scancontexts[:self.nb_items]
            self.scancontexts = grown
            keys = np.zeros((2 * len(self.ringkeys), self.shape[0]))
            keys[:self.nb_items] = self.ringkeys[:self.nb_items]
            self.ringkeys = keys
        self.scancontexts[self.nb_items] = sc
        self.ringkeys[self.nb_items] = sc2rk(sc)
        self.items[self.nb_items] = item
        self.nb_items += 1

    def candidates(self, query):
        """Rows of the nearest ring keys, nearest first (:59-66)."""
        key = sc2rk(np.asarray(query, dtype=np.float64).reshape(self.shape))
        diff = self.ringkeys[:self.nb_items] - key
        d2 = np.zeros(self.nb_items)
        for r in range(self.shape[0]):                         # same accumulation order as the kernel
            d2 = d2 + diff[:, r] * diff[:, r]
        order = np.lexsort((np.arange(self.nb_items), d2))
        return order[:self.num_candidates]

    def search_details(self, query):
        """(row or -1, similarity, yaw shift, candidate rows, candidate distances)."""
        q = np.asarray(query, dtype=np.float64).reshape(self.shape)
        cand = self.candidates(q)
        nn_dist, nn_idx, nn_yaw = 1.0, -1, 0
        dists = []
        for row in cand:                                       # :69-79
            dist, yaw = distance_sc(self.scancontexts[row], q)
            dists.append(dist)
            if dist < nn_dist:
                nn_dist, nn_idx, nn_yaw = dist, int(row), yaw
        sim = 0.0 if nn_idx < 0 else 1 - nn_dist               # :81-87
        return nn_idx, sim, nn_yaw, cand, np.array(dists)

    def search(self, query, k):                                # :48-89
        if self.nb_items < 1:
            return [None], [None]
        row, sim, _, _, _ = self.search_details(query)
        return [self.items[max(row, 0)]], [sim]

    def search_best(self, query):                              # :91-104
        if self.nb_items < 1:
            return None, None
        idxs, sims = self.search(query, 1)
        return idxs[0], sims[0]
