"""tests/golden/scancontext.npz: outputs of the REFERENCE ScanContextMatching
(cslam/lidar_pr/scancontext_matching.py) on the seeded pools of oracle/inputs.py, plus the
intermediate results of its two steps (KD-tree candidates, distance_sc per candidate) so that
the oracle and the CUDA path can be checked stage by stage.  Build container only.

    python oracle/make_golden_sc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference checkout not present")
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from scipy import spatial
    from cslam.lidar_pr.scancontext_matching import ScanContextMatching
    import cslam.lidar_pr.scancontext_utils as sc_utils
    from oracle.inputs import SC_CASES, sc_case
    out = {}
    for tag in SC_CASES:
        pool, items, queries = sc_case(tag)
        m = ScanContextMatching()
        for row, item in zip(pool, items):
            m.add_item(row, item)
        got_items, got_sims, cands, cdist, cyaw = [], [], [], [], []
        for q in queries:
            a, s = m.search(q, 1)
            b, s2 = m.search_best(q)
            assert a[0] == b and s[0] == s2
            got_items.append(a[0])
            got_sims.append(s[0])
            # the two steps on their own, as search() performs them (:59-79)
            tree = spatial.KDTree(np.array(m.ringkeys[:m.nb_items]))
            _, idx = tree.query(sc_utils.sc2rk(q.reshape(m.shape)), k=m.num_candidates)
            idx = np.asarray(idx)
            d, y = [], []
            for c in idx:
                if c >= m.nb_items:                      # fewer entries than candidates
                    d.append(np.nan)
                    y.append(-1)
                    continue
                dist, yaw = sc_utils.distance_sc(m.scancontexts[c], q.reshape(m.shape))
                d.append(dist)
                y.append(yaw)
            cands.append(np.where(idx >= m.nb_items, -1, idx))
            cdist.append(d)
            cyaw.append(y)
        out[tag + "_items"] = np.array(got_items)
        out[tag + "_sims"] = np.array(got_sims, dtype=np.float64)
        out[tag + "_cand"] = np.array(cands, dtype=np.int64)
        out[tag + "_cand_dist"] = np.array(cdist, dtype=np.float64)
        out[tag + "_cand_yaw"] = np.array(cyaw, dtype=np.int64)
        out[tag + "_ringkeys"] = np.array(m.ringkeys[:m.nb_items])
        out[tag + "_capacity"] = np.array(len(m.ringkeys))
    empty = ScanContextMatching()
    assert empty.search(np.zeros(1200), 1) == ([None], [None]) and empty.search_best(np.zeros(1200)) == (None, None)
    path = os.path.join(ROOT, "tests", "golden", "scancontext.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
