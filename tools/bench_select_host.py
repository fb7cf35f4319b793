"""Host-side cost of AlgebraicConnectivityMaximization at the C5 graph (8 x 12500 poses, 1M
candidates): bulk insert + set-up of one selection, columnar table vs the reference-style
edge-by-edge walk (plain dict).  `--solver stub` replaces the GPU solver (runs anywhere);
`--solver gpu` runs the real select_candidates end to end.
"""
import argparse
import json
import time

import numpy as np

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cslam_b200.algebraic_connectivity_maximization import (
    AlgebraicConnectivityMaximization as ACM, EdgeInterRobot)


def c5_matches(m, R=8, poses=12500, seed=0):
    rng = np.random.default_rng(seed)
    r0 = rng.integers(0, R, m)
    r1 = (r0 + rng.integers(1, R, m)) % R
    return r0, rng.integers(0, poses, m), r1, rng.integers(0, poses, m), rng.random(m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--budget", type=int, default=1000)
    ap.add_argument("--solver", default="stub", choices=["stub", "gpu"])
    ap.add_argument("--legacy", type=int, default=1, help="also time the edge-by-edge path")
    a = ap.parse_args()
    R, poses = 8, 12500
    if a.solver == "stub":
        def stub(self, fixed, candidates, w_init, budget):
            return w_init
        ACM.run_mac_solver = stub
    m = c5_matches(a.m, R, poses)
    out = {"candidates": a.m, "budget": a.budget, "solver": a.solver}
    for name in (["columnar", "edge_by_edge"] if a.legacy else ["columnar"]):
        acm = ACM(0, R)
        if name == "edge_by_edge":
            acm.candidate_edges = {}
        for r in range(R):
            acm.nb_poses[r] = poses
        for r in range(R - 1):
            acm.add_fixed_edge(EdgeInterRobot(r, poses - 1, r + 1, poses - 1, 1.0))
        t0 = time.perf_counter()
        if name == "columnar":
            acm.add_matches(*m)
        else:
            for t in range(a.m):
                acm.add_match(EdgeInterRobot(int(m[0][t]), int(m[1][t]), int(m[2][t]), int(m[3][t]), float(m[4][t])))
        t1 = time.perf_counter()
        sel = acm.select_candidates(a.budget, {r: True for r in range(R)})
        t2 = time.perf_counter()
        sel2 = acm.select_candidates(a.budget, {r: True for r in range(R)})
        t3 = time.perf_counter()
        out[name] = {"insert_s": round(t1 - t0, 4), "select_first_s": round(t2 - t1, 4),
                     "select_second_s": round(t3 - t2, 4), "stored": len(acm.candidate_edges) + 2 * a.budget,
                     "selected": [len(sel), len(sel2)]}
        if a.solver == "gpu" and acm.last_mac is not None:
            out[name]["mac_stats"] = acm.last_mac.stats()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
