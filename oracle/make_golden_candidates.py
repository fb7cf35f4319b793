"""tests/golden/candidates.npz: the candidate dictionary of the REFERENCE
AlgebraicConnectivityMaximization (cslam/algebraic_connectivity_maximization.py) after seeded
sequences of add_match / remove_candidate_edges / candidate_edges_to_fixed calls, so that the bulk
`add_matches` of this package and its candidate table are pinned to the reference's bookkeeping
(including the reversed-key lookup of :565-569) on machines without the reference checkout.
Build container only.

    python oracle/make_golden_candidates.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


def scenario(seed):
    """Rounds of (matches [n, 5] as r0, k0, r1, k1, weight; how many of the oldest candidates are
    then removed; how many of the next oldest become fixed).  Shared with the test."""
    rng = np.random.default_rng(1000 + seed)
    R = int(rng.integers(2, 6))
    rounds = []
    for _ in range(4):
        n = int(rng.integers(20, 200))
        r0 = rng.integers(0, R, n)
        r1 = (r0 + rng.integers(1, R, n)) % R
        m = np.stack([r0, rng.integers(0, 8, n), r1, rng.integers(0, 8, n), np.round(rng.random(n), 1)], axis=1)
        rounds.append((m, int(rng.integers(0, 5)), int(rng.integers(0, 3))))
    return R, rounds


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference checkout not present")
    sys.path.insert(0, REF)
    from cslam.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization as RefACM, EdgeInterRobot as RefEdge)
    out = {}
    for seed in range(8):
        R, rounds = scenario(seed)
        ref = RefACM(robot_id=0, max_nb_robots=R)
        for rnd, (m, n_remove, n_fix) in enumerate(rounds):
            for row in m:
                ref.add_match(RefEdge(int(row[0]), int(row[1]), int(row[2]), int(row[3]), float(row[4])))
            snap = [list(k) + list(v) for k, v in ref.candidate_edges.items()]
            out[f"s{seed}_r{rnd}_after_add"] = np.array(snap, dtype=np.float64).reshape(-1, 9)
            out[f"s{seed}_r{rnd}_nb_poses"] = np.array([ref.nb_poses[r] for r in range(R)])
            oldest = list(ref.candidate_edges.values())
            ref.remove_candidate_edges(oldest[:n_remove])
            ref.candidate_edges_to_fixed(list(oldest[n_remove:n_remove + n_fix]))
            snap = [list(k) + list(v) for k, v in ref.candidate_edges.items()]
            out[f"s{seed}_r{rnd}_after_edit"] = np.array(snap, dtype=np.float64).reshape(-1, 9)
            out[f"s{seed}_r{rnd}_considered"] = np.array(sorted(ref.already_considered_matches), dtype=np.int64).reshape(-1, 4)
            out[f"s{seed}_r{rnd}_fixed"] = np.array([list(e) for e in ref.fixed_edges], dtype=np.float64).reshape(-1, 5)
    path = os.path.join(ROOT, "tests", "golden", "candidates.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
