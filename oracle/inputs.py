"""Seeded input generators shared by oracle/make_golden.py and the tests, so that the
golden files only need to store OUTPUTS of the reference (inputs are re-created from
the seed).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

NNS_CASES = {"a": (300, 64, 10, np.float64), "b": (1000, 256, 30, np.float64),
             "c": (300, 64, 10, np.float32)}


def nns_case(tag):
    n, d, k, qdtype = NNS_CASES[tag]
    rng = np.random.default_rng(1234 + ord(tag))
    pool = rng.random((n, d))
    pool /= np.linalg.norm(pool, axis=1, keepdims=True)
    qs = rng.random((6, d))
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    items = [7 * i + 1 for i in range(n)]
    return pool, qs.astype(qdtype), items, k


MAC_CASES = {"g0": (3, 20, 60, 8, 0), "g1": (4, 250, 1500, 40, 1), "g2": (5, 60, 400, 20, 2),
             "g3": (2, 10, 12, 4, 3)}


def multi_robot_graph(R, P, m, seed):
    """Reference-test style graph (tests/test_algebraic_connectivity.py:38-74 of the
    reference): R chains of P poses (implicit odometry), R-1 fixed edges joining
    consecutive robots' last poses, m distinct random inter-robot candidates with
    U(0,1) weights.  Tuples are (robot0, kf0, robot1, kf1, weight)."""
    rng = np.random.default_rng(seed)
    fixed = [(r, P - 1, r + 1, P - 1, 1.0) for r in range(R - 1)]
    seen = set()
    cand = []
    while len(cand) < m:
        r0, r1 = (int(x) for x in rng.choice(R, size=2, replace=False))
        k0, k1 = int(rng.integers(0, P)), int(rng.integers(0, P))
        key = (r0, k0, r1, k1) if r0 < r1 else (r1, k1, r0, k0)
        if key in seen:
            continue
        seen.add(key)
        cand.append((r0, k0, r1, k1, float(rng.random())))
    return fixed, cand


def mac_scale_graph(R=8, P=12500, m=1000000, seed=0):
    """BASELINE.json configs[4]: R robots x P poses (odometry
    chains), R-1 fixed bridges between consecutive robots' last poses, m inter-robot
    candidates with uniform endpoints and U(0,1) weights.  Returns rekeyed arrays
    (fixed_i, fixed_j, fixed_w), (cand_i, cand_j, cand_w), n."""
    rng = np.random.default_rng(seed)
    n = R * P
    fi = np.concatenate([np.arange(r * P, r * P + P - 1) for r in range(R)] +
                        [np.array([(r + 1) * P - 1 for r in range(R - 1)])]).astype(np.int32)
    fj = np.concatenate([np.arange(r * P + 1, r * P + P) for r in range(R)] +
                        [np.array([(r + 2) * P - 1 for r in range(R - 1)])]).astype(np.int32)
    fw = np.ones(len(fi))
    r0 = rng.integers(0, R, m)
    r1 = (r0 + rng.integers(1, R, m)) % R
    ci = (r0 * P + rng.integers(0, P, m)).astype(np.int32)
    cj = (r1 * P + rng.integers(0, P, m)).astype(np.int32)
    cw = rng.random(m)
    return (fi, fj, fw), (ci, cj, cw), n
