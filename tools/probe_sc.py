"""Scan Context matcher on the GPU: timing of the two stages at a given pool size, against
their HBM floors, with the numpy oracle timed beside it on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--q", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--oracle-n", type=int, default=2000)
    a = ap.parse_args()
    from cslam_b200.lidar_pr.scancontext_matching import ScanContextMatching
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    hbm = float(peaks.get("hbm_gbs", 6551.4))
    rng = np.random.default_rng(0)
    R, S = 20, 60
    m = ScanContextMatching()
    t0 = time.perf_counter()
    for lo in range(0, a.n, 50_000):
        k = min(50_000, a.n - lo)
        blk = (rng.random((k, R * S), dtype=np.float32) * 5) * (rng.random((k, R * S), dtype=np.float32) > 0.3)
        if lo == 0:
            first = blk[:a.q].copy()
        m.add_items(blk, range(lo, lo + k))
    t_add = time.perf_counter() - t0
    queries = np.stack([np.roll(first[i].reshape(R, S), 1 + i % (S - 1), axis=1).ravel() for i in range(a.q)])
    knn, dist, wall = [], [], []
    for rep in range(a.reps + 3):
        t0 = time.perf_counter()
        rows, sims, yaw = m.search_batch(queries)
        t1 = time.perf_counter()
        if rep >= 3:
            k_ms, d_ms = m.last_timing()
            knn.append(k_ms); dist.append(d_ms); wall.append((t1 - t0) * 1e3)
    assert np.array_equal(rows, np.arange(a.q)), rows[:8]
    knn_ms, dist_ms = float(np.median(knn)), float(np.median(dist))
    # step 1 streams the ring-key table (rings*8 B per entry) once per tile of 8 queries
    tiles = (a.q + 7) // 8
    knn_bytes = a.n * R * 8 * tiles
    dist_bytes = a.q * m.num_candidates * R * S * 8 * 2
    out = {"pool": a.n, "queries": a.q, "add_s": round(t_add, 2),
           "knn_ms": round(knn_ms, 4), "distance_ms": round(dist_ms, 4), "search_wall_ms": round(float(np.median(wall)), 3),
           "queries_per_s": round(a.q / (float(np.median(wall)) * 1e-3), 1),
           "knn_algorithmic_bytes": knn_bytes, "knn_GBps": round(knn_bytes / (knn_ms * 1e-3) / 1e9, 1),
           "knn_frac_of_hbm_peak": round(knn_bytes / (knn_ms * 1e-3) / 1e9 / hbm, 4),
           "knn_fp64_GFLOPs": round(3.0 * R * a.n * a.q / (knn_ms * 1e-3) / 1e9, 1),
           "hbm_peak_GBps": hbm, "distance_algorithmic_bytes": dist_bytes,
           "distance_GBps": round(dist_bytes / (dist_ms * 1e-3) / 1e9, 1)}
    # the oracle (reference arithmetic, numpy) on a bounded pool
    from oracle.scancontext import ScanContextMatchingOracle
    o = ScanContextMatchingOracle()
    for i in range(a.oracle_n):
        o.add_item(first[i % len(first)].astype(np.float64) + (i // len(first)) * 1e-3, i)
    t0 = time.perf_counter()
    o.search(queries[0].astype(np.float64), 1)
    out["oracle_s_per_query_at_pool_%d" % a.oracle_n] = round(time.perf_counter() - t0, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
