"""Golden for BASELINE.json configs[4] at FULL size (8 x 12 500
poses, 1 000 000 candidate edges, budget 1000, 20 Frank-Wolfe iterations), produced by
EXECUTING the reference's `MAC.fw_subset` (cslam/mac/mac.py:191-233) once in the build
container (about a minute and a half; the networkx TraceMIN/SuperLU Fiedler solve dominates).

Stored (tests/golden/mac_c5.npz, ~100 KB): for every Frank-Wolfe iteration the index set the
reference's `round_solution(grad, k)` picked (sorted int32 [20, 1000]) and its lambda_2; the
final rounded selection, the dual bound u and the support/values of the unrounded w.  The
final w is a fixed rational combination of the per-iteration sets, so those sets ARE the
result.  Inputs are re-created from the seed by
`oracle.inputs.mac_scale_graph` (identical to bench.py's `mac_graph`).

    python oracle/make_golden_c5.py [--poses 12500 --candidates 1000000 --budget 1000]

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied: it is imported and run.
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, default=8)
    ap.add_argument("--poses", type=int, default=12500)
    ap.add_argument("--candidates", type=int, default=1000000)
    ap.add_argument("--budget", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "mac_c5.npz"))
    a = ap.parse_args()
    from oracle.make_golden import _import_reference
    _import_reference()
    from cslam.mac.mac import MAC
    from cslam.mac.utils import Edge
    from oracle.inputs import mac_scale_graph

    (fi, fj, fw), (ci, cj, cw), n = mac_scale_graph(a.robots, a.poses, a.candidates, 0)
    t0 = time.time()
    fixed = [Edge(int(i), int(j), float(w)) for i, j, w in zip(fi, fj, fw)]
    cand = [Edge(int(i), int(j), float(w)) for i, j, w in zip(ci, cj, cw)]
    mac = MAC(fixed, cand, n)
    print(f"reference MAC built in {time.time() - t0:.1f} s", flush=True)

    k = a.budget
    w0 = np.zeros(len(cw))
    w0[np.argpartition(cw, -k)[-k:]] = 1.0   # greedy_initialization, acm.py:205-218

    sets, lams, gaps = [], [], []
    orig_round, orig_eval = mac.round_solution, mac.evaluate_fiedler_pair

    def rec_round(w, kk):
        r = orig_round(w, kk)
        idx = np.flatnonzero(r).astype(np.int32)
        # margin between the k-th and (k+1)-th gradient: how robust this set is to rounding
        part = np.partition(w, [-kk - 1, -kk])
        gaps.append(float(part[-kk] - part[-kk - 1]))
        sets.append(idx)
        print(f"  iteration {len(sets)}: lambda2 {lams[-1]:.12e}  k-th gap {gaps[-1]:.3e}  "
              f"({time.time() - t0:.0f} s)", file=sys.stderr, flush=True)
        return r

    def rec_eval(w, *args, **kw):
        f, v = orig_eval(w, *args, **kw)
        lams.append(float(f))
        return f, v

    mac.round_solution = rec_round
    mac.evaluate_fiedler_pair = rec_eval
    t1 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        rounded, w, u = mac.fw_subset(w0.copy(), k, max_iters=a.iters)
    t_fw = time.time() - t1
    sup = np.flatnonzero(w).astype(np.int32)
    np.savez_compressed(
        a.out, robots=a.robots, poses=a.poses, candidates=a.candidates, budget=k, iters=a.iters,
        sel_iter=np.stack(sets), lambda2_iter=np.array(lams), kth_gap_iter=np.array(gaps),
        rounded_idx=np.flatnonzero(rounded).astype(np.int32), u=np.array(u),
        w_support=sup, w_values=w[sup], fw_subset_seconds=np.array(t_fw))
    print(f"wrote {a.out}: {len(sets)} iterations, fw_subset {t_fw:.1f} s, u = {u:.12e}")


if __name__ == "__main__":
    main()
