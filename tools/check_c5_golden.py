"""C5 (100k poses / 1M candidates / budget 1000) against tests/golden/mac_c5.npz, iteration by
iteration.  Two comparisons:

  run     our own 20-iteration fw_subset vs the reference's sets (errors may propagate: one
          differing boundary edge changes every later iterate a little)
  replay  every iteration on the REFERENCE's iterate w_i (rebuilt from its stored sets, the
          update rule of cslam/mac/mac.py:229-230): our Fiedler pair -> gradient -> top-k against
          the reference's s_i.  Iterations are independent here; a difference can only come
          from the eigen-solvers disagreeing about edges at the k-th/(k+1)-th boundary.

For every differing edge the distance of its gradient from the k-th largest one is printed
relative to the largest gradient: that is the accuracy either eigen-solver would need to
decide it.  Needs a GPU; prints a table (kept under profiles/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import greedy_w_init, mac_graph
    from cslam_b200.mac.mac import MAC
    g = np.load(os.path.join(ROOT, "tests", "golden", "mac_c5.npz"))
    fixed, cand, n = mac_graph(int(g["robots"]), int(g["poses"]), int(g["candidates"]))
    k = int(g["budget"])
    mac = MAC(fixed, cand, n)
    w0 = greedy_w_init(cand[2], k)
    sel, w, u = mac.fw_subset(w0.copy(), k, max_iters=int(g["iters"]), trace=True)
    tsel, tf = mac.last_trace
    print("it  lambda2(ref)      lambda2(run)     rel.diff  run:diff  replay:diff  "
          "max |g_e - g_kth| / g_max of differing edges   ref kth gap / g_max")
    w_i = w0.copy()
    for it, ref in enumerate(g["sel_iter"]):
        lam, vec = mac.evaluate_fiedler_pair(w_i)
        grad = mac.grad_from_fiedler(vec)
        ours = set(np.argpartition(grad, -k)[-k:].tolist())
        d_replay = ours ^ set(ref.tolist())
        d_run = set(tsel[it].tolist()) ^ set(ref.tolist())
        kth = np.partition(grad, -k)[-k]
        band = max([abs(grad[e] - kth) for e in d_replay], default=0.0) / grad.max()
        print(f"{it:2d}  {float(g['lambda2_iter'][it]):.9e}  {tf[it]:.9e}  "
              f"{abs(tf[it] - g['lambda2_iter'][it]) / g['lambda2_iter'][it]:.1e}  {len(d_run):8d}  "
              f"{len(d_replay):11d}  {band:.3e}   (replay lambda2 rel.diff "
              f"{abs(lam - g['lambda2_iter'][it]) / g['lambda2_iter'][it]:.1e})   "
              f"{float(g['kth_gap_iter'][it]) / grad.max():.2e}")
        s_i = np.zeros(len(w0))
        s_i[ref] = 1.0
        w_i = w_i + 2.0 / (it + 2.0) * (s_i - w_i)
    print("final selection identical:", np.array_equal(np.flatnonzero(sel), g["rounded_idx"]),
          " dual bound rel.diff:", abs(u - float(g["u"])) / abs(float(g["u"])))


if __name__ == "__main__":
    main()
