"""cProfile of AlgebraicConnectivityMaximization.select_candidates at C5 (second and third call on a warm process)."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cslam_b200.algebraic_connectivity_maximization import AlgebraicConnectivityMaximization as ACM, EdgeInterRobot
from tools.bench_select_host import c5_matches

R, poses = 8, 12500
m = c5_matches(1_000_000, R, poses)
acm = ACM(0, R)
for r in range(R):
    acm.nb_poses[r] = poses
for r in range(R - 1):
    acm.add_fixed_edge(EdgeInterRobot(r, poses - 1, r + 1, poses - 1, 1.0))
t0 = time.perf_counter(); acm.add_matches(*m); t1 = time.perf_counter()
print("insert (first GPU touch of the process included)", round(t1 - t0, 3))
acm.select_candidates(1000, {r: True for r in range(R)})
for rep in range(2):
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    sel = acm.select_candidates(1000, {r: True for r in range(R)})
    pr.disable(); dt = time.perf_counter() - t0
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
    print(f"select_candidates call {rep + 2}: {dt * 1e3:.1f} ms, {len(sel)} selected")
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[6:24]))
t0 = time.perf_counter(); acm.add_matches(*c5_matches(1_000_000, R, poses, seed=5)); print("second bulk insert of 1M", round(time.perf_counter() - t0, 3))
