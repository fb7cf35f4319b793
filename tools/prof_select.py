import cProfile, pstats, sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np
from bench_select_host import c5_matches
from cslam_b200.algebraic_connectivity_maximization import AlgebraicConnectivityMaximization as ACM, EdgeInterRobot
R, poses = 8, 12500
m = c5_matches(1000000, R, poses)
acm = ACM(0, R)
for r in range(R): acm.nb_poses[r] = poses
for r in range(R - 1): acm.add_fixed_edge(EdgeInterRobot(r, poses - 1, r + 1, poses - 1, 1.0))
acm.add_matches(*m)
acm.select_candidates(1000, {r: True for r in range(R)})
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
acm.select_candidates(1000, {r: True for r in range(R)})
print("select", time.perf_counter() - t0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
