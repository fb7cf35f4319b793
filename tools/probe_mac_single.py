"""Smallest program that launches k_lobpcg_persist on the C5 graph: one cold Fiedler solve
(greedy start, 1000 active candidates, ~180 LOBPCG iterations in ONE launch).  For ncu
(`--replay-mode application` re-runs the whole program once per pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.inputs import mac_scale_graph
from cslam_b200.mac.mac import MAC

fixed, cand, n = mac_scale_graph(8, 12500, 1000000)
mac = MAC(fixed, cand, n)
w0 = np.zeros(len(cand[2]))
w0[np.argpartition(cand[2], -1000)[-1000:]] = 1.0
lam, vec = mac.evaluate_fiedler_pair(w0)
print("lambda2", lam, "iterations", mac.last_lobpcg_iters, mac.solver_timing())
