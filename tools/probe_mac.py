"""GPU probe: MAC.fw_subset at a given scale, timing + (optionally) oracle comparison."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.inputs import mac_scale_graph
from cslam_b200.mac.mac import MAC

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=8)
ap.add_argument("--P", type=int, default=12500)
ap.add_argument("--m", type=int, default=1000000)
ap.add_argument("--k", type=int, default=1000)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--oracle", type=int, default=0)
ap.add_argument("--bs", type=int, default=1)
a = ap.parse_args()
fixed, cand, n = mac_scale_graph(a.R, a.P, a.m)
t0 = time.time()
mac = MAC(fixed, cand, n)
mac.set_options(block_size=a.bs)
print("create", time.time() - t0)
w0 = np.zeros(a.m)
w0[np.argpartition(cand[2], -a.k)[-a.k:]] = 1.0
for r in range(a.reps):
    s0 = mac.stats()
    t0 = time.time()
    rounded, w, u = mac.fw_subset(w0.copy(), a.k, max_iters=a.iters, trace=True)
    dt = time.time() - t0
    s1 = mac.stats()
    print(f"rep {r}: fw_subset {dt*1e3:.1f} ms, fw iters {mac.last_fw_iters}, lobpcg iters {s1['lobpcg_iters']-s0['lobpcg_iters']}, "
          f"spmv cols {s1['spmv_columns']-s0['spmv_columns']}, u {u:.6e}, f0 {mac.last_trace[1][0]:.6e} f_last {np.nanmax(mac.last_trace[1]):.6e} jacobi {s1['jacobi_fallback']}")
t0 = time.time(); lam, vec = mac.evaluate_fiedler_pair(w0); print("single fiedler", time.time() - t0, lam, mac.last_lobpcg_iters)
if a.oracle:
    from oracle.mac import MACOracle, Edge, topk_boundary_gap
    orc = MACOracle.__new__(MACOracle)
    from oracle.mac import laplacian_from_arrays
    orc.L_odom = laplacian_from_arrays(*fixed, n); orc.num_poses = n
    orc.weights = cand[2]; orc.edge_list = np.stack([cand[0], cand[1]], axis=1)
    trace = []
    t0 = time.time(); r_ref, w_ref, u_ref = orc.fw_subset(w0.copy(), a.k, max_iters=a.iters, trace=trace); t_ref = time.time() - t0
    print(f"oracle fw_subset {t_ref:.1f} s, u {u_ref:.6e}")
    tsel, tf = mac.last_trace
    for it, t in enumerate(trace):
        ref_set = set(np.nonzero(t['s'])[0]); ours = set(tsel[it])
        gap = topk_boundary_gap(t['grad'], a.k) / t['grad'].max()
        print(f" it {it}: f ours {tf[it]:.9e} ref {t['f']:.9e}  |s_ours ^ s_ref| = {len(ours ^ ref_set)}  rel gap {gap:.2e}")
    print("final set diff", int(np.abs(rounded - r_ref).sum()), "objective ours", orc.evaluate_objective(rounded), "ref", orc.evaluate_objective(r_ref))
