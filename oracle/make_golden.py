"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (lajoiepy/cslam) on seeded
inputs.  Runs only in the build container, where /root/reference is mounted
read-only; the GPU box never needs it (tests read the committed .npz files).

    python oracle/make_golden.py            # regenerate everything

Nothing from the reference is copied: its modules are imported, called, and only
their numerical outputs are stored.  `ament_index_python` (ROS, absent here) is
stubbed because cslam/vpr/*.py import it at module scope.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference checkout not present; golden files can only be "
                         "regenerated in the build container")
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    stub = types.ModuleType("ament_index_python")
    pk = types.ModuleType("ament_index_python.packages")
    pk.get_package_share_directory = lambda name: "/nonexistent"
    stub.packages = pk
    sys.modules["ament_index_python"] = stub
    sys.modules["ament_index_python.packages"] = pk


# --------------------------------------------------------------------------
def gen_nns():
    from cslam.nns_matching import NearestNeighborsMatching
    from oracle.inputs import NNS_CASES, nns_case
    out = {}
    for tag in NNS_CASES:
        pool, qs, items, k = nns_case(tag)
        nn = NearestNeighborsMatching()
        for i in range(len(pool)):
            nn.add_item(pool[i], items[i])
        ids, sims = [], []
        for q in qs:
            a, s = nn.search(q, k)
            ids.append(a)
            sims.append(s)
        best = [nn.search_best(q) for q in qs]
        out[f"{tag}_ids"] = np.array(ids)
        out[f"{tag}_sims"] = np.array(sims)
        out[f"{tag}_best_id"] = np.array([b[0] for b in best])
        out[f"{tag}_best_sim"] = np.array([b[1] for b in best])
    np.savez_compressed(os.path.join(GOLD, "nns.npz"), **out)


def gen_mac():
    from cslam.algebraic_connectivity_maximization import (AlgebraicConnectivityMaximization,
                                                           EdgeInterRobot)
    from cslam.mac.mac import MAC
    import io
    import contextlib
    out = {}
    from oracle.inputs import MAC_CASES, multi_robot_graph
    for tag, (R, P, m, k, seed) in MAC_CASES.items():
        fixed, cand = multi_robot_graph(R, P, m, seed)
        acm = AlgebraicConnectivityMaximization(robot_id=0, max_nb_robots=R)
        acm.set_graph([EdgeInterRobot(*e) for e in fixed], [EdgeInterRobot(*e) for e in cand])
        inc = {r: True for r in range(R)}
        # MAC-level goldens on the rekeyed graph
        acm.compute_offsets(inc)
        rf = acm.rekey_edges(acm.fixed_edges, inc)
        rf.extend(acm.fill_odometry())
        rc = acm.rekey_edges(acm.candidate_edges.values(), inc)
        n = sum(acm.nb_poses.values())
        mac = MAC(rf, rc, n)
        w0 = acm.greedy_initialization(k, rc)
        lam, vec = mac.evaluate_fiedler_pair(w0)
        grad = mac.grad_from_fiedler(vec)
        with contextlib.redirect_stdout(io.StringIO()):
            rounded, w, u = mac.fw_subset(w0.copy(), k, max_iters=20)
        out[f"{tag}_rekey_fixed"] = np.array([(e.i, e.j, e.weight) for e in rf])
        out[f"{tag}_rekey_cand"] = np.array([(e.i, e.j, e.weight) for e in rc])
        out[f"{tag}_w0"] = w0
        out[f"{tag}_lambda2"] = np.array(lam)
        out[f"{tag}_fiedler"] = vec
        out[f"{tag}_grad"] = grad
        out[f"{tag}_fw_rounded"] = rounded
        out[f"{tag}_fw_w"] = w
        out[f"{tag}_fw_u"] = np.array(u)
        # end-to-end select_candidates through the reference class
        with contextlib.redirect_stdout(io.StringIO()):
            sel = acm.select_candidates(k, inc, greedy_initialization=True)
        out[f"{tag}_selected"] = np.array([tuple(e) for e in sel])
        out[f"{tag}_remaining"] = np.array(len(acm.candidate_edges))
    np.savez_compressed(os.path.join(GOLD, "mac.npz"), **out)


def gen_frontend():
    """Drives the reference LoopClosureSparseMatching with the call sequence of the ROS
    wrapper (global_descriptor_loop_closure_detection.py:148-174 for local keyframes,
    :407-422 for remote descriptors, :309-325 for the periodic selection) on the seeded
    stream of oracle.inputs.frontend_scenario."""
    import collections
    import contextlib
    import io
    from cslam.loop_closure_sparse_matching import LoopClosureSparseMatching
    from cslam.algebraic_connectivity_maximization import EdgeInterRobot
    from oracle.inputs import FRONTEND_PARAMS, frontend_scenario
    Msg = collections.namedtuple("GlobalDescriptor", ["keyframe_id", "robot_id", "descriptor"])
    lcm = LoopClosureSparseMatching(dict(FRONTEND_PARAMS))
    intra, matches = [], []
    for ev in frontend_scenario():
        if ev[0] == 'local':
            for kf, d in zip(ev[1], ev[2]):
                kf_match, _ = lcm.match_local_loop_closures(d, kf)      # detect_intra
                intra.append((kf, -1 if kf_match is None else kf_match))
                matches.extend(lcm.add_local_global_descriptor(d, kf))
        else:
            for kf, d in zip(ev[2], ev[3]):
                m = lcm.add_other_robot_global_descriptor(Msg(kf, ev[1], d.tolist()))
                if m is not None:
                    matches.append(m)
    out = {"intra": np.array(intra), "matches": np.array([tuple(m) for m in matches])}
    cand = lcm.candidate_selector.candidate_edges
    out["cand_keys"] = np.array(sorted(cand.keys()))
    out["cand_weights"] = np.array([cand[k].weight for k in sorted(cand.keys())])
    # geometric verification feedback: one fixed edge per robot pair so that MAC runs
    for r0, r1 in ((0, 1), (1, 2)):
        lcm.candidate_selector.add_fixed_edge(EdgeInterRobot(r0, 0, r1, 0, 1.0))
    with contextlib.redirect_stdout(io.StringIO()):
        sel = lcm.select_candidates(FRONTEND_PARAMS['frontend.inter_robot_loop_closure_budget'],
                                    {0: True, 1: True, 2: True})
    out["selected"] = np.array([tuple(e) for e in sel])
    out["remaining"] = np.array(len(lcm.candidate_selector.candidate_edges))
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **out)


GENERATORS = {"nns": gen_nns, "mac": gen_mac, "frontend": gen_frontend}


def main(argv):
    _import_reference()
    os.makedirs(GOLD, exist_ok=True)
    try:
        from oracle import make_golden_heads
        GENERATORS.update(make_golden_heads.GENERATORS)
    except ImportError:
        pass
    which = argv[1:] or list(GENERATORS)
    for name in which:
        GENERATORS[name]()
        print("wrote golden:", name)


if __name__ == "__main__":
    main(sys.argv)
