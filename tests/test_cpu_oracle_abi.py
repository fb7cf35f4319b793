"""CPU-only tests: the oracle against the reference's golden outputs, the C ABI surface,
and host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD_DIR = os.path.join(ROOT, "tests", "golden")


# ---- oracle vs golden (reference outputs) ------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nns_oracle_matches_reference_golden(tag):
    from oracle.inputs import nns_case
    from oracle.nns import NNSOracle
    gold = np.load(os.path.join(GOLD_DIR, "nns.npz"))
    pool, qs, items, k = nns_case(tag)
    orc = NNSOracle()
    orc.add_items(pool, items)
    for qi, q in enumerate(qs):
        for fn in (orc.search_loop, orc.search_vec):
            ids, sims = fn(q, k)
            assert list(ids) == list(gold[f"{tag}_ids"][qi])
            np.testing.assert_allclose(sims, gold[f"{tag}_sims"][qi], atol=1e-6 if tag == "c" else 1e-12)
        bid, bsim = orc.search_best(q)
        assert bid == gold[f"{tag}_best_id"][qi]


@pytest.mark.parametrize("tag", ["g0", "g1", "g2", "g3"])
def test_mac_oracle_matches_reference_golden(tag):
    from oracle.inputs import MAC_CASES
    from oracle.mac import Edge, MACOracle
    gold = np.load(os.path.join(GOLD_DIR, "mac.npz"))
    R, P, m, k, seed = MAC_CASES[tag]
    fixed = [Edge(int(a), int(b), float(c)) for a, b, c in gold[f"{tag}_rekey_fixed"]]
    cand = [Edge(int(a), int(b), float(c)) for a, b, c in gold[f"{tag}_rekey_cand"]]
    orc = MACOracle(fixed, cand, R * P)
    w0 = gold[f"{tag}_w0"]
    lam, vec = orc.evaluate_fiedler_pair(w0)
    # restated TraceMIN (same seed, same SuperLU options) reproduces networkx to rounding
    assert abs(lam - float(gold[f"{tag}_lambda2"])) < 1e-12
    np.testing.assert_allclose(vec, gold[f"{tag}_fiedler"], atol=1e-9)
    assert np.array_equal(orc.grad_from_fiedler(gold[f"{tag}_fiedler"]), gold[f"{tag}_grad"])
    rounded, w, u = orc.fw_subset(w0.copy(), k, max_iters=20)
    assert np.array_equal(rounded, gold[f"{tag}_fw_rounded"])
    np.testing.assert_allclose(w, gold[f"{tag}_fw_w"], atol=1e-12)
    assert abs(u - float(gold[f"{tag}_fw_u"])) < 1e-9


# ---- C ABI surface -----------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cslam_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cslam_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cslam_b200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared
    assert b"sm_100a" in lib.cslam_version()


def test_no_cpu_fallback_without_device():
    from cslam_b200 import _lib
    lib = _lib.load()
    if lib.cslam_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    assert lib.cslam_nns_create(8, 0, ctypes.byref(h)) == _lib.ERR_CUDA
    assert b"no CUDA device" in lib.cslam_last_error()
    from cslam_b200.nns_matching import NearestNeighborsMatching
    with pytest.raises(_lib.CslamError):
        NearestNeighborsMatching(8)
    from cslam_b200.mac.mac import MAC
    from cslam_b200.mac.utils import Edge
    with pytest.raises(_lib.CslamError):
        MAC([Edge(0, 1, 1.0)], [Edge(0, 1, 0.5)], 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


# ---- host logic that needs no GPU ----------------------------------------------------------
def test_edge_inter_robot_equality_and_acm_bookkeeping_on_cpu():
    from cslam_b200.algebraic_connectivity_maximization import (
        AlgebraicConnectivityMaximization, EdgeInterRobot)
    a, b = EdgeInterRobot(0, 1, 2, 3, 0.5), EdgeInterRobot(2, 3, 0, 1, 0.9)
    assert a == b and a in [b] and not (a != b)
    assert EdgeInterRobot(0, 1, 2, 4, 0.5) != a
    ac = AlgebraicConnectivityMaximization(robot_id=1, max_nb_robots=3)
    assert ac.edge_key(b) == (0, 1, 2, 3)
    ac.add_match(a)
    ac.add_match(EdgeInterRobot(0, 1, 2, 3, 0.1))
    assert ac.candidate_edges[(0, 1, 2, 3)].weight == 0.5
    ac.add_fixed_edge(EdgeInterRobot(0, 5, 1, 7, 1.0))
    assert ac.nb_poses == {0: 6, 1: 8, 2: 4}
    inc = ac.check_graph_disconnections({0: True, 1: True, 2: False})
    assert inc == {0: True, 1: True, 2: False}
    ac.compute_offsets(inc)
    assert ac.offsets == {0: 0, 1: 6, 2: 0}
    # connection-biased greedy needs no solver: robot 2 excluded -> nothing to choose
    assert ac.select_candidates(3, {0: True, 1: True, 2: False}) == []
    w = ac.greedy_initialization(1, [type("E", (), {"weight": x})() for x in (0.1, 0.7, 0.3)])
    assert list(w) == [0.0, 1.0, 0.0]


# ---- descriptor-head oracles vs the reference's own modules --------------------------------
def test_head_oracles_match_reference_golden():
    from oracle import heads
    from oracle.inputs import gem_case, keyframe_image, pca_case, subsample, vlad_case
    gold = np.load(os.path.join(GOLD_DIR, "heads.npz"))
    pre = heads.preprocess(keyframe_image(), 376)
    assert pre.shape == (3, 224, 224)
    assert np.array_equal(subsample(pre, 13), gold["pre_sub"])
    x, conv_w, cent = vlad_case()
    v = heads.netvlad_layer(x, conv_w, cent)
    np.testing.assert_allclose(subsample(v, 37), gold["vlad_sub"], atol=1e-7)
    np.testing.assert_allclose([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()],
                               gold["vlad_sum"], rtol=1e-6)
    px, comp, mean, ev, whiten = pca_case()
    np.testing.assert_allclose(heads.pca_project_normalize(px, comp, mean, ev, whiten),
                               gold["pca_out"], atol=2e-6)
    gx, p, eps, w, b = gem_case()
    np.testing.assert_allclose(heads.gem_head(gx, p, eps, w, b), gold["gem_out"], atol=1e-6)
