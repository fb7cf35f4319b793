"""NCCL multi-robot path on real GPUs (needs >= 2 devices; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swarm_step_over_nccl_matches_per_robot_pools():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "check_swarm.py")],
                         capture_output=True, text=True, timeout=600)
    assert "SWARM_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_round_filters_on_the_device_equal_the_host_restatement():
    """`cslam_swarm_hits` / `cslam_swarm_intra` (the similarity gate and the intra-robot filter of
    cslam/loop_closure_sparse_matching.py:45-53,74-92 applied to the all-gathered top-k on the
    device) against the numpy form of the same filters that the gloo tests exercise: same hits
    in the same (query robot, keyframe, pool robot) order, same kept matches."""
    import numpy as np
    import torch
    from cslam_b200.swarm import SwarmLoopClosureMatching
    rng = np.random.default_rng(3)
    dev = torch.device("cuda", 0)
    for R, B, kx, k_intra, k_search, thr in ((8, 64, 30, 30, 94, 0.9), (3, 5, 1, 4, 9, 0.5), (2, 1, 2, 1, 2, 0.0),
                                            (8, 200, 3, 5, 40, 0.97)):
        g_kf = rng.integers(-1, 1000, (R, R * B, kx))
        g_sims = rng.random((R, R * B, kx))
        g_sims[g_kf < 0] = np.nan                                    # empty pools answer (-1, NaN)
        all_ids = rng.integers(0, 10 ** 6, (R, B))
        rows_before = 500
        own_idx = rng.integers(-1, rows_before + B, (B, k_search))
        own_kf = rng.integers(0, 10 ** 6, (B, k_search))
        own_sims = -np.sort(-rng.random((B, k_search)), axis=1)
        t = [torch.from_numpy(np.ascontiguousarray(a)) for a in (g_kf, g_sims, all_ids, own_idx, own_kf, own_sims)]
        ref_hits, ref_intra = SwarmLoopClosureMatching._filter_round_host(t[0], t[1], t[2], thr, t[3], t[4], t[5],
                                                                          rows_before, k_intra)
        obj = SwarmLoopClosureMatching.__new__(SwarmLoopClosureMatching)
        d = [x.to(dev) for x in t]
        hits, intra = obj._filter_round(d[0], d[1], d[2], thr, d[3], d[4], d[5], rows_before, k_intra)
        assert hits.shape == ref_hits.shape and np.array_equal(hits, ref_hits)
        assert len(hits) > 0 or thr > 0.95
        assert np.array_equal(intra[:, 0], ref_intra[:, 0])
        for b in range(B):
            c = int(intra[b, 0])
            assert np.array_equal(intra[b, 1:1 + c], ref_intra[b, 1:1 + c])
            assert np.array_equal(intra[b, 1 + k_intra:1 + k_intra + c], ref_intra[b, 1 + k_intra:1 + k_intra + c])
