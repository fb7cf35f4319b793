"""Stream ordering of the device-resident entry points — the path bench.py times.

The reference computes a descriptor and then adds / searches it in the same Python thread
(cslam/global_descriptor_loop_closure_detection.py:148-174,388-405), so "the search sees the
descriptor" is trivially true there.  Here the descriptor comes out of GPU kernels that are
still RUNNING when `add_items_device` / `search_batch_device` are called: the C ABI must order
its own kernels (and pool growth copies) behind the producer on the caller's stream.  Every
test below keeps the producer's stream busy for tens of milliseconds first, so that a library
that runs on a private stream reads the descriptor buffer before it is written.

Checked against the oracle (oracle/nns.py, oracle/frontend.py): the appended pool rows are the
descriptors bit for bit, and the matches are the reference arithmetic's matches.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM, K = 512, 30


def _busy(dev, reps=4):
    """~20 ms per repetition of fp32 matmul on the CURRENT stream; returns a small tensor that
    depends on all of it (bounded in [-1, 1])."""
    import torch
    a = torch.randn(8192, 8192, device=dev) * 0.01
    b = a
    for _ in range(reps):
        b = torch.tanh(b @ a)
    return b[:64, :DIM].contiguous()


def _filled_pool(rows, dev, seed=3):
    import torch
    from cslam_b200.nns_matching import NearestNeighborsMatching
    g = torch.Generator(device=dev).manual_seed(seed)
    pool = NearestNeighborsMatching()
    for s in range(0, rows, 65536):
        m = min(65536, rows - s)
        x = torch.rand((m, DIM), generator=g, device=dev)
        pool.add_items_device(x / x.norm(dim=1, keepdim=True), range(s, s + m))
    torch.cuda.synchronize()
    return pool


def _oracle_of(pool, rows):
    from oracle.nns import NNSOracle
    orc = NNSOracle(DIM)
    orc.data, orc.n = pool.read_rows(0, rows), rows
    orc.items = pool.items
    return orc


# 200 000 rows: no growth on the racing append; 262 144 - 32: the append must grow the pool
# (capacity doubles in multiples of 256 rows), i.e. the growth copies race as well
@pytest.mark.parametrize("rows", [200000, 262144 - 32])
@pytest.mark.parametrize("side_stream", [False, True])
def test_append_and_search_wait_for_the_producer(rows, side_stream):
    import torch
    from oracle.nns import lists_match_modulo_ties
    dev = torch.device("cuda", 0)
    pool = _filled_pool(rows, dev)
    g = torch.Generator(device=dev).manual_seed(11)
    base = torch.rand((64, DIM), generator=g, device=dev)
    stream = torch.cuda.Stream(dev) if side_stream else torch.cuda.default_stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        z = _busy(dev)
        emb = torch.nn.functional.normalize(base + 0.05 * z)       # ready only when the stream drains
        pool.add_items_device(emb, range(rows, rows + 64))
        idx, sims = pool.search_batch_device(emb, K)
        # a host-pointer call right behind it runs on the handle's own stream and must be
        # ordered behind the device-pointer append as well
        idx_h, sims_h = pool.search_batch(np.zeros((1, DIM), np.float32) + 1.0, 1)
    torch.cuda.synchronize()
    emb_h = emb.cpu().numpy()
    assert np.isfinite(emb_h).all()
    got = pool.read_rows(rows, 64)
    assert np.array_equal(got, emb_h), "appended rows are not the descriptors (read before written?)"
    orc = _oracle_of(pool, rows + 64)
    idx, sims = idx.cpu().numpy(), sims.cpu().numpy()
    for q in range(64):
        full = orc.similarities_vec(emb_h[q])
        ref = np.argsort(full)[::-1][:K]
        assert ref[0] == rows + q                                   # the descriptor finds itself
        assert lists_match_modulo_ties(list(idx[q]), list(ref), full), f"query {q}"
        assert np.abs(sims[q] - full[idx[q]]).max() < 1e-6
    full = orc.similarities_vec(np.ones(DIM, np.float32))
    assert abs(sims_h[0, 0] - full.max()) < 1e-6 and full[idx_h[0, 0]] >= full.max() - 1e-6


def _wrapper(dev):
    import argparse
    import torch
    import bench
    from cslam_b200.global_descriptor_loop_closure_detection import GlobalDescriptorLoopClosureDetection
    from cslam_b200.local_node import LocalNode
    from cslam_b200.vpr.cosplace import CosPlace
    args = argparse.Namespace(k=K, mac_budget=100, dim=DIM, backbone="resnet18", precision="fp32")
    params = bench.frontend_params(args, 0, 1)
    net = CosPlace(params, None, state_dict=bench.cosplace_state_dict("resnet18", DIM), device=0)
    bench.centre_head_bias(net, dev)
    node = LocalNode(None, "/r0", {'frontend.global_descriptors_topic': 'global_descriptors',
                                   'frontend.inter_robot_matches_topic': 'inter_robot_matches'})
    return GlobalDescriptorLoopClosureDetection(params, node, global_descriptor=net), node, net, params


@pytest.mark.parametrize("side_stream", [False, True])
@pytest.mark.parametrize("pinned_host_images", [False, True])
def test_receive_keyframes_on_a_busy_stream(side_stream, pinned_host_images):
    """`receive_keyframes` (north_star's add_keyframe, batched) on 64 640x480 images, ResNet-18
    trunk, 200k-row pool, with the stream busy beforehand: pool rows == descriptors bit for bit,
    intra-robot matches == the oracle front end fed the same descriptors one by one."""
    import torch
    from cslam_b200 import msgs as M
    from oracle.nns import NNSOracle
    dev = torch.device("cuda", 0)
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        glcd, node, net, params = _wrapper(dev)
        rows = 200000
        g = torch.Generator(device=dev).manual_seed(5)
        # pool = descriptors of the same distribution the network produces for synthetic
        # keyframes would need 200k forward passes; uniform rows plus 512 perturbed copies of
        # the batch's own descriptors (so that the threshold 0.9 has something to accept)
        imgs = torch.randint(0, 256, (64, 480, 640, 3), generator=g, device=dev, dtype=torch.uint8)
        ref_emb = net.compute_embeddings_device(imgs)
        torch.cuda.synchronize()
        pool = glcd.lcm.local_nnsm
        for s in range(0, rows, 50000):
            x = torch.rand((50000, DIM), generator=g, device=dev) - 0.5
            if s == 0:
                x[1000:1512] = ref_emb.repeat(8, 1) + 0.01 * torch.randn((512, DIM), generator=g, device=dev)
            pool.add_items_device(torch.nn.functional.normalize(x), range(s, s + 50000))
        torch.cuda.synchronize()
        local_matches = []
        node.create_subscription(None, 'cslam/local_keyframe_match', local_matches.append)
        kf_ids = list(range(rows, rows + 64))
        images = imgs.cpu().numpy() if pinned_host_images else imgs
        stream = torch.cuda.Stream(dev) if side_stream else torch.cuda.default_stream(dev)
        with torch.cuda.stream(stream):
            _busy(dev)
            glcd.receive_keyframes([M.KeyframeRGB(id=k, image=images[b]) for b, k in enumerate(kf_ids)])
        torch.cuda.synchronize()
        emb_h = ref_emb.cpu().numpy()
        assert np.array_equal(pool.read_rows(rows, 64), emb_h), \
            "pool rows differ from compute_embeddings of the same images"
        buffered = np.stack([np.asarray(glcd.global_descriptors_buffer[k].descriptor) for k in kf_ids])
        assert np.array_equal(buffered.astype(np.float32), emb_h)

        # oracle: the reference's per-keyframe sequence on the same descriptors
        # (detect_intra BEFORE the keyframe is added; lcsm.py:74-92)
        orc = NNSOracle(DIM)
        orc.data, orc.n = pool.read_rows(0, rows + 64), rows + 64
        expect = {}
        for b, kf in enumerate(kf_ids):
            full = orc.similarities_vec(emb_h[b])[:rows + b]   # the pool as it was before keyframe b
            order = np.argsort(full)[::-1][:K]
            for r in order:
                if abs(int(r) - kf) < params['frontend.intra_loop_min_inbetween_keyframes']:
                    continue
                if full[r] < params['frontend.similarity_threshold']:
                    continue
                expect[kf] = int(r)
                break
        assert len(expect) >= 32, "the scenario should produce intra-robot matches"
        assert {m.keyframe0_id: m.keyframe1_id for m in local_matches} == expect
    finally:
        torch.backends.cudnn.deterministic = det
