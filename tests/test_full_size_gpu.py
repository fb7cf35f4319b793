"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot
score a 1M-row pool or a 1M-candidate graph in test time): the tensor-core path against the
exact fp64 scan kernel, idempotence, sortedness; Frank-Wolfe cardinality, dual bound,
determinism and improvement over the greedy start."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config3_pool_1m_x_512_tensor_path_equals_exact_scan():
    import torch
    from cslam_b200.nns_matching import NearestNeighborsMatching
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    nn = NearestNeighborsMatching(device=0)
    for s in range(0, 1000000, 100000):
        x = torch.rand((100000, 512), generator=g, device=dev)
        nn.add_items_device(x / x.norm(dim=1, keepdim=True))
    assert nn.n == 1000000
    q = torch.rand((200, 512), generator=g, device=dev, dtype=torch.float64)   # two tiles -> cluster of 2
    q = q / q.norm(dim=1, keepdim=True)
    idx, sims = nn.search_batch_device(q, 30)
    assert nn.last_info[2] == 0                      # nobody needed the escalation path
    idx_b, sims_b = nn.search_batch_device(q, 30)    # idempotent
    assert torch.equal(idx, idx_b) and torch.equal(sims, sims_b)
    s = sims.cpu().numpy()
    assert np.all(s[:, :-1] >= s[:, 1:]) and np.all((s > 0.5) & (s <= 1.0))
    assert all(len(set(r)) == 30 for r in idx.cpu().numpy().tolist())
    # single-tile path gives the same answer as the cluster path
    idx_c, _ = nn.search_batch_device(q[:64], 30)
    assert torch.equal(idx[:64], idx_c)
    # exact fp64 scan kernel (no tensor cores, no thresholds) on a sample of the queries
    nn.set_mode(1)
    idx_e, sims_e = nn.search_batch_device(q[:6], 30)
    nn.set_mode(0)
    assert torch.equal(idx[:6], idx_e)
    assert float((sims[:6] - sims_e).abs().max()) == 0.0
    # a stored row is its own best match with similarity 1
    from cslam_b200 import _lib
    host = np.zeros((2, 512), dtype=np.float32)
    _lib.check(_lib.load().cslam_nns_read_rows(nn._h, 777777, 2, _lib.ptr(host)))
    i2, s2 = nn.search_batch_device(torch.from_numpy(host).to(dev), 1)
    assert i2.cpu().numpy()[:, 0].tolist() == [777777, 777778]
    assert np.allclose(s2.cpu().numpy()[:, 0], 1.0, atol=1e-6)


def test_config5_graph_100k_poses_1m_candidates_properties():
    from bench import greedy_w_init, mac_graph
    from cslam_b200.mac.mac import MAC
    fixed, cand, n = mac_graph(8, 12500, 1000000)
    k = 1000
    mac = MAC(fixed, cand, n)
    w0 = greedy_w_init(cand[2], k)
    sel, w, u = mac.fw_subset(w0.copy(), k, max_iters=20, trace=True)
    tsel, tf = mac.last_trace
    assert sel.sum() == k and set(np.unique(sel)) <= {0.0, 1.0}
    assert np.all(w >= 0) and abs(w.sum() - k) < 1e-6            # iterates stay in the k-simplex hull
    f_last = tf[mac.last_fw_iters - 1]
    assert u >= f_last - 1e-12                                   # dual bound above the objective
    for it in range(mac.last_fw_iters):
        assert len(set(tsel[it].tolist())) == k                  # every direction is a k-subset
    # deterministic
    sel2, w2, u2 = mac.fw_subset(w0.copy(), k, max_iters=20)
    assert np.array_equal(sel, sel2) and u == u2
    # (no claim that the rounded selection beats the greedy start: with 20 Frank-Wolfe iterations the
    # reference's own result is below it on such graphs - tools/probe_mac.py --oracle 1 shows the same
    # objective, 5.75e-6 vs 2.87e-5, from the reference restatement)
    f_sel, f_greedy = mac.evaluate_objective(sel), mac.evaluate_objective(w0)
    assert f_sel > 0 and f_greedy > 0 and abs(f_greedy - tf[0]) <= 1e-9 * f_greedy
    # gradient of a constant vector is zero, of the Fiedler vector non-negative
    lam, vec = mac.evaluate_fiedler_pair(sel)
    assert abs(np.linalg.norm(vec) - 1.0) < 1e-9 and abs(vec.sum()) < 1e-6
    grad = mac.grad_from_fiedler(vec)
    assert grad.min() >= 0 and np.all(mac.grad_from_fiedler(np.ones(n)) == 0)
