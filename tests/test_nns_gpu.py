"""GPU parity tests for NearestNeighborsMatching (through the C ABI) against
the numpy oracle (oracle/nns.py, pinned to cslam/nns_matching.py)."""
import numpy as np
import pytest

from oracle.nns import NNSOracle, lists_match_modulo_ties

pytestmark = pytest.mark.gpu


def _unit(rng, n, d, dtype=np.float64):
    x = rng.random((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def _build(pool, items=None):
    from cslam_b200.nns_matching import NearestNeighborsMatching
    gpu = NearestNeighborsMatching()
    orc = NNSOracle()
    items = list(range(len(pool))) if items is None else items
    gpu.add_items(pool, items)
    orc.add_items(pool, items)
    return gpu, orc


def _check_queries(gpu, orc, queries, k, sim_tol=1e-6):
    idx, sims = gpu.search_batch(queries, k)
    for qi, q in enumerate(queries):
        full = orc.similarities_vec(q)
        order = np.argsort(full)[::-1][:k]
        assert lists_match_modulo_ties(list(idx[qi]), list(order), full), (qi, idx[qi], order)
        # similarity values: north_star tolerance is 1e-3; we hold 1e-6
        np.testing.assert_allclose(sims[qi], full[idx[qi]], rtol=0, atol=sim_tol)
        assert np.all(sims[qi][:-1] >= sims[qi][1:])
    return idx, sims


def test_reference_test_similarity_shape():
    # tests/test_sparse_matching.py:51-81 (100 x 100-d, k=100) through the reference API
    from cslam_b200.nns_matching import NearestNeighborsMatching
    rng = np.random.default_rng(0)
    nnsm = NearestNeighborsMatching()
    pool = _unit(rng, 100, 100)
    for i in range(100):
        nnsm.add_item(pool[i], i)
    assert nnsm.n == 100 and nnsm.data.shape == (1000, 100)
    np.testing.assert_allclose(nnsm.data[:100], pool.astype(np.float32))
    for _ in range(20):
        query = _unit(rng, 1, 100)[0]
        ds = np.linalg.norm(query[np.newaxis, :] - nnsm.data[:nnsm.n], axis=1)
        ns_dist = np.argsort(ds)[:100]
        ns_sim, sims = nnsm.search(query, 100)
        assert np.all(sims[:-1] >= sims[1:])
        for j in range(100):
            if ns_dist[j] != ns_sim[j]:
                assert abs(ds[ns_dist[j]] - ds[ns_sim[j]]) < 1e-6
        best, sim_best = nnsm.search_best(query)
        assert best == ns_sim[0] and sim_best == sims[0]


@pytest.mark.parametrize("k", [1, 10, 30])
def test_config1_1k_by_4096(k):
    # BASELINE.json configs[0]: 1k random 4096-d descriptors
    rng = np.random.default_rng(1234)
    pool = _unit(rng, 1000, 4096)
    queries = _unit(rng, 256, 4096)
    gpu, orc = _build(pool, items=[3 * i for i in range(1000)])
    idx, _ = _check_queries(gpu, orc, queries, k)
    assert gpu.last_info[2] == 0, "tensor-core path escalated to the exact scan"
    items, sims = gpu.search(queries[0], k)
    ref_items, ref_sims = orc.search_loop(queries[0], k)
    assert items == ref_items
    np.testing.assert_allclose(sims, ref_sims, atol=1e-9)


@pytest.mark.parametrize("n,d", [(1, 10), (7, 10), (255, 64), (257, 100), (5000, 512), (40000, 128)])
def test_ragged_shapes(n, d):
    rng = np.random.default_rng(n * 1000 + d)
    pool = _unit(rng, n, d)
    queries = _unit(rng, 9, d)
    gpu, orc = _build(pool)
    _check_queries(gpu, orc, queries, min(30, n))
    # k larger than the pool returns the whole pool, ranked
    idx, sims = gpu.search_batch(queries[:2], n + 5 if n < 1000 else 1024)
    assert idx.shape[1] == min(n, n + 5 if n < 1000 else 1024)


def test_sampled_threshold_path_large_pool():
    # > 2*sample rows => sampled tau + filtered full pass
    rng = np.random.default_rng(5)
    pool = _unit(rng, 70000, 192, np.float32)
    queries = _unit(rng, 130, 192)  # two query tiles
    gpu, orc = _build(pool)
    _check_queries(gpu, orc, queries, 30)
    assert gpu.last_info[2] == 0


def test_exact_scan_mode_agrees():
    rng = np.random.default_rng(6)
    pool = _unit(rng, 3000, 256)
    queries = _unit(rng, 16, 256)
    gpu, orc = _build(pool)
    a_idx, a_sims = gpu.search_batch(queries, 30)
    gpu.set_mode(1)
    b_idx, b_sims = gpu.search_batch(queries, 30)
    assert np.array_equal(a_idx, b_idx)
    assert np.array_equal(a_sims, b_sims)
    _check_queries(gpu, orc, queries, 30)


def test_float32_queries_and_unnormalised_rows():
    rng = np.random.default_rng(7)
    pool = (rng.random((2000, 512)) * 3.0).astype(np.float32)   # not unit norm
    queries = (rng.random((8, 512)) * 0.5).astype(np.float32)
    gpu, orc = _build(pool)
    idx, sims = gpu.search_batch(queries, 30)
    for qi, q in enumerate(queries):
        full = orc.similarities_vec(q.astype(np.float64))
        order = np.argsort(full)[::-1][:30]
        assert lists_match_modulo_ties(list(idx[qi]), list(order), full)
        np.testing.assert_allclose(sims[qi], full[idx[qi]], atol=1e-6)


def test_duplicates_tie_order_and_clustered_pool():
    rng = np.random.default_rng(8)
    base = _unit(rng, 50, 64)
    pool = np.concatenate([base, base[:10], base[:10]])  # exact duplicates
    gpu, orc = _build(pool)
    q = base[3]
    items, sims = gpu.search(q, 5)
    # three identical rows (3, 53, 63): we return exact ties by descending row id; the
    # reference's np.argsort(...)[::-1] leaves their order unspecified (observed [53, 63, 3])
    assert items[:3] == [63, 53, 3]
    ref_items, ref_sims = orc.search_loop(q, 5)
    assert sorted(items[:3]) == sorted(ref_items[:3]) and items[3:] == ref_items[3:]
    np.testing.assert_allclose(sims, ref_sims, atol=1e-9)
    # near-duplicate cluster larger than the re-rank window still resolves exactly
    centre = _unit(rng, 1, 64)[0]
    cluster = centre[None, :] + 1e-4 * rng.standard_normal((600, 64))
    gpu2, orc2 = _build(np.concatenate([cluster, _unit(rng, 400, 64)]))
    _check_queries(gpu2, orc2, centre[None, :], 30)


def test_empty_and_growth():
    from cslam_b200.nns_matching import NearestNeighborsMatching
    nnsm = NearestNeighborsMatching()
    assert nnsm.search(np.ones(4), 3) == ([], [])
    assert nnsm.search_best(np.ones(4)) == (None, None)
    rng = np.random.default_rng(9)
    pool = _unit(rng, 2100, 32)
    orc = NNSOracle()
    for i in range(2100):
        nnsm.add_item(pool[i], i)
        orc.add_item(pool[i], i)
        if i in (0, 999, 1000, 2099):
            items, sims = nnsm.search(pool[0], 3)
            ref_items, ref_sims = orc.search_loop(pool[0], 3)
            assert items == ref_items
            np.testing.assert_allclose(sims, ref_sims, atol=1e-9)
    assert nnsm.data.shape == (4000, 32)


@pytest.mark.parametrize("nq", [129, 256, 300, 512, 700])
def test_query_groups_run_as_clusters(nq):
    """Batches wider than one 128-query tile take the thread-block-cluster path (2 or 4 CTAs
    share every pool tile through TMA multicast; > 512 queries = several groups)."""
    rng = np.random.default_rng(40 + nq)
    pool = _unit(rng, 70000, 192, np.float32)     # sampled threshold + filtered full pass
    gpu, orc = _build(pool)
    queries = _unit(rng, nq, 192)
    idx, sims = gpu.search_batch(queries, 30)
    assert gpu.last_info[0] + gpu.last_info[2] == nq
    for qi in list(range(0, nq, 37)) + [127, 128, nq - 1]:
        full = orc.similarities_vec(queries[qi])
        order = np.argsort(full)[::-1][:30]
        assert lists_match_modulo_ties(list(idx[qi]), list(order), full), qi
        np.testing.assert_allclose(sims[qi], full[idx[qi]], rtol=0, atol=1e-6)
    # identical to the same queries searched one tile at a time
    idx1 = np.concatenate([gpu.search_batch(queries[s:s + 128], 30)[0] for s in range(0, nq, 128)])
    assert np.array_equal(idx, idx1)
    # small pool: unfiltered (exhaustive) pass, one pool tile per cluster
    small, orc2 = _build(_unit(rng, 3000, 192, np.float32))
    idx2, sims2 = small.search_batch(queries, 5)
    for qi in (0, 128, nq - 1):
        full = orc2.similarities_vec(queries[qi])
        assert lists_match_modulo_ties(list(idx2[qi]), list(np.argsort(full)[::-1][:5]), full)
