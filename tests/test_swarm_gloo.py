"""world_size-2/3 gloo tests (CPU) of the multi-robot host logic in cslam_b200/swarm.py:
the two all-gathers, the edge construction and the intra-robot filter.  The GPU pool is
replaced by a CPU pool backed by the oracle, so only the host/collective logic is under
test here; the GPU pool itself is covered by tests/test_nns_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OraclePool(object):
    """CPU stand-in with the device-pool interface (add_items_device/search_batch_device)."""

    def __init__(self):
        from oracle.nns import NNSOracle
        self.o = NNSOracle()
        self.items = self.o.items

    @property
    def n(self):
        return self.o.n

    def add_items_device(self, rows, items):
        for r, it in zip(rows.numpy(), items):
            self.o.add_item(r, it)

    def search_batch_device(self, queries, k):
        k = min(k, self.o.n)
        idx = np.full((queries.shape[0], k), -1, dtype=np.int32)
        sims = np.full((queries.shape[0], k), np.nan)
        for i, q in enumerate(queries.numpy()):
            full = self.o.similarities_vec(q)
            order = np.argsort(-full, kind="stable")[:k]
            idx[i], sims[i] = order, full[order]
        return torch.from_numpy(idx), torch.from_numpy(sims)


def _stream(world, rounds, B, dim, seed=3):
    rng = np.random.default_rng(seed)
    places = rng.random((6, dim))
    out = []
    for t in range(rounds):
        x = places[rng.integers(0, 6, (world, B))] + 0.3 * rng.random((world, B, dim))
        x /= np.linalg.norm(x, axis=2, keepdims=True)
        out.append(x.astype(np.float32))
    return out


def _params(rank, world):
    return {'robot_id': rank, 'max_nb_robots': world, 'frontend.similarity_threshold': 0.85,
            'frontend.nb_best_matches': 4, 'frontend.enable_intra_robot_loop_closures': True,
            'frontend.enable_sparsification': True, 'frontend.sensor_type': 'stereo',
            'evaluation.enable_sparsification_comparison': False}


def _expected(world, stream, B, thr, k_intra):
    """Sequential restatement: pools as numpy arrays, cosine in float64."""
    pools = [np.zeros((0, stream[0].shape[2]), np.float32) for _ in range(world)]
    edges, intra = [], [[] for _ in range(world)]
    for t, x in enumerate(stream):
        before = [len(p) for p in pools]
        pools = [np.concatenate([pools[r], x[r]]) for r in range(world)]
        for q in range(world):
            for b in range(B):
                d = x[q, b].astype(np.float64)
                for g in range(world):
                    P = pools[g].astype(np.float64)
                    s = (P @ d) / np.sqrt((d @ d) * np.einsum("ij,ij->i", pools[g], pools[g]).astype(np.float64))
                    if g == q:
                        vis = s[:before[g] + b]
                        order = np.argsort(-vis, kind="stable")[:k_intra]
                        intra[q].append((t * B + b, order.tolist()))
                    else:
                        j = int(np.argmax(s))
                        if s[j] >= thr:
                            edges.append((q, t * B + b, g, j, float(s[j])))
    return edges, intra


def _worker(rank, world, port, B, rounds, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cslam_b200.swarm import SwarmExchange, SwarmLoopClosureMatching
        ex = SwarmExchange()
        sw = SwarmLoopClosureMatching(_params(rank, world), ex, pool=OraclePool())
        stream = _stream(world, rounds, B, 16)
        edges, intra = [], []
        for t, x in enumerate(stream):
            e, i = sw.step(torch.from_numpy(x[rank]), list(range(t * B, (t + 1) * B)))
            edges.extend(tuple(v) for v in e)
            intra.extend((kf, ids) for kf, ids, _ in i)
        ret[rank] = (edges, intra, sorted(sw.candidate_selector.candidate_edges.keys()),
                     ex.bytes_gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_swarm_step_over_gloo(world):
    B, rounds = 4, 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, rounds, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    exp_edges, exp_intra = _expected(world, _stream(world, rounds, B, 16), B, 0.85, 4)
    assert len(exp_edges) > 0
    for r in range(world):
        edges, intra, keys, nbytes = ret[r]
        assert [e[:4] for e in edges] == [e[:4] for e in exp_edges]          # same list on every rank
        np.testing.assert_allclose([e[4] for e in edges], [e[4] for e in exp_edges], atol=1e-9)
        assert intra == exp_intra[r]
        assert keys == ret[0][2]
        # descriptors [R,B,D+1] f64 + top-1 pairs [R,R*B,1,2] f64 per round
        assert nbytes == rounds * (world * B * 17 * 8 + world * world * B * 2 * 8)
