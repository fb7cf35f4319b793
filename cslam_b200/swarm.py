"""Multi-robot matching with one robot per GPU (BASELINE.json configs[3]).

In the reference every robot process keeps a copy of every other robot's descriptor pool:
descriptors are broadcast on '/cslam/global_descriptors', each receiver matches them
against its own pool, and the resulting matches are broadcast again on
'/cslam/inter_robot_matches' so that the broker sees the whole candidate graph
(cslam/global_descriptor_loop_closure_detection.py:198-289, :407-433;
cslam/loop_closure_sparse_matching.py:36-72).  On one 8xB200 box the same dataflow is:

    rank r = robot r; its descriptor pool is resident in rank r's HBM only (no copies)
    1. all-gather the B new descriptors of every robot        (the GlobalDescriptors broadcast)
    2. every rank searches ITS pool for all R*B descriptors   (top-1 for the other robots'
       keyframes, top-`nb_best_matches` for its own = intra-robot loop closures)
    3. all-gather the per-shard [R*B, k] (similarity, keyframe id) results over NVLink
       (the InterRobotMatches broadcast) -> every rank holds the same candidate edges

The pool of robot g is searched exactly once per keyframe of robot q, where the reference
does the same search twice (on robot q against its copy of g's pool, and on robot g when q's
descriptor arrives).  Collectives are `torch.distributed` (NCCL on GPUs; gloo in the CPU
tests of this host logic); there is no collective on the single-robot path.
"""
import numpy as np

from .algebraic_connectivity_maximization import (AlgebraicConnectivityMaximization,
                                                  EdgeInterRobot)


class SwarmExchange(object):
    """The two collectives of the multi-robot path."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.bytes_gathered = 0

    def _all_gather(self, x):
        import torch
        x = x.contiguous()
        out = torch.empty(self.world * x.numel(), dtype=x.dtype, device=x.device)
        self.dist.all_gather_into_tensor(out, x.reshape(-1), group=self.group)
        self.bytes_gathered += out.numel() * out.element_size()
        return out.reshape((self.world,) + tuple(x.shape))

    def all_gather_descriptors(self, embeddings, kf_ids):
        """[B, D] float32 + [B] ids (same B on every rank) -> ([R, B, D], [R, B] int64)."""
        import torch
        B, D = embeddings.shape
        packed = torch.empty((B, D + 1), dtype=torch.float64, device=embeddings.device)
        packed[:, :D] = embeddings
        ids = torch.as_tensor(kf_ids, dtype=torch.float64)
        if embeddings.is_cuda:
            ids = ids.pin_memory().to(embeddings.device, non_blocking=True)
        packed[:, D] = ids
        out = self._all_gather(packed)
        return out[:, :, :D].float().contiguous(), out[:, :, D].long()

    def all_gather_topk(self, ids, sims):
        """per-shard [Q, k] (int ids, float64 sims) -> ([R, Q, k] int64, [R, Q, k] float64);
        one collective on a packed float64 buffer (ids < 2^53 are exact)."""
        import torch
        packed = torch.stack([sims.double(), ids.double()], dim=-1)
        out = self._all_gather(packed)
        return out[..., 1].long(), out[..., 0]


class SwarmLoopClosureMatching(object):
    """Rank-local robot of a swarm whose pools are sharded one robot per rank."""

    def __init__(self, params, exchange, pool=None, exchange_k=1):
        """
        Args:
            exchange_k (int): matches per (keyframe, pool) carried by the all-gather; the
                reference's semantics need 1 (top-1 per other robot), BASELINE configs[3]
                quotes 30
            params (dict): reference parameter dict; `robot_id` must equal the rank and
                `max_nb_robots` the world size
            exchange (SwarmExchange)
            pool: descriptor pool with add_items_device / search_batch_device / n / items
                (default: the GPU NearestNeighborsMatching)
        """
        assert params['robot_id'] == exchange.rank and params['max_nb_robots'] == exchange.world
        self.params = params
        self.exchange = exchange
        if pool is None:
            from .nns_matching import NearestNeighborsMatching
            pool = NearestNeighborsMatching()
        self.local_nnsm = pool
        self.exchange_k = int(exchange_k)
        self._row_ids = None       # device int64 [capacity]: pool row -> keyframe id
        self.candidate_selector = AlgebraicConnectivityMaximization(
            params['robot_id'], params['max_nb_robots'], extra_params=params)

    def _append_ids(self, kf_ids, device):
        import torch
        new = torch.as_tensor(kf_ids, dtype=torch.int64)
        if torch.device(device).type == "cuda":
            new = new.pin_memory().to(device, non_blocking=True)
        n = self.local_nnsm.n - len(kf_ids)
        if self._row_ids is None or self._row_ids.numel() < n + len(kf_ids):
            grown = torch.empty(max(1024, 2 * (n + len(kf_ids))), dtype=torch.int64, device=device)
            if n > 0:
                grown[:n] = self._row_ids[:n]
            self._row_ids = grown
        self._row_ids[n:n + len(kf_ids)] = new

    def step(self, embeddings, kf_ids):
        """One lock-step round: every robot contributes B new keyframe descriptors.

        Order on every rank: append the own batch to the own pool, then search the pool for
        all R*B descriptors in ONE call.  Other robots' keyframes get their top-`exchange_k`
        (column 0 decides the candidate edge, reference lcsm.py:45-53); the own keyframes get
        the intra-robot result of reference lcsm.py:74-92, i.e. keyframe b sees the pool as it
        was just before b was appended (rows appended at or after b are dropped from a
        k + B wide result).

        Returns:
            (new_edges, intra): candidate edges of this round (identical list on every
            rank, query-robot-major) and, for this rank's keyframes, a list of
            (kf_id, [local kf ids], [similarities]) with at most `nb_best_matches` entries.
        """
        import torch
        R, me = self.exchange.world, self.exchange.rank
        B = embeddings.shape[0]
        kf_ids = [int(i) for i in kf_ids]
        intra_on = bool(self.params.get('frontend.enable_intra_robot_loop_closures', False))
        k_intra = int(self.params['frontend.nb_best_matches'])
        kx = self.exchange_k
        all_emb, all_ids = self.exchange.all_gather_descriptors(embeddings, kf_ids)
        rows_before = self.local_nnsm.n
        self.local_nnsm.add_items_device(embeddings.float().contiguous(), kf_ids)
        self._append_ids(kf_ids, embeddings.device)
        k_search = min(max(kx, k_intra + B if intra_on else 0), self.local_nnsm.n)
        idx, sims = self.local_nnsm.search_batch_device(all_emb.reshape(R * B, -1), k_search)
        idx = idx.long()
        kf = torch.where(idx >= 0, self._row_ids[idx.clamp(min=0)], idx)
        x_kf, x_sims = kf[:, :kx], sims[:, :kx]
        if x_kf.shape[1] < kx:      # pool smaller than exchange_k: fixed message width
            pad = kx - x_kf.shape[1]
            x_kf = torch.cat([x_kf, x_kf.new_full((R * B, pad), -1)], dim=1)
            x_sims = torch.cat([x_sims, x_sims.new_full((R * B, pad), float('nan'))], dim=1)
        g_kf, g_sims = self.exchange.all_gather_topk(x_kf, x_sims)      # [R(pool), R*B, kx]
        thr = float(self.params['frontend.similarity_threshold'])
        # best match of every descriptor in every OTHER robot's pool, in the order the reference
        # meets them (query robot, keyframe, pool robot): thresholded and compacted on the
        # device, only the hits travel to the host; one bulk insert into the candidate table
        own = slice(me * B, (me + 1) * B)
        hits, intra_raw = self._filter_round(g_kf, g_sims, all_ids, thr, idx[own], kf[own], sims[own],
                                             rows_before, k_intra if intra_on else 0)
        qq = hits[:, 0].astype(np.int64)
        m_kf0, gg, m_kf1, m_s = (hits[:, 1].astype(np.int64), hits[:, 2].astype(np.int64),
                                 hits[:, 3].astype(np.int64), hits[:, 4])
        self.candidate_selector.add_matches(qq, m_kf0, gg, m_kf1, m_s)
        edges = [EdgeInterRobot(*e) for e in zip(qq.tolist(), m_kf0.tolist(), gg.tolist(),
                                                 m_kf1.tolist(), m_s.tolist())]
        intra = []
        if intra_on:
            for b in range(B):
                cnt = int(intra_raw[b, 0])
                intra.append((kf_ids[b], intra_raw[b, 1:1 + cnt].astype(np.int64).tolist(),
                              intra_raw[b, 1 + k_intra:1 + k_intra + cnt].tolist()))
        self._last = (all_ids, g_kf, g_sims, R, B, kx)
        return edges, intra

    # the exchange buffers of the last round, materialised on the host only when asked for
    @property
    def last_all_ids(self):
        return self._last[0].cpu().numpy()

    @property
    def last_exchange(self):
        _, g_kf, g_sims, R, B, kx = self._last
        return (g_kf.cpu().numpy().reshape(R, R, B, kx), g_sims.cpu().numpy().reshape(R, R, B, kx))

    def _filter_round(self, g_kf, g_sims, all_ids, thr, own_idx, own_kf, own_sims, rows_before, k_intra):
        """-> (hits float64 [n, 5] = (query robot, query kf, pool robot, matched kf, similarity),
        intra float64 [B, 1 + 2 k_intra] = (count, kfs, sims)) on the host.
        CUDA tensors: `cslam_swarm_hits` / `cslam_swarm_intra` on the current stream and ONE
        device->host copy through a pinned buffer.  CPU tensors (the gloo tests of the exchange
        logic, which run the collectives without a GPU): the same filters in numpy."""
        import torch
        R, RB, kx = g_kf.shape
        B = RB // R
        if not g_kf.is_cuda:
            return self._filter_round_host(g_kf, g_sims, all_ids, thr, own_idx, own_kf, own_sims,
                                           rows_before, k_intra)
        import ctypes
        from . import _lib
        lib = _lib.load()
        cap = R * B * max(R - 1, 1)
        n_hits = 1 + 5 * cap
        n_intra = B * (1 + 2 * k_intra) if k_intra > 0 else 0
        dev = g_kf.device
        buf = getattr(self, "_round_buf", None)
        if buf is None or buf[0].numel() < n_hits + n_intra or buf[0].device != dev:
            buf = (torch.empty(n_hits + n_intra, dtype=torch.float64, device=dev),
                   torch.empty(n_hits + n_intra, dtype=torch.float64).pin_memory())
            self._round_buf = buf
        d_out, h_out = buf
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        g_kf, g_sims, all_ids = g_kf.contiguous(), g_sims.contiguous(), all_ids.contiguous()
        _lib.check(lib.cslam_swarm_hits(R, B, kx, _lib.ptr(g_kf), _lib.ptr(g_sims), _lib.ptr(all_ids), thr,
                                        _lib.ptr(d_out), cap, stream))
        if k_intra > 0:
            own_idx, own_kf, own_sims = own_idx.contiguous(), own_kf.contiguous(), own_sims.contiguous()
            _lib.check(lib.cslam_swarm_intra(B, own_idx.shape[1], k_intra, int(rows_before), _lib.ptr(own_idx),
                                             _lib.ptr(own_kf), _lib.ptr(own_sims),
                                             ctypes.c_void_p(d_out.data_ptr() + 8 * n_hits), stream))
        h_out[:n_hits + n_intra].copy_(d_out[:n_hits + n_intra], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = h_out.numpy()
        n = int(host[0])
        hits = host[1:1 + 5 * n].reshape(n, 5).copy()
        intra = host[n_hits:n_hits + n_intra].reshape(B, 1 + 2 * k_intra).copy() if k_intra > 0 else None
        return hits, intra

    @staticmethod
    def _filter_round_host(g_kf, g_sims, all_ids, thr, own_idx, own_kf, own_sims, rows_before, k_intra):
        R, RB, kx = g_kf.shape
        B = RB // R
        kf = g_kf.numpy().reshape(R, R, B, kx)[:, :, :, 0].transpose(1, 2, 0)      # [query robot, b, pool]
        sm = g_sims.numpy().reshape(R, R, B, kx)[:, :, :, 0].transpose(1, 2, 0)
        ids = all_ids.numpy()
        robots = np.arange(R)
        with np.errstate(invalid="ignore"):
            hit = (kf >= 0) & (sm >= thr) & (robots[:, None, None] != robots[None, None, :])
        qq, bb, gg = np.nonzero(hit)
        hits = np.stack([qq, ids[qq, bb], gg, kf[qq, bb, gg], sm[qq, bb, gg]], axis=1).astype(np.float64) \
            if len(qq) else np.zeros((0, 5))
        intra = None
        if k_intra > 0:
            oi, ok, os_ = own_idx.numpy(), own_kf.numpy(), own_sims.numpy()
            intra = np.zeros((B, 1 + 2 * k_intra))
            for b in range(B):
                keep = (oi[b] >= 0) & (oi[b] < rows_before + b)
                kk, ss = ok[b][keep][:k_intra], os_[b][keep][:k_intra]
                intra[b, 0] = len(kk)
                intra[b, 1:1 + len(kk)] = kk
                intra[b, 1 + k_intra:1 + k_intra + len(kk)] = ss
        return hits, intra

    def select_candidates(self, number_of_candidates, is_neighbor_in_range,
                          greedy_initialization=True):
        """Broker-side sparsification (reference loop_closure_sparse_matching.py:94-110);
        every rank holds the same candidate graph, the broker (lowest rank) calls this."""
        return self.candidate_selector.select_candidates(number_of_candidates,
                                                         is_neighbor_in_range,
                                                         greedy_initialization)
