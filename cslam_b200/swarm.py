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
        packed[:, D] = torch.as_tensor(kf_ids, dtype=torch.float64, device=embeddings.device)
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
        new = torch.as_tensor(kf_ids, dtype=torch.int64, device=device)
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

        g_kf = g_kf.cpu().numpy().reshape(R, R, B, kx)                  # [pool, query robot, b, j]
        g_sims = g_sims.cpu().numpy().reshape(R, R, B, kx)
        all_ids = all_ids.cpu().numpy()
        thr = self.params['frontend.similarity_threshold']
        # best match of every descriptor in every OTHER robot's pool, in the order the reference
        # meets them (query robot, keyframe, pool robot); one bulk insert into the candidate table
        top_kf = g_kf[:, :, :, 0].transpose(1, 2, 0)                    # [query robot, b, pool]
        top_s = g_sims[:, :, :, 0].transpose(1, 2, 0)
        robots = np.arange(R)
        with np.errstate(invalid="ignore"):
            hit = (top_kf >= 0) & (top_s >= thr) & (robots[:, None, None] != robots[None, None, :])
        qq, bb, gg = np.nonzero(hit)
        m_kf0, m_kf1, m_s = all_ids[qq, bb], top_kf[qq, bb, gg], top_s[qq, bb, gg]
        self.candidate_selector.add_matches(qq, m_kf0, gg, m_kf1, m_s)
        edges = [EdgeInterRobot(*e) for e in zip(qq.tolist(), m_kf0.tolist(), gg.tolist(),
                                                 m_kf1.tolist(), m_s.tolist())]
        intra = []
        if intra_on:
            own_idx = idx[me * B:(me + 1) * B].cpu().numpy()
            own_kf = kf[me * B:(me + 1) * B].cpu().numpy()
            own_sims = sims[me * B:(me + 1) * B].cpu().numpy()
            for b in range(B):
                keep = (own_idx[b] >= 0) & (own_idx[b] < rows_before + b)
                intra.append((kf_ids[b], own_kf[b][keep][:k_intra].tolist(),
                              own_sims[b][keep][:k_intra].tolist()))
        self.last_all_ids = all_ids
        self.last_exchange = (g_kf, g_sims)
        return edges, intra

    def select_candidates(self, number_of_candidates, is_neighbor_in_range,
                          greedy_initialization=True):
        """Broker-side sparsification (reference loop_closure_sparse_matching.py:94-110);
        every rank holds the same candidate graph, the broker (lowest rank) calls this."""
        return self.candidate_selector.select_candidates(number_of_candidates,
                                                         is_neighbor_in_range,
                                                         greedy_initialization)
