"""LoopClosureSparseMatching — routes global descriptors to the per-robot descriptor
pools, turns top-1 matches above the similarity threshold into inter-robot candidate
edges and hands them to the algebraic-connectivity selector.  Same API as the reference
(cslam/loop_closure_sparse_matching.py:12-110); pools and searches live on the GPU.
"""
import numpy as np

from .algebraic_connectivity_maximization import (AlgebraicConnectivityMaximization,
                                                  EdgeInterRobot)
from .nns_matching import NearestNeighborsMatching


class LoopClosureSparseMatching(object):
    """Sparse matching for loop closure detection."""

    def __init__(self, params):
        """
        Args:
            params (dict): the reference's flat ROS 2 parameter dict
        """
        self.params = params
        # reference :21-31: Scan Context matchers for lidar, cosine matchers otherwise
        self._lidar = self.params["frontend.sensor_type"] == "lidar"
        if self._lidar:
            from .lidar_pr.scancontext_matching import ScanContextMatching as matcher
        else:
            matcher = NearestNeighborsMatching
        self.local_nnsm = matcher()
        self.other_robots_nnsm = {}
        for i in range(self.params['max_nb_robots']):
            if i != self.params['robot_id']:
                self.other_robots_nnsm[i] = matcher()
        self.candidate_selector = AlgebraicConnectivityMaximization(
            self.params['robot_id'], self.params['max_nb_robots'], extra_params=self.params)

    def add_local_global_descriptor(self, embedding, keyframe_id):
        """Add a local keyframe; match it against every other robot's pool (reference :36-54).

        Returns:
            list(EdgeInterRobot): new candidate edges
        """
        matches = []
        self.local_nnsm.add_item(embedding, keyframe_id)
        for i in range(self.params['max_nb_robots']):
            if i == self.params['robot_id']:
                continue
            kf, similarity = self.other_robots_nnsm[i].search_best(embedding)
            if kf is not None and similarity >= self.params['frontend.similarity_threshold']:
                match = EdgeInterRobot(self.params['robot_id'], keyframe_id, i, kf, similarity)
                self.candidate_selector.add_match(match)
                matches.append(match)
        return matches

    def add_local_global_descriptors(self, embeddings, keyframe_ids):
        """Batched form of `add_local_global_descriptor` for a [B, D] float32 CUDA tensor
        (or array) of descriptors of consecutive local keyframes.  Identical result to B
        sequential calls: a local keyframe is only matched against the OTHER robots' pools,
        which do not change while the batch is added.  One top-1 search per other robot for
        the whole batch, one device->host copy of the [R-1, B] results.

        Returns:
            list(EdgeInterRobot): new candidate edges, keyframe-major like the reference
        """
        if self._lidar:   # Scan Context pools: the reference's one-by-one routing (same result by definition)
            rows = embeddings.detach().cpu().numpy() if hasattr(embeddings, "detach") else np.asarray(embeddings)
            return [m for row, k in zip(rows, keyframe_ids)
                    for m in self.add_local_global_descriptor(row, int(k))]
        import torch
        keyframe_ids = [int(k) for k in keyframe_ids]
        if not torch.is_tensor(embeddings):
            embeddings = torch.as_tensor(np.asarray(embeddings, dtype=np.float32))
        if not embeddings.is_cuda:
            embeddings = embeddings.to(torch.device("cuda", self.local_nnsm._resolve_device()))
        embeddings = embeddings.float().contiguous()
        self.local_nnsm.add_items_device(embeddings, keyframe_ids)
        robots = [i for i in range(self.params['max_nb_robots'])
                  if i != self.params['robot_id'] and self.other_robots_nnsm[i].n > 0]
        if not robots:
            return []
        B = embeddings.shape[0]
        idx = torch.empty((len(robots), B, 1), dtype=torch.int32, device=embeddings.device)
        sims = torch.empty((len(robots), B, 1), dtype=torch.float64, device=embeddings.device)
        for j, i in enumerate(robots):
            self.other_robots_nnsm[i].search_batch_device(embeddings, 1, out=(idx[j], sims[j]))
        idx_h = idx.cpu().numpy()[:, :, 0]
        sims_h = sims.cpu().numpy()[:, :, 0]
        thr = self.params['frontend.similarity_threshold']
        matches = []
        for b in range(B):
            for j, i in enumerate(robots):
                if sims_h[j, b] >= thr:
                    kf = self.other_robots_nnsm[i].items[int(idx_h[j, b])]
                    match = EdgeInterRobot(self.params['robot_id'], keyframe_ids[b], i, kf,
                                           float(sims_h[j, b]))
                    self.candidate_selector.add_match(match)
                    matches.append(match)
        return matches

    def add_other_robot_global_descriptors(self, msgs):
        """Batched form of `add_other_robot_global_descriptor` for descriptors of ONE other
        robot (a `GlobalDescriptors` message, reference
        global_descriptor_loop_closure_detection.py:407-422).  Remote descriptors are matched
        against the local pool, which does not change meanwhile, so one batched top-1
        search gives the sequential result.

        Returns:
            list(EdgeInterRobot or None): one entry per message, like the scalar method
        """
        if len(msgs) == 0:
            return []
        if self._lidar:
            return [self.add_other_robot_global_descriptor(m) for m in msgs]
        robot_id = msgs[0].robot_id
        assert all(m.robot_id == robot_id for m in msgs)
        desc = np.stack([np.asarray(m.descriptor) for m in msgs])
        self.other_robots_nnsm[robot_id].add_items(desc, [m.keyframe_id for m in msgs])
        if self.local_nnsm.n == 0:
            return [None] * len(msgs)
        idx, sims = self.local_nnsm.search_batch(desc, 1)
        out = []
        for b, m in enumerate(msgs):
            match = None
            if sims[b, 0] >= self.params['frontend.similarity_threshold']:
                match = EdgeInterRobot(self.params['robot_id'],
                                       self.local_nnsm.items[int(idx[b, 0])], robot_id,
                                       m.keyframe_id, float(sims[b, 0]))
                self.candidate_selector.add_match(match)
            out.append(match)
        return out

    def match_local_loop_closures_batch(self, embeddings, keyframe_ids, rows_before):
        """Intra-robot matches for a batch that `add_local_global_descriptors` has ALREADY
        appended as pool rows [rows_before, rows_before + B).  The reference matches keyframe
        b against the pool as it was just before b was added (detect_intra precedes
        add_local_global_descriptor, global_descriptor_loop_closure_detection.py:157-160):
        search k + B neighbours in the grown pool and drop, per query, the rows that were
        appended at or after it.

        Returns:
            list((kf_match or None, kfs or None)): per keyframe, as match_local_loop_closures
        """
        if self._lidar:
            raise NotImplementedError("batched intra-robot matching is written for the cosine pools; "
                                      "use match_local_loop_closures per keyframe with Scan Context")
        import torch
        k = self.params['frontend.nb_best_matches']
        B = len(keyframe_ids)
        if not torch.is_tensor(embeddings):
            embeddings = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)).to(
                torch.device("cuda", self.local_nnsm._resolve_device()))
        idx, sims = self.local_nnsm.search_batch_device(embeddings.float().contiguous(), k + B)
        idx, sims = idx.cpu().numpy(), sims.cpu().numpy()
        out = []
        for b in range(B):
            keep = (idx[b] >= 0) & (idx[b] < rows_before + b)
            kfs = [self.local_nnsm.items[int(r)] for r in idx[b][keep][:k]]
            out.append(self._pick_local_match(kfs, sims[b][keep][:k], keyframe_ids[b]))
        return out

    def add_other_robot_global_descriptor(self, msg):
        """Add another robot's keyframe descriptor; match it against the local pool
        (reference :56-72).  `msg` needs .robot_id, .keyframe_id, .descriptor."""
        descriptor = np.asarray(msg.descriptor)
        self.other_robots_nnsm[msg.robot_id].add_item(descriptor, msg.keyframe_id)
        match = None
        kf, similarity = self.local_nnsm.search_best(descriptor)
        if kf is not None and similarity >= self.params['frontend.similarity_threshold']:
            match = EdgeInterRobot(self.params['robot_id'], kf, msg.robot_id, msg.keyframe_id,
                                   similarity)
            self.candidate_selector.add_match(match)
        return match

    def match_local_loop_closures(self, descriptor, kf_id):
        """Intra-robot loop closure: best of the top-k local matches that is far enough in
        time and similar enough (reference :74-92)."""
        kfs, similarities = self.local_nnsm.search(descriptor,
                                                   k=self.params['frontend.nb_best_matches'])
        return self._pick_local_match(kfs, similarities, kf_id)

    def _pick_local_match(self, kfs, similarities, kf_id):
        if len(kfs) > 0 and kfs[0] == kf_id:
            kfs, similarities = kfs[1:], similarities[1:]
        if len(kfs) == 0 or kfs[0] is None:
            return None, None
        for kf, similarity in zip(kfs, similarities):
            if abs(kf - kf_id) < self.params['frontend.intra_loop_min_inbetween_keyframes']:
                continue
            if similarity < self.params['frontend.similarity_threshold']:
                continue
            return kf, kfs
        return None, None

    def select_candidates(self, number_of_candidates, is_neighbor_in_range,
                          greedy_initialization=True):
        """Select inter-robot loop closure candidates within the budget (reference :94-110)."""
        return self.candidate_selector.select_candidates(number_of_candidates,
                                                         is_neighbor_in_range,
                                                         greedy_initialization)
