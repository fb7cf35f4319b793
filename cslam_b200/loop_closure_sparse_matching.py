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
        if self.params["frontend.sensor_type"] == "lidar":
            # reference :21-22,28-29 uses ScanContextMatching here; the lidar modality is
            # outside the scope of this GPU front end
            raise NotImplementedError("cslam_b200 covers the visual (global descriptor) path; "
                                      "sensor_type 'lidar' is not supported")
        self.local_nnsm = NearestNeighborsMatching()
        self.other_robots_nnsm = {}
        for i in range(self.params['max_nb_robots']):
            if i != self.params['robot_id']:
                self.other_robots_nnsm[i] = NearestNeighborsMatching()
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
