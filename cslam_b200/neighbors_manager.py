"""Neighbour bookkeeping for the loop-closure front end: which robots are in communication
range, who is the broker, and from which keyframe / match index the next broadcast must
start so that no neighbour misses data.  Same classes and methods as the reference
(cslam/neighbors_manager.py:8-185, cslam/neighbor_monitor.py:4-53), host code, duck-typed on
the node handle (rclpy node or `cslam_b200.local_node.LocalNode`).
"""
from time import time

from .msgs import RobotIdsAndOrigin, String, UInt32


class NeighborMonitor(object):
    """Liveness of one neighbouring robot from its heartbeat topic."""

    def __init__(self, node, rid, is_enabled, init_delay_sec, max_delay_sec):
        self.node = node
        self.robot_id = rid
        self.is_enabled = is_enabled
        self.origin_robot_id = rid
        self.init_delay_sec = init_delay_sec
        self.max_delay_sec = max_delay_sec
        self.first_heartbeat_received = False
        self.init_time = time()
        self.latest_time_stamp = self.init_time
        self.last_keyframe_received = -1
        self.last_keyframe_sent = -1
        self.last_match_sent = -1
        try:
            from std_msgs.msg import UInt32 as RosUInt32
        except ImportError:
            RosUInt32 = UInt32
        self.heartbeat_subscriber = node.create_subscription(
            RosUInt32, '/r' + str(rid) + '/cslam/heartbeat', self.heartbeat_callback, 10)

    def heartbeat_callback(self, msg):
        self.origin_robot_id = msg.data
        self.latest_time_stamp = time()
        if not self.first_heartbeat_received:
            self.first_heartbeat_received = True
            self.init_time = time()

    def is_alive(self):
        """True when heartbeats are recent.  With monitoring disabled the reference falls off
        the end of the function and returns None, i.e. "not alive" to every caller
        (neighbor_monitor.py:46-53); reproduced."""
        if self.is_enabled:
            now = time()
            return (self.first_heartbeat_received and now - self.init_time > self.init_delay_sec
                    and now - self.latest_time_stamp < self.max_delay_sec)
        return None


class NeighborManager(object):
    """Keeps track of which other robots are in communication range."""

    def __init__(self, node, params):
        self.node = node
        self.params = params
        self.robot_id = params['robot_id']
        self.max_nb_robots = params['max_nb_robots']
        self.neighbors_monitors = {}
        for rid in range(self.max_nb_robots):
            if rid != self.robot_id:
                self.neighbors_monitors[rid] = NeighborMonitor(
                    node, rid, params['neighbor_management.enable_neighbor_monitoring'],
                    params['neighbor_management.init_delay_sec'],
                    params['neighbor_management.max_heartbeat_delay_sec'])
        # the pose-graph back end asks who is in range before it starts an optimisation
        # (src/back_end/decentralized_pgo.cpp:134-141) and waits for the answer (:25-29)
        self.subscriber = node.create_subscription(
            String, 'cslam/get_current_neighbors', self.get_current_neighbors_callback, 100)
        self.neighbors_publisher = node.create_publisher(
            RobotIdsAndOrigin, 'cslam/current_neighbors', 100)

    def _alive(self):
        return [rid for rid, m in self.neighbors_monitors.items() if m.is_alive()]

    def check_neighbors_in_range(self):
        """-> ({robot: in range?}, [robots in range]); the local robot always is."""
        alive = set(self._alive())
        flags = {i: (i == self.robot_id or i in alive) for i in range(self.max_nb_robots)}
        return flags, [i for i in range(self.max_nb_robots) if flags[i]]

    def local_robot_is_broker(self):
        """The lowest id among the robots in range is the broker (neighbors_manager.py:48-64)."""
        return all(self.robot_id < rid for rid in self._alive())

    def _send_window(self, latest, attr):
        alive = self._alive()
        start = latest
        for rid in alive:
            start = min(getattr(self.neighbors_monitors[rid], attr), start)
        for rid in alive:
            setattr(self.neighbors_monitors[rid], attr, latest)
        return start + 1

    def select_from_which_kf_to_send(self, latest_local_id):
        """First keyframe id some neighbour in range has not been sent yet (:66-86)."""
        return self._send_window(latest_local_id, 'last_keyframe_sent')

    def select_from_which_match_to_send(self, latest_local_match_idx):
        """Same for the inter-robot match buffer (:88-108)."""
        return self._send_window(latest_local_match_idx, 'last_match_sent')

    def useless_descriptors(self, last_kf_id):
        """Descriptors below the returned id were sent to every robot (:110-122)."""
        return min([last_kf_id] + [m.last_keyframe_sent for m in self.neighbors_monitors.values()])

    def useless_matches(self, last_match_id):
        return min([last_match_id] + [m.last_match_sent for m in self.neighbors_monitors.values()])

    def update_received_kf_id(self, other_robot_id, kf_id):
        self.neighbors_monitors[other_robot_id].last_keyframe_received = kf_id

    def get_unknown_range(self, descriptors):
        """Indices of the descriptors of a `GlobalDescriptors` message that are newer than
        what was already received from that robot (:147-169)."""
        other = descriptors[0].robot_id
        mon = self.neighbors_monitors[other]
        fresh = [i for i, d in enumerate(descriptors) if d.keyframe_id > mon.last_keyframe_received]
        mon.last_keyframe_received = max(mon.last_keyframe_received,
                                         max(d.keyframe_id for d in descriptors))
        return fresh

    def get_current_neighbors_callback(self, msg):
        """Publish the robots currently in range (without the local one) and the origin robot
        each of them reports in its heartbeat (neighbors_manager.py:171-185)."""
        _, in_range = self.check_neighbors_in_range()
        in_range.remove(self.robot_id)
        out = RobotIdsAndOrigin()
        out.robots.ids = list(in_range)
        for i in in_range:
            out.origins.ids.append(self.neighbors_monitors[i].origin_robot_id)
        self.neighbors_publisher.publish(out)
