"""In-process stand-in for the ROS 2 node handle the reference classes receive
(`rclpy.node.Node`: cslam/loop_closure_detection_node.py:15-101).

The loop-closure front end only needs six node services — create_publisher,
create_subscription, create_timer, get_logger, get_parameter, declare_parameters — and a
transport that delivers a published message to the subscribers of the same topic.  `LocalBus`
is that transport inside one process (the reference uses DDS between processes; on the
B200 box the robots of a swarm are ranks/handles of one job), `LocalNode` the handle.  A
real `rclpy` node can be passed to every cslam_b200 class instead: nothing below is
imported by them.
"""
import logging
import time


class LocalBus(object):
    """Topic name -> subscriber callbacks.  Delivery is synchronous and in subscription
    order, which makes multi-robot tests deterministic."""

    def __init__(self):
        self.subscribers = {}
        self.published = {}   # topic -> count (for tests / logs)

    def subscribe(self, topic, callback):
        self.subscribers.setdefault(topic, []).append(callback)

    def publish(self, topic, msg):
        self.published[topic] = self.published.get(topic, 0) + 1
        for cb in list(self.subscribers.get(topic, [])):
            cb(msg)


class _Publisher(object):
    def __init__(self, bus, topic):
        self.bus, self.topic = bus, topic

    def publish(self, msg):
        self.bus.publish(self.topic, msg)


class _Timer(object):
    def __init__(self, period, callback):
        self.period, self.callback = float(period), callback
        self.next_time = time.monotonic() + self.period
        self.cancelled = False

    def cancel(self):
        self.cancelled = True


class _Parameter(object):
    def __init__(self, value):
        self.value = value


class LocalNode(object):
    """Duck-typed node handle bound to a `LocalBus`.  Relative topic names are resolved in
    the node's namespace ('/r<robot_id>' in Swarm-SLAM launch files)."""

    def __init__(self, bus=None, namespace="", parameters=None, name="loop_closure_detection"):
        self.bus = bus if bus is not None else LocalBus()
        self.namespace = namespace.rstrip("/")
        self.parameters = dict(parameters or {})
        self.timers = []
        self.logger = logging.getLogger(f"cslam_b200{self.namespace.replace('/', '.')}.{name}")

    def resolve(self, topic):
        return topic if topic.startswith("/") else f"{self.namespace}/{topic}"

    def create_publisher(self, msg_type, topic, qos=10):
        return _Publisher(self.bus, self.resolve(topic))

    def create_subscription(self, msg_type, topic, callback, qos=10):
        self.bus.subscribe(self.resolve(topic), callback)
        return (self.resolve(topic), callback)

    def create_timer(self, period, callback, clock=None):
        t = _Timer(period, callback)
        self.timers.append(t)
        return t

    def get_logger(self):
        return self.logger

    def declare_parameters(self, namespace="", parameters=()):
        for name, default in parameters:
            self.parameters.setdefault(name, default)

    def get_parameter(self, name):
        return _Parameter(self.parameters[name])

    def spin_once(self, now=None, force=False):
        """Run the timers that are due (all of them when `force`)."""
        now = time.monotonic() if now is None else now
        for t in self.timers:
            if not t.cancelled and (force or now >= t.next_time):
                t.next_time = now + t.period
                t.callback()
