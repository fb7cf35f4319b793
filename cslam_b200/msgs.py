"""Message types of the loop-closure front end.

The reference exchanges ROS 2 messages of the (un-vendored) package
`cslam_common_interfaces` plus `diagnostic_msgs/KeyValue`; the field lists below are the ones
the reference reads and writes (cslam/global_descriptor_loop_closure_detection.py:164-167,
:213-214, :233-238, :276-278, :304-306, :334-337; src/front_end/rgbd_handler.cpp:523-531,
:566-568).  When ROS is installed the real message classes are used, otherwise these
plain-Python stand-ins with the same attribute names: every consumer in cslam_b200 is
duck-typed on the attributes only.
"""
from dataclasses import dataclass, field
from typing import Any, List

try:  # pragma: no cover - ROS is not present in the build container
    from cslam_common_interfaces.msg import (GlobalDescriptor, GlobalDescriptors,  # noqa: F401
                                             InterRobotLoopClosure, InterRobotMatch,
                                             InterRobotMatches, KeyframeRGB,
                                             LocalDescriptorsRequest, LocalKeyframeMatch)
    from diagnostic_msgs.msg import KeyValue  # noqa: F401
    HAVE_ROS_MSGS = True
except ImportError:
    HAVE_ROS_MSGS = False

    @dataclass
    class GlobalDescriptor:
        keyframe_id: int = 0
        robot_id: int = 0
        descriptor: Any = field(default_factory=list)   # float32[] on the wire

    @dataclass
    class GlobalDescriptors:
        descriptors: List[GlobalDescriptor] = field(default_factory=list)

    @dataclass
    class InterRobotMatch:
        robot0_id: int = 0
        robot0_keyframe_id: int = 0
        robot1_id: int = 0
        robot1_keyframe_id: int = 0
        weight: float = 0.0

    @dataclass
    class InterRobotMatches:
        robot_id: int = 0
        matches: List[InterRobotMatch] = field(default_factory=list)

    @dataclass
    class LocalKeyframeMatch:
        keyframe0_id: int = 0
        keyframe1_id: int = 0

    @dataclass
    class LocalDescriptorsRequest:
        keyframe_id: int = 0
        matches_robot_id: List[int] = field(default_factory=list)
        matches_keyframe_id: List[int] = field(default_factory=list)

    @dataclass
    class InterRobotLoopClosure:
        robot0_id: int = 0
        robot0_keyframe_id: int = 0
        robot1_id: int = 0
        robot1_keyframe_id: int = 0
        success: bool = False
        transform: Any = None

    @dataclass
    class KeyframeRGB:
        id: int = 0
        image: Any = None      # sensor_msgs/Image, or an HxWx3 uint8 array

    @dataclass
    class KeyValue:
        key: str = ""
        value: str = ""


try:  # pragma: no cover
    from cslam_common_interfaces.msg import RobotIdsAndOrigin  # noqa: F401
except ImportError:
    @dataclass
    class RobotIds:
        ids: List[int] = field(default_factory=list)

    @dataclass
    class RobotIdsAndOrigin:
        """cslam_common_interfaces/RobotIdsAndOrigin: robots in range and, per robot, the robot
        whose frame it currently expresses its estimates in (cslam/neighbors_manager.py:171-185)."""
        robots: RobotIds = field(default_factory=RobotIds)
        origins: RobotIds = field(default_factory=RobotIds)

try:  # pragma: no cover
    from std_msgs.msg import String  # noqa: F401
except ImportError:
    @dataclass
    class String:
        data: str = ""


@dataclass
class UInt32:
    """std_msgs/UInt32 stand-in (heartbeat payload = origin robot id)."""
    data: int = 0
