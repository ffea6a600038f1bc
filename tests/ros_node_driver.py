"""Drive the UNMODIFIED reference ROS node (scripts/gvom_ros.py) callback by callback against a chosen
`gvom.Gvom` implementation, using the ROS stand-ins in tests/ros_stubs.  Returns everything it published."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def find_node_file():
    for cand in ("/root/reference/scripts/gvom_ros.py", os.path.join(ROOT, "baseline", "_ref", "gvom_ros.py")):
        if os.path.exists(cand):
            return cand
    return None


def run_node(gvom_class, frames, params=None):
    """frames: list of (cloud (N,3) sensor frame, ego xyz, yaw).  -> {topic: [payload, ...]}"""
    stubs = os.path.join(HERE, "ros_stubs")
    if stubs not in sys.path:
        sys.path.insert(0, stubs)
    import rospy
    import tf2_ros
    from nav_msgs.msg import Odometry
    from sensor_msgs.msg import PointCloud2
    rospy.PARAMS.clear(); rospy.PUBLISHERS.clear(); rospy.LOG.clear()
    rospy.PARAMS.update(params or {})
    mod = types.ModuleType("gvom")                       # what `import gvom` in the node resolves to
    mod.Gvom = gvom_class
    saved = sys.modules.get("gvom")
    sys.modules["gvom"] = mod
    try:
        spec = importlib.util.spec_from_file_location("gvom_ros_reference", find_node_file())
        node_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(node_mod)
        node = node_mod.VoxelMapper()
        for i, (pc, ego, yaw) in enumerate(frames):
            tf2_ros.Buffer.POSES[i] = (ego, (0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2))))
            node.cb_odom(Odometry(*ego))
            node.cb_lidar(PointCloud2(points=pc, stamp=i))
            node.cb_timer(None)
    finally:
        if saved is not None:
            sys.modules["gvom"] = saved
        else:
            sys.modules.pop("gvom", None)
    out = {}
    for topic, pub in rospy.PUBLISHERS.items():
        out[topic] = [m.data if hasattr(m, "data") else m for m in pub.messages]
    return out
