"""CPU checks of the ROS-free replay driver (gvom_b200/node.py) against the unmodified
reference node run through the ROS stand-ins, both over the CPU oracle; and of the PointCloud2 payload helper."""
import numpy as np
import pytest

from gvom_b200 import synth
from gvom_b200.node import GRID_TOPICS, PointCloud2Payload, VoxelMapperReplay, from_translation_rotation
from host_grids import host_grids


def frames(n):
    out = []
    for i in range(n):
        pc, ego, _ = synth.frame(i, 8, 128, wall_radius=5.0, ego0=(10.0, 5.0, 1.0), dego=(0.9, 0.5, 0.25))
        pc = pc.copy()
        pc[7::97] = np.nan
        out.append((pc, ego, 0.01 * i))
    return out


def test_payload_round_trip_and_nan_filter():
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(100, 3)).astype(np.float32).astype(np.float64)
    xyz[3] = np.nan; xyz[10, 2] = np.inf
    for step, offs in ((16, (0, 4, 8)), (48, (0, 4, 8)), (28, (8, 16, 24))):
        p = PointCloud2Payload.from_xyz(xyz, step, offs)
        assert p.data.size == 100 * step
        back = p.to_xyz_array()
        keep = np.isfinite(xyz).all(axis=1)
        assert back.dtype == np.float64 and np.array_equal(back, xyz[keep])
        assert p.to_xyz_array(remove_nans=False).shape == (100, 3)


def test_transform_matches_tf_stand_in():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ros_stubs"))
    import tf
    q = (0.1, -0.2, 0.3, 0.9)
    assert np.array_equal(from_translation_rotation((1, 2, 3), q), tf.TransformerROS().fromTranslationRotation((1, 2, 3), q))


def test_host_grids_numpy_semantics():
    pos = np.array([[0, 50], [51, 100]], np.int32); neg = np.array([[100, 0], [0, 0]], np.int32)
    vis = np.array([[1, 0], [1, 1]], np.int32); rough = np.array([[-1.0, -25.0], [3.0, -4.5]])
    g = host_grids(pos, neg, rough, vis, 50, -10, 0)
    assert g["hard"].tolist() == [100, 100, 0, 100] and g["soft"].tolist() == [0, 0, 100, 0]      # Fortran order
    assert g["certainty"].tolist() == [100, 100, 0, 100] and g["negative"].tolist() == [100, 0, 0, 0]
    last = int(((-4.5 + -10) / (0 - -10)) * 100)                 # plain float64 arithmetic, truncated toward zero
    wrap = lambda v: ((v & 0xff) ^ 0x80) - 0x80                  # astype(int8) keeps the low byte: -200 -> 56, -145 -> 111
    assert g["roughness"].tolist() == [-110, -100, 56, wrap(last)]
    assert all(v.dtype == np.int8 for v in g.values())


def test_replay_driver_matches_unmodified_node_on_the_oracle():
    from ros_node_driver import find_node_file, run_node
    if find_node_file() is None:
        pytest.skip("reference gvom_ros.py not available (baseline/_ref)")
    from oracle.gvom_oracle import OracleGvom
    fr = frames(3)
    want = run_node(OracleGvom, fr, {"~width": 32, "~height": 16, "~robot_radius": 2.0, "~buffer_size": 2})
    node = VoxelMapperReplay(OracleGvom, fused=False, host_postprocess=host_grids, width=32, height=16, robot_radius=2.0, buffer_size=2)
    assert node.cb_timer() is None
    for i, (pc, ego, yaw) in enumerate(fr):
        node.cb_odom(ego)
        assert node.cb_lidar(pc, ego, (0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2))))
        out = node.cb_timer()
        for topic in GRID_TOPICS:
            assert np.array_equal(out[topic], want[topic][i]), (i, topic)
        assert np.array_equal(out["~debug/height_map"][:, 2], want["~debug/height_map"][i]["cloud"]["z"])
