"""Stand-in for ros_numpy.point_cloud2: the fake PointCloud2 carries its points as an (N,3) array."""
import numpy as np


def pointcloud2_to_xyz_array(msg, remove_nans=True):
    pts = np.asarray(msg.points, dtype=np.float64)      # ros_numpy returns float64 xyz
    if remove_nans:
        pts = pts[~np.isnan(pts).any(axis=1)]
    return pts


def array_to_pointcloud2(rec, stamp=None, frame_id=None):
    return {"cloud": rec, "frame_id": frame_id}
