"""ROS-free replay driver: the callbacks of the reference ROS node without ROS.

Mirrors `VoxelMapper` of the reference's scripts/gvom_ros.py:14-190 -- same parameter names and defaults
(gvom_ros.py:23-41), same callback sequence:

  cb_odom   (gvom_ros.py:79-80)    remember the latest odometry position
  cb_lidar  (gvom_ros.py:82-109)   TF (translation + quaternion) -> 4x4 matrix, NaN filter, Process_pointcloud
  cb_timer  (gvom_ros.py:113-190)  combine_maps, thresholding, Fortran-order int8 OccupancyGrid payloads for the
                                   seven map topics, the three debug clouds

so that recorded (cloud, odom, TF) sequences can be replayed against any `Gvom`-compatible class, and the
"runs unchanged" claim of the drop-in class can be tested without a ROS installation.

With `fused=True` (B200 class only) the two host-side stages either side of the path move to the device:
PointCloud2 byte payloads are ingested directly (`Process_pointcloud2`, replacing ros_numpy's float64
conversion) and the OccupancyGrid payloads come from `combine_maps_grids` (the thresholding of
gvom_ros.py:142-164 as one kernel; 5 instead of 20 bytes per cell cross PCIe).
"""
import numpy as np

# topic -> key of the grid it carries (the node publishes the certainty grid on two topics, gvom_ros.py:153-155)
GRID_TOPICS = {
    "~hard_obstacle_map": "hard",
    "~soft_obstacle_map": "soft",
    "~ground_certainty_map": "certainty",
    "~all_ground_certainty_map": "certainty",
    "~negative_obstacle_map": "negative",
    "~roughness_map": "roughness",
}

DEFAULTS = {                                  # gvom_ros.py:23-41
    "odom_frame": "/camera_init", "xy_resolution": 0.40, "z_resolution": 0.2, "width": 256, "height": 64,
    "buffer_size": 4, "min_point_distance": 1.0, "positive_obstacle_threshold": 0.50,
    "negative_obstacle_threshold": 0.5, "density_threshold": 50, "slope_obsacle_threshold": 0.3,
    "min_roughness": -10, "max_roughness": 0, "robot_height": 2.0, "robot_radius": 4.0,
    "ground_to_lidar_height": 1.0, "freq": 10.0, "xy_eigen_dist": 1, "z_eigen_dist": 1,
}


def quaternion_matrix(q):
    """Homogeneous rotation matrix of the quaternion (x, y, z, w): the published algorithm of
    tf.transformations.quaternion_matrix, which TransformerROS.fromTranslationRotation uses (gvom_ros.py:106)."""
    q = np.array(q, dtype=np.float64)
    nq = float(np.dot(q, q))
    if nq < np.finfo(np.float64).eps * 4.0:
        return np.identity(4)
    q = q * np.sqrt(2.0 / nq)
    o = np.outer(q, q)
    return np.array([[1.0 - o[1, 1] - o[2, 2], o[0, 1] - o[2, 3], o[0, 2] + o[1, 3], 0.0],
                     [o[0, 1] + o[2, 3], 1.0 - o[0, 0] - o[2, 2], o[1, 2] - o[0, 3], 0.0],
                     [o[0, 2] - o[1, 3], o[1, 2] + o[0, 3], 1.0 - o[0, 0] - o[1, 1], 0.0],
                     [0.0, 0.0, 0.0, 1.0]])


def from_translation_rotation(translation, rotation):
    m = quaternion_matrix(rotation)
    m[:3, 3] = np.asarray(translation, dtype=np.float64)
    return m


class PointCloud2Payload:
    """The part of a sensor_msgs/PointCloud2 the path needs: byte payload + layout of the x, y, z fields."""

    def __init__(self, data, n_points, point_step, offsets=(0, 4, 8)):
        self.data, self.n_points, self.point_step, self.offsets = data, int(n_points), int(point_step), tuple(offsets)

    @classmethod
    def from_xyz(cls, xyz, point_step=16, offsets=(0, 4, 8)):
        """Pack an (N, 3) array into PointCloud2 records (float32 fields, remaining bytes zero)."""
        xyz = np.asarray(xyz)
        rec = np.zeros((xyz.shape[0], point_step), np.uint8)
        for k in range(3):
            rec[:, offsets[k]:offsets[k] + 4] = xyz[:, k].astype("<f4").reshape(-1, 1).view(np.uint8)
        return cls(rec.reshape(-1), xyz.shape[0], point_step, offsets)

    def to_xyz_array(self, remove_nans=True):
        """What ros_numpy.point_cloud2.pointcloud2_to_xyz_array returns: float64 (N', 3), non-finite rows dropped."""
        rec = np.frombuffer(self.data, np.uint8, self.n_points * self.point_step).reshape(self.n_points, self.point_step)
        pts = np.empty((self.n_points, 3), np.float64)
        for k in range(3):
            o = self.offsets[k]
            pts[:, k] = np.ascontiguousarray(rec[:, o:o + 4]).view("<f4")[:, 0]
        if remove_nans:
            pts = pts[np.isfinite(pts).all(axis=1)]
        return pts


class VoxelMapperReplay:
    """The reference node's VoxelMapper (gvom_ros.py:14-190) without ROS.

    gvom_class: any class with the Gvom API (default: the B200 class).  fused: use the device-side
    PointCloud2 ingestion and OccupancyGrid post-processing of the B200 class.  debug: also produce the
    three debug exports in cb_timer, as the node does."""

    def __init__(self, gvom_class=None, fused=False, debug=True, host_postprocess=None, **params):
        unknown = set(params) - set(DEFAULTS)
        if unknown:
            raise TypeError(f"unknown node parameter(s): {sorted(unknown)}")
        p = dict(DEFAULTS, **params)
        self.__dict__.update(p)
        self.fused, self.debug = bool(fused), bool(debug)
        self.host_postprocess = host_postprocess      # fused=False: the node's numpy OccupancyGrid post-processing
        self.odom_data = None
        if gvom_class is None:
            from .gvom import Gvom as gvom_class
        self.voxel_mapper = gvom_class(                   # gvom_ros.py:44-59, positional like the node
            self.xy_resolution, self.z_resolution, self.width, self.height, self.buffer_size,
            self.min_point_distance, self.positive_obstacle_threshold, self.negative_obstacle_threshold,
            self.slope_obsacle_threshold, self.robot_height, self.robot_radius, self.ground_to_lidar_height,
            self.xy_eigen_dist, self.z_eigen_dist)

    def cb_odom(self, position):
        self.odom_data = (position[0], position[1], position[2])

    def cb_lidar(self, cloud, translation, rotation):
        """cloud: (N, 3) array (sensor frame; NaN rows are dropped like ros_numpy does) or a PointCloud2Payload.
        translation / rotation (x, y, z, w): the TF lidar -> odom transform at the scan's stamp."""
        if self.odom_data is None:
            print("no odom")
            return False
        odom_data = self.odom_data
        tf_matrix = from_translation_rotation(translation, rotation)
        if isinstance(cloud, PointCloud2Payload):
            if self.fused:
                self.voxel_mapper.Process_pointcloud2(cloud.data, cloud.n_points, cloud.point_step, odom_data, tf_matrix,
                                                      cloud.offsets)
                return True
            pc = cloud.to_xyz_array()
        else:
            pc = np.asarray(cloud, dtype=np.float64)
            pc = pc[~np.isnan(pc).any(axis=1)]
        self.voxel_mapper.Process_pointcloud(pc, odom_data, tf_matrix)
        return True

    def cb_timer(self):
        """-> None when there is no data, else {"origin": (x, y), topic: int8 payload ..., debug topics ...}."""
        vm = self.voxel_mapper
        if self.fused:
            res = vm.combine_maps_grids(self.density_threshold, self.min_roughness, self.max_roughness)
            if res is None:
                return None
            map_origin, grids = res
            obs_map = None
        else:
            map_data = vm.combine_maps()
            if map_data is None:
                return None
            map_origin, obs_map, neg_map, rough_map, cert_map = map_data
            if self.host_postprocess is None:
                raise RuntimeError("VoxelMapperReplay(fused=False) needs host_postprocess= (the node's numpy "
                                   "post-processing lives with the tests: tests/host_grids.py)")
            grids = self.host_postprocess(obs_map, neg_map, rough_map, cert_map, self.density_threshold,
                                          self.min_roughness, self.max_roughness)
        out = {"origin": (float(map_origin[0]), float(map_origin[1])), "resolution": self.xy_resolution,
               "width": self.width, "frame_id": self.odom_frame}
        for topic, key in GRID_TOPICS.items():
            out[topic] = grids[key]
        if self.debug:
            out["~debug/voxel"] = vm.make_debug_voxel_map()
            out["~debug/height_map"] = vm.make_debug_height_map()
            out["~debug/inferred_height_map"] = vm.make_debug_inferred_height_map()
        return out
