"""Synthetic lidar scans and replay scenarios (test / bench input generation).

Pure numpy, no CUDA.  Shared by the golden-vector generator
(tests/golden/make_golden.py, which drives the *reference*), the parity tests
(which drive the oracle and the CUDA path) and bench.py, so that all of them
see bit-identical inputs.

The scan model is the one SURVEY.md section 8(d) defines (OS1-like elevation
fan over rolling ground inside a cylinder wall).  Two refinements make it
bit-reproducible across hosts: every coordinate is quantised to 1/1024 m
(real OS1 data is float32; this also makes the float64 cloud float32-exact),
and the yaw rotation entries are quantised to 2**-24.
"""
import hashlib

import numpy as np

# ctor arguments of gvom_ros.py:24-41 (defaults), in the positional order of
# gvom.py:21-22.  buffer_size is slot 4.
DEFAULT_PARAMS = dict(
    xy_resolution=0.4, z_resolution=0.2, xy_size=256, z_size=64, buffer_size=4,
    min_distance=1.0, positive_obstacle_threshold=0.5,
    negative_obstacle_threshold=0.5, slope_obsacle_threshold=0.3,
    robot_height=2.0, robot_radius=4.0, ground_to_lidar_height=1.0,
    xy_eigen_dist=1, z_eigen_dist=1)

PARAM_ORDER = ("xy_resolution", "z_resolution", "xy_size", "z_size",
               "buffer_size", "min_distance", "positive_obstacle_threshold",
               "negative_obstacle_threshold", "slope_obsacle_threshold",
               "robot_height", "robot_radius", "ground_to_lidar_height",
               "xy_eigen_dist", "z_eigen_dist")


def params_tuple(**over):
    p = dict(DEFAULT_PARAMS)
    p.update(over)
    return tuple(p[k] for k in PARAM_ORDER)


def _q(a, q=1024.0):
    return np.round(np.asarray(a, dtype=np.float64) * q) / q


def synthetic_scan(beams, cols, seed, ego, wall_radius=45.0, lidar_height=1.0,
                   fov_deg=22.5):
    """OS1-like scan in the SENSOR frame: float64 (beams*cols, 3), every beam returns.

    Elevations uniform in [-fov, +fov]; azimuths uniform in [0, 2pi).  Downward
    beams hit rolling ground z = 0.3 sin(0.15 X) cos(0.1 Y) (world X, Y; two
    fixed-point iterations, lidar `lidar_height` above z = 0), everything is
    capped by a cylinder wall of radius `wall_radius`; range noise N(0, 0.02 m)
    from default_rng(seed).  Coordinates quantised to 1/1024 m.
    """
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(-fov_deg, fov_deg, beams))[:, None]
    az = (np.arange(cols) * (2 * np.pi / cols))[None, :]
    ce, se = np.cos(el), np.sin(el)
    r = np.broadcast_to(wall_radius / ce, (beams, cols)).copy()
    down = np.broadcast_to(se < 0, (beams, cols))
    rg = np.where(down, lidar_height / np.maximum(-se, 1e-9), 1e6)
    for _ in range(2):
        X = ego[0] + rg * ce * np.cos(az)
        Y = ego[1] + rg * ce * np.sin(az)
        zg = 0.3 * np.sin(0.15 * X) * np.cos(0.1 * Y)
        rg = np.where(down, (lidar_height - zg) / np.maximum(-se, 1e-9), 1e6)
    r = np.minimum(r, rg)
    r = r + rng.normal(0.0, 0.02, size=r.shape)
    pts = np.stack([r * ce * np.cos(az), r * ce * np.sin(az), r * se], axis=-1)
    return np.ascontiguousarray(_q(pts.reshape(-1, 3)))


def pose_matrix(ego, yaw):
    """4x4 float64 sensor->world transform: yaw about z, translation = ego."""
    T = np.eye(4)
    c, s = _q(np.cos(yaw), 2.0 ** 24), _q(np.sin(yaw), 2.0 ** 24)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[:3, 3] = ego
    return T


def frame(i, beams, cols, wall_radius=45.0, seed_base=0, ego0=(100.0, 50.0, 1.0),
          dego=(0.4, 0.1, 0.0), dyaw=0.01):
    """Frame i of the SURVEY 8(d) stream: (cloud_sensor_frame, ego, T)."""
    ego = tuple(float(_q(ego0[k] + dego[k] * i)) for k in range(3))
    T = pose_matrix(ego, dyaw * i)
    pc = synthetic_scan(beams, cols, seed=seed_base + i, ego=ego, wall_radius=wall_radius)
    return pc, ego, T


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


# --------------------------------------------------------------------------
# Replay scenarios.  A scenario is (ctor params, list of steps); a step is
#   ("scan", cloud (N,3|4) ndarray, ego 3-tuple, T 4x4 or None)
#   ("combine",)           -> record the 5-tuple and the retained state
#   ("debug",)             -> record the three make_debug_* arrays
# --------------------------------------------------------------------------

def _apply(T, pc):
    return np.ascontiguousarray(_q(pc @ T[:3, :3].T + T[:3, 3]))


def scenario(name):
    """Return (params_tuple, steps) for a named scenario."""
    if name == "tiny":
        # CUDASIM-sized (reference simulator: minutes).  Origin moves between scans.
        P = params_tuple(xy_size=16, z_size=8, buffer_size=2, robot_radius=1.1)  # 1.1: no exact tie in the ego disc test
        steps = []
        for i in range(3):
            pc, ego, T = frame(i, 4, 24, wall_radius=2.6, ego0=(10.0, 5.0, 1.0),
                               dego=(0.5, 0.45, 0.25))
            steps += [("scan", pc, ego, T), ("combine",)]
        steps += [("combine",), ("debug",)]
        return P, steps
    if name == "small_moving":
        # 64x64x16 grid, ring buffer wraps (6 scans, B=4), origin shifts in x, y and z,
        # combine after every scan, twice at the end (count re-accumulation).
        P = params_tuple(xy_size=64, z_size=16, buffer_size=4, robot_radius=2.0)
        steps = []
        for i in range(6):
            pc, ego, T = frame(i, 16, 256, wall_radius=9.0, ego0=(10.0, 5.0, 1.0),
                               dego=(0.9, 0.5, 0.25))
            steps += [("scan", pc, ego, T), ("combine",)]
        steps += [("combine",), ("debug",)]
        return P, steps
    if name == "small_quirks":
        # ego near the WORLD origin (min-distance test is world-frame, gvom.py:1145-1149),
        # transform=None with pre-transformed float32 cloud, a zero-length ray, points on
        # exact voxel boundaries, (N,3) float32 input, combines not after every scan.
        P = params_tuple(xy_size=64, z_size=16, buffer_size=3, robot_radius=2.0,
                         xy_eigen_dist=1, z_eigen_dist=1)
        steps = []
        for i in range(4):
            pc, ego, T = frame(i, 16, 256, wall_radius=8.0, seed_base=100,
                               ego0=(0.3, -0.2, 1.0), dego=(0.45, 0.3, -0.2), dyaw=0.2)
            w = _apply(T, pc)
            extra = np.array([ego,                      # zero-length ray
                              (ego[0] + 0.4, ego[1], ego[2]),
                              (0.8, 0.4, 0.2), (-0.4, 0.0, 1.0),   # on voxel faces
                              (0.5, 0.5, 0.5),                      # < min_distance of world origin
                              (ego[0] + 20.0, ego[1], ego[2]),      # leaves the grid in +x
                              (ego[0], ego[1] - 30.0, ego[2] + 0.1),
                              (ego[0] + 1.0, ego[1] + 1.0, ego[2] + 5.0)], dtype=np.float64)
            w = np.concatenate([w, _q(extra)], axis=0)
            if i % 2 == 0:
                steps.append(("scan", np.ascontiguousarray(w, dtype=np.float32), ego, None))
            else:
                steps.append(("scan", w, ego, None))
            if i != 1:
                steps.append(("combine",))
        steps += [("debug",)]
        return P, steps
    if name == "small_eigen2":
        # wider moment neighbourhood (xy_eigen_dist=2, z_eigen_dist=0) and anisotropic sizes
        P = params_tuple(xy_size=48, z_size=24, buffer_size=2, robot_radius=1.5,
                         xy_eigen_dist=2, z_eigen_dist=0, z_resolution=0.4)
        steps = []
        for i in range(3):
            pc, ego, T = frame(i, 16, 128, wall_radius=7.0, seed_base=200,
                               ego0=(-20.0, 33.0, 2.0), dego=(-0.6, 0.7, 0.5), dyaw=-0.1)
            steps += [("scan", pc, ego, T), ("combine",)]
        steps += [("debug",)]
        return P, steps
    if name == "os1_64":
        # BASELINE.json configs[0] at full size: 64x1024 points, 256x256x64 grid, B=4
        P = params_tuple()
        steps = []
        for i in range(3):
            pc, ego, T = frame(i, 64, 1024)
            steps += [("scan", pc, ego, T), ("combine",)]
        return P, steps
    if name == "os1_128":
        # BASELINE.json configs[1]: 128x2048 points, same grid; ring buffer wraps
        P = params_tuple()
        steps = []
        for i in range(6):
            pc, ego, T = frame(i, 128, 2048)
            steps += [("scan", pc, ego, T), ("combine",)]
        steps += [("debug",)]
        return P, steps
    if name == "long_range":
        # BASELINE.json configs[4], primary reading: wall 200 m, B=16, same grid
        P = params_tuple(buffer_size=16)
        steps = []
        for i in range(3):
            pc, ego, T = frame(i, 128, 2048, wall_radius=200.0)
            steps += [("scan", pc, ego, T), ("combine",)]
        return P, steps
    if name == "dense":
        # BASELINE.json configs[3]: 2M-point aggregated cloud (16 x 128x1024 scans), 1024x1024x128 grid @0.1 m
        P = params_tuple(xy_resolution=0.1, z_resolution=0.1, xy_size=1024, z_size=128, buffer_size=4)
        steps = []
        for i in range(2):
            ego = (100.0 + 0.4 * i, 50.0 + 0.1 * i, 1.0)
            T = pose_matrix(ego, 0.01 * i)
            pc = np.concatenate([synthetic_scan(128, 1024, seed=5000 + 16 * i + k, ego=ego) for k in range(16)], axis=0)
            steps += [("scan", np.ascontiguousarray(pc), ego, T), ("combine",)]
        return P, steps
    raise KeyError(name)


SMALL_SCENARIOS = ("tiny", "small_moving", "small_quirks", "small_eigen2")
FULL_SCENARIOS = ("os1_64", "os1_128", "long_range")
