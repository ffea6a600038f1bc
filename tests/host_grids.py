"""The reference node's host-side OccupancyGrid post-processing (gvom_ros.py:142-164) restated with numpy.
Test infrastructure: the checker for the device-side `k_occupancy_grids` / `Gvom.occupancy_grids`, and the
`host_postprocess` of the ROS-free replay driver when it is run against the CPU oracle."""
import numpy as np


def host_grids(obs_map, neg_map, rough_map, cert_map, density_threshold, min_roughness, max_roughness):
    """The node's host-side post-processing (gvom_ros.py:142-164), verbatim in behaviour: dict of int8 payloads."""
    out = {}
    out["hard"] = np.reshape(np.maximum(100 * (obs_map > density_threshold), neg_map), -1, order="F").astype(np.int8)
    out["soft"] = np.reshape(100 * (obs_map <= density_threshold) * (obs_map > 0), -1, order="F").astype(np.int8)
    out["certainty"] = np.reshape(cert_map * 100, -1, order="F").astype(np.int8)
    out["negative"] = np.reshape(neg_map, -1, order="F").astype(np.int8)
    r = ((np.maximum(np.minimum(rough_map, max_roughness), min_roughness) + min_roughness)
         / (max_roughness - min_roughness)) * 100
    out["roughness"] = np.reshape(r, -1, order="F").astype(np.int8)
    return out
