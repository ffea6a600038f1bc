"""The rows SURVEY.md 8(f) marks "next" -- the callers and data formats either side of the path -- through the
C-ABI on the GPU, each compared with the reference-side code it replaces:

  PointCloud2 ingestion   vs  ros_numpy's float64 conversion + Process_pointcloud       (gvom_ros.py:108-109)
  OccupancyGrid payloads  vs  the node's numpy post-processing                           (gvom_ros.py:142-164)
  asynchronous combine    vs  combine_maps
  state save / restore    vs  an uninterrupted run
  ROS-free replay driver  vs  the unmodified node driven through the ROS stand-ins
"""
import numpy as np
import pytest

import canon
from gvom_b200 import synth
from gvom_b200.node import PointCloud2Payload, VoxelMapperReplay
from host_grids import host_grids

pytestmark = pytest.mark.gpu

SMALLP = dict(xy_size=64, z_size=16, buffer_size=3, robot_radius=2.0)


def make(P, **kw):
    from gvom_b200 import Gvom
    return Gvom(*P, **kw)


def small_frames(n, nan=True):
    out = []
    for i in range(n):
        pc, ego, T = synth.frame(i, 16, 256, wall_radius=9.0, ego0=(10.0, 5.0, 1.0), dego=(0.9, 0.5, 0.25))
        pc = pc.copy()
        if nan:
            pc[7::97] = np.nan
            pc[11::131, 1] = np.inf
        out.append((pc, ego, T))
    return out


def same_state(a, b):
    da, db = canon.canon_scan(a.refview()), canon.canon_scan(b.refview())
    for k in ("n_occ", "codes_sha", "ids_sha", "hit_sha", "total_sha", "minh_sha"):
        assert da[k] == db[k], k
    assert np.allclose(da["metrics"], db["metrics"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("layout", [(16, (0, 4, 8)), (48, (0, 4, 8)), (32, (4, 12, 20)), (12, (0, 4, 8))])
@pytest.mark.parametrize("where", ["bytes", "pinned", "device"])
def test_pointcloud2_equals_xyz_array_path(layout, where):
    import torch
    step, offs = layout
    P = synth.params_tuple(**SMALLP)
    a, b = make(P), make(P)
    for pc, ego, T in small_frames(3):
        msg = PointCloud2Payload.from_xyz(pc, step, offs)
        a.Process_pointcloud(msg.to_xyz_array(), ego, T)                # what the node does (ros_numpy, float64)
        raw = msg.data
        if where == "bytes":
            data = raw.tobytes()
        elif where == "pinned":
            data = torch.from_numpy(raw.copy()).pin_memory()
        else:
            data = torch.from_numpy(raw.copy()).cuda()
        b.Process_pointcloud2(data, msg.n_points, step, ego, T, offs)
        same_state(a, b)
        oa, ob = a.combine_maps(), b.combine_maps()
        for x, y in zip(oa, ob):
            assert np.array_equal(x, y)


def test_pointcloud2_full_scan_chunked_and_without_transform():
    """OS1-128 sized payload (the chunked, pipelined staging path) and transform=None."""
    P = synth.params_tuple()
    pc, ego, T = synth.frame(0, 128, 2048)
    pc = pc.copy(); pc[5::1001] = np.nan
    msg = PointCloud2Payload.from_xyz(pc, 48, (0, 4, 8))
    a, b = make(P), make(P)
    a.Process_pointcloud(msg.to_xyz_array(), ego, T)
    b.Process_pointcloud2(msg.data.tobytes(), msg.n_points, 48, ego, T)
    da, db = canon.canon_scan(a.refview(), full=False), canon.canon_scan(b.refview(), full=False)
    for k in ("n_occ", "codes_sha", "ids_sha", "hit_sha", "total_sha", "minh_sha"):
        assert da[k] == db[k], k
    w = (msg.to_xyz_array() @ T[:3, :3].T + T[:3, 3]).astype(np.float32)    # pre-transformed, float32-exact
    msg2 = PointCloud2Payload.from_xyz(w, 16)
    a.Process_pointcloud(msg2.to_xyz_array(), ego, None)
    b.Process_pointcloud2(msg2.data, msg2.n_points, 16, ego, None)
    da, db = canon.canon_scan(a.refview(), full=False), canon.canon_scan(b.refview(), full=False)
    for k in ("n_occ", "codes_sha", "ids_sha", "hit_sha", "total_sha", "minh_sha"):
        assert da[k] == db[k], k


def test_pointcloud2_without_mapped_host_access(monkeypatch):
    """GVOM_H2D=dma: no zero-copy kernel reads; the PointCloud2 payload is widened to float64 x 3 by the staging
    threads and takes the chunked-DMA array path.  Same results."""
    monkeypatch.setenv("GVOM_H2D", "dma")
    P = synth.params_tuple(**SMALLP)
    a, b = make(P), make(P)
    for pc, ego, T in small_frames(3):
        msg = PointCloud2Payload.from_xyz(pc, 24, (0, 8, 16))
        a.Process_pointcloud(msg.to_xyz_array(), ego, T)
        b.Process_pointcloud2(msg.data.tobytes(), msg.n_points, 24, ego, T, (0, 8, 16))
        same_state(a, b)
        for x, y in zip(a.combine_maps(), b.combine_maps()):
            assert np.array_equal(x, y)


def test_pointcloud2_message_object_and_errors():
    class F:
        def __init__(self, name, offset, datatype=7):
            self.name, self.offset, self.datatype = name, offset, datatype

    class Msg:
        pass
    P = synth.params_tuple(**SMALLP)
    pc, ego, T = small_frames(1)[0]
    p = PointCloud2Payload.from_xyz(pc, 24, (4, 8, 16))
    m = Msg()
    m.fields, m.point_step, m.width, m.height, m.row_step = [F("x", 4), F("y", 8), F("intensity", 12), F("z", 16)], 24, pc.shape[0], 1, 24 * pc.shape[0]
    m.data, m.is_bigendian = p.data.tobytes(), False
    a, b = make(P), make(P)
    a.Process_pointcloud(p.to_xyz_array(), ego, T)
    b.process_pointcloud2_msg(m, ego, T)
    same_state(a, b)
    with pytest.raises(RuntimeError):
        b.Process_pointcloud2(p.data, p.n_points, 24, ego, T, (4, 8, 22))       # field runs past the record
    with pytest.raises(RuntimeError):
        b.Process_pointcloud2(p.data, p.n_points, 10, ego, T)                   # point_step not a multiple of 4
    with pytest.raises(ValueError):
        b.Process_pointcloud2(p.data[:100], p.n_points, 24, ego, T, (4, 8, 16))  # payload too short


@pytest.mark.parametrize("thr", [(50, -10, 0), (30.5, -6.0, 2.0), (0, -1, 1)])
def test_occupancy_grids_equal_the_nodes_numpy(thr):
    P = synth.params_tuple(**SMALLP)
    g = make(P)
    assert g.occupancy_grids() is None
    for pc, ego, T in small_frames(4, nan=False):
        g.Process_pointcloud(pc, ego, T)
        origin, pos, neg, rough, vis = g.combine_maps()
        want = host_grids(pos, neg, rough, vis, *thr)
        got = g.occupancy_grids(*thr)
        got_dev = g.occupancy_grids(*thr, device_outputs=True)
        for k, w in want.items():
            assert got[k].dtype == np.int8 and got[k].shape == w.shape
            assert np.array_equal(got[k], w), k
            assert np.array_equal(got_dev[k].cpu().numpy(), w), k
    # after a combine with device-resident outputs the library's own result block must be current too
    g.Process_pointcloud(*small_frames(5, nan=False)[4])
    o = g.combine_maps(device_outputs=True)
    want = host_grids(o[1].cpu().numpy(), o[2].cpu().numpy(), o[3].cpu().numpy(), o[4].cpu().numpy(), *thr)
    got = g.occupancy_grids(*thr)
    for k, w in want.items():
        assert np.array_equal(got[k], w), k


def test_fused_combine_grids_and_full_size():
    P = synth.params_tuple()
    a, b = make(P), make(P, pinned_outputs=False)
    assert b.combine_maps_grids() is None
    for i in range(3):
        pc, ego, T = synth.frame(i, 64, 1024)
        a.Process_pointcloud(pc, ego, T)
        b.Process_pointcloud(pc, ego, T)
        origin, pos, neg, rough, vis = a.combine_maps()
        want = host_grids(pos, neg, rough, vis, 50, -10, 0)
        o2, got = b.combine_maps_grids(50, -10, 0)
        assert np.array_equal(origin, o2)
        for k, w in want.items():
            assert np.array_equal(got[k], w), k
    # the state behind the fused call is the same combined map
    ca, cb = canon.canon_combine(a.refview(), a.combine_maps(), full=False), canon.canon_combine(b.refview(), b.combine_maps(), full=False)
    for k in ("n_occ", "codes_sha", "hit_sha", "total_sha", "minh_sha"):
        assert ca[k] == cb[k], k


def test_async_combine_matches_sync():
    P = synth.params_tuple(**SMALLP)
    a, b = make(P), make(P)
    assert b.combine_maps_async().result() is None
    fr = small_frames(5, nan=False)
    pend = None
    for i, (pc, ego, T) in enumerate(fr):
        a.Process_pointcloud(pc, ego, T)
        want = a.combine_maps()
        b.Process_pointcloud(pc, ego, T)
        pend = b.combine_maps_async()
        if i + 1 < len(fr):
            # the next scan is enqueued while the maps of this combine are still in flight ...
            a.Process_pointcloud(*fr[i + 1]); b.Process_pointcloud(*fr[i + 1])
        got = pend.result()
        assert pend.done()
        for x, y in zip(want, got):
            assert np.array_equal(x, y)
        if i + 1 < len(fr):
            # ... and then combined once more, so both objects see the same call sequence
            want2, got2 = a.combine_maps(), b.combine_maps_async().result()
            for x, y in zip(want2, got2):
                assert np.array_equal(x, y)
    # a pending combine is completed implicitly by any other call that needs its results
    b.Process_pointcloud(*fr[0]); a.Process_pointcloud(*fr[0])
    p2 = b.combine_maps_async()
    w = a.combine_maps()
    assert b.make_debug_voxel_map().shape == a.make_debug_voxel_map().shape
    for x, y in zip(w, p2.result()):
        assert np.array_equal(x, y)


def test_state_save_restore(tmp_path):
    P = synth.params_tuple(**SMALLP)
    fr = small_frames(6, nan=False)
    a = make(P)
    for pc, ego, T in fr[:3]:
        a.Process_pointcloud(pc, ego, T)
        a.combine_maps()
    blob = a.save_state(str(tmp_path / "state.npy"))
    assert blob.dtype == np.uint8 and blob.size > 64 * 64 * 16 * 4
    b = make(P)
    b.load_state(str(tmp_path / "state.npy"))
    # restored object answers the debug exports like the saved one ...
    for f in ("make_debug_voxel_map", "make_debug_height_map", "make_debug_inferred_height_map"):
        x, y = getattr(a, f)(), getattr(b, f)()
        if f == "make_debug_voxel_map":
            x, y = x[np.lexsort(x[:, :3].T[::-1])], y[np.lexsort(y[:, :3].T[::-1])]
        assert np.array_equal(x, y), f
    # ... and continues exactly like it (history-dependent: ring position, previous combined map, counts)
    for pc, ego, T in fr[3:]:
        a.Process_pointcloud(pc, ego, T); b.Process_pointcloud(pc, ego, T)
        oa, ob = a.combine_maps(), b.combine_maps()
        for x, y in zip(oa, ob):
            assert np.array_equal(x, y)
        ca, cb = canon.canon_combine(a.refview(), oa), canon.canon_combine(b.refview(), ob)
        for k in ("n_occ", "codes_sha", "ids_sha", "hit_sha", "total_sha", "minh_sha"):
            assert ca[k] == cb[k], k
        assert np.allclose(ca["metrics"], cb["metrics"], rtol=1e-5, atol=1e-7)
    # a state saved before any data, and parameter mismatch
    e = make(P)
    e.load_state(make(P).save_state())
    assert e.combine_maps() is None
    with pytest.raises(RuntimeError):
        make(synth.params_tuple(**dict(SMALLP, buffer_size=2))).load_state(blob)


def test_replay_driver_fused_matches_host_post_processing():
    params = dict(width=64, height=16, robot_radius=2.0, buffer_size=3, density_threshold=40)
    host = VoxelMapperReplay(fused=False, host_postprocess=host_grids, **params)
    dev = VoxelMapperReplay(fused=True, **params)
    assert host.cb_lidar(np.zeros((4, 3)), (0, 0, 0), (0, 0, 0, 1)) is False        # "no odom"
    assert dev.cb_timer() is None
    for i, (pc, ego, _) in enumerate(small_frames(4)):
        yaw = 0.05 * i
        rot = (0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2)))
        msg = PointCloud2Payload.from_xyz(pc, 32, (0, 4, 8))
        for n in (host, dev):
            n.cb_odom(ego)
            assert n.cb_lidar(msg, ego, rot)
        a, b = host.cb_timer(), dev.cb_timer()
        assert a["origin"] == b["origin"]
        for topic in ("~hard_obstacle_map", "~soft_obstacle_map", "~ground_certainty_map", "~all_ground_certainty_map",
                      "~negative_obstacle_map", "~roughness_map"):
            assert np.array_equal(a[topic], b[topic]), topic
        assert np.array_equal(a["~debug/height_map"], b["~debug/height_map"])


def test_replay_driver_matches_the_unmodified_node():
    from ros_node_driver import find_node_file, run_node
    if find_node_file() is None:
        pytest.skip("reference gvom_ros.py not available (baseline/_ref)")
    from gvom_b200.gvom import Gvom
    fr = []
    for i, (pc, ego, _) in enumerate(small_frames(4)):
        fr.append((pc, ego, 0.01 * i))
    want = run_node(Gvom, fr, {"~width": 64, "~height": 16, "~robot_radius": 2.0, "~buffer_size": 3})
    node = VoxelMapperReplay(fused=True, width=64, height=16, robot_radius=2.0, buffer_size=3)
    for i, (pc, ego, yaw) in enumerate(fr):
        node.cb_odom(ego)
        node.cb_lidar(PointCloud2Payload.from_xyz(pc, 16), ego, (0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2))))
        out = node.cb_timer()
        for topic in ("~hard_obstacle_map", "~soft_obstacle_map", "~ground_certainty_map", "~all_ground_certainty_map",
                      "~negative_obstacle_map", "~roughness_map"):
            assert np.array_equal(out[topic], want[topic][i]), (i, topic)


@pytest.mark.parametrize("mask", [2, 4, 64, 128, 256, 450])
def test_ab_switches_do_not_change_results(mask, monkeypatch):
    """The A/B switches kept for measurements (GVOM_VARIANT bits: generic merge kernel, bulk-copy pipeline build of the row
    merge, DMA outputs, F2I floor in the DDA, per-thread loads instead of the bulk copy of host clouds) stay parity-green."""
    import replay
    monkeypatch.setenv("GVOM_VARIANT", str(mask))
    for name in ("small_moving", "os1_64"):
        bad, _ = replay.replay(make, name, replay.golden(name), view=lambda g: g.refview(), what=f"variant {mask}")
        assert not bad, "\n".join(bad[:20])


def test_concurrent_callers_on_the_new_entry_points():
    """ROS use (README.md:49): a subscriber thread feeds PointCloud2 payloads while a timer thread takes grids,
    asynchronous maps and a state snapshot.  No errors, every result well formed."""
    import threading
    P = synth.params_tuple(**SMALLP)
    g = make(P)
    fr = [synth.frame(i, 16, 256, wall_radius=9.0, ego0=(10.0, 5.0, 1.0), dego=(0.3, 0.2, 0.05)) for i in range(30)]
    msgs = [PointCloud2Payload.from_xyz(f[0], 32) for f in fr]
    errors, grids, maps = [], [], []

    def feeder():
        try:
            for m, (_, ego, T) in zip(msgs, fr):
                g.Process_pointcloud2(m.data, m.n_points, 32, ego, T)
        except Exception as ex:                      # pragma: no cover
            errors.append(ex)

    def timer():
        try:
            for k in range(40):
                r = g.combine_maps_grids(40, -8, 1)
                if r is not None:
                    grids.append(r[1])
                p = g.combine_maps_async()
                if k % 7 == 0:
                    g.save_state()
                o = p.result()
                if o is not None:
                    maps.append(o)
        except Exception as ex:                      # pragma: no cover
            errors.append(ex)

    th = [threading.Thread(target=feeder), threading.Thread(target=timer)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    assert grids and maps
    assert all(set(np.unique(x["hard"])) <= {0, 100} and x["roughness"].dtype == np.int8 for x in grids)
    assert all(o[1].shape == (64, 64) and np.isfinite(o[3]).all() for o in maps)
