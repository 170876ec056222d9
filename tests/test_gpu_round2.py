"""GPU tier, round 2: the host-side machinery added around the kernels -- vectorised patch stores,
uint16 ids, begin / end casts with pinned staging and CUDA-graph replay, the "latest wins" pose
mirror, peer buffers -- each checked against the oracle or against the plain path."""
import ctypes as C
import os

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import _capi, engine, scenes, shard
from mujoco_ros2_simulation_b200.model import Poses

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clutter():
    return scenes.clutter500(n_el=64, n_az=256)


@pytest.fixture()
def gpu_scene(clutter):
    s = engine.Scene(clutter.model, nworld=1)
    s.update_poses(clutter.poses)
    yield s
    s.close()


def test_vector_and_scalar_result_stores_agree(oracle, clutter, gpu_scene):
    """K2b writes a group's results as float4 / int4 lines (B2_VEC_OUT, default) or one scalar per
    lane: same bytes, and both equal the oracle (1 M rays -> the two-kernel cast)."""
    import torch
    cs = clutter
    lid = engine.Lidar(gpu_scene, cs.lidar)
    nscan = 64
    pos, mat = cs.sensor_poses(nscan, seed=5)
    dev = torch.device("cuda", 0)
    d_pos, d_mat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
    outs = {}
    for flag in ("1", "0"):
        os.environ["B2_VEC_OUT"] = flag
        r = torch.full((nscan, lid.nray), -9.0, dtype=torch.float32, device=dev)
        g = torch.full((nscan, lid.nray), -9, dtype=torch.int32, device=dev)
        lid.cast_device(nscan, d_pos, d_mat, None, ranges=r, geomid=g)
        gpu_scene.sync()
        outs[flag] = (r.cpu().numpy(), g.cpu().numpy())
    os.environ.pop("B2_VEC_OUT")
    assert np.array_equal(outs["1"][0], outs["0"][0]) and np.array_equal(outs["1"][1], outs["0"][1])
    o = oracle.OracleScene(cs.model, cs.poses)
    for i in (0, 17, nscan - 1):
        rr, gg, _ = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=4)
        assert np.array_equal(outs["1"][0][i], rr) and np.array_equal(outs["1"][1][i], gg)
    # a ranges-only and a geomid-only launch take the same path with one array missing
    r = torch.full((nscan, lid.nray), -9.0, dtype=torch.float32, device=dev)
    lid.cast_device(nscan, d_pos, d_mat, None, ranges=r)
    g = torch.full((nscan, lid.nray), -9, dtype=torch.int32, device=dev)
    lid.cast_device(nscan, d_pos, d_mat, None, geomid=g)
    gpu_scene.sync()
    assert np.array_equal(r.cpu().numpy(), outs["1"][0]) and np.array_equal(g.cpu().numpy(), outs["1"][1])
    lid.close()


def test_uint16_geom_ids(clutter, gpu_scene):
    import torch
    cs = clutter
    lid = engine.Lidar(gpu_scene, cs.lidar)
    pos, mat = cs.sensor_poses(70, seed=9)
    ref = lid.cast(pos, mat, want=("ranges", "geomid"))
    out = lid.cast_u16(pos, mat)                        # bulk path (> 1 M rays), pageable buffers
    assert np.array_equal(out["ranges"], ref["ranges"])
    assert np.array_equal(out["geomid16"].astype(np.int32), np.where(ref["geomid"] < 0, 0xFFFF, ref["geomid"]))
    assert (out["geomid16"] == 0xFFFF).any() and (out["geomid16"] != 0xFFFF).any()
    small = lid.cast_u16(pos[:2], mat[:2])              # small path
    assert np.array_equal(small["geomid16"], out["geomid16"][:2])
    dev = torch.device("cuda", 0)
    r = torch.empty((70, lid.nray), dtype=torch.float32, device=dev)
    g16 = torch.empty((70, lid.nray), dtype=torch.uint16, device=dev)
    lid.cast_u16_device(70, torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev), None, r, g16)
    gpu_scene.sync()
    assert np.array_equal(g16.cpu().numpy(), out["geomid16"])
    lid.close()


def test_begin_end_pinned_and_pageable_and_graph_replay(oracle, clutter, gpu_scene):
    """b2_lidar_cast = begin + end.  Single scans go through pinned staging and, from the third call
    of a shape on, replay a captured CUDA graph: every replay must see ITS poses."""
    import torch
    cs = clutter
    lid = engine.Lidar(gpu_scene, cs.lidar)
    o = oracle.OracleScene(cs.model, cs.poses)
    pos, mat = cs.sensor_poses(6, seed=21)
    l0 = engine.launch_count()
    for i in range(6):                                   # 1st plain, 2nd captures, then replays
        out = lid.cast(pos[i:i + 1], mat[i:i + 1], want=("ranges", "geomid", "dist"))
        r, g, raw = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=4)
        assert np.array_equal(out["ranges"][0], r) and np.array_equal(out["geomid"][0], g)
        assert np.array_equal(out["dist"][0], raw)
        assert lid.last_device_us() > 0
    per_call = (engine.launch_count() - l0) / 6
    assert per_call >= 1                                 # replays keep counting their kernels
    # a different output set is a different shape: falls back to plain, then captures again
    out = lid.cast(pos[:1], mat[:1], want=("ranges",))
    assert np.array_equal(out["ranges"][0], o.lidar_scan(cs.lidar, pos[0], mat[0], nthread=4)[0])
    # begin .. end with work in between; results only after end
    pend = lid.cast_begin(pos[2:3], mat[2:3], want=("ranges", "geomid"))
    gpu_scene.update_poses(cs.poses)                     # a refresh lands after the enqueued cast
    got = lid.cast_end()
    assert got is pend and np.array_equal(got["geomid"][0], o.lidar_scan(cs.lidar, pos[2], mat[2], nthread=4)[1])
    with pytest.raises(_capi.B2Error):
        lid.cast_end()                                   # nothing in flight
    # bulk path: pinned destination (direct D2H) == pageable destination (staged)
    npos, nmat = cs.sensor_poses(80, seed=22)
    pinned = {"ranges": torch.empty((80, lid.nray), dtype=torch.float32).pin_memory().numpy(),
              "geomid": torch.empty((80, lid.nray), dtype=torch.int32).pin_memory().numpy()}
    a = lid.cast(npos, nmat, out=pinned)
    b = lid.cast(npos, nmat)
    assert np.array_equal(a["ranges"], b["ranges"]) and np.array_equal(a["geomid"], b["geomid"])
    assert np.array_equal(b["geomid"][33], o.lidar_scan(cs.lidar, npos[33], nmat[33], nthread=4)[1])
    lid.close()


def test_graph_survives_buffer_growth(oracle, clutter, gpu_scene):
    """A captured single-scan graph bakes device addresses in.  A bigger cast in between reallocates
    the lidar's and the scene's scratch: the next single scan must not replay the stale graph."""
    cs = clutter
    lid = engine.Lidar(gpu_scene, cs.lidar)
    o = oracle.OracleScene(cs.model, cs.poses)
    pos, mat = cs.sensor_poses(90, seed=41)
    for i in range(4):                                   # plain, capture, replay, replay
        lid.cast(pos[i:i + 1], mat[i:i + 1])
    big = lid.cast(pos, mat)                             # 90 scans: every buffer grows
    assert np.array_equal(big["geomid"][50], o.lidar_scan(cs.lidar, pos[50], mat[50], nthread=4)[1])
    other = engine.Lidar(gpu_scene, scenes.LidarSpec(cs.lidar.local_dir[:4096], tile_w=0, bodyexclude=63))
    other.cast(pos[:300 % 90 + 60], mat[:300 % 90 + 60])  # another sensor grows the scene's scratch too
    for i in (5, 6, 7):
        out = lid.cast(pos[i:i + 1], mat[i:i + 1], want=("ranges", "geomid"))
        r, g, _ = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=4)
        assert np.array_equal(out["ranges"][0], r) and np.array_equal(out["geomid"][0], g)
    other.close()
    lid.close()


def test_latest_wins_pose_mirror(oracle, clutter):
    """After the first update, host updates of all worlds only fill a pinned slot; the next cast
    uploads the NEWEST one.  Interleaved with partial / device updates the order is kept."""
    import torch
    cs = clutter
    P = scenes.clutter_world_block(cs, 0, 4, seed=77)
    s = engine.Scene(cs.model, nworld=1)
    s.update_poses(P.world(0))                           # immediate: builds the topology
    lid = engine.Lidar(s, cs.lidar)
    pos, mat = cs.sensor_poses(1)
    for w in (1, 2, 3):                                  # three refreshes, no cast in between
        s.update_poses(P.world(w))
    got = lid.cast(pos, mat)
    o = oracle.OracleScene(cs.model, P, world=3)
    r, g, _ = o.lidar_scan(cs.lidar, pos[0], mat[0], nthread=4)
    assert np.array_equal(got["ranges"][0], r) and np.array_equal(got["geomid"][0], g)
    # host refresh, then a device update: the device update is the newer one
    dev = torch.device("cuda", 0)
    s.update_poses(P.world(1))
    s.update_poses_device(torch.from_numpy(P.geom_xpos[2:3]).to(dev), torch.from_numpy(P.geom_xmat[2:3]).to(dev))
    got = lid.cast(pos, mat)
    o.set_world(P, 2)
    assert np.array_equal(got["geomid"][0], o.lidar_scan(cs.lidar, pos[0], mat[0], nthread=4)[1])
    # and the other way round
    s.update_poses_device(torch.from_numpy(P.geom_xpos[2:3]).to(dev), torch.from_numpy(P.geom_xmat[2:3]).to(dev))
    s.update_poses(P.world(0))
    s.sync()                                             # sync flushes the pending refresh too
    got = lid.cast(pos, mat)
    o.set_world(P, 0)
    assert np.array_equal(got["geomid"][0], o.lidar_scan(cs.lidar, pos[0], mat[0], nthread=4)[1])
    lid.close()
    s.close()


def test_worlds_not_yet_updated_mirror_the_first(oracle, clutter):
    """A cast against a world no update has covered walks valid boxes: it sees the poses the
    topology was built on (never uninitialised memory)."""
    cs = clutter
    P = scenes.clutter_world_block(cs, 0, 3, seed=78)
    s = engine.Scene(cs.model, nworld=5)
    s.update_poses(Poses(P.geom_xpos[1:2], P.geom_xmat[1:2]), world0=2)   # first update covers world 2 only
    lid = engine.Lidar(s, cs.lidar)
    pos, mat = cs.sensor_poses(5, seed=3)
    got = lid.cast(pos, mat, scan_world=np.arange(5, dtype=np.int32))
    o = oracle.OracleScene(cs.model, P, world=1)
    for w in range(5):
        assert np.array_equal(got["geomid"][w], o.lidar_scan(cs.lidar, pos[w], mat[w], nthread=4)[1])
    s.update_poses(Poses(P.geom_xpos[2:3], P.geom_xmat[2:3]), world0=4)
    got = lid.cast(pos, mat, scan_world=np.arange(5, dtype=np.int32))
    o2 = oracle.OracleScene(cs.model, P, world=2)
    assert np.array_equal(got["geomid"][4], o2.lidar_scan(cs.lidar, pos[4], mat[4], nthread=4)[1])
    assert np.array_equal(got["geomid"][0], o.lidar_scan(cs.lidar, pos[0], mat[0], nthread=4)[1])
    lid.close()
    s.close()


def test_peer_buffer_as_cast_output(clutter, gpu_scene):
    """b2_peer_buffer_create: exportable device memory; the cast writes straight into it (what a
    non-root rank does with the root's buffer over NVLink)."""
    import torch
    cs = clutter
    lib = _capi.load()
    lid = engine.Lidar(gpu_scene, cs.lidar)
    nscan = 8
    n = nscan * lid.nray
    ptrs, handles = [], []
    for nbytes in (n * 4, n * 4):
        p = C.c_void_p()
        h = (C.c_uint8 * _capi.PEER_HANDLE_BYTES)()
        _capi.check(lib.b2_peer_buffer_create(0, nbytes, C.byref(p), h))
        ptrs.append(int(p.value))
        handles.append(bytes(h))
    assert any(handles[0]) and handles[0] != handles[1]
    pos, mat = cs.sensor_poses(nscan, seed=4)
    dev = torch.device("cuda", 0)
    lid.cast_device(nscan, torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev), None,
                    ranges=ptrs[0], geomid=ptrs[1])
    gpu_scene.sync()
    r = torch.as_tensor(shard._DevMem(ptrs[0], (nscan, lid.nray), "<f4"), device=dev).cpu().numpy()
    g = torch.as_tensor(shard._DevMem(ptrs[1], (nscan, lid.nray), "<i4"), device=dev).cpu().numpy()
    ref = lid.cast(pos, mat)
    assert np.array_equal(r, ref["ranges"]) and np.array_equal(g, ref["geomid"])
    for p in ptrs:
        _capi.check(lib.b2_peer_buffer_destroy(0, C.c_void_p(p)))
    lid.close()


def test_camera_hides_collision_groups(oracle):
    """the reference renders with mjv_defaultOption -- geom groups 3-5 (the converter's
    collision geoms) are invisible to the camera while the lidar still sees them."""
    from mujoco_ros2_simulation_b200 import mjcf
    sc = mjcf.load_mjcf(os.path.join(os.path.dirname(__file__), "golden", "plugin_scene.xml"))
    poses = sc.poses(None)
    s = engine.Scene(sc.model, nworld=1)
    s.update_poses(poses)
    spec = scenes.CameraSpec(64, 48, 58.0, sc.znear, sc.zfar, 1.0, 12)
    bpos, bmat, _ = sc.body_poses(None)
    c = sc.cameras[0]
    from mujoco_ros2_simulation_b200.model import quat2mat
    cpos = bpos[c.body] + bmat[c.body] @ c.frame.pos
    cmat = bmat[c.body] @ quat2mat(c.frame.quat)
    gl_def, z_def, _ = engine.DepthCamera(spec, scene=s).render_to_cloud(cpos[None], cmat.reshape(1, 9))
    spec_all = scenes.CameraSpec(64, 48, 58.0, sc.znear, sc.zfar, 1.0, 12, geomgroup=(1,) * 6)
    gl_all, z_all, _ = engine.DepthCamera(spec_all, scene=s).render_to_cloud(cpos[None], cmat.reshape(1, 9))
    K = engine.DepthCamera(spec, scene=s).info()[0]
    vs, us = np.meshgrid(np.arange(48), np.arange(64), indexing="ij")
    l = np.stack([(us + 0.5 - K[2]) / K[0], -(vs + 0.5 - K[5]) / K[4], -np.ones_like(us, float)], -1).reshape(-1, 3)
    vec = l @ cmat.T
    o = oracle.OracleScene(sc.model, poses)
    d_def, g_def = (a.reshape(48, 64) for a in o.multi_ray(cpos, vec, geomgroup=[1, 1, 1, 0, 0, 0]))
    d_all, g_all = (a.reshape(48, 64) for a in o.multi_ray(cpos, vec))
    # pixels where a collision-only geom (the sphere) is nearer than anything the renderer draws; the
    # coincident visual / collision boxes differ in id only, not in depth
    occluded = (d_all >= 0) & (d_all != d_def)
    assert occluded.sum() > 0
    assert (z_def[0][occluded] > z_all[0][occluded]).all()
    same = ~occluded
    assert np.array_equal(z_def[0][same], z_all[0][same])
    inside = (d_def < 0) | (d_def < 40.0)               # hits beyond zfar = 50 m sit on the far plane
    assert np.array_equal((gl_def[0][::-1] < 1.0)[inside], (g_def >= 0)[inside])   # GL buffer is bottom-up
    s.close()


def test_refresh_cost_while_the_gpu_is_busy():
    """A9 / SURVEY B3: SceneMirror::refresh after mj_step must cost the physics thread microseconds
    while a sensor thread casts 131 072-ray scans back to back (1.5 kHz for 2 s, native threads)."""
    from mujoco_ros2_simulation_b200 import mjcf
    from tests import plugin_binding as plug
    from tests.test_plugin import INFO
    plug.build()
    sc = mjcf.load_mjcf(os.path.join(os.path.dirname(__file__), "golden", "plugin_scene.xml"))
    h = plug.PluginHarness(sc, INFO)
    assert h.register() == 0 and h.activate() == 0, h.error()
    res = h.refresh_stress(64, 2048, 1500.0, 2.0)
    print("refresh while casting:", res)
    assert res["refreshes"] > 2500 and res["casts"] > 100
    assert res["p99_us"] < 100.0, res
    h.close()


def test_three_threads_like_the_plugin(oracle, clutter):
    """physics thread refreshing, lidar thread casting single scans (graph capture + replay), camera
    thread rendering through the blocking call and syncing -- all on one scene, as in the plugin
    (ref:src/mujoco_system_interface.cpp threads :595,:724, lidar.cpp:216, cameras.cpp:141).  No
    call may fail (a stream-wide synchronize during another thread's capture would), and the
    results are the oracle's."""
    import threading
    cs = clutter
    s = engine.Scene(cs.model, nworld=1)
    s.update_poses(cs.poses)
    lid = engine.Lidar(s, cs.lidar)
    cam = engine.DepthCamera(scenes.CameraSpec(64, 48, 58.0), scene=s)
    pos, mat = cs.sensor_poses(8, seed=61)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    cp = np.array([[3.0, 0.0, 1.5]])
    errors, scans, frames = [], {}, []
    stop = threading.Event()

    def physics():
        try:
            while not stop.is_set():
                s.update_poses(cs.poses)
        except Exception as e:  # noqa: BLE001
            errors.append(("physics", repr(e)))

    def lidar_thread():
        try:
            for it in range(120):
                i = it % 8
                scans[i] = lid.cast(pos[i:i + 1], mat[i:i + 1], want=("ranges", "geomid"))
        except Exception as e:  # noqa: BLE001
            errors.append(("lidar", repr(e)))

    def camera_thread():
        try:
            for _ in range(40):
                frames.append(cam.render_to_cloud(cp, R, want=("depth",))[1])
                s.sync()
        except Exception as e:  # noqa: BLE001
            errors.append(("camera", repr(e)))

    ts = [threading.Thread(target=f) for f in (physics, lidar_thread, camera_thread)]
    for t in ts:
        t.start()
    ts[1].join()
    ts[2].join()
    stop.set()
    ts[0].join()
    assert not errors, errors
    o = oracle.OracleScene(cs.model, cs.poses)
    for i in range(8):
        r, g, _ = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=4)
        assert np.array_equal(scans[i]["ranges"][0], r) and np.array_equal(scans[i]["geomid"][0], g)
    assert len(frames) == 40 and all(np.array_equal(f, frames[0]) for f in frames)
    cam.close()
    lid.close()
    s.close()
