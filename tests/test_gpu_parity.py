"""GPU tier: the CUDA path through the C ABI (include/b2sense.h) against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): geomid bit-exact, distances within 1e-5 relative --
the implementation computes accepted distances in IEEE double in the oracle's operation order, so
these tests assert BIT-EXACT distances too.  Full-size cases use size-independent properties."""
import ctypes as C

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import _capi, engine, scenes
from mujoco_ros2_simulation_b200.model import GEOM_MESH, GEOM_PLANE, GEOM_SPHERE, Model, Poses

pytestmark = pytest.mark.gpu
EYE = np.eye(3).reshape(9)


@pytest.fixture(scope="module")
def clutter():
    return scenes.clutter500(n_el=32, n_az=256)


@pytest.fixture(scope="module")
def clutter_gpu(clutter):
    s = engine.Scene(clutter.model, nworld=1)
    s.update_poses(clutter.poses)
    yield s
    s.close()


def test_native_library_is_the_path():
    """the extension is loaded from the tree and launches kernels; no eager/CPU stand-in exists"""
    lib = _capi.load()
    assert lib.b2_device_count() >= 1
    n0 = engine.launch_count()
    rs = scenes.robot_scene()
    s = engine.Scene(rs.model)
    s.update_poses(rs.poses)
    s.multiray(rs.lidar_pos, (rs.lidar_mat.reshape(3, 3) @ rs.lidar.local_dir.T).T)
    assert engine.launch_count() >= n0 + 2  # refit + cast


def test_test_robot_known_answer(oracle):
    """cfg1 / S1: the reference's fixture scene, 24 rays as shipped and the 360-ray variant."""
    for n in (None, 360):
        rs = scenes.robot_scene(n)
        s = engine.Scene(rs.model)
        s.update_poses(rs.poses)
        lid = engine.Lidar(s, rs.lidar)
        out = lid.cast(rs.lidar_pos[None], rs.lidar_mat.reshape(1, 9), want=("ranges", "geomid", "dist", "xyz"))
        o = oracle.OracleScene(rs.model, rs.poses)
        r, g, raw = o.lidar_scan(rs.lidar, rs.lidar_pos, rs.lidar_mat)
        assert np.array_equal(out["geomid"][0], g)
        assert np.array_equal(out["dist"][0], raw)
        assert np.array_equal(out["ranges"][0], r)
        ok = (g >= 0) & (r != -1)
        ref_xyz = (rs.lidar.local_dir * raw[:, None]).astype(np.float32)
        assert np.array_equal(out["xyz"][0][ok], ref_xyz[ok]) and np.isnan(out["xyz"][0][~ok]).all()
        if n is None:
            theta = -0.3 + 0.025 * np.arange(24)
            np.testing.assert_allclose(out["dist"][0], 1.95 / np.cos(theta), rtol=2e-6)
            assert out["geomid"][0].tolist() == [0] * 13 + [2] * 11


@pytest.mark.parametrize("gg,fs", [(None, 1), ([1, 0, 1, 1, 0, 0], 1), (None, 0), ([1, 0, 1, 1, 0, 0], 0)])
def test_multiray_clutter_masks(oracle, clutter, clutter_gpu, gg, fs):
    cs = clutter
    o = oracle.OracleScene(cs.model, cs.poses)
    vec = cs.lidar.local_dir
    d, g = o.multi_ray(cs.sensor_pos, vec, geomgroup=gg, flg_static=fs, bodyexclude=63, nthread=8)
    d2, g2 = clutter_gpu.multiray(cs.sensor_pos, vec, geomgroup=gg, flg_static=fs, bodyexclude=63)
    assert (g != g2).sum() == 0
    assert np.array_equal(d, d2)
    if fs == 1:
        assert (cs.model.geom_type[g[g >= 0]] == GEOM_MESH).sum() > 50


def test_multiray_cutoff_and_degenerate(oracle, clutter, clutter_gpu):
    cs = clutter
    o = oracle.OracleScene(cs.model, cs.poses)
    vec = np.concatenate([cs.lidar.local_dir[::5],
                          [[0, 0, 0.0], [1e-20, 0, 0], [0, 0, -1.0], [3.0, 4.0, 0.0], [0, 0, 1.0]]])
    for cutoff in (6.0, 1e10):
        d, g = o.multi_ray(cs.sensor_pos, vec, bodyexclude=63, cutoff=cutoff)
        d2, g2 = clutter_gpu.multiray(cs.sensor_pos, vec, bodyexclude=63, cutoff=cutoff)
        assert np.array_equal(d, d2) and np.array_equal(g, g2)
    d0, g0 = clutter_gpu.multiray(cs.sensor_pos, np.zeros((0, 3)))
    assert d0.shape == (0,)


def test_lidar_batch_host_and_device_paths(oracle, clutter, clutter_gpu):
    """many scans in one launch; host (blocking, chunked read-back) and device entry points agree
    with each other and with the oracle; 2-D tiling (tile_w) does not change results."""
    import torch
    cs = clutter
    pos, mat = cs.sensor_poses(24)
    lid = engine.Lidar(clutter_gpu, cs.lidar)
    spec1d = scenes.LidarSpec(cs.lidar.local_dir, tile_w=0, bodyexclude=63)
    lid1d = engine.Lidar(clutter_gpu, spec1d)
    out = lid.cast(pos, mat, scan_world=np.zeros(24, np.int32), want=("ranges", "geomid", "dist"))
    out1 = lid1d.cast(pos, mat, want=("ranges", "geomid", "dist"))
    for k in ("ranges", "geomid", "dist"):
        assert np.array_equal(out[k], out1[k])
    dev = torch.device("cuda", 0)
    d_pos, d_mat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
    r = torch.empty((24, lid.nray), dtype=torch.float32, device=dev)
    g = torch.empty((24, lid.nray), dtype=torch.int32, device=dev)
    lid.cast_device(24, d_pos, d_mat, None, ranges=r, geomid=g)
    clutter_gpu.sync()
    assert np.array_equal(r.cpu().numpy(), out["ranges"]) and np.array_equal(g.cpu().numpy(), out["geomid"])
    o = oracle.OracleScene(cs.model, cs.poses)
    for i in (0, 7, 23):
        rr, gg, raw = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=8)
        assert np.array_equal(out["geomid"][i], gg)
        assert np.array_equal(out["dist"][i], raw)
        assert np.array_equal(out["ranges"][i], rr)


def test_lidar_cutoff_and_range_filter(oracle, clutter, clutter_gpu):
    cs = clutter
    spec = scenes.LidarSpec(cs.lidar.local_dir, tile_w=cs.n_az, bodyexclude=63, cutoff=12.0,
                            range_min=1.0, range_max=11.0)
    lid = engine.Lidar(clutter_gpu, spec)
    pos, mat = cs.sensor_poses(3)
    out = lid.cast(pos, mat, want=("ranges", "geomid", "dist", "xyz"))
    o = oracle.OracleScene(cs.model, cs.poses)
    for i in range(3):
        r, g, raw = o.lidar_scan(spec, pos[i], mat[i], nthread=8)
        assert np.array_equal(out["ranges"][i], r) and np.array_equal(out["geomid"][i], g)
        assert np.array_equal(out["dist"][i], raw)
        assert np.isnan(out["xyz"][i][r == -1]).all()
    assert (out["ranges"] == -1).any() and (out["dist"] == 12.0).any()


def test_arm_worlds_refit_per_step(oracle):
    """cfg3 shape (reduced): W replicated worlds, per-world poses refit every step, one lidar each."""
    W = 64
    ws = scenes.arm_worlds(nworld=W, n_rays=1024)
    s = engine.Scene(ws.model, nworld=W)
    lid = engine.Lidar(s, ws.lidar)
    for step in range(3):
        ws = scenes.arm_worlds(nworld=W, n_rays=1024, step=step)
        s.update_poses(ws.poses)
        out = lid.cast(ws.sensor_pos, ws.sensor_mat, scan_world=ws.scan_world, want=("ranges", "geomid", "dist"))
        for w in (0, 1, 17, W - 1):
            o = oracle.OracleScene(ws.model, ws.poses, world=w)
            r, g, raw = o.lidar_scan(ws.lidar, ws.sensor_pos[w], ws.sensor_mat[w], nthread=4)
            assert np.array_equal(out["geomid"][w], g), (step, w)
            assert np.array_equal(out["dist"][w], raw)
            assert np.array_equal(out["ranges"][w], r)
    # partial update: only worlds [10, 20) change
    ws2 = scenes.arm_worlds(nworld=W, n_rays=1024, step=9)
    sub = Poses(ws2.poses.geom_xpos[10:20], ws2.poses.geom_xmat[10:20])
    s.update_poses(sub, world0=10)
    out2 = lid.cast(ws.sensor_pos, ws.sensor_mat, scan_world=ws.scan_world, want=("geomid", "dist"))
    assert np.array_equal(out2["dist"][:10], out["dist"][:10]) and np.array_equal(out2["dist"][20:], out["dist"][20:])
    o = oracle.OracleScene(ws2.model, ws2.poses, world=13)
    r, g, raw = o.lidar_scan(ws.lidar, ws.sensor_pos[13], ws.sensor_mat[13])
    assert np.array_equal(out2["dist"][13], raw) and np.array_equal(out2["geomid"][13], g)


def test_clutter_worlds_refit_and_rebuild(oracle, clutter):
    """S5 shape (reduced): shared BVH topology from world 0, per-world refit; device pose upload;
    b2_scene_rebuild leaves results unchanged."""
    import torch
    cs = clutter
    W = 4
    P = scenes.clutter_worlds(cs, W)
    s = engine.Scene(cs.model, nworld=W)
    dev = torch.device("cuda", 0)
    s.update_poses_device(torch.from_numpy(P.geom_xpos).to(dev), torch.from_numpy(P.geom_xmat).to(dev))
    s.sync()
    lid = engine.Lidar(s, cs.lidar)
    pos, mat = cs.sensor_poses(W)
    out = lid.cast(pos, mat, scan_world=np.arange(W, dtype=np.int32), want=("geomid", "dist"))
    for w in range(W):
        o = oracle.OracleScene(cs.model, P, world=w)
        r, g, raw = o.lidar_scan(cs.lidar, pos[w], mat[w], nthread=8)
        assert np.array_equal(out["geomid"][w], g)
        assert np.array_equal(out["dist"][w], raw)
    s.rebuild(world=W - 1)
    out2 = lid.cast(pos, mat, scan_world=np.arange(W, dtype=np.int32), want=("geomid", "dist"))
    assert np.array_equal(out["geomid"], out2["geomid"]) and np.array_equal(out["dist"], out2["dist"])


def test_full_size_properties():
    """cfg2 at full size (64x2048 rays x 64 poses): properties that need no oracle.
    (a) determinism; (b) exact homogeneity: doubling every ray vector halves every distance bit for
    bit and keeps every geomid (distances are in units of |vec|, scaling by 2 is exact in IEEE);
    (c) rays reported on the ground plane satisfy the plane equation; (d) xyz = dir * dist."""
    import torch
    cs = scenes.clutter500()
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    lid = engine.Lidar(s, cs.lidar)
    spec2 = scenes.LidarSpec(cs.lidar.local_dir * 2.0, tile_w=cs.n_az, bodyexclude=63)
    lid2 = engine.Lidar(s, spec2)
    n = 64
    pos, mat = cs.sensor_poses(n)
    dev = torch.device("cuda", 0)
    d_pos, d_mat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
    mk = lambda dt, *sh: torch.empty(sh, dtype=dt, device=dev)
    g1, g1b, g2 = (mk(torch.int32, n, lid.nray) for _ in range(3))
    d1, d1b, d2 = (mk(torch.float64, n, lid.nray) for _ in range(3))
    xyz = mk(torch.float32, n, lid.nray, 3)
    lid.cast_device(n, d_pos, d_mat, None, geomid=g1, dist=d1, xyz=xyz)
    lid.cast_device(n, d_pos, d_mat, None, geomid=g1b, dist=d1b)
    lid2.cast_device(n, d_pos, d_mat, None, geomid=g2, dist=d2)
    s.sync()
    assert torch.equal(g1, g1b) and torch.equal(d1, d1b)
    assert torch.equal(g1, g2)
    hit = g1 >= 0
    assert torch.equal(d1[hit], d2[hit] * 2.0)
    assert float(hit.double().mean()) > 0.95
    # ground plane z = 0: origin_z + dist * dir_z == 0 for rays whose geomid is the plane (id 0)
    ld = torch.from_numpy(cs.lidar.local_dir).to(dev)
    R = d_mat.view(n, 3, 3)
    wz = torch.einsum("sk,rk->sr", R[:, 2, :], ld)
    on_plane = g1 == 0
    resid = (d_pos[:, 2:3] + d1 * wz)[on_plane].abs().max()
    assert float(resid) < 1e-12 and int(on_plane.sum()) > 0.3 * g1.numel()
    ref = (ld[None] * d1[..., None]).to(torch.float32)
    assert torch.equal(xyz[hit], ref[hit]) and torch.isnan(xyz[~hit]).all()


@pytest.mark.parametrize("w,h,ps,ncam", [(640, 480, 12, 3), (640, 480, 16, 2), (1280, 720, 12, 1),
                                         (322, 7, 12, 2), (8, 4, 16, 1)])
def test_depth_to_cloud_bit_exact(oracle, w, h, ps, ncam):
    """K3 vs ref:src/mujoco_cameras.cpp:222-240 (float32 linearise + flip) and
    depth_image_proc::point_cloud_xyz on the reference's intrinsics (:119-124): bit-exact."""
    spec = scenes.CameraSpec(w, h, 58.0, 0.01, 50.0, 1.0, ps)
    cam = engine.DepthCamera(spec)
    gl = scenes.gl_depth(ncam, w, h)
    depth, cloud = cam.depth_to_cloud(gl)
    ref_d = oracle.depth_linearise_flip(gl, 0.01, 50.0, 1.0)
    K, P = cam.info()
    assert np.array_equal(K, oracle.camera_intrinsics(58.0, w, h))
    assert P[0] == K[0] and P[2] == K[2] and P[5] == K[4] and P[6] == K[5] and P[10] == 1.0
    ref_c = oracle.depth_to_cloud(ref_d, K, ps // 4)
    assert np.array_equal(depth, ref_d)
    assert np.array_equal(cloud, ref_c, equal_nan=True)
    # depth only / cloud only
    d_only, none = cam.depth_to_cloud(gl, want_cloud=False)
    assert none is None and np.array_equal(d_only, ref_d)
    none, c_only = cam.depth_to_cloud(gl, want_depth=False)
    assert none is None and np.array_equal(c_only, ref_c, equal_nan=True)


def test_depth_nonfinite_is_nan_point(oracle):
    spec = scenes.CameraSpec(16, 4, 58.0, 0.01, 50.0, 1.0, 12)
    cam = engine.DepthCamera(spec)
    gl = np.full((1, 4, 16), 0.5, np.float32)
    gl[0, 0, 0] = np.float32(1.0) / (np.float32(1.0) - np.float32(0.01) / np.float32(50.0))  # 1 - d*s -> 0
    gl[0, 1, 1] = np.nan
    depth, cloud = cam.depth_to_cloud(gl)
    ref_d = oracle.depth_linearise_flip(gl, 0.01, 50.0, 1.0)
    ref_c = oracle.depth_to_cloud(ref_d, cam.info()[0], 3)
    assert np.array_equal(depth, ref_d, equal_nan=True) and np.array_equal(cloud, ref_c, equal_nan=True)
    assert np.isnan(cloud[0, 2, 1]).all()


def test_render_depth_roundtrip(oracle, clutter, clutter_gpu):
    """N1: ray-cast GL depth -> K3 linearisation reproduces the oracle's ray distance along the
    optical axis (float32 depth encoding: 1e-3 relative near the far plane is the GL format's
    resolution, so compare where z < 10 m at 2e-4)."""
    import torch
    cs = clutter
    spec = scenes.CameraSpec(64, 48, 58.0, 0.01, 50.0, 1.0, 12)
    cam = engine.DepthCamera(spec, scene=clutter_gpu)
    dev = torch.device("cuda", 0)
    # camera at the lidar origin looking along +X: MuJoCo camera looks along -Z, +Y up
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]])  # columns: cam x,y,z axes in world
    pos = np.array([[3.0, 0.0, 1.5]])
    d_pos, d_mat = torch.from_numpy(pos).to(dev), torch.from_numpy(R.reshape(1, 9)).to(dev)
    gl = torch.empty((1, 48, 64), dtype=torch.float32, device=dev)
    cam.render_depth_device(1, d_pos, d_mat, gl)
    depth = torch.empty_like(gl)
    cam.depth_to_cloud_device(1, gl, depth, None)
    clutter_gpu.sync()
    z = depth.cpu().numpy()[0]
    K, _ = cam.info()
    o = oracle.OracleScene(cs.model, cs.poses)
    vs, us = np.meshgrid(np.arange(48), np.arange(64), indexing="ij")
    l = np.stack([(us + 0.5 - K[2]) / K[0], -(vs + 0.5 - K[5]) / K[4], -np.ones_like(us, float)], -1).reshape(-1, 3)
    vec = l @ R.T
    # the renderer draws what mjv_defaultOption shows (ref:src/mujoco_cameras.cpp:165): groups 0-2
    d, g = o.multi_ray(pos[0], vec, geomgroup=[1, 1, 1, 0, 0, 0], nthread=8)
    d = d.reshape(48, 64)
    near = (d > 0) & (d < 10)
    assert near.sum() > 500
    np.testing.assert_allclose(z[near], d[near], rtol=2e-4)
    assert np.all(z[d < 0] > 49.9)   # misses sit on the far plane
    # an explicit mask (all six groups) shows the hidden geoms as well
    cam_all = engine.DepthCamera(scenes.CameraSpec(64, 48, 58.0, 0.01, 50.0, 1.0, 12, geomgroup=(1,) * 6),
                                 scene=clutter_gpu)
    cam_all.render_depth_device(1, d_pos, d_mat, gl)
    cam_all.depth_to_cloud_device(1, gl, depth, None)
    clutter_gpu.sync()
    z_all = depth.cpu().numpy()[0]
    d_all = o.multi_ray(pos[0], vec, nthread=8)[0].reshape(48, 64)
    near = (d_all > 0) & (d_all < 10)
    np.testing.assert_allclose(z_all[near], d_all[near], rtol=2e-4)
    assert (d_all != d).any()        # groups 3-5 do occlude something in this view


def test_empty_and_tiny_scenes():
    m = Model(np.zeros(0, np.int32), [], [], [], [], np.zeros((0, 3)), np.zeros((0, 4)), np.zeros((0, 4)), [0])
    s = engine.Scene(m)
    d, g = s.multiray(np.zeros(3), np.array([[0, 0, 1.0], [1, 0, 0]]))
    assert np.all(d == -1) and np.all(g == -1)
    m1 = Model([GEOM_SPHERE], [0], [0], [-1], [-1], [[0.5, 0, 0]], [[1, 1, 1, 1]], np.zeros((0, 4)), [0])
    s1 = engine.Scene(m1)
    with pytest.raises(_capi.B2Error):
        s1.multiray(np.zeros(3), np.array([[0, 0, 1.0]]))   # cast before any pose update
    s1.update_poses(Poses(np.array([[0.0, 0, 3.0]]), EYE[None]))
    d, g = s1.multiray(np.zeros(3), np.array([[0, 0, 1.0], [0, 0, -1.0], [0, 0, 2.0]]))
    assert d.tolist() == [2.5, -1.0, 1.25] and g.tolist() == [0, -1, 0]
    mp = Model([GEOM_PLANE], [0], [0], [-1], [-1], [[0, 0, 0.01]], [[1, 1, 1, 1]], np.zeros((0, 4)), [0])
    sp = engine.Scene(mp)
    sp.update_poses(Poses(np.zeros((1, 3)), EYE[None]))
    d, g = sp.multiray(np.array([0, 0, 2.0]), np.array([[0, 0, -1.0], [0, 0, 1.0]]))
    assert d.tolist() == [2.0, -1.0] and g.tolist() == [0, -1]
    lid = engine.Lidar(s1, scenes.LidarSpec(np.zeros((0, 3))))
    out = lid.cast(np.zeros((2, 3)), np.tile(EYE, (2, 1)))
    assert out["ranges"].shape == (2, 0)
