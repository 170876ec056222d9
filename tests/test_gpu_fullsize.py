"""GPU tier at BASELINE.json's full sizes, through size-independent properties (the oracle would
take minutes here) plus oracle spot checks on sampled worlds."""
import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import engine, scenes
from mujoco_ros2_simulation_b200.model import Poses

pytestmark = pytest.mark.gpu


def test_cfg3_4096_worlds_refit_per_step(oracle):
    """cfg3: 4096 replicated worlds x 1024-ray lidar, poses refit per step.  Properties: every
    world's scan equals the scan of the same world cast alone (world independence); refitting to
    new poses then back reproduces the first result bit for bit (no state leaks through the refit);
    sampled worlds match the oracle."""
    import torch
    W = 4096
    dev = torch.device("cuda", 0)
    ws0 = scenes.arm_worlds(nworld=W, n_rays=1024, step=0)
    ws1 = scenes.arm_worlds(nworld=W, n_rays=1024, step=1)
    s = engine.Scene(ws0.model, nworld=W)
    lid = engine.Lidar(s, ws0.lidar)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    spos, smat, sw = d(ws0.sensor_pos), d(ws0.sensor_mat), d(ws0.scan_world)
    mk = lambda dt: torch.empty((W, 1024), dtype=dt, device=dev)

    def run(ws):
        s.update_poses_device(d(ws.poses.geom_xpos), d(ws.poses.geom_xmat))
        g, dist = mk(torch.int32), mk(torch.float64)
        lid.cast_device(W, spos, smat, sw, geomid=g, dist=dist)
        s.sync()
        return g, dist

    g0, d0 = run(ws0)
    g1, d1 = run(ws1)
    g0b, d0b = run(ws0)
    assert torch.equal(g0, g0b) and torch.equal(d0, d0b)
    assert not torch.equal(d0, d1)
    assert float((g0 >= 0).double().mean()) > 0.3
    # world independence: a single-world scene gives the same scan
    for w in (0, 1234, W - 1):
        s1 = engine.Scene(ws0.model, nworld=1)
        s1.update_poses(Poses(ws0.poses.geom_xpos[w], ws0.poses.geom_xmat[w]))
        out = engine.Lidar(s1, ws0.lidar).cast(ws0.sensor_pos[w][None], ws0.sensor_mat[w][None],
                                               want=("geomid", "dist"))
        assert np.array_equal(out["dist"][0], d0[w].cpu().numpy())
        assert np.array_equal(out["geomid"][0], g0[w].cpu().numpy())
        o = oracle.OracleScene(ws0.model, ws0.poses, world=w)
        r, g, raw = o.lidar_scan(ws0.lidar, ws0.sensor_pos[w], ws0.sensor_mat[w], nthread=4)
        assert np.array_equal(raw, out["dist"][0]) and np.array_equal(g, out["geomid"][0])


def test_cfg4_depth_1024_cameras_vs_torch_fp32():
    """cfg4 depth half: 1024 x 640x480 GL depth buffers -> 32FC1 + PointCloud2 in one launch,
    compared with a plain PyTorch fp32 restatement of ref:src/mujoco_cameras.cpp:222-240 and
    depth_image_proc's projection (torch CUDA elementwise ops are unfused IEEE fp32: bit-exact)."""
    import torch
    dev = torch.device("cuda", 0)
    n, H, Wd = 1024, 480, 640
    cam = engine.DepthCamera(scenes.depth_camera())
    gen = torch.Generator(device=dev).manual_seed(7)
    gl = torch.rand((n, H, Wd), dtype=torch.float32, device=dev, generator=gen)
    gl[:, 0, 0] = 0.0
    gl[:, 0, 1] = 1.0
    depth = torch.empty_like(gl)
    cloud = torch.empty((n, H, Wd, 3), dtype=torch.float32, device=dev)
    cam.depth_to_cloud_device(n, gl, depth, cloud)
    torch.cuda.synchronize()
    near, far = torch.tensor(0.01, dtype=torch.float32, device=dev), torch.tensor(50.0, dtype=torch.float32, device=dev)
    scale = 1.0 - near / far
    ref = (near / (1.0 - gl * scale)).flip(1)
    assert torch.equal(depth, ref)
    K, _ = cam.info()
    u = torch.arange(Wd, dtype=torch.float32, device=dev)[None, None, :]
    v = torch.arange(H, dtype=torch.float32, device=dev)[None, :, None]
    cx, cy = torch.tensor(np.float32(K[2]), device=dev), torch.tensor(np.float32(K[5]), device=dev)
    ifx = torch.tensor(np.float32(1.0) / np.float32(K[0]), device=dev)
    ify = torch.tensor(np.float32(1.0) / np.float32(K[4]), device=dev)
    for lo in range(0, n, 128):  # in slabs: the reference expression needs temporaries
        z = ref[lo:lo + 128]
        assert torch.equal(cloud[lo:lo + 128, ..., 2], z)
        assert torch.equal(cloud[lo:lo + 128, ..., 0], (u - cx) * z * ifx)
        assert torch.equal(cloud[lo:lo + 128, ..., 1], (v - cy) * z * ify)


def test_cfg5_shape_worlds_on_clutter_topology(oracle):
    """cfg5 per-GPU shape, reduced count: 16 worlds of the cluttered mesh scene with per-world body
    jitter on ONE shared BVH topology, each with the full 64x2048 lidar and a depth camera whose
    buffer is ray-cast from the same scene mirror.  Lidar: oracle on sampled scan lines; camera:
    depth linearisation of the ray-cast buffer is consistent with the cloud's z."""
    import torch
    dev = torch.device("cuda", 0)
    cs = scenes.clutter500()
    W = 16
    P = scenes.clutter_worlds(cs, W)
    s = engine.Scene(cs.model, nworld=W)
    s.update_poses(P)
    lid = engine.Lidar(s, cs.lidar)
    pos, mat = cs.sensor_poses(W)
    out = lid.cast(pos, mat, scan_world=np.arange(W, dtype=np.int32), want=("geomid", "dist"))
    sub = scenes.LidarSpec(cs.lidar.local_dir[::37], bodyexclude=63)
    for w in (0, 7, W - 1):
        o = oracle.OracleScene(cs.model, P, world=w)
        r, g, raw = o.lidar_scan(sub, pos[w], mat[w], nthread=8)
        assert np.array_equal(out["dist"][w][::37], raw)
        assert np.array_equal(out["geomid"][w][::37], g)
    cam = engine.DepthCamera(scenes.depth_camera(160, 120), scene=s)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    cpos = torch.from_numpy(np.tile([[3.0, 0.0, 1.5]], (W, 1))).to(dev)  # outside the sensor housing
    cmat = torch.from_numpy(np.tile(R, (W, 1))).to(dev)
    cw = torch.arange(W, dtype=torch.int32, device=dev)
    gl = torch.empty((W, 120, 160), dtype=torch.float32, device=dev)
    cam.render_depth_device(W, cpos, cmat, gl, d_world=cw)
    depth = torch.empty_like(gl)
    cloud = torch.empty((W, 120, 160, 3), dtype=torch.float32, device=dev)
    cam.depth_to_cloud_device(W, gl, depth, cloud)
    s.sync()
    assert torch.equal(cloud[..., 2], depth) and bool((gl <= 1).all()) and bool((gl >= 0).all())
    assert not torch.equal(gl[0], gl[5])     # the worlds differ
