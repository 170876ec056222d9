"""GPU tier: the launch shapes of K2 against each other and the oracle -- the two-kernel cast
(collect -> drain, with its fused fall-back) vs the fused kernel, and the persistent (dynamic group
fetch) launch vs the static tiling.  Single-world batches above ~1.2 M rays take the persistent launch (B2_DYNAMIC=0 forces
the static one); both must give the same bits, with and without the split mesh path, in every
mode (lidar / multiRay / camera rays)."""
import os

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import engine, scenes

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _cast_dev(lid, s, pos, mat, xyz=False):
    import torch
    dev = torch.device("cuda", 0)
    n = pos.shape[0]
    g = torch.full((n, lid.nray), -7, dtype=torch.int32, device=dev)
    d = torch.full((n, lid.nray), -7.0, dtype=torch.float64, device=dev)
    r = torch.full((n, lid.nray), -7.0, dtype=torch.float32, device=dev)
    x = torch.zeros((n, lid.nray, 3), dtype=torch.float32, device=dev) if xyz else None
    lid.cast_device(n, torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev), None, geomid=g, dist=d,
                    ranges=r, xyz=x)
    s.sync()
    return g.cpu().numpy(), d.cpu().numpy(), r.cpu().numpy(), None if x is None else x.cpu().numpy()


@pytest.mark.parametrize("split", [1, 0])
def test_lidar_dynamic_equals_static_and_oracle(oracle, split):
    cs = scenes.clutter500()                       # 64 x 2048 rays per scan
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    lid = engine.Lidar(s, cs.lidar)
    n = 12 if split else 10                        # 1.57 M / 1.31 M rays: above the persistent threshold
    pos, mat = cs.sensor_poses(n)
    with _env(B2_DYNAMIC=1, B2_SPLIT=split):
        dyn = _cast_dev(lid, s, pos, mat, xyz=True)
    with _env(B2_DYNAMIC=0, B2_SPLIT=split):
        sta = _cast_dev(lid, s, pos, mat, xyz=True)
    with _env(B2_TWO_KERNEL=0, B2_SPLIT=split):
        fused = _cast_dev(lid, s, pos, mat, xyz=True)             # the fused K2 instead of collect -> drain
    for a, b, c in zip(dyn, sta, fused):
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a, c, equal_nan=True)
    assert (dyn[0] != -7).all() and (dyn[1] != -7.0).all()      # every ray was written
    o = oracle.OracleScene(cs.model, cs.poses)
    for sc in (0, n - 1):
        r, g, raw = o.lidar_scan(cs.lidar, pos[sc], mat[sc], nthread=8)
        assert np.array_equal(raw, dyn[1][sc]) and np.array_equal(g, dyn[0][sc])
        assert np.array_equal(r, dyn[2][sc])


def test_ragged_scan_dynamic(oracle):
    """nray not a multiple of the group (4 tiles x 32 rays) nor of the tile: the last group of every
    scan is short and its last tile is ragged."""
    cs = scenes.clutter500(n_el=32, n_az=256)
    nray = 32 * 256 - 77
    spec = scenes.LidarSpec(cs.lidar.local_dir[:nray], bodyexclude=63)
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    lid = engine.Lidar(s, spec)
    pos, mat = cs.sensor_poses(160)                # 1.3 M rays
    with _env(B2_DYNAMIC=1):
        dyn = _cast_dev(lid, s, pos, mat)
    with _env(B2_DYNAMIC=0):
        sta = _cast_dev(lid, s, pos, mat)
    for a, b in zip(dyn[:3], sta[:3]):
        assert np.array_equal(a, b, equal_nan=True)
    assert (dyn[0] != -7).all()
    o = oracle.OracleScene(cs.model, cs.poses)
    r, g, raw = o.lidar_scan(spec, pos[159], mat[159], nthread=8)
    assert np.array_equal(raw, dyn[1][159]) and np.array_equal(g, dyn[0][159])


def test_multiray_dynamic(oracle):
    cs = scenes.clutter500(n_el=32, n_az=256)
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    rng = np.random.default_rng(11)
    n = 1_400_000
    vec = rng.normal(size=(n, 3))
    pnt = cs.sensor_pos
    with _env(B2_DYNAMIC=1):
        d1, g1 = s.multiray(pnt, vec, bodyexclude=63)
    with _env(B2_DYNAMIC=0):
        d0, g0 = s.multiray(pnt, vec, bodyexclude=63)
    assert np.array_equal(d1, d0) and np.array_equal(g1, g0)
    o = oracle.OracleScene(cs.model, cs.poses)
    sel = rng.choice(n, 4000, replace=False)
    d, g = o.multi_ray(pnt, vec[sel], bodyexclude=63, nthread=8)
    assert np.array_equal(d, d1[sel]) and np.array_equal(g, g1[sel])


def test_camera_rays_dynamic():
    """camera mode (N1): 6 cameras x 640x480 = 1.8 M rays in one world -- persistent == static."""
    import torch
    cs = scenes.clutter500(n_el=32, n_az=256)
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    cam = engine.DepthCamera(scenes.CameraSpec(640, 480, 58.0, 0.01, 50.0, 1.0, 12), scene=s)
    dev = torch.device("cuda", 0)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]])
    n = 6
    pos = np.array([[3.0 + 0.3 * i, 0.2 * i, 1.5] for i in range(n)])
    d_pos = torch.from_numpy(pos).to(dev)
    d_mat = torch.from_numpy(np.tile(R.reshape(1, 9), (n, 1))).to(dev)
    outs = []
    for dyn in (1, 0):
        gl = torch.full((n, 480, 640), -1.0, dtype=torch.float32, device=dev)
        with _env(B2_DYNAMIC=dyn):
            cam.render_depth_device(n, d_pos, d_mat, gl)
        s.sync()
        outs.append(gl.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    assert (outs[0] >= 0).all() and (outs[0] < 1.0).mean() > 0.2
    assert not np.array_equal(outs[0][0], outs[0][5])


def test_two_kernel_vs_fused_with_overflowing_groups(oracle):
    """a dense cluster (every group's candidate list overflows -> fused fall-back kernel), rays whose
    length is outside the float range (group handed to the fall-back too) and ordinary groups, in one
    launch above the two-kernel threshold: two-kernel == fused == oracle."""
    from mujoco_ros2_simulation_b200.model import GEOM_BOX, GEOM_SPHERE, Model, Poses, random_rotations
    rng = np.random.default_rng(25)
    n = 400
    t = rng.choice([GEOM_BOX, GEOM_SPHERE], size=n)
    pos = rng.normal(size=(n, 3)) * 0.6
    pos[300:] += np.array([0.0, 30.0, 0.0])          # a second, sparser cluster: ordinary groups
    pos[300:] *= np.array([1.0, 1.0, 8.0])
    m = Model(t, np.zeros(n, int), np.zeros(n, int), -np.ones(n, int), -np.ones(n, int),
              rng.uniform(0.5, 1.5, size=(n, 3)), np.ones((n, 4)), np.zeros((0, 4)), [0])
    m.geom_size[300:] *= 0.2
    P = Poses(pos, random_rotations(rng, n).reshape(n, 9))
    s = engine.Scene(m)
    s.update_poses(P)
    nray = 98304
    az = np.linspace(0, 2 * np.pi, nray, endpoint=False)
    vec = np.stack([np.cos(az), np.sin(az), 0.05 * np.sin(7 * az)], 1)
    vec[5000:5040] *= 1e-25                           # |vec|^2 underflows float: no culling for that group
    vec[70000:70100] *= 1e22
    pnt = np.array([6.0, 0.5, 0.2])
    with _env(B2_TWO_KERNEL=1, B2_TWO_MIN_RAYS=1):    # (the default threshold is 1 M rays)
        d2, g2 = s.multiray(pnt, vec)
    with _env(B2_TWO_KERNEL=0):
        d1, g1 = s.multiray(pnt, vec)
    assert np.array_equal(d1, d2) and np.array_equal(g1, g2)
    sel = np.concatenate([np.arange(0, nray, 37), np.arange(4990, 5050), np.arange(69990, 70110)])
    od, og = oracle.OracleScene(m, P).multi_ray(pnt, vec[sel], nthread=8)
    assert np.array_equal(od, d2[sel]) and np.array_equal(og, g2[sel])
    assert (g2 >= 0).mean() > 0.2


def test_two_kernel_fall_back_in_several_worlds(oracle):
    """overflowing groups with nworld > 1: the unstaged fall-back kernel picks the world per group."""
    import torch
    from mujoco_ros2_simulation_b200.model import GEOM_BOX, GEOM_SPHERE, Model, Poses, random_rotations
    rng = np.random.default_rng(31)
    n, W = 300, 3
    t = rng.choice([GEOM_BOX, GEOM_SPHERE], size=n)
    m = Model(t, np.zeros(n, int), np.zeros(n, int), -np.ones(n, int), -np.ones(n, int),
              rng.uniform(0.4, 1.2, size=(n, 3)), np.ones((n, 4)), np.zeros((0, 4)), [0])
    pos = rng.normal(size=(W, n, 3)) * 0.7                      # three different dense clusters
    mat = np.stack([random_rotations(rng, n).reshape(n, 9) for _ in range(W)])
    P = Poses(pos, mat)
    s = engine.Scene(m, nworld=W)
    s.update_poses(P)
    az = np.linspace(0, 2 * np.pi, 4096, endpoint=False)
    spec = scenes.LidarSpec(np.stack([np.cos(az), np.sin(az), 0.1 * np.sin(5 * az)], 1))
    lid = engine.Lidar(s, spec)
    S = 9
    spos = np.tile([[5.0, 0.3, 0.1]], (S, 1)) + rng.normal(size=(S, 3)) * 0.2
    smat = np.tile(np.eye(3).reshape(1, 9), (S, 1))
    sw = (np.arange(S) % W).astype(np.int32)
    dev = torch.device("cuda", 0)
    res = []
    for two in (1, 0):
        g = torch.full((S, 4096), -7, dtype=torch.int32, device=dev)
        d = torch.full((S, 4096), -7.0, dtype=torch.float64, device=dev)
        with _env(B2_TWO_KERNEL=two, B2_TWO_MIN_RAYS=1):
            lid.cast_device(S, torch.from_numpy(spos).to(dev), torch.from_numpy(smat).to(dev),
                            torch.from_numpy(sw).to(dev), geomid=g, dist=d)
        s.sync()
        res.append((g.cpu().numpy(), d.cpu().numpy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for sc in (0, 4, 8):
        o = oracle.OracleScene(m, P, world=int(sw[sc]))
        r, gg, raw = o.lidar_scan(spec, spos[sc], smat[sc], nthread=8)
        assert np.array_equal(raw, res[0][1][sc]) and np.array_equal(gg, res[0][0][sc])
    assert (res[0][0] >= 0).mean() > 0.1


@pytest.mark.parametrize("n_az,nscan", [(128, 136), (256, 4)])
def test_group_width_follows_the_ray_table(oracle, n_az, nscan):
    """coarse ray tables get narrower collect groups (runtime.cu choose_group_tiles): 4, 2 and 1
    tiles per group and the automatic choice give the same bits as the oracle, in the two-kernel cast
    (1.1 M rays) and in the fused kernel (65 536 rays)."""
    cs = scenes.clutter500(n_el=64, n_az=n_az)
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    pos, mat = cs.sensor_poses(nscan)
    outs = []
    for gt in (0, 4, 2, 1):
        with _env(B2_GROUP_TILES_RT=gt):
            lid = engine.Lidar(s, cs.lidar)
        outs.append(_cast_dev(lid, s, pos, mat))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
        assert np.array_equal(o[2], outs[0][2])
    oc = oracle.OracleScene(cs.model, cs.poses)
    for sc in (0, nscan - 1):
        r, g, raw = oc.lidar_scan(cs.lidar, pos[sc], mat[sc], nthread=8)
        assert np.array_equal(raw, outs[0][1][sc]) and np.array_equal(g, outs[0][0][sc])


def test_many_tiny_scans(oracle):
    """50 000 poses of the reference fixture's 24-ray lidar in one call (1.2 M rays, one ragged tile
    per scan, one group per scan): two-kernel == fused == oracle."""
    rs = scenes.robot_scene()
    s = engine.Scene(rs.model)
    s.update_poses(rs.poses)
    lid = engine.Lidar(s, rs.lidar)
    S = 50_000
    rng = np.random.default_rng(2)
    pos = rs.lidar_pos[None] + rng.normal(size=(S, 3)) * 0.3
    mat = np.tile(rs.lidar_mat.reshape(1, 9), (S, 1))
    with _env(B2_TWO_KERNEL=1, B2_TWO_MIN_RAYS=1):
        two = _cast_dev(lid, s, pos, mat)
    with _env(B2_TWO_KERNEL=0):
        fused = _cast_dev(lid, s, pos, mat)
    for a, b in zip(two[:3], fused[:3]):
        assert np.array_equal(a, b, equal_nan=True)
    assert (two[0] != -7).all()
    o = oracle.OracleScene(rs.model, rs.poses)
    for sc in (0, 777, S - 1):
        r, g, raw = o.lidar_scan(rs.lidar, pos[sc], mat[sc])
        assert np.array_equal(raw, two[1][sc]) and np.array_equal(g, two[0][sc]) and np.array_equal(r, two[2][sc])
