"""CPU tier: pin the oracle (oracle/b2_oracle.c) -- the reference holds no golden vector for this
path ("parity unpinned"), so the pins are (i) the survey's hand-derived known
answer for the reference's own fixture scene, (ii) analytic distances, (iii) an independent
numpy implementation, (iv) the reference's own arithmetic quirks (ref:src/mujoco_lidar.cpp:92-93,
269-271; ref:src/mujoco_cameras.cpp:119-124,222-240)."""
import json
import os

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import scenes
from mujoco_ros2_simulation_b200.model import (GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER,
                                               GEOM_ELLIPSOID, GEOM_PLANE, GEOM_SPHERE, Model, Poses,
                                               random_rotations)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
EYE = np.eye(3).reshape(9)


def test_known_answer_test_robot(oracle):
    """range[i] = 1.95/cos(theta_i) within 2e-6, geomid 0 (i<=12) / 2 (i>=13)."""
    fx = json.load(open(os.path.join(GOLDEN, "test_robot_scene.json")))
    rs = scenes.robot_scene()
    o = oracle.OracleScene(rs.model, rs.poses)
    ranges, gid, raw = o.lidar_scan(rs.lidar, rs.lidar_pos, rs.lidar_mat)
    ka = fx["known_answer"]
    assert ka["num_rangefinders"] == 24
    np.testing.assert_allclose(raw, ka["range"], rtol=2e-6)
    assert gid.tolist() == ka["geomid"]
    # the survey's spot values
    for i, v in ((0, 2.041167942), (11, 1.950609713), (12, 1.95), (13, 1.950609355), (23, 2.026129390)):
        assert abs(raw[i] - v) < 5e-6
    # LaserScan filter: range_min 0.05, range_max 10 keep everything here
    np.testing.assert_array_equal(ranges, raw.astype(np.float32))


def test_num_rangefinders_truncation(oracle):
    """ref:src/mujoco_lidar.cpp:92-93: 0.6/0.025 = 23.999.. -> 23 + 1 = 24 (not 25)."""
    assert oracle.lidar_num_rangefinders(-0.3, 0.3, 0.025) == 24
    assert oracle.lidar_num_rangefinders(0.0, 1.0, 0.25) == 5
    assert oracle.lidar_num_rangefinders(-np.pi, np.pi, 2 * np.pi / 360) in (360, 361)


def test_range_filter_on_float(oracle):
    """ref:src/mujoco_lidar.cpp:269-271: the comparison is on the float-rounded range."""
    m = Model([GEOM_SPHERE], [0], [0], [-1], [-1], [[0.5, 0, 0]], [[1, 1, 1, 1]], np.zeros((0, 4)), [0])
    P = Poses(np.array([[0.0, 0, 3.0]]), EYE[None])
    o = oracle.OracleScene(m, P)
    spec = scenes.LidarSpec(np.array([[0, 0, 1.0]]), range_min=2.5, range_max=10.0)
    r, g, raw = o.lidar_scan(spec, np.zeros(3), EYE)
    assert raw[0] == 2.5 and r[0] == np.float32(2.5) and g[0] == 0
    spec = scenes.LidarSpec(np.array([[0, 0, 1.0]]), range_min=2.5000001, range_max=10.0)
    r, g, raw = o.lidar_scan(spec, np.zeros(3), EYE)
    assert r[0] == -1.0 and g[0] == 0  # filtered, geomid kept
    spec = scenes.LidarSpec(np.array([[0, 0, 1.0]]), cutoff=1.0)
    r, g, raw = o.lidar_scan(spec, np.zeros(3), EYE)
    assert raw[0] == 1.0 and r[0] == 1.0  # rangefinder cutoff: min(cutoff, v)
    spec = scenes.LidarSpec(np.array([[0, 0, -1.0]]))
    r, g, raw = o.lidar_scan(spec, np.zeros(3), EYE)
    assert raw[0] == -1.0 and r[0] == -1.0 and g[0] == -1  # miss


@pytest.mark.parametrize("gtype,size,pnt,vec,expect", [
    (GEOM_SPHERE, (0.5, 0, 0), (0, 0, -3), (0, 0, 1), 2.5),
    (GEOM_SPHERE, (0.5, 0, 0), (0, 0, 0), (0, 0, 2), 0.25),          # inside: far wall, units of |vec|
    (GEOM_SPHERE, (0.5, 0, 0), (0, 0, 3), (0, 0, 1), -1),            # behind
    (GEOM_PLANE, (0, 0, 0.01), (1, 2, 3), (0, 0, -1), 3.0),
    (GEOM_PLANE, (0, 0, 0.01), (1, 2, -3), (0, 0, 1), -1),           # one-sided
    (GEOM_PLANE, (1, 1, 0.01), (2, 0, 3), (0, 0, -1), -1),           # outside the rendered rectangle
    (GEOM_BOX, (0.5, 0.25, 0.125), (-3, 0, 0), (1, 0, 0), 2.5),
    (GEOM_BOX, (0.5, 0.25, 0.125), (0, 0, 0), (0, 1, 0), 0.25),      # inside: exit face
    (GEOM_BOX, (0.5, 0.25, 0.125), (-3, 0.3, 0), (1, 0, 0), -1),
    (GEOM_CAPSULE, (0.2, 0.5, 0), (-2, 0, 0), (1, 0, 0), 1.8),       # round side
    (GEOM_CAPSULE, (0.2, 0.5, 0), (0, 0, 3), (0, 0, -1), 2.3),       # top cap
    (GEOM_CAPSULE, (0.2, 0.5, 0), (0, 0, -3), (0, 0, 1), 2.3),       # bottom cap
    (GEOM_CYLINDER, (0.2, 0.5, 0), (-2, 0, 0), (1, 0, 0), 1.8),
    (GEOM_CYLINDER, (0.2, 0.5, 0), (0.1, 0, 3), (0, 0, -1), 2.5),    # flat cap
    (GEOM_CYLINDER, (0.2, 0.5, 0), (0.3, 0, 3), (0, 0, -1), -1),
    (GEOM_ELLIPSOID, (0.5, 0.25, 0.125), (0, 0, -2), (0, 0, 1), 1.875),
    (GEOM_ELLIPSOID, (0.5, 0.25, 0.125), (-2, 0, 0), (1, 0, 0), 1.5),
])
def test_primitives_analytic(oracle, gtype, size, pnt, vec, expect):
    x = oracle.ray_geom(np.zeros(3), EYE, np.array(size, float), np.array(pnt, float),
                        np.array(vec, float), gtype)
    if expect < 0:
        assert x == -1
    else:
        assert abs(x - expect) < 1e-12


def _implicit_residual(gtype, size, p):
    """|f(p)| for the implicit surface of a primitive at the origin (local frame)."""
    x, y, z = p
    if gtype == GEOM_SPHERE:
        return abs(np.linalg.norm(p) - size[0])
    if gtype == GEOM_ELLIPSOID:
        return abs((x / size[0]) ** 2 + (y / size[1]) ** 2 + (z / size[2]) ** 2 - 1)
    if gtype == GEOM_BOX:
        q = np.abs(p) - size
        return abs(max(q))
    if gtype == GEOM_CAPSULE:
        zc = np.clip(z, -size[1], size[1])
        return abs(np.linalg.norm([x, y, z - zc]) - size[0])
    if gtype == GEOM_CYLINDER:
        return abs(max(np.hypot(x, y) - size[0], abs(z) - size[1]))
    raise ValueError


@pytest.mark.parametrize("gtype", [GEOM_SPHERE, GEOM_ELLIPSOID, GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER])
def test_primitives_random_vs_implicit(oracle, gtype):
    """Independent check: every reported hit lies on the implicit surface, is the first crossing
    along the ray (dense sampling finds no earlier sign change), and misses are real misses."""
    rng = np.random.default_rng(5)
    rots = random_rotations(rng, 60)
    nhit = 0
    for k in range(60):
        size = rng.uniform(0.2, 1.0, 3)
        pos = rng.uniform(-1, 1, 3)
        mat = rots[k]
        pnt = pos + rng.normal(size=3) * 3.0
        vec = (pos + rng.normal(size=3) * 0.4) - pnt
        vec *= rng.uniform(0.3, 2.0)
        x = oracle.ray_geom(pos, mat.reshape(9), size, pnt, vec, gtype)

        def inside(t):
            p = mat.T @ (pnt + t * vec - pos)
            if gtype == GEOM_SPHERE:
                return np.linalg.norm(p) < size[0]
            if gtype == GEOM_ELLIPSOID:
                return ((p / size) ** 2).sum() < 1
            if gtype == GEOM_BOX:
                return np.all(np.abs(p) < size)
            if gtype == GEOM_CAPSULE:
                zc = np.clip(p[2], -size[1], size[1])
                return np.linalg.norm([p[0], p[1], p[2] - zc]) < size[0]
            return np.hypot(p[0], p[1]) < size[0] and abs(p[2]) < size[1]

        ts = np.linspace(0, 4.0 / np.linalg.norm(vec) * 3, 4000)
        ins = np.array([inside(t) for t in ts])
        if x >= 0:
            nhit += 1
            p = mat.T @ (pnt + x * vec - pos)
            assert _implicit_residual(gtype, size, p) < 1e-9
            if not ins[0]:
                first = ts[np.argmax(ins)] if ins.any() else np.inf
                assert x <= first + 1e-9 and x >= first - (ts[1] - ts[0]) - 1e-9
        else:
            assert not ins.any()
    assert nhit > 10


def test_tie_lowest_geom_id_wins(oracle):
    """A.5: coincident geoms (converter's visual + collision duplicates) -> lowest id."""
    m = Model([GEOM_BOX, GEOM_BOX, GEOM_BOX], [0, 0, 0], [3, 2, 0], [-1] * 3, [-1] * 3,
              [[0.5, 0.5, 0.5]] * 3, np.ones((3, 4)), np.zeros((0, 4)), [0])
    P = Poses(np.array([[0.0, 0, 3.0]] * 2 + [[0, 0, 6.0]]), np.tile(EYE, (3, 1)))
    o = oracle.OracleScene(m, P)
    d, g = o.ray(np.zeros(3), np.array([0, 0, 1.0]))
    assert d == 2.5 and g == 0
    # group mask removes geom 0 -> geom 1 wins the (now single) closest hit
    d, g = o.ray(np.zeros(3), np.array([0, 0, 1.0]), geomgroup=[1, 1, 1, 0, 1, 1])
    assert d == 2.5 and g == 1


def test_eliminate_rules(oracle):
    """A.2: bodyexclude, invisible rgba / material alpha, flg_static, geomgroup."""
    m = Model([GEOM_SPHERE] * 5, [1, 2, 3, 3, 0], [0, 1, 2, 7, -3], [-1, -1, 0, -1, -1], [-1] * 5,
              [[0.2, 0, 0]] * 5, [[1, 1, 1, 1], [1, 1, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 1, 1]],
              [[1, 1, 1, 0]], [0, 1, 0, 3])
    pos = np.array([[0, 0, 1.0], [0, 0, 2.0], [0, 0, 3.0], [0, 0, 4.0], [0, 0, 5.0]])
    o = oracle.OracleScene(m, Poses(pos, np.tile(EYE, (5, 1))))
    z = np.array([0, 0, 1.0])
    assert o.ray(np.zeros(3), z)[1] == 0
    assert o.ray(np.zeros(3), z, bodyexclude=1)[1] == 3          # 1: alpha 0, 2: material alpha 0
    assert o.ray(np.zeros(3), z, flg_static=0)[1] == 0           # body 1 welds to itself
    assert o.ray(np.zeros(3), z, flg_static=0, bodyexclude=1)[1] == 3
    assert o.ray(np.zeros(3), z, bodyexclude=1, geomgroup=[1, 1, 1, 1, 1, 0])[1] == 4  # group 7 -> 5
    assert o.ray(np.zeros(3), z, bodyexclude=1, geomgroup=[0, 1, 1, 1, 1, 0])[1] == -1  # group -3 -> 0


def test_mesh_bvh_equals_brute_force(oracle):
    cs = scenes.clutter500(n_el=8, n_az=64)
    o = oracle.OracleScene(cs.model, cs.poses)
    rng = np.random.default_rng(3)
    mesh_geoms = np.nonzero(cs.model.geom_type == 7)[0]
    nhit = 0
    for g in mesh_geoms[:20]:
        c = cs.poses.geom_xpos[0, g]
        for _ in range(40):
            pnt = c + rng.normal(size=3) * 3
            vec = (c + rng.normal(size=3) * 0.3) - pnt
            a, b = o.ray_mesh(int(g), pnt, vec), o.ray_mesh(int(g), pnt, vec, brute=True)
            assert a == b
            nhit += a >= 0
    assert nhit > 200


def test_multiray_equals_ray_loop(oracle):
    cs = scenes.clutter500(n_el=16, n_az=128)
    o = oracle.OracleScene(cs.model, cs.poses)
    vec = cs.lidar.local_dir[::7]
    for gg, fs in ((None, 1), ([1, 0, 1, 1, 0, 0], 0)):
        d, g = o.multi_ray(cs.sensor_pos, vec, geomgroup=gg, flg_static=fs, bodyexclude=63, nthread=2)
        for i in range(0, vec.shape[0], 5):
            di, gi = o.ray(cs.sensor_pos, vec[i], geomgroup=gg, flg_static=fs, bodyexclude=63)
            assert (di, gi) == (d[i], g[i])
    d2, g2 = o.multi_ray(cs.sensor_pos, vec, bodyexclude=63, cutoff=5.0)
    d, g = o.multi_ray(cs.sensor_pos, vec, bodyexclude=63)
    far = d > 5.0
    assert np.all(d2[far] == -1) and np.all(g2[far] == -1)
    assert np.array_equal(d2[~far], d[~far])


def test_multiray_culls_leave_the_result_unchanged(oracle):
    """mj_multiRay's per-body azimuth / elevation rectangle (and the per-geom
    bounding-sphere cull) are pure accelerators -- every variant must return mj_ray's answer, from
    origins inside and outside bodies' boxes, with masks and a cutoff, and on a scene whose bodies
    are compact (one object per body: the cull really prunes) as well as scattered (clutter500)."""
    cs = scenes.clutter500(n_el=16, n_az=128)
    o = oracle.OracleScene(cs.model, cs.poses)
    rng = np.random.default_rng(3)
    origins = [cs.sensor_pos, np.array([7.0, -3.0, 0.4]), np.array([-15.0, 12.0, 5.5]),
               cs.poses.geom_xpos[0, 40] + 0.01]  # the last one sits inside a geom
    vec = np.concatenate([cs.lidar.local_dir[::5], rng.normal(size=(300, 3)),
                          [[0, 0, 1.0], [0, 0, -1.0], [1.0, 0, 0], [-1.0, 1e-12, 0]]])
    for pnt in origins:
        for gg, fs, be, cut in ((None, 1, 63, 1e10), ([1, 0, 1, 1, 0, 0], 0, -1, 1e10), (None, 1, 5, 9.0)):
            ref = o.multi_ray_cull(pnt, vec, oracle.CULL_NONE, geomgroup=gg, flg_static=fs, bodyexclude=be,
                                   cutoff=cut)
            for cull in (oracle.CULL_BODY_ANGULAR, oracle.CULL_GEOM_SPHERE):
                d, g = o.multi_ray_cull(pnt, vec, cull, geomgroup=gg, flg_static=fs, bodyexclude=be,
                                        cutoff=cut, nthread=3)
                assert np.array_equal(d, ref[0]) and np.array_equal(g, ref[1]), (pnt, cull)
        for i in range(0, vec.shape[0], 37):
            di, gi = o.ray(pnt, vec[i], bodyexclude=63)
            d, g = o.multi_ray_cull(pnt, vec[i:i + 1], oracle.CULL_BODY_ANGULAR, bodyexclude=63)
            assert (di, gi) == (d[0], g[0])
    # one object per body: arm worlds' fixture bodies are compact, the rectangle prunes there
    ws = scenes.arm_worlds(nworld=2, n_rays=256)
    o2 = oracle.OracleScene(ws.model, ws.poses, world=1)
    for cull in (oracle.CULL_BODY_ANGULAR, oracle.CULL_GEOM_SPHERE):
        a = o2.lidar_scan_multiray(ws.lidar, ws.sensor_pos[1], ws.sensor_mat[1], cull=cull)
        b = o2.lidar_scan(ws.lidar, ws.sensor_pos[1], ws.sensor_mat[1])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # lidar_scan_multiray == lidar_scan on the cluttered scene incl. cutoff / range filter
    spec = scenes.LidarSpec(cs.lidar.local_dir, bodyexclude=63, cutoff=12.0, range_min=1.0, range_max=11.0)
    pos, mat = cs.sensor_poses(2)
    b = o.lidar_scan(spec, pos[1], mat[1], nthread=2)
    for cull in (oracle.CULL_NONE, oracle.CULL_BODY_ANGULAR, oracle.CULL_GEOM_SPHERE):
        a = o.lidar_scan_multiray(spec, pos[1], mat[1], cull=cull, nthread=2)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_camera_intrinsics_and_depth(oracle):
    """ref:src/mujoco_cameras.cpp:119-124 (f = 1/tan(fovy/2) * H/2) and :222-240."""
    K = oracle.camera_intrinsics(58.0, 1280, 720)
    assert abs(K[0] - 649.46) < 0.01 and K[0] == K[4] and K[2] == 640 and K[5] == 360 and K[8] == 1
    K = oracle.camera_intrinsics(58.0, 640, 480)
    assert abs(K[0] - 432.97) < 0.01
    gl = scenes.gl_depth(1, 64, 48)
    z = oracle.depth_linearise_flip(gl, 0.01, 50.0, 1.0)
    near, far = np.float32(0.01), np.float32(50.0)
    scale = np.float32(1.0) - near / far
    ref = (near / (np.float32(1.0) - gl[0] * scale))[::-1]
    np.testing.assert_array_equal(z[0], ref.astype(np.float32))
    assert z[0][-1, 0] == near                       # d = 0 -> near plane (GL row 0 = bottom row)
    assert abs(z[0][-1, 2] - 50.0) < 0.05            # d = 1 -> far plane
    cloud = oracle.depth_to_cloud(z, K, 3)
    u, v = 10, 7
    zz = z[0, v, u]
    fx = np.float32(K[0])
    np.testing.assert_allclose(cloud[0, v, u], [(u - 320) * zz / fx, (v - 240) * zz / fx, zz], rtol=1e-6)
    z2 = z.copy()
    z2[0, 3, 3] = np.inf
    c2 = oracle.depth_to_cloud(z2, K, 4)
    assert np.isnan(c2[0, 3, 3, :3]).all() and c2[0, 3, 3, 3] == 0
