"""CPU tier: the product's traversal code (csrc/ray_core.cuh + scene_build/bvh_build), compiled for
the host by tests/emu with a 1-lane warp, against the oracle -- bit-exact distances and geom ids.
This checks the LOGIC that runs on the B200 (BVH build, refit boxes, conservative fp32 culling,
packet traversal control flow, deferred mesh candidates, tie rule, epilogues) before any GPU time
is spent; the GPU tier (test_gpu_parity.py) repeats it through the C ABI on the device."""
import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import scenes
from mujoco_ros2_simulation_b200.model import GEOM_MESH


@pytest.fixture(scope="module")
def clutter():
    return scenes.clutter500(n_el=32, n_az=256)


def test_emu_test_robot(oracle, emu):
    rs = scenes.robot_scene()
    o = oracle.OracleScene(rs.model, rs.poses)
    r, g, raw = o.lidar_scan(rs.lidar, rs.lidar_pos, rs.lidar_mat)
    r2, g2, raw2, xyz = emu.lidar(rs.model, rs.poses, rs.lidar, rs.lidar_pos, rs.lidar_mat)
    assert np.array_equal(r, r2) and np.array_equal(g, g2) and np.array_equal(raw, raw2)
    assert g2.tolist() == [0] * 13 + [2] * 11          # ties: collision box wins over visual copy
    ref_xyz = (rs.lidar.local_dir * raw[:, None]).astype(np.float32)
    assert np.array_equal(xyz, ref_xyz)
    rs360 = scenes.robot_scene(360)                    # cfg1
    r, g, raw = o.lidar_scan(rs360.lidar, rs360.lidar_pos, rs360.lidar_mat)
    r2, g2, raw2, xyz = emu.lidar(rs360.model, rs360.poses, rs360.lidar, rs360.lidar_pos, rs360.lidar_mat)
    assert np.array_equal(r, r2) and np.array_equal(g, g2) and np.array_equal(raw, raw2)
    assert (g >= 0).sum() > 30 and (r == -1).sum() > 200
    assert np.isnan(xyz[r2 == -1]).all()


@pytest.mark.parametrize("gg,fs", [(None, 1), ([1, 0, 1, 1, 0, 0], 1), (None, 0), ([1, 0, 1, 1, 0, 0], 0)])
def test_emu_clutter_masks(oracle, emu, clutter, gg, fs):
    cs = clutter
    o = oracle.OracleScene(cs.model, cs.poses)
    vec = cs.lidar.local_dir
    d, g = o.multi_ray(cs.sensor_pos, vec, geomgroup=gg, flg_static=fs, bodyexclude=63, nthread=4)
    d2, g2 = emu.multiray(cs.model, cs.poses, cs.sensor_pos, vec, geomgroup=gg, flg_static=fs,
                          bodyexclude=63)
    assert np.array_equal(d, d2)
    assert np.array_equal(g, g2)
    assert (cs.model.geom_type[g[g >= 0]] == GEOM_MESH).sum() > 50 or fs == 0


def test_emu_ties_resolved_to_lowest_id(oracle, emu, clutter):
    """every 10th clutter geom has a coincident copy with the next id: it must never be reported."""
    cs = clutter
    m = cs.model
    dup = np.zeros(m.ngeom, bool)
    for i in range(1, m.ngeom):
        dup[i] = (m.geom_type[i] == m.geom_type[i - 1]
                  and np.array_equal(cs.poses.geom_xpos[0, i], cs.poses.geom_xpos[0, i - 1])
                  and m.geom_group[i] == m.geom_group[i - 1])
    assert dup.sum() > 20
    d2, g2 = emu.multiray(m, cs.poses, cs.sensor_pos, cs.lidar.local_dir, bodyexclude=63)
    hit = g2[g2 >= 0]
    assert not dup[hit].any()
    assert dup[hit + 1 - (hit + 1 >= m.ngeom)].any()   # the originals of some duplicates are hit


def test_emu_bodyexclude_matters(oracle, emu, clutter):
    cs = clutter
    d_in, g_in = emu.multiray(cs.model, cs.poses, cs.sensor_pos, cs.lidar.local_dir[::9], bodyexclude=-1)
    assert np.isin(g_in, np.nonzero(cs.model.geom_bodyid == 63)[0]).all()   # housing hit from inside
    o = oracle.OracleScene(cs.model, cs.poses)
    d, g = o.multi_ray(cs.sensor_pos, cs.lidar.local_dir[::9], bodyexclude=-1)
    assert np.array_equal(d, d_in) and np.array_equal(g, g_in)


def test_emu_refit_world_with_foreign_topology(oracle, emu, clutter):
    """S5: BVH topology built on world 0, boxes refit to a jittered world."""
    cs = clutter
    P = scenes.clutter_worlds(cs, 3)
    o = oracle.OracleScene(cs.model, P, world=2)
    vec = cs.lidar.local_dir[::3]
    d, g = o.multi_ray(cs.sensor_pos, vec, bodyexclude=63, nthread=4)
    d2, g2 = emu.multiray(cs.model, P, cs.sensor_pos, vec, bodyexclude=63, world=2, topo_world=0)
    assert np.array_equal(d, d2) and np.array_equal(g, g2)


def test_emu_jittered_sensor_and_cutoff(oracle, emu, clutter):
    cs = clutter
    pos, mat = cs.sensor_poses(3)
    o = oracle.OracleScene(cs.model, cs.poses)
    spec = scenes.LidarSpec(cs.lidar.local_dir, bodyexclude=63, cutoff=12.0, range_min=1.0, range_max=11.0)
    for i in (1, 2):
        r, g, raw = o.lidar_scan(spec, pos[i], mat[i], nthread=4)
        r2, g2, raw2, _ = emu.lidar(cs.model, cs.poses, spec, pos[i], mat[i])
        assert np.array_equal(r, r2) and np.array_equal(g, g2) and np.array_equal(raw, raw2)
        assert (r == -1).any() and (raw == 12.0).any()
    d, g = o.multi_ray(pos[1], cs.lidar.local_dir[::5], bodyexclude=63, cutoff=6.0)
    d2, g2 = emu.multiray(cs.model, cs.poses, pos[1], cs.lidar.local_dir[::5], bodyexclude=63, cutoff=6.0)
    assert np.array_equal(d, d2) and np.array_equal(g, g2)


def test_emu_arm_worlds(oracle, emu):
    ws = scenes.arm_worlds(nworld=6, n_rays=256)
    for w in range(6):
        o = oracle.OracleScene(ws.model, ws.poses, world=w)
        r, g, raw = o.lidar_scan(ws.lidar, ws.sensor_pos[w], ws.sensor_mat[w])
        r2, g2, raw2, _ = emu.lidar(ws.model, ws.poses, ws.lidar, ws.sensor_pos[w], ws.sensor_mat[w], world=w)
        assert np.array_equal(r, r2) and np.array_equal(g, g2) and np.array_equal(raw, raw2)
    assert (g == 0).any()   # the tilted scan plane reaches the ground plane


def test_emu_degenerate_inputs(oracle, emu, clutter):
    cs = clutter
    vec = np.array([[0, 0, 0.0], [1e-20, 0, 0], [0, 0, -1.0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0],
                    [1e-9, 0, -1.0], [3.0, 4.0, 0.0]])
    o = oracle.OracleScene(cs.model, cs.poses)
    d, g = o.multi_ray(cs.sensor_pos, vec, bodyexclude=63)
    d2, g2 = emu.multiray(cs.model, cs.poses, cs.sensor_pos, vec, bodyexclude=63)
    assert np.array_equal(d, d2) and np.array_equal(g, g2)
    assert d[0] == -1 and g[0] == -1 and d[2] == 1.5 and g[2] == 0      # straight down: plane at 1.5
    assert abs(d[7] * 5.0 - d[3] * 1.0) < 30                            # non-unit vec: units of |vec|
    # empty ray list
    d3, g3 = emu.multiray(cs.model, cs.poses, cs.sensor_pos, np.zeros((0, 3)), bodyexclude=63)
    assert d3.shape == (0,)


def test_emu_mesh_leaf_sizes_agree(emu, clutter):
    cs = clutter
    vec = cs.lidar.local_dir[::2]
    ref = emu.multiray(cs.model, cs.poses, cs.sensor_pos, vec, bodyexclude=63, mesh_leaf=1)
    for leaf in (2, 4):
        out = emu.multiray(cs.model, cs.poses, cs.sensor_pos, vec, bodyexclude=63, mesh_leaf=leaf)
        assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])


def test_tile_cone_is_conservative(emu):
    """collect phase: a box touched by any ray of a tile is never rejected by the tile cone, and
    its ordering key never exceeds a ray's entry distance (random tiles x random boxes, incl.
    boxes containing the origin, behind it, huge walls and far-from-origin coordinates)."""
    rng = np.random.default_rng(11)
    total = 0
    for trial in range(60):
        origin = rng.uniform(-30, 30, 3) * (100.0 if trial % 7 == 0 else 1.0)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        spread = 10 ** rng.uniform(-3, -0.3)          # 0.001 .. 0.5 rad
        n = 32 if trial % 5 else 7
        vec = axis[None] + rng.normal(size=(n, 3)) * spread
        vec *= 10 ** rng.uniform(-2, 2, size=(n, 1))   # non-unit lengths
        nb = 400
        cen = origin[None] + rng.normal(size=(nb, 3)) * 10 ** rng.uniform(-1, 1.5, size=(nb, 1))
        # half of the boxes are placed along the cone so that many are grazed
        t = 10 ** rng.uniform(-1, 1.5, size=(nb // 2, 1))
        cen[: nb // 2] = origin[None] + axis[None] * t + rng.normal(size=(nb // 2, 3)) * t * spread * 2
        half = 10 ** rng.uniform(-2, 1, size=(nb, 3))
        boxes = np.concatenate([cen - half, cen + half], axis=1).astype(np.float32)
        total += emu.cone_violations(origin, vec, boxes)
    assert total == 0


def test_group_cone_on_lidar_shaped_groups(emu):
    """the device's group shape (32 azimuths x 4 scan lines of the 64x2048 lidar, 128 rays) against
    the cluttered scene's own geom boxes: the cone never drops a touched box and keeps a bounded
    surplus (measured: ~10 kept for ~5 touched per group; a 4-plane pyramid would keep ~7 but cost
    as many instructions as it saved -- profiles/r01_sweeps.txt)."""
    cs = scenes.clutter500()
    m, P = cs.model, cs.poses
    pos = P.geom_xpos.reshape(-1, m.ngeom, 3)[0]
    R = np.abs(P.geom_xmat.reshape(-1, m.ngeom, 3, 3)[0])
    ext = np.einsum("gij,gj->gi", R, np.abs(m.geom_size).max(1, keepdims=True).repeat(3, 1) * 1.8)
    sel = m.geom_type != 0
    boxes = np.concatenate([pos - ext, pos + ext], 1)[sel].astype(np.float32)
    D = cs.lidar.local_dir.reshape(64, 2048, 3)
    rng = np.random.default_rng(5)
    bad = kept = touched = 0
    for _ in range(120):
        row, col = int(rng.integers(0, 16)), int(rng.integers(0, 64))
        vec = D[row * 4:row * 4 + 4, col * 32:col * 32 + 32].reshape(-1, 3)
        b, k, t = emu.cone_counts(cs.sensor_pos, vec, boxes)
        bad, kept, touched = bad + b, kept + k, touched + t
    assert bad == 0
    assert touched > 0 and kept <= 3.0 * touched + 120


def test_two_kernel_cast_emulated(emu, oracle, monkeypatch):
    """the two-kernel cast (split_trace.cuh: collect_group -> drain_tile, fused trace_group for the
    groups the single collect pass gives up on) against the oracle: the clutter scene, and a scene
    whose candidate lists overflow (200 spheres strung along the rays) so the fall-back runs."""
    monkeypatch.setenv("B2_EMU_TWO_KERNEL", "1")
    cs = scenes.clutter500(n_el=16, n_az=128)
    pos, mat = cs.sensor_poses(2)
    o = oracle.OracleScene(cs.model, cs.poses)
    before = emu.fallbacks()
    for i in range(2):
        r, g, raw, _ = emu.lidar(cs.model, cs.poses, cs.lidar, pos[i], mat[i])
        rr, gg, rraw = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=4)
        assert np.array_equal(raw, rraw) and np.array_equal(g, gg) and np.array_equal(r, rr)
    # a string of pearls: every ray's cone holds far more than the candidate cap
    from mujoco_ros2_simulation_b200.model import GEOM_SPHERE, Model, Poses
    n = 200
    gpos = np.stack([1.0 + 0.1 * np.arange(n), 0.02 * np.sin(np.arange(n)), 0.02 * np.cos(np.arange(n))], 1)
    model = Model(np.full(n, GEOM_SPHERE), np.zeros(n, int), np.zeros(n, int), -np.ones(n, int),
                  -np.ones(n, int), np.tile([0.03, 0.0, 0.0], (n, 1)), np.ones((n, 4)), np.zeros((0, 4)), [0])
    poses = Poses(gpos, np.tile(np.eye(3).reshape(1, 9), (n, 1)))
    rng = np.random.default_rng(3)
    vec = np.array([1.0, 0, 0])[None] + rng.normal(size=(64, 3)) * 0.01
    d, g = emu.multiray(model, poses, np.zeros(3), vec)
    od, og = oracle.OracleScene(model, poses).multi_ray(np.zeros(3), vec)
    assert np.array_equal(d, od) and np.array_equal(g, og)
    assert emu.fallbacks() > before
