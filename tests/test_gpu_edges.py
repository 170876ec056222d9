"""GPU tier: edge cases of the ray path against the oracle (bit-exact), and the small N2 kernel."""
import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import engine, scenes
from mujoco_ros2_simulation_b200.model import (GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID,
                                               GEOM_PLANE, GEOM_SPHERE, Model, Poses, random_rotations)

pytestmark = pytest.mark.gpu
EYE = np.eye(3).reshape(9)


def _check(oracle, model, poses, pnt, vec, **kw):
    s = engine.Scene(model)
    s.update_poses(poses)
    d2, g2 = s.multiray(pnt, vec, **kw)
    o = oracle.OracleScene(model, poses)
    d, g = o.multi_ray(pnt, vec, nthread=4, **kw)
    assert np.array_equal(g, g2), np.nonzero(g != g2)[0][:5]
    assert np.array_equal(d, d2)
    s.close()
    return d, g


def _prims(rng, n, types, centre_scale=5.0, offset=0.0):
    t = rng.choice(types, size=n)
    size = rng.uniform(0.1, 1.0, size=(n, 3))
    pos = rng.normal(size=(n, 3)) * centre_scale + offset
    mat = random_rotations(rng, n).reshape(n, 9)
    m = Model(t, np.zeros(n, int), np.zeros(n, int), -np.ones(n, int), -np.ones(n, int), size,
              np.ones((n, 4)), np.zeros((0, 4)), [0])
    return m, Poses(pos, mat)


def test_origin_inside_geoms(oracle):
    """rays that START inside primitives hit the far wall (ray_quad's second root, box exit face)"""
    rng = np.random.default_rng(21)
    types = [GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX]
    m, P = _prims(rng, 40, types, centre_scale=3.0)
    for g in (0, 7, 19, 33):
        pnt = P.geom_xpos[0, g] + rng.normal(size=3) * 0.02      # well inside geom g
        vec = rng.normal(size=(512, 3))
        d, gid = _check(oracle, m, P, pnt, vec)
        assert (gid >= 0).all() and (d > 0).all()


def test_far_from_world_origin(oracle):
    """coordinates around 1e4 m: the fp32 culling must stay conservative (origin slack, box pad)"""
    rng = np.random.default_rng(22)
    m, P = _prims(rng, 120, [GEOM_SPHERE, GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID],
                  centre_scale=8.0, offset=np.array([1.0e4, -2.0e4, 3.0e3]))
    pnt = np.array([1.0e4, -2.0e4, 3.0e3]) + np.array([0.3, -0.2, 0.1])
    el = rng.uniform(-0.5, 0.5, 4096)
    az = np.sort(rng.uniform(-np.pi, np.pi, 4096))
    vec = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], 1)
    d, g = _check(oracle, m, P, pnt, vec)
    assert (g >= 0).mean() > 0.05


def test_tiny_and_huge_geoms_and_nonunit_rays(oracle):
    rng = np.random.default_rng(23)
    m, P = _prims(rng, 60, [GEOM_SPHERE, GEOM_BOX, GEOM_CYLINDER], centre_scale=4.0)
    m.geom_size[:10] *= 1e-3         # millimetre geoms
    m.geom_size[10:14] *= 30.0       # geoms that swallow the origin
    vec = rng.normal(size=(2048, 3)) * 10 ** rng.uniform(-3, 3, size=(2048, 1))
    _check(oracle, m, P, np.array([0.1, 0.2, -0.1]), vec)


def test_planes_finite_tilted_and_many(oracle):
    """several planes (kept outside the BVH), finite and tilted, one-sided test"""
    rng = np.random.default_rng(24)
    n = 6
    R = random_rotations(rng, n).reshape(n, 9)
    R[0] = EYE
    size = np.array([[0, 0, .01], [2, 3, .01], [0, 1.5, .01], [4, 0, .01], [1, 1, .01], [0, 0, .01]], float)
    pos = rng.normal(size=(n, 3)) * 2
    pos[0] = (0, 0, -1.0)
    m = Model([GEOM_PLANE] * n, np.zeros(n, int), np.zeros(n, int), -np.ones(n, int), -np.ones(n, int),
              size, np.ones((n, 4)), np.zeros((0, 4)), [0])
    vec = rng.normal(size=(4096, 3))
    d, g = _check(oracle, m, Poses(pos, R), np.array([0.2, -0.1, 0.5]), vec)
    assert len(set(g.tolist())) >= 4


def test_incoherent_rays_and_overflowing_candidates(oracle):
    """random directions (cone test disabled per group) on a dense cluster of overlapping boxes:
    more candidates than the per-group list holds -> drained in several passes"""
    rng = np.random.default_rng(25)
    n = 400
    m, P = _prims(rng, n, [GEOM_BOX, GEOM_SPHERE], centre_scale=0.6)
    m.geom_size[:] = rng.uniform(0.5, 1.5, size=(n, 3))        # everything overlaps everything
    vec = rng.normal(size=(1024, 3))
    _check(oracle, m, P, np.array([6.0, 0.5, 0.2]), vec)
    _check(oracle, m, P, np.array([6.0, 0.5, 0.2]), vec, cutoff=5.0)
    # coherent fan through the same cluster
    az = np.linspace(np.pi - 0.4, np.pi + 0.4, 2048)
    fan = np.stack([np.cos(az), np.sin(az), 0 * az], 1)
    _check(oracle, m, P, np.array([6.0, 0.5, 0.2]), fan)


def test_mesh_heavy_scene_in_place_and_split_paths(oracle):
    """only mesh geoms, overlapping: small launch (in-place mesh tracing) and large launch (pair
    kernels, many pairs per ray, ties between coincident mesh instances) agree with the oracle"""
    cs = scenes.clutter500(n_el=16, n_az=128)
    m = cs.model
    keep = np.nonzero(m.geom_type == 7)[0][:40]
    n = len(keep) + 10
    dup = np.concatenate([keep, keep[:10]])                     # coincident copies -> geomid ties
    model = Model(m.geom_type[dup], np.zeros(n, int), np.zeros(n, int), -np.ones(n, int), m.geom_dataid[dup],
                  m.geom_size[dup], np.ones((n, 4)), np.zeros((0, 4)), [0], m.mesh_vertadr, m.mesh_vertnum,
                  m.mesh_faceadr, m.mesh_facenum, m.mesh_vert, m.mesh_face)
    rng = np.random.default_rng(26)
    pos = rng.uniform(-2.5, 2.5, size=(n, 3))
    pos[len(keep):] = pos[:10]
    mat = random_rotations(rng, n).reshape(n, 9)
    mat[len(keep):] = mat[:10]
    P = Poses(pos, mat)
    pnt = np.array([7.0, 0.0, 0.0])
    az = np.linspace(np.pi - 0.5, np.pi + 0.5, 400)
    el = np.linspace(-0.4, 0.4, 200)
    A, E = np.meshgrid(az, el)
    vec = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    d_small, g_small = _check(oracle, model, P, pnt, vec[:4000])     # < 65536 rays: in-place path
    d_big, g_big = _check(oracle, model, P, pnt, vec)                # 80000 rays: split path
    assert np.array_equal(d_big[:4000], d_small) and np.array_equal(g_big[:4000], g_small)
    assert (g_big >= 0).mean() > 0.3
    hit = g_big[g_big >= 0]
    assert hit.max() < len(keep)          # a coincident copy (higher id) never wins a tie


@pytest.mark.parametrize("env", [{"B2_MESH_COOP_G": "8"}, {"B2_MESH_COOP_G": "16"}, {"B2_MESH_COOP_G": "32"},
                                 {"B2_MESH_COOP_RAYS": "0"}])
def test_mesh_pairs_of_a_small_launch_every_group_width(oracle, env):
    """a launch of one or two scans traces its (ray, mesh) pairs with a GROUP of lanes per pair (shared
    stack, several nodes / triangles per round): every group width, and the one-lane kernel, give the
    oracle's bits -- on the cluttered scene (one 64x2048 scan) and on the mesh-only scene with
    coincident instances"""
    import os
    old = {k: os.environ.get(k) for k in ("B2_MESH_COOP_G", "B2_MESH_COOP_RAYS")}
    os.environ.update(env)
    try:
        cs = scenes.clutter500()
        sc = engine.Scene(cs.model, nworld=1)
        sc.update_poses(cs.poses)
        lid = engine.Lidar(sc, cs.lidar)
        pos, mat = cs.sensor_poses(3, seed=11)
        out = lid.cast(pos[2:3], mat[2:3], want=("ranges", "geomid", "dist"))
        o = oracle.OracleScene(cs.model, cs.poses)
        r, g, raw = o.lidar_scan(cs.lidar, pos[2], mat[2], nthread=8)
        assert np.array_equal(out["geomid"][0], g) and np.array_equal(out["dist"][0], raw)
        assert np.array_equal(out["ranges"][0], r)
        assert (cs.model.geom_type[g[g >= 0]] == 7).mean() > 0.02     # meshes do win rays here
        lid.close()
        sc.close()
        test_mesh_heavy_scene_in_place_and_split_paths(oracle)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_camera_frame_same_bits_on_every_mesh_path():
    """a 320x240 ray-cast depth frame (76 800 primary rays: the group mesh kernel in camera mode) has the
    bits of the one-lane mesh kernel and of the in-place mesh tracing"""
    import os

    import torch
    cs = scenes.clutter500(n_el=16, n_az=128)
    sc = engine.Scene(cs.model, nworld=1)
    sc.update_poses(cs.poses)
    cam = engine.DepthCamera(scenes.CameraSpec(320, 240, 58.0, 0.01, 50.0, 1.0, 12, geomgroup=(1,) * 6), scene=sc)
    dev = torch.device("cuda", 0)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    d_pos = torch.from_numpy(np.array([[3.0, 0.0, 1.5]])).to(dev)
    d_mat = torch.from_numpy(R).to(dev)
    frames = {}
    keys = ("B2_MESH_COOP_RAYS", "B2_MESH_COOP_G", "B2_SPLIT")
    old = {k: os.environ.get(k) for k in keys}
    try:
        for name, env in (("group", {}), ("group16", {"B2_MESH_COOP_G": "16"}), ("one_lane", {"B2_MESH_COOP_RAYS": "0"}),
                          ("in_place", {"B2_SPLIT": "0"})):
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            gl = torch.full((1, 240, 320), -1.0, dtype=torch.float32, device=dev)
            cam.render_depth_device(1, d_pos, d_mat, gl)
            sc.sync()
            frames[name] = gl.cpu().numpy()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    f = frames["group"]
    assert 0.05 < (f < 1.0).mean() and f.min() >= 0.0 and f.max() <= 1.0
    for name in ("group16", "one_lane", "in_place"):
        assert np.array_equal(frames[name], f), name
    cam.close()
    sc.close()


def test_rgb_flip_kernel():
    import torch
    dev = torch.device("cuda", 0)
    for (w, h) in ((640, 480), (322, 7)):
        cam = engine.DepthCamera(scenes.CameraSpec(w, h, 58.0))
        gen = torch.Generator(device=dev).manual_seed(3)
        rgb = torch.randint(0, 256, (3, h, w, 3), dtype=torch.uint8, device=dev, generator=gen)
        out = torch.empty_like(rgb)
        cam.rgb_flip_device(3, rgb, out)
        torch.cuda.synchronize()
        assert torch.equal(out, rgb.flip(1))


def test_large_scene_unstaged_bvh_and_deep_tree(oracle, emu):
    """6000 geoms: the geom BVH (3 levels of fanout-32 nodes, ~200 KB) no longer fits the per-CTA
    shared-memory stage, so the cast reads it through the read-only path; several lidars share
    one scene."""
    rng = np.random.default_rng(31)
    n = 6000
    m, P = _prims(rng, n, [GEOM_SPHERE, GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID],
                  centre_scale=25.0)
    m.geom_size[:] *= 0.6
    st = emu.bvh_stats(m, P)
    assert st["depth"] >= 3 and st["bvh_slots"] * 32 > 96 * 1024
    s = engine.Scene(m)
    s.update_poses(P)
    o = oracle.OracleScene(m, P)
    specs = [scenes.LidarSpec(scenes.spinning_lidar_dirs(8, 512), tile_w=512),
             scenes.LidarSpec(scenes.planar_lidar_dirs(0.0, 2 * np.pi / 1000, 1000), range_min=0.5, range_max=40.0),
             scenes.LidarSpec(scenes.spinning_lidar_dirs(4, 64), tile_w=64, geomgroup=[1, 0, 0, 0, 0, 0])]
    lids = [engine.Lidar(s, sp) for sp in specs]
    pos = np.array([[0.5, -0.3, 0.2], [10.0, 5.0, -3.0]])
    mat = np.stack([EYE, random_rotations(rng, 1)[0].reshape(9)])
    for sp, lid in zip(specs, lids):
        out = lid.cast(pos, mat, want=("ranges", "geomid", "dist"))
        for i in range(2):
            r, g, raw = o.lidar_scan(sp, pos[i], mat[i], nthread=8)
            assert np.array_equal(out["geomid"][i], g)
            assert np.array_equal(out["dist"][i], raw) and np.array_equal(out["ranges"][i], r)
    assert (out["geomid"] >= 0).any()


def test_two_threads_share_one_scene(oracle):
    """the plugin's lidar and camera threads call into the same scene concurrently"""
    import threading
    cs = scenes.clutter500(n_el=16, n_az=128)
    s = engine.Scene(cs.model)
    s.update_poses(cs.poses)
    lid_a = engine.Lidar(s, cs.lidar)
    lid_b = engine.Lidar(s, scenes.LidarSpec(cs.lidar.local_dir[::3], bodyexclude=63))
    pos, mat = cs.sensor_poses(40)     # 40 x 2048 rays: the split mesh path with shared scratch
    ref_a = lid_a.cast(pos, mat, want=("geomid", "dist"))
    ref_b = lid_b.cast(pos, mat, want=("geomid", "dist"))
    errs = []

    def work(lid, ref):
        try:
            for _ in range(6):
                out = lid.cast(pos, mat, want=("geomid", "dist"))
                assert np.array_equal(out["dist"], ref["dist"]) and np.array_equal(out["geomid"], ref["geomid"])
                s.update_poses(cs.poses)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(lid_a, ref_a)), threading.Thread(target=work, args=(lid_b, ref_b))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    o = oracle.OracleScene(cs.model, cs.poses)
    r, g, raw = o.lidar_scan(cs.lidar, pos[3], mat[3], nthread=8)
    assert np.array_equal(ref_a["dist"][3], raw) and np.array_equal(ref_a["geomid"][3], g)
