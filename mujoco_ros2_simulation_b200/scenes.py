"""Seeded synthetic scenes for the BASELINE.json configs.

S1 ``test_robot``   the reference's own fixture scene (ref:test/test_resources/test_robot.xml),
                    read from tests/golden/test_robot_scene.json (made by
                    tests/golden/make_test_robot_fixture.py from the reference XML);
S2 ``clutter500``   1 ground plane + 4 walls + primitives + mesh instances, 64 bodies, seed
                    20261018; 64x2048 3-D lidar at (0,0,1.5);
S3 ``arm_worlds``   S1 + ground plane replicated over W worlds with per-world joint angles;
S4 ``gl_depth``     GL-style depth buffers for the depth camera path;
S5                  S2 topology with per-world body jitter (``clutter_worlds``).

Everything is numpy and deterministic for a given seed; the same arrays feed the oracle and the
CUDA path.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

from .model import (GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID, GEOM_MESH, GEOM_PLANE,
                    GEOM_SPHERE, CameraSpec, LidarSpec, Model, Poses, axis_angle_quat, quat2mat,
                    quat_mul, random_rotations)

SEED = 20261018
_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ------------------------------------------------------------------------------------------------
# lidar ray tables
# ------------------------------------------------------------------------------------------------
def planar_lidar_dirs(min_angle: float, angle_increment: float, count: int) -> np.ndarray:
    """Ray i of a reference-style planar lidar in the lidar *body* frame.

    The converter emits replica i rotated by i*inc about body Y with site quat (0,0,0,1), so the
    ray (site +Z) is (sin(i*inc), 0, cos(i*inc)) in the body frame; ``min_angle`` lives in the body
    quaternion (ref:scripts/make_mjcf_from_robot_description.py:736-750,1120-1140)."""
    del min_angle
    a = np.arange(count, dtype=np.float64) * angle_increment
    return np.stack([np.sin(a), np.zeros_like(a), np.cos(a)], axis=1)


def lidar_body_quat(site_quat, min_angle: float) -> np.ndarray:
    """site_quat (x) [.5,.5,.5,.5] (x) RotY(min_angle)  (converter :1120-1140)."""
    q = quat_mul(np.asarray(site_quat, dtype=np.float64), np.array([0.5, 0.5, 0.5, 0.5]))
    return quat_mul(q, axis_angle_quat((0, 1, 0), min_angle))


def spinning_lidar_dirs(n_el: int = 64, n_az: int = 2048, el_min_deg: float = -25.0,
                        el_max_deg: float = 15.0) -> np.ndarray:
    """3-D lidar table: ray e*n_az + a, sensor frame Z up, azimuth CCW from +X."""
    el = np.deg2rad(np.linspace(el_min_deg, el_max_deg, n_el))
    az = np.arange(n_az, dtype=np.float64) * (2.0 * np.pi / n_az)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :],
                  np.broadcast_to(se, (n_el, n_az))], axis=2)
    return np.ascontiguousarray(d.reshape(-1, 3))


# ------------------------------------------------------------------------------------------------
# S1: the reference fixture
# ------------------------------------------------------------------------------------------------
@dataclass
class RobotScene:
    model: Model
    poses: Poses
    lidar: LidarSpec
    lidar_pos: np.ndarray
    lidar_mat: np.ndarray
    camera: CameraSpec
    tree: dict  # kinematic description (for S3 FK)


def _load_fixture() -> dict:
    with open(os.path.join(_GOLDEN, "test_robot_scene.json")) as f:
        return json.load(f)


def _arm_fk(tree: dict, q: np.ndarray, with_plane: bool):
    """Poses of the fixture's geoms for joint angles q[W,2] (both hinges about body Z at the body
    origin, ref:test/test_resources/test_robot.xml:6,9).  Returns xpos[W,G,3], xmat[W,G,9]."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    w = q.shape[0]
    gl_body = np.array(tree["geom_body"])
    gl_pos = np.array(tree["geom_local_pos"], dtype=np.float64)
    ng = gl_body.shape[0]
    off = 1 if with_plane else 0
    xpos = np.zeros((w, ng + off, 3))
    xmat = np.zeros((w, ng + off, 9))
    if with_plane:
        xpos[:, 0] = (0.0, 0.0, float(tree.get("plane_z", -0.5)))
        xmat[:, 0] = np.eye(3).reshape(9)
    forearm_off = np.array(tree["forearm_pos"], dtype=np.float64)
    for i in range(w):
        r1 = quat2mat(axis_angle_quat((0, 0, 1), q[i, 0]))
        p1 = np.zeros(3)
        p2 = p1 + r1 @ forearm_off
        r2 = r1 @ quat2mat(axis_angle_quat((0, 0, 1), q[i, 1]))
        for g in range(ng):
            if gl_body[g] == 1:
                xpos[i, g + off] = p1 + r1 @ gl_pos[g]
                xmat[i, g + off] = r1.reshape(9)
            else:
                xpos[i, g + off] = p2 + r2 @ gl_pos[g]
                xmat[i, g + off] = r2.reshape(9)
    return xpos, xmat


def robot_scene(n_rays: int | None = None) -> RobotScene:
    """S1.  ``n_rays=None``: the 24-ray lidar as shipped (min -0.3, inc 0.025);
    ``n_rays=360``: cfg1 (min_angle -pi, inc 2pi/360, range [0.05, 10])."""
    fx = _load_fixture()
    m = fx["model"]
    model = Model(geom_type=m["geom_type"], geom_bodyid=m["geom_bodyid"], geom_group=m["geom_group"],
                  geom_matid=m["geom_matid"], geom_dataid=m["geom_dataid"], geom_size=m["geom_size"],
                  geom_rgba=m["geom_rgba"], mat_rgba=np.zeros((0, 4), np.float32),
                  body_weldid=m["body_weldid"])
    poses = Poses(np.array(fx["geom_xpos"]), np.array(fx["geom_xmat"]))
    lid = fx["lidar"]
    if n_rays is None:
        dirs = np.array(lid["local_dir"], dtype=np.float64)
        mat = np.array(lid["body_xmat"], dtype=np.float64).reshape(3, 3)
        spec = LidarSpec(dirs, bodyexclude=int(lid["bodyid"]), range_min=lid["range_min"],
                         range_max=lid["range_max"])
    else:
        inc = 2.0 * np.pi / n_rays
        dirs = planar_lidar_dirs(-np.pi, inc, n_rays)
        mat = quat2mat(lidar_body_quat(lid["site_quat"], -np.pi))
        spec = LidarSpec(dirs, bodyexclude=int(lid["bodyid"]), range_min=0.05, range_max=10.0)
    cam = fx["camera"]
    camera = CameraSpec(cam["width"], cam["height"], cam["fovy"], fx["znear"], fx["zfar"],
                        fx["extent"])
    return RobotScene(model, poses, spec, np.array(lid["body_xpos"], dtype=np.float64), mat, camera,
                      fx["tree"])


# ------------------------------------------------------------------------------------------------
# S3: replicated arm worlds
# ------------------------------------------------------------------------------------------------
@dataclass
class WorldsScene:
    model: Model
    poses: Poses  # [W,...]
    lidar: LidarSpec
    sensor_pos: np.ndarray  # [W,3]
    sensor_mat: np.ndarray  # [W,9]
    scan_world: np.ndarray  # [W] i32


def arm_worlds(nworld: int = 4096, n_rays: int = 1024, seed: int = SEED, step: int = 0) -> WorldsScene:
    """S3: S1 + a ground plane (geom 0, z = -0.5) in ``nworld`` worlds, joint angles
    q ~ U(-pi, pi) per world (seed + step), one ``n_rays`` 2-pi planar lidar per world."""
    base = robot_scene()
    bm = base.model
    g = bm.ngeom + 1
    model = Model(
        geom_type=np.concatenate([[GEOM_PLANE], bm.geom_type]),
        geom_bodyid=np.concatenate([[0], bm.geom_bodyid]),
        geom_group=np.concatenate([[0], bm.geom_group]),
        geom_matid=-np.ones(g, np.int32), geom_dataid=-np.ones(g, np.int32),
        geom_size=np.concatenate([[[0.0, 0.0, 0.01]], bm.geom_size]),
        geom_rgba=np.concatenate([[[0.5, 0.5, 0.5, 1.0]], bm.geom_rgba]).astype(np.float32),
        mat_rgba=np.zeros((0, 4), np.float32), body_weldid=bm.body_weldid)
    rng = np.random.default_rng(seed + 7919 * step)
    q = rng.uniform(-np.pi, np.pi, size=(nworld, 2))
    xpos, xmat = _arm_fk(base.tree, q, with_plane=True)
    inc = 2.0 * np.pi / n_rays
    dirs = planar_lidar_dirs(-np.pi, inc, n_rays)
    fx = _load_fixture()
    # tilt the scan plane slightly so part of every scan reaches the ground plane
    quat = quat_mul(axis_angle_quat((1, 0, 0), np.deg2rad(4.0)),
                    lidar_body_quat(fx["lidar"]["site_quat"], -np.pi))
    mat = quat2mat(quat)
    spec = LidarSpec(dirs, bodyexclude=int(fx["lidar"]["bodyid"]), range_min=0.05, range_max=30.0)
    return WorldsScene(model, Poses(xpos, xmat), spec,
                       np.tile(base.lidar_pos[None], (nworld, 1)),
                       np.tile(mat.reshape(1, 9), (nworld, 1)),
                       np.arange(nworld, dtype=np.int32))


# ------------------------------------------------------------------------------------------------
# procedural meshes
# ------------------------------------------------------------------------------------------------
def icosphere(subdiv: int):
    """Unit icosphere: 20*4**subdiv faces."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
                 dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdiv):
        nv = v.shape[0]
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        key = np.sort(edges, axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid])
        inv = inv.reshape(-1)
        nf = f.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return v, f.astype(np.int32)


def noisy_icosphere(subdiv: int, radius: float, seed: int):
    """Noise-displaced icosphere, float32 vertices (mjModel.mesh_vert is float)."""
    v, f = icosphere(subdiv)
    rng = np.random.default_rng(seed)
    r = np.ones(v.shape[0])
    for _ in range(6):
        k = rng.normal(size=3) * rng.uniform(1.0, 4.0)
        r += 0.06 * np.sin(v @ k + rng.uniform(0, 2 * np.pi))
    scale = rng.uniform(0.6, 1.0, size=3)
    vv = (v * r[:, None]) * scale[None, :] * radius
    return vv.astype(np.float32), f


# ------------------------------------------------------------------------------------------------
# S2: clutter500
# ------------------------------------------------------------------------------------------------
@dataclass
class ClutterScene:
    model: Model
    poses: Poses  # world 0
    lidar: LidarSpec
    sensor_pos: np.ndarray  # nominal (0,0,1.5)
    body_of_geom_local: tuple  # (geom local pos [G,3], geom local mat [G,3,3]) for S5 jitter
    body_pos: np.ndarray  # [B,3]
    body_mat: np.ndarray  # [B,3,3]
    n_el: int
    n_az: int

    def sensor_poses(self, n: int, seed: int = SEED + 1):
        """n sensor frames: SE(3) jitter about the nominal pose (yaw uniform, roll/pitch +-5 deg,
        position +-0.3 m).  Returns pos[n,3], mat[n,9]."""
        rng = np.random.default_rng(seed)
        pos = self.sensor_pos[None] + rng.uniform(-0.3, 0.3, size=(n, 3))
        mat = np.zeros((n, 9))
        for i in range(n):
            q = axis_angle_quat((0, 0, 1), rng.uniform(-np.pi, np.pi))
            q = quat_mul(q, axis_angle_quat((1, 0, 0), np.deg2rad(rng.uniform(-5, 5))))
            q = quat_mul(q, axis_angle_quat((0, 1, 0), np.deg2rad(rng.uniform(-5, 5))))
            mat[i] = quat2mat(q).reshape(9)
        if n > 0:  # scan 0 is the nominal pose
            pos[0] = self.sensor_pos
            mat[0] = np.eye(3).reshape(9)
        return pos, mat


def clutter500(seed: int = SEED, n_el: int = 64, n_az: int = 2048, n_other: int = 495,
               n_mesh_geoms: int = 100, mesh_subdivs=(2, 2, 3, 3, 3, 4, 4, 4, 5, 5),
               half_extent: float = 20.0) -> ClutterScene:
    """S2.  Geoms are body-major (as the MuJoCo compiler orders them): body 0 = world (plane +
    4 walls), bodies 1..62 carry the clutter, body 63 is the sensor body with 3 small housing
    geoms around the lidar origin (the rangefinder's bodyexclude)."""
    rng = np.random.default_rng(seed)
    nbody = 64
    sensor_body = nbody - 1
    sensor_pos = np.array([0.0, 0.0, 1.5])

    # ---- meshes --------------------------------------------------------------------------------
    verts, faces, vadr, vnum, fadr, fnum = [], [], [], [], [], []
    mesh_half = []
    for k, sd in enumerate(mesh_subdivs):
        v, f = noisy_icosphere(sd, float(rng.uniform(0.35, 1.0)), seed + k)
        vadr.append(sum(vnum))
        fadr.append(sum(fnum))
        vnum.append(v.shape[0])
        fnum.append(f.shape[0])
        verts.append(v)
        faces.append(f)
        mesh_half.append(np.abs(v.astype(np.float64)).max(axis=0))  # geom_size of a mesh geom
    nmesh = len(mesh_subdivs)

    # ---- unique clutter geoms, then every 10th slot is a coincident copy of its predecessor ------
    n_unique = n_other - n_other // 10
    n_mesh_unique = int(round(n_mesh_geoms * n_unique / n_other))
    n_prim_unique = n_unique - n_mesh_unique
    prim_types = np.array([GEOM_BOX, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID])
    prim_w = np.array([100, 80, 80, 70, 65], dtype=np.float64)
    types = np.concatenate([rng.choice(prim_types, size=n_prim_unique, p=prim_w / prim_w.sum()),
                            np.full(n_mesh_unique, GEOM_MESH)])
    rng.shuffle(types)

    def sample_centre():
        while True:
            c = np.array([rng.uniform(-half_extent, half_extent),
                          rng.uniform(-half_extent, half_extent), rng.uniform(0.0, 6.0)])
            if np.linalg.norm(c - sensor_pos) > 2.5:
                return c

    recs = []  # (type, size, dataid, pos, mat)
    rots = random_rotations(rng, n_unique)
    for i in range(n_unique):
        t = int(types[i])
        size = rng.uniform(0.1, 1.0, size=3)
        dataid = -1
        if t == GEOM_MESH:
            dataid = int(rng.integers(0, nmesh))
            size = mesh_half[dataid].copy()
        elif t == GEOM_SPHERE:
            size[1:] = 0.0
        elif t in (GEOM_CAPSULE, GEOM_CYLINDER):
            size[2] = 0.0
        recs.append((t, size, dataid, sample_centre(), rots[i]))
    full = []
    for i, r in enumerate(recs):
        full.append(r)
        if len(full) % 10 == 9 and len(full) < n_other:
            full.append(r)  # coincident duplicate: exact tie, the lower id must win
    full = full[:n_other]

    # ---- assemble body-major -----------------------------------------------------------------------
    gtype, gbody, gsize, gdata, gpos, gmat = [], [], [], [], [], []

    def add(t, size, dataid, pos, mat, body):
        gtype.append(t)
        gbody.append(body)
        gsize.append(np.asarray(size, dtype=np.float64))
        gdata.append(dataid)
        gpos.append(np.asarray(pos, dtype=np.float64))
        gmat.append(np.asarray(mat, dtype=np.float64))

    eye = np.eye(3)
    add(GEOM_PLANE, (0.0, 0.0, 0.01), -1, (0, 0, 0), eye, 0)
    he, wh, wt = half_extent + 1.0, 4.0, 0.25
    add(GEOM_BOX, (he, wt, wh), -1, (0, he, wh), eye, 0)
    add(GEOM_BOX, (he, wt, wh), -1, (0, -he, wh), eye, 0)
    add(GEOM_BOX, (wt, he, wh), -1, (he, 0, wh), eye, 0)
    add(GEOM_BOX, (wt, he, wh), -1, (-he, 0, wh), eye, 0)
    per_body = int(np.ceil(len(full) / (nbody - 2)))
    for i, (t, size, dataid, pos, mat) in enumerate(full):
        add(t, size, dataid, pos, mat, 1 + i // per_body)
    add(GEOM_BOX, (0.08, 0.08, 0.05), -1, sensor_pos + (0, 0, -0.06), eye, sensor_body)
    add(GEOM_CYLINDER, (0.07, 0.06, 0.0), -1, sensor_pos, eye, sensor_body)
    add(GEOM_SPHERE, (0.09, 0.0, 0.0), -1, sensor_pos + (0, 0, 0.02), eye, sensor_body)

    g = len(gtype)
    group = rng.integers(0, 6, size=g).astype(np.int32)
    group[:5] = 0
    rgba = np.ones((g, 4), np.float32)
    rgba[:, :3] = rng.uniform(0.1, 0.9, size=(g, 3))
    invisible = rng.uniform(size=g) < 0.05
    invisible[:5] = False
    rgba[invisible, 3] = 0.0
    # materials: 4, material 3 is fully transparent; ~10% of geoms reference one
    mat_rgba = np.array([[0.8, 0.2, 0.2, 1], [0.2, 0.8, 0.2, 1], [0.2, 0.2, 0.8, 0.5],
                         [0.5, 0.5, 0.5, 0.0]], np.float32)
    matid = -np.ones(g, np.int32)
    has_mat = rng.uniform(size=g) < 0.10
    has_mat[:5] = False
    matid[has_mat] = rng.integers(0, 4, size=int(has_mat.sum()))
    # coincident duplicates share visibility / group so the tie is really observed
    for i in range(1, g):
        if gtype[i] == gtype[i - 1] and np.array_equal(gpos[i], gpos[i - 1]) and gbody[i] == gbody[i - 1]:
            rgba[i, 3] = rgba[i - 1, 3]
            matid[i] = matid[i - 1]
            group[i] = (group[i - 1] + 1) % 6 if rng.uniform() < 0.3 else group[i - 1]
    # some bodies are welded to the world (static): flg_static = 0 must drop their geoms too
    weld = np.arange(nbody, dtype=np.int32)
    weld[0] = 0
    weld[rng.choice(np.arange(1, nbody - 1), size=8, replace=False)] = 0
    weld[sensor_body] = sensor_body

    model = Model(geom_type=gtype, geom_bodyid=gbody, geom_group=group, geom_matid=matid,
                  geom_dataid=gdata, geom_size=np.array(gsize), geom_rgba=rgba, mat_rgba=mat_rgba,
                  body_weldid=weld, mesh_vertadr=vadr, mesh_vertnum=vnum, mesh_faceadr=fadr,
                  mesh_facenum=fnum, mesh_vert=np.concatenate(verts), mesh_face=np.concatenate(faces))
    xpos = np.array(gpos)
    xmat = np.array(gmat).reshape(g, 9)

    # body frames for S5: body origin = mean of its geoms, identity orientation
    gb = np.array(gbody)
    body_pos = np.zeros((nbody, 3))
    for b in range(nbody):
        if np.any(gb == b) and b != 0:
            body_pos[b] = xpos[gb == b].mean(axis=0)
    body_mat = np.tile(np.eye(3)[None], (nbody, 1, 1))
    lpos = xpos - body_pos[gb]
    lmat = np.array(gmat)

    lidar = LidarSpec(spinning_lidar_dirs(n_el, n_az), tile_w=n_az, bodyexclude=sensor_body,
                      range_min=0.0, range_max=1000.0)
    return ClutterScene(model, Poses(xpos, xmat), lidar, sensor_pos, (lpos, lmat), body_pos,
                        body_mat, n_el, n_az)


def clutter_worlds(scene: ClutterScene, nworld: int, seed: int = SEED + 2, step: int = 0) -> Poses:
    """S5 poses: every non-world body of S2 gets a seeded SE(3) jitter per world (+-0.5 m, +-20 deg
    about a random axis); world 0 is the unperturbed scene."""
    rng = np.random.default_rng(seed + 104729 * step)
    gb = scene.model.geom_bodyid
    lpos, lmat = scene.body_of_geom_local
    nb = scene.body_pos.shape[0]
    g = scene.model.ngeom
    xpos = np.zeros((nworld, g, 3))
    xmat = np.zeros((nworld, g, 9))
    for w in range(nworld):
        dp = rng.uniform(-0.5, 0.5, size=(nb, 3))
        ang = np.deg2rad(rng.uniform(-20, 20, size=nb))
        ax = rng.normal(size=(nb, 3))
        if w == 0:
            dp[:] = 0
            ang[:] = 0
        dp[0] = 0
        ang[0] = 0
        dp[nb - 1] = 0
        ang[nb - 1] = 0
        rb = np.stack([quat2mat(axis_angle_quat(ax[b], ang[b])) for b in range(nb)])
        pb = scene.body_pos + dp
        xpos[w] = pb[gb] + np.einsum("gij,gj->gi", rb[gb], lpos)
        xmat[w] = np.einsum("gij,gjk->gik", rb[gb], lmat).reshape(g, 9)
    return Poses(xpos, xmat)


def clutter_world_block(scene: ClutterScene, world0: int, nworld: int, seed: int = SEED + 5,
                        step: int = 0) -> Poses:
    """S5 poses of worlds [world0, world0 + nworld) of a job of any size: world w's body jitter
    (+-0.5 m, +-20 deg about a random axis; the world body and the sensor body stay put) depends only
    on (seed, step, w), so a rank's block is the same whether the job runs on 1 or 8 GPUs
    (shard.world_block).  Vectorised over worlds: 4096 worlds x 503 geoms in about a second."""
    gb = scene.model.geom_bodyid
    lpos, lmat = scene.body_of_geom_local
    nb = scene.body_pos.shape[0]
    g = scene.model.ngeom
    dp = np.empty((nworld, nb, 3))
    ang = np.empty((nworld, nb))
    ax = np.empty((nworld, nb, 3))
    for i in range(nworld):
        rng = np.random.default_rng([seed, step, world0 + i])
        dp[i] = rng.uniform(-0.5, 0.5, size=(nb, 3))
        ang[i] = np.deg2rad(rng.uniform(-20, 20, size=nb))
        ax[i] = rng.normal(size=(nb, 3))
    for b in (0, nb - 1):
        dp[:, b] = 0
        ang[:, b] = 0
    ax /= np.linalg.norm(ax, axis=2, keepdims=True) + 1e-300
    # Rodrigues: R = I + sin(a) K + (1 - cos(a)) K^2
    K = np.zeros((nworld, nb, 3, 3))
    K[..., 0, 1], K[..., 0, 2] = -ax[..., 2], ax[..., 1]
    K[..., 1, 0], K[..., 1, 2] = ax[..., 2], -ax[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -ax[..., 1], ax[..., 0]
    sa, ca = np.sin(ang)[..., None, None], np.cos(ang)[..., None, None]
    rb = np.eye(3) + sa * K + (1 - ca) * (K @ K)
    pb = scene.body_pos[None] + dp
    xpos = pb[:, gb] + np.einsum("wgij,gj->wgi", rb[:, gb], lpos)
    xmat = np.einsum("wgij,gjk->wgik", rb[:, gb], lmat).reshape(nworld, g, 9)
    return Poses(xpos, xmat)


# ------------------------------------------------------------------------------------------------
# S4: GL-style depth buffers
# ------------------------------------------------------------------------------------------------
def gl_depth(ncam: int, width: int = 640, height: int = 480, seed: int = SEED + 3) -> np.ndarray:
    """Raw depth as mjr_readPixels returns it: float32 in [0,1], bottom-left origin.  U(0,1) plus
    the edge values 0, 1-2^-24 and 1.0 (far plane)."""
    rng = np.random.default_rng(seed)
    d = rng.random(size=(ncam, height, width), dtype=np.float32)
    flat = d.reshape(ncam, -1)
    flat[:, 0] = 0.0
    flat[:, 1] = np.float32(1.0) - np.float32(2.0 ** -24)
    flat[:, 2] = 1.0
    flat[:, -1] = 1.0
    return d


def depth_camera(width: int = 640, height: int = 480, fovy: float = 58.0, point_step: int = 12) -> CameraSpec:
    return CameraSpec(width, height, fovy, 0.01, 50.0, 1.0, point_step)
