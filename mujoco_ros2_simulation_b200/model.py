"""Flat numpy mirror of the mjModel / mjData fields the sensor path reads.

Field names are MuJoCo's, so a scene built here is laid out exactly like the arrays the plugin
would hand to ``b2_scene_create`` (include/b2sense.h) from a real ``mjModel``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE = 0, 1, 2, 3
GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 4, 5, 6, 7
NGROUP = 6
MINVAL = 1e-15


def _a(x, dtype, shape=None):
    arr = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    if shape is not None:
        arr = arr.reshape(shape)
    return arr


@dataclass
class Model:
    geom_type: np.ndarray
    geom_bodyid: np.ndarray
    geom_group: np.ndarray
    geom_matid: np.ndarray
    geom_dataid: np.ndarray
    geom_size: np.ndarray  # [G,3] f64
    geom_rgba: np.ndarray  # [G,4] f32
    mat_rgba: np.ndarray  # [M,4] f32
    body_weldid: np.ndarray  # [B] i32
    mesh_vertadr: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mesh_vertnum: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mesh_faceadr: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mesh_facenum: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mesh_vert: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    mesh_face: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))

    def __post_init__(self):
        self.geom_type = _a(self.geom_type, np.int32)
        g = self.geom_type.shape[0]
        self.geom_bodyid = _a(self.geom_bodyid, np.int32, (g,))
        self.geom_group = _a(self.geom_group, np.int32, (g,))
        self.geom_matid = _a(self.geom_matid, np.int32, (g,))
        self.geom_dataid = _a(self.geom_dataid, np.int32, (g,))
        self.geom_size = _a(self.geom_size, np.float64, (g, 3))
        self.geom_rgba = _a(self.geom_rgba, np.float32, (g, 4))
        self.mat_rgba = _a(self.mat_rgba, np.float32, (-1, 4))
        self.body_weldid = _a(self.body_weldid, np.int32)
        self.mesh_vertadr = _a(self.mesh_vertadr, np.int32)
        self.mesh_vertnum = _a(self.mesh_vertnum, np.int32)
        self.mesh_faceadr = _a(self.mesh_faceadr, np.int32)
        self.mesh_facenum = _a(self.mesh_facenum, np.int32)
        self.mesh_vert = _a(self.mesh_vert, np.float32, (-1, 3))
        self.mesh_face = _a(self.mesh_face, np.int32, (-1, 3))

    @property
    def ngeom(self) -> int:
        return int(self.geom_type.shape[0])

    @property
    def nbody(self) -> int:
        return int(self.body_weldid.shape[0])

    @property
    def nmesh(self) -> int:
        return int(self.mesh_vertadr.shape[0])

    @property
    def nmat(self) -> int:
        return int(self.mat_rgba.shape[0])

    @property
    def nface(self) -> int:
        return int(self.mesh_face.shape[0])


@dataclass
class Poses:
    """mjData geom_xpos / geom_xmat for ``nworld`` worlds: [W,G,3] and [W,G,9] float64."""

    geom_xpos: np.ndarray
    geom_xmat: np.ndarray

    def __post_init__(self):
        self.geom_xpos = _a(self.geom_xpos, np.float64)
        self.geom_xmat = _a(self.geom_xmat, np.float64)
        if self.geom_xpos.ndim == 2:
            self.geom_xpos = self.geom_xpos[None]
            self.geom_xmat = self.geom_xmat[None]
        w, g = self.geom_xpos.shape[:2]
        self.geom_xmat = self.geom_xmat.reshape(w, g, 9)

    @property
    def nworld(self) -> int:
        return int(self.geom_xpos.shape[0])

    def world(self, w: int) -> "Poses":
        return Poses(self.geom_xpos[w : w + 1].copy(), self.geom_xmat[w : w + 1].copy())


@dataclass
class LidarSpec:
    """Static description of one lidar (include/b2sense.h ``b2_lidar_desc``)."""

    local_dir: np.ndarray  # [N,3] f64, sensor frame
    tile_w: int = 0
    geomgroup: np.ndarray | None = None  # [6] u8 or None
    flg_static: int = 1
    bodyexclude: int = -1
    cutoff: float = 0.0
    range_min: float = 0.0
    range_max: float = 1000.0

    def __post_init__(self):
        self.local_dir = _a(self.local_dir, np.float64, (-1, 3))
        if self.geomgroup is not None:
            self.geomgroup = _a(self.geomgroup, np.uint8, (NGROUP,))

    @property
    def nray(self) -> int:
        return int(self.local_dir.shape[0])


@dataclass
class CameraSpec:
    width: int
    height: int
    fovy: float
    znear: float = 0.01
    zfar: float = 50.0
    extent: float = 1.0
    point_step: int = 12
    geomgroup: tuple | None = None  # geom groups the ray-cast renderer draws; None = mjv_defaultOption


# ------------------------------------------------------------------------------------------------
# small rotation helpers (MuJoCo conventions: quat = (w,x,y,z), mat row-major, columns = axes)
# ------------------------------------------------------------------------------------------------
def quat2mat(q) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ],
        dtype=np.float64,
    )


def quat_mul(a, b) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        dtype=np.float64,
    )


def axis_angle_quat(axis, angle) -> np.ndarray:
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = np.sin(angle / 2.0)
    return np.array([np.cos(angle / 2.0), axis[0] * s, axis[1] * s, axis[2] * s])


def random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    """n rotation matrices uniform on SO(3) (normalised Gaussian quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.stack([quat2mat(qi) for qi in q])
