"""ctypes binding of the CPU ORACLE (oracle/b2_oracle.c).  TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (mujoco_ros2_simulation_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2oracle.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


class _OrcModel(C.Structure):
    _fields_ = [
        ("ngeom", C.c_int32), ("nbody", C.c_int32), ("nmesh", C.c_int32), ("nmat", C.c_int32),
        ("geom_type", _i32p), ("geom_bodyid", _i32p), ("geom_group", _i32p), ("geom_matid", _i32p),
        ("geom_dataid", _i32p), ("geom_size", _f64p), ("geom_rgba", _f32p), ("mat_rgba", _f32p),
        ("body_weldid", _i32p), ("mesh_vertadr", _i32p), ("mesh_vertnum", _i32p),
        ("mesh_faceadr", _i32p), ("mesh_facenum", _i32p), ("mesh_vert", _f32p), ("mesh_face", _i32p),
        ("mesh_bvhadr", _i32p), ("mesh_bvhnum", _i32p), ("bvh_aabb", _f64p), ("bvh_child", _i32p),
        ("bvh_nodeid", _i32p),
    ]


class _OrcData(C.Structure):
    _fields_ = [("geom_xpos", _f64p), ("geom_xmat", _f64p)]


def build(force: bool = False) -> str:
    """Compile oracle/_build/libb2oracle.so with the committed Makefile (gcc)."""
    src = os.path.join(_HERE, "b2_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "b2_oracle.h")))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_ray_geom.restype = C.c_double
        L.orc_ray_geom.argtypes = [_f64p, _f64p, _f64p, _f64p, _f64p, C.c_int]
        L.orc_ray_geom_batch.restype = None
        L.orc_ray_geom_batch.argtypes = [C.c_int] + [_f64p] * 5 + [_i32p, _f64p]
        L.orc_ray.restype = C.c_double
        L.orc_ray.argtypes = [C.POINTER(_OrcModel), C.POINTER(_OrcData), _f64p, _f64p, _u8p, C.c_uint8,
                              C.c_int, _i32p]
        L.orc_ray_mesh.restype = C.c_double
        L.orc_ray_mesh.argtypes = [C.POINTER(_OrcModel), C.POINTER(_OrcData), C.c_int, _f64p, _f64p]
        L.orc_ray_mesh_brute.restype = C.c_double
        L.orc_ray_mesh_brute.argtypes = L.orc_ray_mesh.argtypes
        L.orc_multi_ray.restype = None
        L.orc_multi_ray.argtypes = [C.POINTER(_OrcModel), C.POINTER(_OrcData), _f64p, _f64p, _u8p,
                                    C.c_uint8, C.c_int, _i32p, _f64p, C.c_int, C.c_double, C.c_int]
        L.orc_multi_ray_cull.restype = None
        L.orc_multi_ray_cull.argtypes = L.orc_multi_ray.argtypes + [C.c_int]
        L.orc_lidar_scan_multiray.restype = None
        L.orc_lidar_scan_multiray.argtypes = [C.POINTER(_OrcModel), C.POINTER(_OrcData), _f64p, _f64p,
                                              _f64p, C.c_int, _u8p, C.c_uint8, C.c_int, C.c_double,
                                              C.c_double, C.c_double, _f32p, _i32p, _f64p, C.c_int,
                                              C.c_int]
        L.orc_lidar_scan.restype = None
        L.orc_lidar_scan.argtypes = [C.POINTER(_OrcModel), C.POINTER(_OrcData), _f64p, _f64p, _f64p,
                                     C.c_int, _u8p, C.c_uint8, C.c_int, C.c_double, C.c_double,
                                     C.c_double, _f32p, _i32p, _f64p, C.c_int]
        L.orc_lidar_num_rangefinders.restype = C.c_int
        L.orc_lidar_num_rangefinders.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_mesh_bvh_size.restype = C.c_int
        L.orc_mesh_bvh_size.argtypes = [C.c_int]
        L.orc_build_mesh_bvh.restype = C.c_int
        L.orc_build_mesh_bvh.argtypes = [_f32p, _i32p, C.c_int, _f64p, _i32p, _i32p]
        L.orc_camera_intrinsics.restype = None
        L.orc_camera_intrinsics.argtypes = [C.c_double, C.c_int, C.c_int, _f64p]
        L.orc_depth_linearise_flip.restype = None
        L.orc_depth_linearise_flip.argtypes = [_f32p, C.c_int, C.c_int, C.c_double, C.c_double,
                                               C.c_double, _f32p]
        L.orc_depth_to_cloud.restype = None
        L.orc_depth_to_cloud.argtypes = [_f32p, C.c_int, C.c_int, _f64p, C.c_int, _f32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size > 0 else C.cast(None, t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


CULL_NONE, CULL_BODY_ANGULAR, CULL_GEOM_SPHERE = 0, 1, 2


class OracleScene:
    """An orc_model + one world's orc_data.  Builds the mjModel-layout mesh BVHs on construction."""

    def __init__(self, model, poses, world: int = 0):
        self.model = model
        L = lib()
        self._keep = []
        m = _OrcModel()
        m.ngeom, m.nbody, m.nmesh, m.nmat = model.ngeom, model.nbody, model.nmesh, model.nmat
        for name, t in (("geom_type", _i32p), ("geom_bodyid", _i32p), ("geom_group", _i32p),
                        ("geom_matid", _i32p), ("geom_dataid", _i32p), ("geom_size", _f64p),
                        ("geom_rgba", _f32p), ("mat_rgba", _f32p), ("body_weldid", _i32p),
                        ("mesh_vertadr", _i32p), ("mesh_vertnum", _i32p), ("mesh_faceadr", _i32p),
                        ("mesh_facenum", _i32p), ("mesh_vert", _f32p), ("mesh_face", _i32p)):
            arr = getattr(model, name)
            self._keep.append(arr)
            setattr(m, name, _p(arr, t))
        # per-mesh BVH in the mjModel layout
        nm = model.nmesh
        bvhadr = np.zeros(nm, np.int32)
        bvhnum = np.zeros(nm, np.int32)
        tot = 0
        for k in range(nm):
            bvhadr[k] = tot
            tot += L.orc_mesh_bvh_size(int(model.mesh_facenum[k]))
        aabb = np.zeros((max(tot, 1), 6), np.float64)
        child = -np.ones((max(tot, 1), 2), np.int32)
        nodeid = -np.ones(max(tot, 1), np.int32)
        for k in range(nm):
            va, fa, fn = int(model.mesh_vertadr[k]), int(model.mesh_faceadr[k]), int(model.mesh_facenum[k])
            v = model.mesh_vert[va:va + int(model.mesh_vertnum[k])]
            f = model.mesh_face[fa:fa + fn]
            a = int(bvhadr[k])
            n = L.orc_build_mesh_bvh(_p(v, _f32p), _p(f, _i32p), fn, _p(aabb[a:], _f64p),
                                     _p(child[a:], _i32p), _p(nodeid[a:], _i32p))
            bvhnum[k] = n
        self._keep += [bvhadr, bvhnum, aabb, child, nodeid]
        m.mesh_bvhadr, m.mesh_bvhnum = _p(bvhadr, _i32p), _p(bvhnum, _i32p)
        m.bvh_aabb, m.bvh_child, m.bvh_nodeid = _p(aabb, _f64p), _p(child, _i32p), _p(nodeid, _i32p)
        self._m = m
        self.set_world(poses, world)

    def set_world(self, poses, world: int = 0):
        self._xpos = _c(poses.geom_xpos[world], np.float64)
        self._xmat = _c(poses.geom_xmat[world], np.float64)
        d = _OrcData()
        d.geom_xpos, d.geom_xmat = _p(self._xpos, _f64p), _p(self._xmat, _f64p)
        self._d = d

    # ---- A.5 ---------------------------------------------------------------------------------
    def ray(self, pnt, vec, geomgroup=None, flg_static=1, bodyexclude=-1):
        pnt, vec = _c(pnt, np.float64), _c(vec, np.float64)
        gg = None if geomgroup is None else _c(geomgroup, np.uint8)
        gid = C.c_int32(-1)
        dist = lib().orc_ray(C.byref(self._m), C.byref(self._d), _p(pnt, _f64p), _p(vec, _f64p),
                             _p(gg, _u8p), flg_static, bodyexclude, C.byref(gid))
        return dist, gid.value

    def ray_mesh(self, geomid, pnt, vec, brute=False):
        pnt, vec = _c(pnt, np.float64), _c(vec, np.float64)
        fn = lib().orc_ray_mesh_brute if brute else lib().orc_ray_mesh
        return fn(C.byref(self._m), C.byref(self._d), geomid, _p(pnt, _f64p), _p(vec, _f64p))

    # ---- A.6 ---------------------------------------------------------------------------------
    def multi_ray(self, pnt, vec, geomgroup=None, flg_static=1, bodyexclude=-1, cutoff=1e10,
                  nthread=1):
        pnt, vec = _c(pnt, np.float64), _c(vec, np.float64).reshape(-1, 3)
        n = vec.shape[0]
        gg = None if geomgroup is None else _c(geomgroup, np.uint8)
        gid = np.empty(n, np.int32)
        dist = np.empty(n, np.float64)
        lib().orc_multi_ray(C.byref(self._m), C.byref(self._d), _p(pnt, _f64p), _p(vec, _f64p),
                            _p(gg, _u8p), flg_static, bodyexclude, _p(gid, _i32p), _p(dist, _f64p),
                            n, cutoff, nthread)
        return dist, gid

    def multi_ray_cull(self, pnt, vec, cull, geomgroup=None, flg_static=1, bodyexclude=-1, cutoff=1e10,
                       nthread=1):
        """orc_multi_ray with the cull chosen: CULL_NONE / CULL_BODY_ANGULAR / CULL_GEOM_SPHERE."""
        pnt, vec = _c(pnt, np.float64), _c(vec, np.float64).reshape(-1, 3)
        n = vec.shape[0]
        gg = None if geomgroup is None else _c(geomgroup, np.uint8)
        gid = np.empty(n, np.int32)
        dist = np.empty(n, np.float64)
        lib().orc_multi_ray_cull(C.byref(self._m), C.byref(self._d), _p(pnt, _f64p), _p(vec, _f64p),
                                 _p(gg, _u8p), flg_static, bodyexclude, _p(gid, _i32p),
                                 _p(dist, _f64p), n, cutoff, nthread, int(cull))
        return dist, gid

    def lidar_scan_multiray(self, spec, sensor_pos, sensor_mat, cull=1, nthread=1):
        """lidar_scan computed the mj_multiRay way (one cull preparation per scan)."""
        pos, mat = _c(sensor_pos, np.float64).reshape(3), _c(sensor_mat, np.float64).reshape(9)
        n = spec.nray
        ranges = np.empty(n, np.float32)
        gid = np.empty(n, np.int32)
        raw = np.empty(n, np.float64)
        lib().orc_lidar_scan_multiray(C.byref(self._m), C.byref(self._d), _p(pos, _f64p), _p(mat, _f64p),
                                      _p(spec.local_dir, _f64p), n, _p(spec.geomgroup, _u8p),
                                      spec.flg_static, spec.bodyexclude, spec.cutoff, spec.range_min,
                                      spec.range_max, _p(ranges, _f32p), _p(gid, _i32p),
                                      _p(raw, _f64p), nthread, int(cull))
        return ranges, gid, raw

    # ---- rangefinder + MujocoLidar::update -----------------------------------------------------
    def lidar_scan(self, spec, sensor_pos, sensor_mat, nthread=1):
        pos, mat = _c(sensor_pos, np.float64).reshape(3), _c(sensor_mat, np.float64).reshape(9)
        n = spec.nray
        ranges = np.empty(n, np.float32)
        gid = np.empty(n, np.int32)
        raw = np.empty(n, np.float64)
        lib().orc_lidar_scan(C.byref(self._m), C.byref(self._d), _p(pos, _f64p), _p(mat, _f64p),
                             _p(spec.local_dir, _f64p), n, _p(spec.geomgroup, _u8p), spec.flg_static,
                             spec.bodyexclude, spec.cutoff, spec.range_min, spec.range_max,
                             _p(ranges, _f32p), _p(gid, _i32p), _p(raw, _f64p), nthread)
        return ranges, gid, raw


def ray_geom(pos, mat, size, pnt, vec, geomtype) -> float:
    a = [_c(x, np.float64) for x in (pos, mat, size, pnt, vec)]
    return lib().orc_ray_geom(*[_p(x, _f64p) for x in a], int(geomtype))


def ray_geom_batch(pos, mat, size, pnt, vec, geomtype) -> np.ndarray:
    """n independent (geom, ray) cases through orc_ray_geom: arrays [n,3] / [n,9] / [n]."""
    a = [_c(x, np.float64) for x in (pos, mat, size, pnt, vec)]
    t = _c(geomtype, np.int32)
    out = np.empty(t.size, np.float64)
    lib().orc_ray_geom_batch(int(t.size), *[_p(x, _f64p) for x in a], _p(t, _i32p), _p(out, _f64p))
    return out


def lidar_num_rangefinders(min_angle, max_angle, inc) -> int:
    return lib().orc_lidar_num_rangefinders(min_angle, max_angle, inc)


def camera_intrinsics(fovy, width, height) -> np.ndarray:
    k = np.zeros(9, np.float64)
    lib().orc_camera_intrinsics(fovy, width, height, _p(k, _f64p))
    return k


def depth_linearise_flip(gl_depth, znear, zfar, extent) -> np.ndarray:
    d = _c(gl_depth, np.float32)
    h, w = d.shape[-2:]
    out = np.empty_like(d)
    for i in range(d.reshape(-1, h, w).shape[0]):
        lib().orc_depth_linearise_flip(_p(d.reshape(-1, h, w)[i], _f32p), w, h, znear, zfar, extent,
                                       _p(out.reshape(-1, h, w)[i], _f32p))
    return out


def depth_to_cloud(depth_flipped, K, point_floats=3) -> np.ndarray:
    d = _c(depth_flipped, np.float32)
    h, w = d.shape[-2:]
    n = d.reshape(-1, h, w).shape[0]
    K = _c(K, np.float64)
    out = np.empty((n, h, w, point_floats), np.float32)
    for i in range(n):
        lib().orc_depth_to_cloud(_p(d.reshape(-1, h, w)[i], _f32p), w, h, _p(K, _f64p), point_floats,
                                 _p(out[i], _f32p))
    return out.reshape(d.shape + (point_floats,))
