"""ctypes binding of tests/emu/emu.cpp (CPU emulation of the kernel traversal; TEST ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mujoco_ros2_simulation_b200 import _capi
from mujoco_ros2_simulation_b200.engine import _model_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
_CSRC = os.path.join(_ROOT, "mujoco_ros2_simulation_b200", "csrc")
_SO = os.path.join(_HERE, "_build", "libb2emu.so")
_SRCS = [os.path.join(_HERE, "emu.cpp"), os.path.join(_CSRC, "scene_build.cpp"),
         os.path.join(_CSRC, "bvh_build.cpp"), os.path.join(_CSRC, "bvh_collapse4.cpp")]
_DEPS = _SRCS + [os.path.join(_CSRC, f) for f in ("ray_core.cuh", "tile_trace.cuh", "b2_types.h", "bvh_build.h",
                                                    "scene_build.h")]


def build(force: bool = False) -> str:
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in _DEPS)
    if force or stale:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -x c++ so the .cuh header is accepted; no FMA contraction, as in the CUDA build
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-march=x86-64-v2",
               "-Wall", "-Wno-unknown-pragmas", "-o", _SO] + _SRCS
        subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.emu_multiray.restype = C.c_int
        L.emu_multiray.argtypes = [C.POINTER(_capi.ModelDesc)] + [C.c_void_p] * 7 + [
            C.c_uint8, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.emu_lidar.restype = C.c_int
        L.emu_lidar.argtypes = [C.POINTER(_capi.ModelDesc), C.c_void_p, C.c_void_p,
                                C.POINTER(_capi.LidarDesc)] + [C.c_void_p] * 6 + [C.c_int]
        L.emu_bvh_stats.restype = C.c_int
        L.emu_bvh_stats.argtypes = [C.POINTER(_capi.ModelDesc), C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.emu_prim_batch.restype = None
        L.emu_prim_batch.argtypes = [C.c_int] + [C.c_void_p] * 7
        L.emu_sphere_miss_batch.restype = None
        L.emu_sphere_miss_batch.argtypes = [C.c_int] + [C.c_void_p] * 5
        _lib = L
    return _lib


def sphere_miss_batch(pos, radius, pnt, vec) -> np.ndarray:
    """the kernels' sign-only bounding-sphere pre-test (ray_core.cuh sphere_miss) on n cases."""
    a = [np.ascontiguousarray(x, np.float64) for x in (pos, radius, pnt, vec)]
    out = np.empty(a[1].size, np.uint8)
    lib().emu_sphere_miss_batch(int(a[1].size), *[C.c_void_p(x.ctypes.data) for x in a],
                                C.c_void_p(out.ctypes.data))
    return out.astype(bool)


def prim_batch(pos, mat, size, pnt, vec, geomtype) -> np.ndarray:
    """n independent (geom, ray) cases through the kernels' leaf routine prim_test."""
    a = [np.ascontiguousarray(x, np.float64) for x in (pos, mat, size, pnt, vec)]
    t = np.ascontiguousarray(geomtype, np.int32)
    out = np.empty(t.size, np.float64)
    lib().emu_prim_batch(int(t.size), *[C.c_void_p(x.ctypes.data) for x in a],
                         C.c_void_p(t.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def multiray(model, poses, pnt, vec, geomgroup=None, flg_static=1, bodyexclude=-1, cutoff=1e10,
             world=0, topo_world=None, mesh_leaf=2):
    desc = _model_desc(model)
    xpos = np.ascontiguousarray(poses.geom_xpos[world], np.float64)
    xmat = np.ascontiguousarray(poses.geom_xmat[world], np.float64)
    tpos = tmat = None
    if topo_world is not None:
        tpos = np.ascontiguousarray(poses.geom_xpos[topo_world], np.float64)
        tmat = np.ascontiguousarray(poses.geom_xmat[topo_world], np.float64)
    pnt = np.ascontiguousarray(pnt, np.float64).reshape(3)
    vec = np.ascontiguousarray(vec, np.float64).reshape(-1, 3)
    n = vec.shape[0]
    gg = None if geomgroup is None else np.ascontiguousarray(geomgroup, np.uint8)
    gid = np.empty(n, np.int32)
    dist = np.empty(n, np.float64)
    rc = lib().emu_multiray(C.byref(desc), _p(xpos), _p(xmat), _p(tpos), _p(tmat), _p(pnt), _p(vec),
                            _p(gg), flg_static, bodyexclude, _p(gid), _p(dist), n, float(cutoff),
                            mesh_leaf)
    assert rc == 0
    return dist, gid


def lidar(model, poses, spec, pos, mat, world=0, mesh_leaf=2):
    desc = _model_desc(model)
    xpos = np.ascontiguousarray(poses.geom_xpos[world], np.float64)
    xmat = np.ascontiguousarray(poses.geom_xmat[world], np.float64)
    d = _capi.LidarDesc()
    d.nray = spec.nray
    d.local_dir = spec.local_dir.ctypes.data_as(_capi.f64p)
    d.tile_w = spec.tile_w
    d.geomgroup = (spec.geomgroup.ctypes.data_as(_capi.u8p) if spec.geomgroup is not None
                   else C.cast(None, _capi.u8p))
    d.flg_static, d.bodyexclude = spec.flg_static, spec.bodyexclude
    d.cutoff, d.range_min, d.range_max = spec.cutoff, spec.range_min, spec.range_max
    pos = np.ascontiguousarray(pos, np.float64).reshape(3)
    mat = np.ascontiguousarray(mat, np.float64).reshape(9)
    n = spec.nray
    ranges, gid = np.empty(n, np.float32), np.empty(n, np.int32)
    dist, xyz = np.empty(n, np.float64), np.empty((n, 3), np.float32)
    rc = lib().emu_lidar(C.byref(desc), _p(xpos), _p(xmat), C.byref(d), _p(pos), _p(mat), _p(ranges),
                         _p(gid), _p(dist), _p(xyz), mesh_leaf)
    assert rc == 0
    return ranges, gid, dist, xyz


def bvh_stats(model, poses, world=0, mesh_leaf=2):
    desc = _model_desc(model)
    xpos = np.ascontiguousarray(poses.geom_xpos[world], np.float64)
    xmat = np.ascontiguousarray(poses.geom_xmat[world], np.float64)
    nn, dp, mn = C.c_int32(), C.c_int32(), C.c_int64()
    rc = lib().emu_bvh_stats(C.byref(desc), _p(xpos), _p(xmat), mesh_leaf, C.byref(nn), C.byref(dp),
                             C.byref(mn))
    assert rc == 0
    return {"bvh_slots": nn.value, "depth": dp.value, "mesh_nodes": mn.value}


def cone_violations(pnt, vec, boxes) -> int:
    """number of boxes a ray's slab test touches but the tile cone (built from `vec`) rejects,
    plus ordering keys that exceed a ray's entry distance.  Must be 0."""
    pnt = np.ascontiguousarray(pnt, np.float64).reshape(3)
    vec = np.ascontiguousarray(vec, np.float64).reshape(-1, 3)
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    L = lib()
    L.emu_cone_violations.restype = C.c_int
    L.emu_cone_violations.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return L.emu_cone_violations(_p(pnt), _p(vec), vec.shape[0], _p(boxes), boxes.shape[0])


def cone_counts(pnt, vec, boxes):
    """(violations, kept, touched): boxes the group bound keeps vs boxes some ray's slab test touches."""
    pnt = np.ascontiguousarray(pnt, np.float64).reshape(3)
    vec = np.ascontiguousarray(vec, np.float64).reshape(-1, 3)
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    L = lib()
    kept, touched = C.c_int(0), C.c_int(0)
    L.emu_cone_check.restype = C.c_int
    L.emu_cone_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    bad = L.emu_cone_check(_p(pnt), _p(vec), vec.shape[0], _p(boxes), boxes.shape[0], C.byref(kept),
                           C.byref(touched))
    return bad, kept.value, touched.value


def fallbacks() -> int:
    """groups the two-kernel emulation (B2_EMU_TWO_KERNEL=1) handed to the fused trace_group so far"""
    L = lib()
    L.emu_fallbacks.restype = C.c_longlong
    return int(L.emu_fallbacks())
