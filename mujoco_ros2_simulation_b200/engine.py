"""Thin Python host layer over the C ABI (include/b2sense.h) used by tests and bench.py.

``Scene`` / ``Lidar`` / ``DepthCamera`` own the opaque handles.  Methods ending in ``_device``
take torch CUDA tensors (device memory + streams are PyTorch's job here, nothing else is) and
enqueue on the scene's stream; the others take numpy arrays (or pinned torch CPU tensors) and
block.  All compute happens in libb2sense.so's CUDA kernels; there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _capi
from .model import CameraSpec, LidarSpec, Model, Poses


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _ptr(x):
    """void* of a numpy array, a torch tensor (CPU or CUDA), a raw address (int) or None."""
    if x is None:
        return None
    if isinstance(x, int):  # a raw device pointer (e.g. a slice of a peer buffer, shard.PeerGather.ptr)
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())


def _model_desc(model: Model):
    d = _capi.ModelDesc()
    d.ngeom, d.nbody, d.nmesh, d.nmat = model.ngeom, model.nbody, model.nmesh, model.nmat
    for name, t in (("geom_type", _capi.i32p), ("geom_bodyid", _capi.i32p),
                    ("geom_group", _capi.i32p), ("geom_matid", _capi.i32p),
                    ("geom_dataid", _capi.i32p), ("geom_size", _capi.f64p),
                    ("geom_rgba", _capi.f32p), ("mat_rgba", _capi.f32p),
                    ("body_weldid", _capi.i32p), ("mesh_vertadr", _capi.i32p),
                    ("mesh_vertnum", _capi.i32p), ("mesh_faceadr", _capi.i32p),
                    ("mesh_facenum", _capi.i32p), ("mesh_vert", _capi.f32p),
                    ("mesh_face", _capi.i32p)):
        arr = getattr(model, name)
        setattr(d, name, arr.ctypes.data_as(t) if arr.size else C.cast(None, t))
    return d


def device_count() -> int:
    return int(_capi.load().b2_device_count())


def launch_count() -> int:
    return int(_capi.load().b2_launch_count())


class Scene:
    """GPU scene mirror of mjModel / mjData (``b2_scene``)."""

    def __init__(self, model: Model, nworld: int = 1, device: int = 0, stream: int | None = None):
        self.lib = _capi.load()
        self.model = model  # keeps the arrays alive during create
        self.nworld = int(nworld)
        self.device = int(device)
        h = C.c_void_p()
        desc = _model_desc(model)
        _capi.check(self.lib.b2_scene_create(C.byref(desc), self.nworld, self.device,
                                             C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self._children = weakref.WeakSet()  # lidars / cameras bound to this scene

    def close(self):
        """Destroys the scene; lidars and cameras still bound to it are closed first (the C ABI
        refuses to destroy a scene with live dependents)."""
        if getattr(self, "_h", None):
            for child in list(self._children):
                child.close()
            _capi.check(self.lib.b2_scene_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_poses(self, poses: Poses, world0: int = 0):
        xpos = np.ascontiguousarray(poses.geom_xpos, np.float64)
        xmat = np.ascontiguousarray(poses.geom_xmat, np.float64)
        _capi.check(self.lib.b2_scene_update_poses(self._h, world0, poses.nworld, _np_ptr(xpos),
                                                   _np_ptr(xmat)))

    def update_poses_device(self, xpos, xmat, world0: int = 0):
        """xpos [W,G,3] / xmat [W,G,9] float64 CUDA tensors."""
        _capi.check(self.lib.b2_scene_update_poses_device(self._h, world0, int(xpos.shape[0]),
                                                          _ptr(xpos), _ptr(xmat)))

    def rebuild(self, world: int = 0):
        _capi.check(self.lib.b2_scene_rebuild(self._h, world))

    def sync(self):
        _capi.check(self.lib.b2_scene_sync(self._h))

    def copy_async(self, dst, src, nbytes: int):
        """copy-engine copy on the scene's copy stream, ordered after the scene's stream (dst / src:
        device tensors or raw addresses, e.g. a slice of a peer buffer)"""
        _capi.check(self.lib.b2_scene_copy_async(self._h, _ptr(dst), _ptr(src), int(nbytes)))

    def join(self):
        """the scene's stream waits for the copies enqueued so far"""
        _capi.check(self.lib.b2_scene_join(self._h))

    def copy_fence(self):
        """streaming loop: mark the copies so far as this step's; wait for the PREVIOUS step's only"""
        _capi.check(self.lib.b2_scene_copy_fence(self._h))

    def stats(self) -> dict:
        nn, npl = C.c_int32(), C.c_int32()
        mn, mt = C.c_int64(), C.c_int64()
        _capi.check(self.lib.b2_scene_stats(self._h, C.byref(nn), C.byref(npl), C.byref(mn),
                                            C.byref(mt)))
        return {"bvh_nodes": nn.value, "planes": npl.value, "mesh_nodes": mn.value,
                "mesh_tris": mt.value}

    def multiray(self, pnt, vec, geomgroup=None, flg_static=1, bodyexclude=-1, cutoff=1e10,
                 world: int = 0):
        """mj_multiRay-shaped call: returns (dist f64[n], geomid i32[n])."""
        pnt = np.ascontiguousarray(pnt, np.float64).reshape(3)
        vec = np.ascontiguousarray(vec, np.float64).reshape(-1, 3)
        n = vec.shape[0]
        gg = None if geomgroup is None else np.ascontiguousarray(geomgroup, np.uint8)
        dist = np.empty(n, np.float64)
        gid = np.empty(n, np.int32)
        _capi.check(self.lib.b2_multiray(self._h, world, _np_ptr(pnt), _np_ptr(vec), _np_ptr(gg),
                                         flg_static, bodyexclude, _np_ptr(gid), _np_ptr(dist), n,
                                         float(cutoff)))
        return dist, gid


class Lidar:
    """One lidar sensor model bound to a scene (``b2_lidar``)."""

    def __init__(self, scene: Scene, spec: LidarSpec):
        self.scene = scene
        self.spec = spec
        self.lib = scene.lib
        d = _capi.LidarDesc()
        d.nray = spec.nray
        d.local_dir = spec.local_dir.ctypes.data_as(_capi.f64p)
        d.tile_w = int(spec.tile_w)
        d.geomgroup = (spec.geomgroup.ctypes.data_as(_capi.u8p) if spec.geomgroup is not None
                       else C.cast(None, _capi.u8p))
        d.flg_static = int(spec.flg_static)
        d.bodyexclude = int(spec.bodyexclude)
        d.cutoff = float(spec.cutoff)
        d.range_min = float(spec.range_min)
        d.range_max = float(spec.range_max)
        h = C.c_void_p()
        _capi.check(self.lib.b2_lidar_create(scene._h, C.byref(d), C.byref(h)))
        self._h = h
        scene._children.add(self)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b2_lidar_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nray(self) -> int:
        return self.spec.nray

    def cast(self, sensor_pos, sensor_mat, scan_world=None, want=("ranges", "geomid"), out=None):
        """Host path (blocking).  sensor_pos [S,3], sensor_mat [S,9] -> dict of numpy arrays.
        ``out`` may supply preallocated (e.g. pinned) buffers keyed like ``want``."""
        pos = np.ascontiguousarray(sensor_pos, np.float64).reshape(-1, 3)
        mat = np.ascontiguousarray(sensor_mat, np.float64).reshape(-1, 9)
        s = pos.shape[0]
        sw = None if scan_world is None else np.ascontiguousarray(scan_world, np.int32)
        n = s * self.nray
        out = dict(out or {})
        shapes = {"ranges": ((s, self.nray), np.float32), "geomid": ((s, self.nray), np.int32),
                  "dist": ((s, self.nray), np.float64), "xyz": ((s, self.nray, 3), np.float32)}
        for k in want:
            if k not in out:
                out[k] = np.empty(shapes[k][0], shapes[k][1])
        del n
        _capi.check(self.lib.b2_lidar_cast(self._h, s, _ptr(sw), _ptr(pos), _ptr(mat),
                                           _ptr(out.get("ranges")), _ptr(out.get("geomid")),
                                           _ptr(out.get("dist")), _ptr(out.get("xyz"))))
        return out

    def cast_begin(self, sensor_pos, sensor_mat, scan_world=None, want=("ranges", "geomid"), out=None):
        """b2_lidar_cast_begin: enqueue only.  Returns the dict ``cast_end`` will have filled."""
        pos = np.ascontiguousarray(sensor_pos, np.float64).reshape(-1, 3)
        mat = np.ascontiguousarray(sensor_mat, np.float64).reshape(-1, 9)
        s = pos.shape[0]
        sw = None if scan_world is None else np.ascontiguousarray(scan_world, np.int32)
        out = dict(out or {})
        shapes = {"ranges": ((s, self.nray), np.float32), "geomid": ((s, self.nray), np.int32),
                  "dist": ((s, self.nray), np.float64), "xyz": ((s, self.nray, 3), np.float32)}
        for k in want:
            if k not in out:
                out[k] = np.empty(shapes[k][0], shapes[k][1])
        self._pending = (pos, mat, sw, out)  # the arrays must outlive the call
        _capi.check(self.lib.b2_lidar_cast_begin(self._h, s, _ptr(sw), _ptr(pos), _ptr(mat),
                                                 _ptr(out.get("ranges")), _ptr(out.get("geomid")),
                                                 _ptr(out.get("dist")), _ptr(out.get("xyz"))))
        return out

    def cast_end(self):
        _capi.check(self.lib.b2_lidar_cast_end(self._h))
        out = self._pending[3]
        self._pending = None
        return out

    def cast_u16(self, sensor_pos, sensor_mat, scan_world=None, out=None):
        """ranges f32 + geom ids as uint16 (0xFFFF = miss): 6 B/ray over the host link."""
        pos = np.ascontiguousarray(sensor_pos, np.float64).reshape(-1, 3)
        mat = np.ascontiguousarray(sensor_mat, np.float64).reshape(-1, 9)
        s = pos.shape[0]
        sw = None if scan_world is None else np.ascontiguousarray(scan_world, np.int32)
        out = dict(out or {})
        if "ranges" not in out:
            out["ranges"] = np.empty((s, self.nray), np.float32)
        if "geomid16" not in out:
            out["geomid16"] = np.empty((s, self.nray), np.uint16)
        _capi.check(self.lib.b2_lidar_cast_u16(self._h, s, _ptr(sw), _ptr(pos), _ptr(mat),
                                               _ptr(out["ranges"]), _ptr(out["geomid16"])))
        return out

    def cast_u16_device(self, nscan, d_pos, d_mat, d_world=None, ranges=None, geomid16=None):
        _capi.check(self.lib.b2_lidar_cast_u16_device(self._h, int(nscan), _ptr(d_world), _ptr(d_pos),
                                                      _ptr(d_mat), _ptr(ranges), _ptr(geomid16)))

    def last_device_us(self) -> float:
        """device time of the last host cast (first enqueue .. last kernel), microseconds"""
        return float(self.lib.b2_lidar_last_device_us(self._h))

    def cast_device(self, nscan, d_pos, d_mat, d_world=None, ranges=None, geomid=None, dist=None,
                    xyz=None):
        """Device path (async): all arguments are torch CUDA tensors (or None)."""
        _capi.check(self.lib.b2_lidar_cast_device(self._h, int(nscan), _ptr(d_world), _ptr(d_pos),
                                                  _ptr(d_mat), _ptr(ranges), _ptr(geomid),
                                                  _ptr(dist), _ptr(xyz)))


class DepthCamera:
    """One camera model (``b2_camera``): depth linearisation + PointCloud2 projection."""

    def __init__(self, spec: CameraSpec, scene: Scene | None = None, device: int = 0):
        self.lib = _capi.load()
        self.spec = spec
        self.scene = scene
        d = _capi.CameraDesc(spec.width, spec.height, spec.fovy, spec.znear, spec.zfar, spec.extent,
                             spec.point_step)
        gg = getattr(spec, "geomgroup", None)  # None: mjv_defaultOption (groups 0-2 visible)
        if gg is not None:
            for i in range(6):
                d.geomgroup[i] = int(gg[i])
            d.geomgroup_set = 1
        h = C.c_void_p()
        _capi.check(self.lib.b2_camera_create(scene._h if scene else None, device, C.byref(d),
                                              C.byref(h)))
        self._h = h
        if scene is not None:
            scene._children.add(self)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b2_camera_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        k = np.zeros(9, np.float64)
        p = np.zeros(12, np.float64)
        _capi.check(self.lib.b2_camera_info(self._h, k.ctypes.data_as(_capi.f64p),
                                            p.ctypes.data_as(_capi.f64p)))
        return k, p

    def depth_to_cloud(self, gl_depth, want_depth=True, want_cloud=True, out=None):
        """Host path.  gl_depth [N,H,W] float32 (GL layout) -> (depth32fc1 [N,H,W], cloud
        [N,H,W,point_step/4])."""
        d = gl_depth if not isinstance(gl_depth, np.ndarray) else np.ascontiguousarray(gl_depth, np.float32)
        h, w = self.spec.height, self.spec.width
        n = int(np.prod(d.shape)) // (h * w)
        pf = self.spec.point_step // 4
        out = dict(out or {})
        if want_depth and "depth" not in out:
            out["depth"] = np.empty((n, h, w), np.float32)
        if want_cloud and "cloud" not in out:
            out["cloud"] = np.empty((n, h, w, pf), np.float32)
        _capi.check(self.lib.b2_depth_to_cloud(self._h, n, _ptr(d), _ptr(out.get("depth")),
                                               _ptr(out.get("cloud"))))
        return out.get("depth"), out.get("cloud")

    def depth_to_cloud_device(self, ncam, d_gl_depth, d_depth=None, d_cloud=None):
        _capi.check(self.lib.b2_depth_to_cloud_device(self._h, int(ncam), _ptr(d_gl_depth),
                                                      _ptr(d_depth), _ptr(d_cloud)))

    def rgb_flip_device(self, ncam, d_rgb_gl, d_rgb_out):
        """uint8 CUDA tensors [ncam,H,W,3]: GL bottom-up rows -> Image top-down rows."""
        _capi.check(self.lib.b2_rgb_flip_device(self._h, int(ncam), _ptr(d_rgb_gl), _ptr(d_rgb_out)))

    def render_to_cloud(self, cam_pos, cam_mat, cam_world=None, want=("gl", "depth", "cloud")):
        """Host path: ray-cast GL depth of the camera poses, then linearise + flip + cloud."""
        pos = np.ascontiguousarray(cam_pos, np.float64).reshape(-1, 3)
        mat = np.ascontiguousarray(cam_mat, np.float64).reshape(-1, 9)
        n, h, w = pos.shape[0], self.spec.height, self.spec.width
        sw = None if cam_world is None else np.ascontiguousarray(cam_world, np.int32)
        gl = np.empty((n, h, w), np.float32) if "gl" in want else None
        depth = np.empty((n, h, w), np.float32) if "depth" in want else None
        cloud = np.empty((n, h, w, self.spec.point_step // 4), np.float32) if "cloud" in want else None
        _capi.check(self.lib.b2_camera_render_to_cloud(self._h, n, _ptr(sw), _ptr(pos), _ptr(mat),
                                                       _ptr(gl), _ptr(depth), _ptr(cloud)))
        return gl, depth, cloud

    def render_depth_device(self, ncam, d_pos, d_mat, d_gl_depth, d_world=None):
        _capi.check(self.lib.b2_camera_render_depth_device(self._h, int(ncam), _ptr(d_world),
                                                           _ptr(d_pos), _ptr(d_mat),
                                                           _ptr(d_gl_depth)))
