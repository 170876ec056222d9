"""ctypes harness for the C++ wrapper classes (mujoco_ros2_simulation_b200/plugin): builds the
stand-in mjModel / mjData of b2_compat.hpp from the flat numpy scene + the MJCF-subset reader's
sites / sensors / cameras, and drives MujocoLidar / MujocoCameras through plugin/test_api.cpp the way
MujocoSystemInterface drives the reference classes."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
_PLUGIN_DIR = os.path.join(_ROOT, "mujoco_ros2_simulation_b200", "plugin")
_SO = os.path.join(_ROOT, "mujoco_ros2_simulation_b200", "libb2plugin.so")

ip = C.POINTER(C.c_int)
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)


class MjModel(C.Structure):
    """field order of `struct mjModel` in plugin/b2_compat.hpp"""
    _fields_ = [(n, C.c_int) for n in ("ngeom", "nbody", "nmesh", "nmat", "nsite", "ncam", "nsensor",
                                       "nsensordata")] + [
        ("geom_type", ip), ("geom_bodyid", ip), ("geom_group", ip), ("geom_matid", ip),
        ("geom_dataid", ip), ("geom_size", dp), ("geom_rgba", fp), ("mat_rgba", fp),
        ("body_weldid", ip), ("mesh_vertadr", ip), ("mesh_vertnum", ip), ("mesh_faceadr", ip),
        ("mesh_facenum", ip), ("mesh_vert", fp), ("mesh_face", ip),
        ("site_bodyid", ip), ("site_pos", dp), ("site_quat", dp),
        ("sensor_type", ip), ("sensor_objid", ip), ("sensor_adr", ip), ("sensor_cutoff", dp),
        ("name_sensoradr", ip), ("name_camadr", ip), ("names", C.c_char_p),
        ("cam_resolution", ip), ("cam_fovy", dp),
        ("znear", C.c_float), ("zfar", C.c_float), ("extent", C.c_double),
    ]


class MjData(C.Structure):
    _fields_ = [("time", C.c_double), ("geom_xpos", dp), ("geom_xmat", dp), ("xpos", dp), ("xmat", dp),
                ("cam_xpos", dp), ("cam_xmat", dp), ("sensordata", dp)]


def build() -> str:
    subprocess.run(["make", "-C", _PLUGIN_DIR, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.b2p_create.restype = C.c_void_p
        L.b2p_create.argtypes = [C.POINTER(MjModel), C.POINTER(MjData), C.c_char_p, C.c_double, C.c_double]
        for name in ("b2p_register", "b2p_activate", "b2p_update", "b2p_lidar_count", "b2p_camera_count"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.b2p_run_threads.restype = C.c_int
        L.b2p_run_threads.argtypes = [C.c_void_p, C.c_int]
        L.b2p_error.restype = C.c_char_p
        L.b2p_error.argtypes = [C.c_void_p]
        L.b2p_destroy.restype = None
        L.b2p_destroy.argtypes = [C.c_void_p]
        L.b2p_lidar_info.restype = C.c_int
        L.b2p_lidar_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.b2p_lidar_scan.restype = C.c_int
        L.b2p_lidar_scan.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        L.b2p_camera_info.restype = C.c_int
        L.b2p_camera_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.b2p_camera_frames.restype = C.c_int
        L.b2p_camera_frames.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
        L.b2p_camera_set_gl_depth.restype = C.c_int
        L.b2p_camera_set_gl_depth.argtypes = [C.c_void_p, C.c_void_p]
        L.b2p_camera_rgb.restype = C.c_int
        L.b2p_camera_rgb.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.b2p_wire_bytes.restype = C.c_long
        L.b2p_wire_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long]
        L.b2p_logs.restype = C.c_int
        L.b2p_logs.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.b2p_camera_pub_counts.restype = C.c_int
        L.b2p_camera_pub_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2p_refresh_stress.restype = C.c_int
        L.b2p_refresh_stress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.b2p_sizeof_mjmodel.restype = C.c_size_t
        L.b2p_sizeof_mjdata.restype = C.c_size_t
        assert L.b2p_sizeof_mjmodel() == C.sizeof(MjModel), "mjModel stand-in layout drifted"
        assert L.b2p_sizeof_mjdata() == C.sizeof(MjData), "mjData stand-in layout drifted"
        _lib = L
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size else C.cast(None, t)


class PluginHarness:
    """An MJCF-subset scene (mjcf.MjcfScene) presented to the wrappers as mjModel / mjData."""

    def __init__(self, scene, hardware_info: str, qpos=None, lidar_rate=5.0, camera_rate=5.0,
                 extent=1.0):
        from mujoco_ros2_simulation_b200.model import quat2mat
        self.L = lib()
        m = scene.model
        self.scene = scene
        k = self._keep = {}
        for name, dt in (("geom_type", np.int32), ("geom_bodyid", np.int32), ("geom_group", np.int32),
                         ("geom_matid", np.int32), ("geom_dataid", np.int32), ("geom_size", np.float64),
                         ("geom_rgba", np.float32), ("mat_rgba", np.float32), ("body_weldid", np.int32),
                         ("mesh_vertadr", np.int32), ("mesh_vertnum", np.int32), ("mesh_faceadr", np.int32),
                         ("mesh_facenum", np.int32), ("mesh_vert", np.float32), ("mesh_face", np.int32)):
            k[name] = np.ascontiguousarray(getattr(m, name), dt)
        k["site_bodyid"] = np.array([s.body for s in scene.sites], np.int32)
        k["site_pos"] = np.array([s.frame.pos for s in scene.sites], np.float64).reshape(-1, 3)
        k["site_quat"] = np.array([s.frame.quat for s in scene.sites], np.float64).reshape(-1, 4)
        # rangefinder sensors in MJCF order, one sensordata slot each; names packed like mjModel.names
        rfs = scene.rangefinders
        ns = len(rfs)
        k["sensor_type"] = np.full(ns, 7, np.int32)
        k["sensor_objid"] = np.array([r.site for r in rfs], np.int32)
        k["sensor_adr"] = np.arange(ns, dtype=np.int32)
        k["sensor_cutoff"] = np.array([r.cutoff for r in rfs], np.float64)
        blob, sadr, cadr = b"\0", [], []
        for r in rfs:
            sadr.append(len(blob))
            blob += r.name.encode() + b"\0"
        for c in scene.cameras:
            cadr.append(len(blob))
            blob += c.name.encode() + b"\0"
        k["name_sensoradr"] = np.array(sadr, np.int32)
        k["name_camadr"] = np.array(cadr, np.int32)
        self._names = C.create_string_buffer(blob, len(blob))
        k["cam_resolution"] = np.array([c.resolution for c in scene.cameras], np.int32).reshape(-1, 2)
        k["cam_fovy"] = np.array([c.fovy for c in scene.cameras], np.float64)

        mm = MjModel()
        mm.ngeom, mm.nbody, mm.nmesh, mm.nmat = m.ngeom, m.nbody, m.nmesh, m.nmat
        mm.nsite, mm.ncam, mm.nsensor, mm.nsensordata = len(scene.sites), len(scene.cameras), ns, ns
        for name, t in (("geom_type", ip), ("geom_bodyid", ip), ("geom_group", ip), ("geom_matid", ip),
                        ("geom_dataid", ip), ("geom_size", dp), ("geom_rgba", fp), ("mat_rgba", fp),
                        ("body_weldid", ip), ("mesh_vertadr", ip), ("mesh_vertnum", ip),
                        ("mesh_faceadr", ip), ("mesh_facenum", ip), ("mesh_vert", fp), ("mesh_face", ip),
                        ("site_bodyid", ip), ("site_pos", dp), ("site_quat", dp), ("sensor_type", ip),
                        ("sensor_objid", ip), ("sensor_adr", ip), ("sensor_cutoff", dp),
                        ("name_sensoradr", ip), ("name_camadr", ip), ("cam_resolution", ip),
                        ("cam_fovy", dp)):
            setattr(mm, name, _ptr(k[name], t))
        mm.names = C.cast(self._names, C.c_char_p)
        mm.znear, mm.zfar, mm.extent = scene.znear, scene.zfar, extent
        self.mjmodel = mm

        # mjData at qpos
        poses = scene.poses(qpos)
        bpos, bmat, _ = scene.body_poses(qpos)
        k["geom_xpos"] = np.ascontiguousarray(poses.geom_xpos[0])
        k["geom_xmat"] = np.ascontiguousarray(poses.geom_xmat[0])
        k["xpos"] = np.ascontiguousarray(bpos)
        k["xmat"] = np.ascontiguousarray(bmat.reshape(-1, 9))
        cpos = np.zeros((len(scene.cameras), 3))
        cmat = np.zeros((len(scene.cameras), 9))
        for i, c in enumerate(scene.cameras):
            cpos[i] = bpos[c.body] + bmat[c.body] @ c.frame.pos
            cmat[i] = (bmat[c.body] @ quat2mat(c.frame.quat)).reshape(9)
        k["cam_xpos"], k["cam_xmat"] = cpos, cmat
        k["sensordata"] = np.zeros(max(ns, 1))
        dd = MjData()
        for name in ("geom_xpos", "geom_xmat", "xpos", "xmat", "cam_xpos", "cam_xmat", "sensordata"):
            setattr(dd, name, _ptr(k[name], dp))
        self.mjdata = dd
        self.poses = poses
        self.h = C.c_void_p(self.L.b2p_create(C.byref(mm), C.byref(dd), hardware_info.encode(),
                                              lidar_rate, camera_rate))

    def close(self):
        if self.h:
            self.L.b2p_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self) -> str:
        return self.L.b2p_error(self.h).decode()

    def register(self) -> int:
        return self.L.b2p_register(self.h)

    def activate(self) -> int:
        return self.L.b2p_activate(self.h)

    def update(self) -> int:
        return self.L.b2p_update(self.h)

    def run_threads(self, ms: int) -> int:
        return self.L.b2p_run_threads(self.h, ms)

    def wire_bytes(self, which: str, k=0) -> bytes:
        """last published message as classic CDR (plugin/cdr.hpp): 'scan', 'depth', 'info', 'points', 'image'"""
        code = {"scan": 0, "depth": 1, "info": 2, "points": 3, "image": 4}[which]
        n = self.L.b2p_wire_bytes(self.h, code, k, None, 0)
        buf = (C.c_uint8 * n)()
        self.L.b2p_wire_bytes(self.h, code, k, buf, n)
        return bytes(buf)

    def logs(self) -> list:
        buf = C.create_string_buffer(1 << 16)
        self.L.b2p_logs(self.h, buf, len(buf))
        return [l for l in buf.value.decode().split("\n") if l]

    def camera_pub_counts(self, k=0) -> dict:
        out = np.zeros(4, np.int32)
        self.L.b2p_camera_pub_counts(self.h, k, out.ctypes.data)
        return dict(zip(("image", "depth", "info", "points"), (int(v) for v in out)))

    def refresh_stress(self, n_el=64, n_az=2048, hz=1500.0, seconds=2.0) -> dict:
        """physics-thread refresh at `hz` while a sensor thread casts back to back (test_api.cpp)"""
        out = np.zeros(5, np.float64)
        rc = self.L.b2p_refresh_stress(self.h, n_el, n_az, hz, seconds, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"b2p_refresh_stress {rc}: {self.error()}")
        return {"p50_us": out[0], "p99_us": out[1], "max_us": out[2], "casts": int(out[3]),
                "refreshes": int(out[4])}

    def lidar_info(self, k=0) -> dict:
        n = C.c_int()
        f7 = np.zeros(7, np.float32)
        body = C.c_int()
        frame, topic = C.create_string_buffer(128), C.create_string_buffer(128)
        idx = np.zeros(4096, np.int32)
        ld = np.zeros(4096 * 3, np.float64)
        self.L.b2p_lidar_info(self.h, k, C.byref(n), f7.ctypes.data, C.byref(body), frame, topic,
                              idx.ctypes.data, ld.ctypes.data)
        nn = n.value
        return {"num_rangefinders": nn, "angle_min": f7[0], "angle_max": f7[1], "angle_increment": f7[2],
                "time_increment": f7[3], "scan_time": f7[4], "range_min": f7[5], "range_max": f7[6],
                "bodyid": body.value, "frame_id": frame.value.decode(), "topic": topic.value.decode(),
                "sensor_indexes": idx[:nn].copy(), "local_dir": ld[:3 * nn].reshape(nn, 3).copy()}

    def lidar_scan(self, k=0):
        ranges = np.zeros(4096, np.float32)
        npub, nint = C.c_int(), C.c_int()
        n = self.L.b2p_lidar_scan(self.h, k, ranges.ctypes.data, C.byref(npub), C.byref(nint))
        return ranges[:n].copy(), npub.value, nint.value

    def camera_info(self, k=0) -> dict:
        wh = np.zeros(2, np.int32)
        K, R, P = np.zeros(9), np.ones(9), np.zeros(12)
        nd = C.c_int()
        frame = C.create_string_buffer(128)
        topics = C.create_string_buffer(512)
        enc = C.create_string_buffer(32)
        steps = np.zeros(4, np.int32)
        self.L.b2p_camera_info(self.h, k, wh.ctypes.data, K.ctypes.data, R.ctypes.data, P.ctypes.data,
                               C.byref(nd), frame, topics, enc, steps.ctypes.data)
        raw = topics.raw
        tp = [raw[i * 128:(i + 1) * 128].split(b"\0")[0].decode() for i in range(4)]
        return {"width": int(wh[0]), "height": int(wh[1]), "K": K, "R": R, "P": P, "nd": nd.value,
                "frame_id": frame.value.decode(), "info_topic": tp[0], "image_topic": tp[1],
                "depth_topic": tp[2], "points_topic": tp[3],
                "image_encoding": enc.raw[:16].split(b"\0")[0].decode(),
                "depth_encoding": enc.raw[16:].split(b"\0")[0].decode(),
                "image_step": int(steps[0]), "depth_step": int(steps[1]), "point_step": int(steps[2]),
                "row_step": int(steps[3])}

    def camera_frames(self, k=0):
        ci = self.camera_info(k)
        n = ci["width"] * ci["height"]
        gl, depth, cloud = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n * 3, np.float32)
        npub = self.L.b2p_camera_frames(self.h, k, gl.ctypes.data, depth.ctypes.data, cloud.ctypes.data)
        shp = (ci["height"], ci["width"])
        return gl.reshape(shp), depth.reshape(shp), cloud.reshape(shp + (3,)), npub

    def set_gl_depth(self, gl_depth):
        self._gl = np.ascontiguousarray(gl_depth, np.float32)
        self.L.b2p_camera_set_gl_depth(self.h, self._gl.ctypes.data)

    def camera_rgb(self, k=0):
        ci = self.camera_info(k)
        n = ci["width"] * ci["height"] * 3
        a, b = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.L.b2p_camera_rgb(self.h, k, a.ctypes.data, b.ctypes.data)
        shp = (ci["height"], ci["width"], 3)
        return a.reshape(shp), b.reshape(shp)
