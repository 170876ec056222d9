"""ctypes declarations of the C ABI in include/b2sense.h (libb2sense.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.  There is no
fallback: if the shared object is missing, importing the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2_LIB_PATH") or os.path.join(_HERE, "libb2sense.so")  # override: A/B builds

i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)

B2_OK, B2_ERR_INVALID, B2_ERR_CUDA, B2_ERR_UNSUPPORTED, B2_ERR_NOMEM = 0, -1, -2, -3, -4


class ModelDesc(C.Structure):
    _fields_ = [
        ("ngeom", C.c_int32), ("nbody", C.c_int32), ("nmesh", C.c_int32), ("nmat", C.c_int32),
        ("geom_type", i32p), ("geom_bodyid", i32p), ("geom_group", i32p), ("geom_matid", i32p),
        ("geom_dataid", i32p), ("geom_size", f64p), ("geom_rgba", f32p), ("mat_rgba", f32p),
        ("body_weldid", i32p), ("mesh_vertadr", i32p), ("mesh_vertnum", i32p),
        ("mesh_faceadr", i32p), ("mesh_facenum", i32p), ("mesh_vert", f32p), ("mesh_face", i32p),
    ]


class LidarDesc(C.Structure):
    _fields_ = [
        ("nray", C.c_int32), ("local_dir", f64p), ("tile_w", C.c_int32), ("geomgroup", u8p),
        ("flg_static", C.c_uint8), ("bodyexclude", C.c_int32), ("cutoff", C.c_double),
        ("range_min", C.c_double), ("range_max", C.c_double),
    ]


class CameraDesc(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("fovy", C.c_double), ("znear", C.c_double),
        ("zfar", C.c_double), ("extent", C.c_double), ("point_step", C.c_int32),
        ("geomgroup", C.c_uint8 * 6), ("geomgroup_set", C.c_uint8),
    ]


# every symbol include/b2sense.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b2_version": (C.c_int, []),
    "b2_device_count": (C.c_int, []),
    "b2_last_error": (C.c_char_p, []),
    "b2_launch_count": (C.c_longlong, []),
    "b2_scene_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_void_p,
                                  C.POINTER(C.c_void_p)]),
    "b2_scene_destroy": (C.c_int, [C.c_void_p]),
    "b2_scene_update_poses": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "b2_scene_update_poses_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p]),
    "b2_scene_rebuild": (C.c_int, [C.c_void_p, C.c_int32]),
    "b2_scene_sync": (C.c_int, [C.c_void_p]),
    "b2_scene_stats": (C.c_int, [C.c_void_p, i32p, i32p, C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64)]),
    "b2_lidar_create": (C.c_int, [C.c_void_p, C.POINTER(LidarDesc), C.POINTER(C.c_void_p)]),
    "b2_lidar_destroy": (C.c_int, [C.c_void_p]),
    "b2_lidar_cast": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    "b2_lidar_cast_device": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    "b2_lidar_cast_begin": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    "b2_lidar_cast_end": (C.c_int, [C.c_void_p]),
    "b2_lidar_cast_u16": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 5),
    "b2_lidar_cast_u16_device": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 5),
    "b2_lidar_last_device_us": (C.c_double, [C.c_void_p]),
    "b2_camera_render_to_cloud_begin": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 6),
    "b2_camera_render_to_cloud_end": (C.c_int, [C.c_void_p]),
    "b2_multiray": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint8,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_double]),
    "b2_camera_create": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(CameraDesc),
                                   C.POINTER(C.c_void_p)]),
    "b2_camera_destroy": (C.c_int, [C.c_void_p]),
    "b2_camera_info": (C.c_int, [C.c_void_p, f64p, f64p]),
    "b2_depth_to_cloud": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2_depth_to_cloud_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "b2_camera_render_depth_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "b2_camera_render_to_cloud": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 6),
    "b2_rgb_flip_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "b2_scene_copy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b2_scene_join": (C.c_int, [C.c_void_p]),
    "b2_scene_copy_fence": (C.c_int, [C.c_void_p]),
    "b2_peer_buffer_create": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p), u8p]),
    "b2_peer_buffer_open": (C.c_int, [C.c_int32, u8p, C.POINTER(C.c_void_p)]),
    "b2_peer_buffer_close": (C.c_int, [C.c_int32, C.c_void_p]),
    "b2_peer_buffer_destroy": (C.c_int, [C.c_int32, C.c_void_p]),
}
PEER_HANDLE_BYTES = 64

_lib = None


def load():
    """dlopen libb2sense.so and declare its prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). This engine has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class B2Error(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"b2sense error {code}: {text}")
        self.code = code


def check(rc: int):
    if rc != B2_OK:
        raise B2Error(rc, load().b2_last_error().decode("utf-8", "replace"))
