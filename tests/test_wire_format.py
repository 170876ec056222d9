"""N3: the published messages as they go on the wire.  plugin/cdr.hpp serialises what the stand-in
publisher last received as classic CDR (what rmw_fastrtps / rmw_cyclonedds put on the wire for
sensor_msgs); here an INDEPENDENT parser, written from the .msg definitions, reads the bytes back:
field order, alignment, string / sequence framing and the payloads must come out as the reference
fills them (ref:src/mujoco_lidar.cpp:101-109,275-279; ref:src/mujoco_cameras.cpp:91-124,253-264)."""
import os
import struct

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import mjcf
from tests.test_plugin import GOLDEN, INFO


class Cdr:
    """classic CDR, little endian: alignment relative to the end of the 4-byte encapsulation header"""

    def __init__(self, data: bytes):
        assert data[:4] == b"\x00\x01\x00\x00", "CDR_LE encapsulation header"
        self.d, self.p = data, 4

    def _align(self, n):
        self.p += (n - (self.p - 4) % n) % n

    def prim(self, fmt):
        n = struct.calcsize(fmt)
        self._align(n)
        v = struct.unpack_from("<" + fmt, self.d, self.p)[0]
        self.p += n
        return v

    def string(self):
        n = self.prim("I")
        s = self.d[self.p:self.p + n - 1].decode()
        assert self.d[self.p + n - 1] == 0
        self.p += n
        return s

    def seq(self, dtype):
        n = self.prim("I")
        return self.array(dtype, n)

    def array(self, dtype, n):
        dt = np.dtype(dtype)
        if n:
            self._align(dt.itemsize)
        a = np.frombuffer(self.d, dt, n, self.p).copy()
        self.p += n * dt.itemsize
        return a

    def header(self):
        return {"sec": self.prim("i"), "nanosec": self.prim("I"), "frame_id": self.string()}

    def done(self):
        return self.p == len(self.d)


def parse_laserscan(b):
    c = Cdr(b)
    m = {"header": c.header()}
    for k in ("angle_min", "angle_max", "angle_increment", "time_increment", "scan_time", "range_min", "range_max"):
        m[k] = c.prim("f")
    m["ranges"], m["intensities"] = c.seq("<f4"), c.seq("<f4")
    assert c.done()
    return m


def parse_image(b):
    c = Cdr(b)
    m = {"header": c.header(), "height": c.prim("I"), "width": c.prim("I"), "encoding": c.string(),
         "is_bigendian": c.prim("B"), "step": c.prim("I")}
    m["data"] = c.seq("u1")
    assert c.done()
    return m


def parse_camera_info(b):
    c = Cdr(b)
    m = {"header": c.header(), "height": c.prim("I"), "width": c.prim("I"), "distortion_model": c.string()}
    m["d"] = c.seq("<f8")
    m["k"], m["r"], m["p"] = c.array("<f8", 9), c.array("<f8", 9), c.array("<f8", 12)
    m["binning"] = (c.prim("I"), c.prim("I"))
    m["roi"] = tuple(c.prim("I") for _ in range(4)) + (c.prim("B"),)
    assert c.done()
    return m


def parse_pointcloud2(b):
    c = Cdr(b)
    m = {"header": c.header(), "height": c.prim("I"), "width": c.prim("I")}
    m["fields"] = [(c.string(), c.prim("I"), c.prim("B"), c.prim("I")) for _ in range(c.prim("I"))]
    m["is_bigendian"], m["point_step"], m["row_step"] = c.prim("B"), c.prim("I"), c.prim("I")
    m["data"] = c.seq("u1")
    m["is_dense"] = c.prim("B")
    assert c.done()
    return m


@pytest.fixture(scope="module")
def scene():
    return mjcf.load_mjcf(os.path.join(GOLDEN, "plugin_scene.xml"))


@pytest.fixture(scope="module")
def plug():
    from tests import plugin_binding
    plugin_binding.build()
    return plugin_binding


def test_static_fields_on_the_wire(plug, scene):
    """before anything is published the stand-in publisher holds a default message: serialise the
    registered message templates instead by publishing nothing -- the layout test needs no GPU"""
    h = plug.PluginHarness(scene, INFO, lidar_rate=4.0)
    assert h.register() == 0, h.error()
    # nothing published yet: an all-default LaserScan is 4 (encapsulation) + 8 (stamp) + 4 + 1 (empty
    # frame_id) padded to 4 + 7 floats + two empty sequences
    empty = parse_laserscan(h.wire_bytes("scan"))
    assert empty["header"]["frame_id"] == "" and empty["ranges"].size == 0
    assert len(h.wire_bytes("scan")) == 4 + 8 + 4 + 1 + 3 + 28 + 4 + 4
    h.close()


@pytest.mark.gpu
def test_published_messages_on_the_wire(plug, scene, oracle):
    h = plug.PluginHarness(scene, INFO, lidar_rate=4.0)
    assert h.register() == 0 and h.activate() == 0, h.error()
    gl_in = np.random.default_rng(0).random((48, 64), dtype=np.float32)
    h.set_gl_depth(gl_in)
    assert h.update() == 0, h.error()
    # ---- LaserScan (ref:src/mujoco_lidar.cpp:101-109,175,275-279)
    m = parse_laserscan(h.wire_bytes("scan"))
    ranges, npub, _ = h.lidar_scan()
    assert m["header"]["frame_id"] == "lidar_sensor_frame" and m["header"]["sec"] > 0
    assert m["angle_min"] == np.float32(-0.3) and m["angle_max"] == np.float32(0.3)
    assert m["angle_increment"] == np.float32(0.025) and m["time_increment"] == 0.0
    assert m["scan_time"] == np.float32(0.25)
    assert m["range_min"] == np.float32(0.05) and m["range_max"] == np.float32(10.0)
    assert np.array_equal(m["ranges"], ranges) and m["ranges"].size == 24 and m["intensities"].size == 0
    # ---- depth Image 32FC1 (ref:src/mujoco_cameras.cpp:99-107,240)
    gl, depth, cloud, _ = h.camera_frames()
    d = parse_image(h.wire_bytes("depth"))
    assert (d["height"], d["width"], d["encoding"], d["step"], d["is_bigendian"]) == (48, 64, "32FC1", 256, 0)
    assert d["header"]["frame_id"] == "camera_color_mujoco_frame"
    assert np.array_equal(d["data"].view("<f4").reshape(48, 64), depth)
    # ---- colour Image RGB8 (:91-97)
    i = parse_image(h.wire_bytes("image"))
    assert (i["height"], i["width"], i["encoding"], i["step"]) == (48, 64, "rgb8", 192) and i["data"].size == 48 * 64 * 3
    # ---- CameraInfo (:109-124): plumb_bob, five zero coefficients, K / P from fovy, R all zero
    ci = parse_camera_info(h.wire_bytes("info"))
    info = h.camera_info()
    assert ci["distortion_model"] == "plumb_bob" and ci["d"].tolist() == [0.0] * 5
    assert np.array_equal(ci["k"], info["K"]) and np.array_equal(ci["p"], info["P"]) and not ci["r"].any()
    assert (ci["height"], ci["width"]) == (48, 64) and ci["binning"] == (0, 0) and ci["roi"] == (0, 0, 0, 0, 0)
    # ---- PointCloud2: organised, x / y / z FLOAT32 at 0 / 4 / 8, point_step 12, not dense
    pc = parse_pointcloud2(h.wire_bytes("points"))
    assert (pc["height"], pc["width"], pc["point_step"], pc["row_step"]) == (48, 64, 12, 768)
    assert pc["fields"] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1)]
    assert pc["is_bigendian"] == 0 and pc["is_dense"] == 0
    assert np.array_equal(pc["data"].view("<f4").reshape(48, 64, 3), cloud, equal_nan=True)
    assert pc["header"]["frame_id"] == "camera_color_mujoco_frame"
    h.close()
