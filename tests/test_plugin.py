"""The C++ wrapper classes (plugin/mujoco_lidar.cpp, plugin/mujoco_cameras.cpp) driven the way
MujocoSystemInterface drives the reference's (ref:src/mujoco_system_interface.cpp:707-717,938-939):
CPU tier = registration / message layout / error behaviour; GPU tier = one publish cycle checked
against the oracle."""
import os

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import mjcf

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
INFO = ("lidar|frame_name=lidar_sensor_frame;angle_increment=0.025;min_angle=-0.3;max_angle=0.3;"
        "range_min=0.05;range_max=10;laserscan_topic=/scan\n"
        "camera|frame_name=camera_color_mujoco_frame;info_topic=/camera/color/camera_info;"
        "image_topic=/camera/color/image_raw;depth_topic=/camera/aligned_depth_to_color/image_raw")


@pytest.fixture(scope="module")
def scene():
    return mjcf.load_mjcf(os.path.join(GOLDEN, "plugin_scene.xml"))


@pytest.fixture(scope="module")
def plug():
    from tests import plugin_binding
    plugin_binding.build()
    return plugin_binding


def test_register_lidar_message_layout(plug, scene):
    h = plug.PluginHarness(scene, INFO, lidar_rate=4.0)
    assert h.register() == 0, h.error()
    assert h.L.b2p_lidar_count(h.h) == 1
    li = h.lidar_info()
    assert li["num_rangefinders"] == 24                       # int(0.6/0.025)+1 with f64 truncation
    assert li["angle_min"] == np.float32(-0.3) and li["angle_max"] == np.float32(0.3)  # configured max
    assert li["angle_increment"] == np.float32(0.025) and li["time_increment"] == 0.0
    assert li["scan_time"] == np.float32(0.25)                # 1 / lidar_publish_rate
    assert li["range_min"] == np.float32(0.05) and li["range_max"] == np.float32(10.0)
    assert li["frame_id"] == "lidar_sensor_frame" and li["topic"] == "/scan"
    assert li["sensor_indexes"].tolist() == list(range(24))   # lidar-00 .. lidar-23 -> sensor_adr
    assert li["bodyid"] == scene.body_id("scanner_lidar_body")
    a = 0.025 * np.arange(24)
    np.testing.assert_allclose(li["local_dir"], np.stack([np.sin(a), 0 * a, np.cos(a)], 1), atol=1e-12)
    h.close()


def test_register_lidar_defaults_and_failure(plug, scene):
    h = plug.PluginHarness(scene, "lidar|frame_name=f;angle_increment=0.1;min_angle=0;max_angle=1.0")
    assert h.register() == 0
    li = h.lidar_info()
    assert li["topic"] == "/scan" and li["range_min"] == 0.0 and li["range_max"] == 1000.0
    assert li["num_rangefinders"] == 11
    h.close()
    # a required parameter is missing -> register_lidar false -> on_init FAILURE (ref :713-717)
    h = plug.PluginHarness(scene, "lidar|frame_name=f;min_angle=0;max_angle=1.0")
    assert h.register() == -1 and "lidar" in h.error()
    h.close()
    # no <sensor name="lidar"> at all
    h = plug.PluginHarness(scene, "")
    assert h.register() == -1
    h.close()


def test_register_cameras_message_layout(plug, scene):
    h = plug.PluginHarness(scene, INFO)
    assert h.register() == 0
    ci = h.camera_info()
    assert (ci["width"], ci["height"]) == (64, 48)
    f = (1.0 / np.tan(np.deg2rad(58.0) / 2.0)) * 48 / 2.0
    assert ci["K"][0] == f and ci["K"][4] == f and ci["K"][2] == 32.0 and ci["K"][5] == 24.0 and ci["K"][8] == 1.0
    assert ci["P"][0] == f and ci["P"][5] == f and ci["P"][2] == 32.0 and ci["P"][6] == 24.0 and ci["P"][10] == 1.0
    assert not ci["R"].any() and ci["nd"] == 5                # reference quirk: R left all-zero
    assert ci["image_encoding"] == "rgb8" and ci["depth_encoding"] == "32FC1"
    assert ci["image_step"] == 64 * 3 and ci["depth_step"] == 64 * 4
    assert ci["point_step"] == 12 and ci["row_step"] == 64 * 12
    assert ci["frame_id"] == "camera_color_mujoco_frame"
    assert ci["depth_topic"] == "/camera/aligned_depth_to_color/image_raw"
    assert ci["points_topic"] == ci["depth_topic"] + "/points"
    h.close()
    # defaults when the URDF has no <sensor> for the camera
    h = plug.PluginHarness(scene, INFO.split("\n")[0])
    assert h.register() == 0
    ci = h.camera_info()
    assert ci["frame_id"] == "camera_frame" and ci["info_topic"] == "camera/camera_info"
    assert ci["image_topic"] == "camera/color" and ci["depth_topic"] == "camera/depth"
    h.close()
    # all-or-nothing camera parameters: a missing one throws out of register_cameras (ref :63-66)
    h = plug.PluginHarness(scene, INFO.split("\n")[0] + "\ncamera|frame_name=x;info_topic=y")
    assert h.register() == -2
    h.close()


def test_registration_is_logged_and_strict(plug, scene):
    """The reference logs what it registers and warns about what it skips (ref:src/mujoco_lidar.cpp
    :140-150,178-185); so does the drop-in.  New: replicas beyond the URDF's ray count are reported,
    and rangefinders of one lidar with different cutoffs are refused (MuJoCo applies sensor_cutoff
    per sensor, the cast takes one per lidar)."""
    h = plug.PluginHarness(scene, INFO)
    assert h.register() == 0, h.error()
    logs = h.logs()
    assert any(l.startswith("INFO: lidar 'lidar'") and "24 rays" in l for l in logs)
    assert any(l.startswith("INFO: camera 'camera'") for l in logs)
    h.close()
    # the URDF describes fewer rays than the MJCF replicates: the surplus is dropped with a warning
    short = INFO.replace("max_angle=0.3", "max_angle=0.2")          # int(0.5 / 0.025) + 1 = 21 rays
    h = plug.PluginHarness(scene, short)
    assert h.register() == 0, h.error()
    assert h.lidar_info()["num_rangefinders"] == 21
    assert sum("lies beyond the 21 rays" in l for l in h.logs()) == 3
    h.close()
    # differing cutoffs inside one lidar
    import copy
    sc2 = copy.deepcopy(scene)
    sc2.rangefinders[5].cutoff = 3.0
    h = plug.PluginHarness(sc2, INFO)
    assert h.register() == -1 and "different cutoffs" in h.error()
    assert any(l.startswith("ERROR:") and "different cutoffs" in l for l in h.logs())
    h.close()


def test_no_cpu_fallback_in_plugin(plug, scene, has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    h = plug.PluginHarness(scene, INFO)
    assert h.register() == 0
    assert h.activate() == -1 and "no CPU fallback" in h.error()
    h.close()


@pytest.mark.gpu
def test_lidar_publish_cycle_matches_oracle(plug, scene, oracle):
    from mujoco_ros2_simulation_b200 import scenes
    for q in ([0.0, 0.0], [0.4, -0.9]):
        h = plug.PluginHarness(scene, INFO, qpos=q)
        assert h.register() == 0 and h.activate() == 0, h.error()
        assert h.update() == 0, h.error()
        ranges, npub, nint = h.lidar_scan()
        assert npub == 1 and nint == 0 and ranges.shape == (24,)
        li = h.lidar_info()
        spec = scenes.LidarSpec(li["local_dir"], bodyexclude=li["bodyid"], range_min=0.05, range_max=10.0)
        bpos, bmat, _ = scene.body_poses(q)
        o = oracle.OracleScene(scene.model, h.poses)
        r, g, raw = o.lidar_scan(spec, bpos[li["bodyid"]], bmat[li["bodyid"]].reshape(9))
        assert np.array_equal(ranges, r)
        if q == [0.0, 0.0]:  # 1.95 / cos(theta) on the arm's side faces
            theta = -0.3 + 0.025 * np.arange(24)
            np.testing.assert_allclose(ranges, 1.95 / np.cos(theta), rtol=3e-6)
        assert h.update() == 0 and h.lidar_scan()[1] == 2
        h.close()


@pytest.mark.gpu
def test_camera_publish_cycle(plug, scene, oracle):
    from mujoco_ros2_simulation_b200 import scenes
    h = plug.PluginHarness(scene, INFO)
    assert h.register() == 0 and h.activate() == 0, h.error()
    # (a) GL readback path: step 3 of MujocoCameras::update on the GPU, bit-exact vs the oracle
    gl_in = scenes.gl_depth(1, 64, 48)[0]
    h.set_gl_depth(gl_in)
    assert h.update() == 0, h.error()
    gl, depth, cloud, npub = h.camera_frames()
    assert npub == 1 and np.array_equal(gl, gl_in)
    ref_d = oracle.depth_linearise_flip(gl_in[None], scene.znear, scene.zfar, 1.0)[0]
    assert np.array_equal(depth, ref_d)
    ref_c = oracle.depth_to_cloud(ref_d[None], h.camera_info()["K"], 3)[0]
    assert np.array_equal(cloud, ref_c, equal_nan=True)
    pub, raw = h.camera_rgb()
    assert np.array_equal(pub, raw[::-1])                     # RGB rows flipped (ref :243-249)
    h.close()
    # (b) no OpenGL: the depth buffer is ray-cast from the scene mirror
    h = plug.PluginHarness(scene, INFO)
    assert h.register() == 0 and h.activate() == 0, h.error()
    assert h.update() == 0, h.error()
    gl, depth, cloud, _ = h.camera_frames()
    assert 0.0 <= gl.min() and gl.max() <= 1.0 and (gl < 1.0).mean() > 0.3   # arm + floor in view
    assert np.array_equal(depth, oracle.depth_linearise_flip(gl[None], scene.znear, scene.zfar, 1.0)[0])
    assert np.isfinite(cloud[..., 2]).all() and np.array_equal(cloud[..., 2], depth)
    # ray casting renders depth only: no colour image goes out, everything else does
    assert h.camera_pub_counts() == {"image": 0, "depth": 1, "info": 1, "points": 1}
    h.close()


@pytest.mark.gpu
def test_threads_publish_at_rate(plug, scene):
    h = plug.PluginHarness(scene, INFO, lidar_rate=50.0, camera_rate=50.0)
    assert h.register() == 0 and h.activate() == 0, h.error()
    h.run_threads(400)   # init() -> threads at the publish rate -> close() joins
    assert h.lidar_scan()[1] >= 3 and h.camera_frames()[3] >= 3
    h.close()
