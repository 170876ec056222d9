"""N3: the scene conventions the reference's URDF->MJCF converter produces
(ref:scripts/make_mjcf_from_robot_description.py), authored as MJCF in the
converter's own style and pushed through mjcf.py -> oracle (CPU tier) and mjcf.py -> GPU -> oracle
(GPU tier):

  * every visual geom appears twice, coincident: class "collision" (group 3) FIRST, then class
    "visual" (group 2) (ref :635-652 appends collision, then visual) -- a rangefinder (geomgroup NULL)
    sees both and must report the LOWER id, i.e. the collision geom; a camera (mjv_defaultOption)
    sees the visual one only;
  * a lidar is a massless body at the reference site with quat  site (x) [.5,.5,.5,.5] (x) RotY(min)
    (ref :1120-1140) holding <replicate count euler="0 inc 0"> of a site with quat "0 0 0 1"
    (ref :736-750): ray i must point along (cos th_i, sin th_i, 0) of the URDF lidar frame;
  * a camera sits at its optical-frame site rotated by pi about X (ref :1080-1092): the MuJoCo camera
    looks along the optical +Z, its +X is the optical +X and its +Y the optical -Y, so the organised
    cloud comes out in the optical frame (x right, y down, z forward).
"""
import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import mjcf, scenes
from mujoco_ros2_simulation_b200.model import axis_angle_quat, quat2mat, quat_mul

MIN_ANGLE, INC, COUNT = -0.6, 0.05, 25
SITE_QUAT = axis_angle_quat((0.3, -0.2, 0.9), 0.7)        # an arbitrary URDF lidar frame
SITE_POS = np.array([0.4, -0.3, 0.6])
OPT_QUAT = quat_mul(axis_angle_quat((0, 0, 1), 0.4), axis_angle_quat((1, 0, 0), -1.9))  # optical frame
OPT_POS = np.array([-0.2, 0.1, 0.9])


def _q(v):
    return " ".join(repr(float(x)) for x in v)


def _scene_xml():
    lidar_quat = scenes.lidar_body_quat(SITE_QUAT, MIN_ANGLE)
    cam_quat = quat_mul(OPT_QUAT, np.array([0.0, 1.0, 0.0, 0.0]))     # ref :1080 x_rotation
    # targets placed in the URDF lidar frame so that different rays meet different objects
    Rs = quat2mat(SITE_QUAT)
    box_pos = SITE_POS + Rs @ np.array([2.0, 0.0, 0.0])
    sph_pos = SITE_POS + Rs @ (1.5 * np.array([np.cos(0.45), np.sin(0.45), 0.0]))
    # and one box on the camera's optical axis, 2 m ahead, shifted 0.3 m to the optical +X (right)
    Ro = quat2mat(OPT_QUAT)
    cam_box = OPT_POS + Ro @ np.array([0.3, 0.0, 2.0])
    return f"""
<mujoco model="converter_style">
  <compiler angle="radian"/>
  <default>
    <default class="visual"><geom group="2" contype="0" conaffinity="0"/></default>
    <default class="collision"><geom group="3"/></default>
  </default>
  <worldbody>
    <body name="link" pos="0 0 0">
      <geom type="box" size="0.2 0.6 0.4" pos="{_q(box_pos)}" quat="{_q(SITE_QUAT)}" class="collision"/>
      <geom type="box" size="0.2 0.6 0.4" pos="{_q(box_pos)}" quat="{_q(SITE_QUAT)}" class="visual" rgba="0.8 0.2 0.2 1"/>
      <geom type="sphere" size="0.25" pos="{_q(sph_pos)}" class="collision"/>
      <geom type="sphere" size="0.25" pos="{_q(sph_pos)}" class="visual" rgba="0.2 0.8 0.2 1"/>
      <geom type="box" size="0.15 0.15 0.05" pos="{_q(cam_box)}" quat="{_q(OPT_QUAT)}" class="collision"/>
      <geom type="box" size="0.15 0.15 0.05" pos="{_q(cam_box)}" quat="{_q(OPT_QUAT)}" class="visual" rgba="0.2 0.2 0.8 1"/>
      <site name="lidar_site" pos="{_q(SITE_POS)}" quat="{_q(SITE_QUAT)}"/>
      <site name="camera_optical" pos="{_q(OPT_POS)}" quat="{_q(OPT_QUAT)}"/>
      <camera name="camera" fovy="58" resolution="64 48" pos="{_q(OPT_POS)}" quat="{_q(cam_quat)}"/>
      <body name="lidar_site_lidar_body" pos="{_q(SITE_POS)}" quat="{_q(lidar_quat)}">
        <replicate count="{COUNT}" sep="-" offset="0 0 0" euler="0 {INC} 0">
          <site name="rf" pos="0 0 0" quat="0 0 0 1"/>
        </replicate>
      </body>
    </body>
  </worldbody>
  <sensor>
    <rangefinder name="scan" site="rf"/>
  </sensor>
</mujoco>
"""


@pytest.fixture(scope="module")
def conv():
    sc = mjcf.load_mjcf_string(_scene_xml())
    return sc


def _lidar_rays(sc):
    """origin and world directions of the replicated rangefinders, in sensor order"""
    sx, sm = sc.site_poses(None)
    ids = [r.site for r in sc.rangefinders]
    assert [r.name for r in sc.rangefinders] == [f"scan-{i:02d}" for i in range(COUNT)]   # <replicate sep="-">
    return sx[ids[0]], np.stack([sm[i][:, 2] for i in ids]), sc.sites[ids[0]].body


def test_lidar_body_frame_gives_laserscan_angles(conv):
    origin, dirs, _ = _lidar_rays(conv)
    np.testing.assert_allclose(origin, SITE_POS, atol=1e-12)
    theta = MIN_ANGLE + INC * np.arange(COUNT)
    want = (quat2mat(SITE_QUAT) @ np.stack([np.cos(theta), np.sin(theta), np.zeros(COUNT)])).T
    np.testing.assert_allclose(dirs, want, atol=1e-12)     # ray i along (cos th_i, sin th_i, 0), CCW, Z up


def test_duplicates_resolve_to_the_collision_geom(conv, oracle):
    m = conv.model
    assert m.geom_group.tolist() == [3, 2, 3, 2, 3, 2]       # collision first, then visual (ref :651-652)
    origin, dirs, body = _lidar_rays(conv)
    o = oracle.OracleScene(m, conv.poses(None))
    d, g = o.multi_ray(origin, dirs, bodyexclude=body)       # the rangefinder's call: geomgroup NULL
    assert set(g.tolist()) <= {-1, 0, 2} and {0, 2} <= set(g.tolist())    # never the visual twin
    theta = MIN_ANGLE + INC * np.arange(COUNT)
    face = np.abs(theta) < 0.25                               # rays meeting the box's near face x = 1.8
    np.testing.assert_allclose(d[face], 1.8 / np.cos(theta[face]), rtol=1e-12)
    # the camera's view (groups 0-2): same distances, the visual ids
    dv, gv = o.multi_ray(origin, dirs, geomgroup=[1, 1, 1, 0, 0, 0], bodyexclude=body)
    assert np.array_equal(dv, d) and np.array_equal(gv[g >= 0], g[g >= 0] + 1)


def test_camera_is_the_optical_frame_turned_about_x(conv):
    bpos, bmat, _ = conv.body_poses(None)
    c = conv.cameras[0]
    cmat = bmat[c.body] @ quat2mat(c.frame.quat)
    Ro = quat2mat(OPT_QUAT)
    np.testing.assert_allclose(cmat[:, 0], Ro[:, 0], atol=1e-12)      # +X right in both
    np.testing.assert_allclose(cmat[:, 1], -Ro[:, 1], atol=1e-12)     # MuJoCo +Y up = optical -Y
    np.testing.assert_allclose(-cmat[:, 2], Ro[:, 2], atol=1e-12)     # looks along -Z = optical +Z


@pytest.mark.gpu
def test_converter_scene_on_the_gpu(conv, oracle):
    from mujoco_ros2_simulation_b200 import engine
    m, poses = conv.model, conv.poses(None)
    s = engine.Scene(m, nworld=1)
    s.update_poses(poses)
    # lidar exactly as the plugin registers it: rays in the lidar body's frame, bodyexclude = that body
    bpos, bmat, _ = conv.body_poses(None)
    origin, dirs, body = _lidar_rays(conv)
    local = dirs @ bmat[body]                                  # R^T d, row form
    spec = scenes.LidarSpec(local, bodyexclude=body, range_min=0.05, range_max=30.0)
    lid = engine.Lidar(s, spec)
    out = lid.cast(bpos[body][None], bmat[body].reshape(1, 9), want=("ranges", "geomid", "dist"))
    o = oracle.OracleScene(m, poses)
    r, g, raw = o.lidar_scan(spec, bpos[body], bmat[body].reshape(9))
    assert np.array_equal(out["ranges"][0], r) and np.array_equal(out["geomid"][0], g)
    assert np.array_equal(out["dist"][0], raw)
    assert set(g.tolist()) <= {-1, 0, 2} and {0, 2} <= set(g.tolist())
    # camera: ray-cast depth -> organised cloud in the OPTICAL frame
    c = conv.cameras[0]
    cpos = bpos[c.body] + bmat[c.body] @ c.frame.pos
    cmat = bmat[c.body] @ quat2mat(c.frame.quat)
    cam = engine.DepthCamera(scenes.CameraSpec(64, 48, 58.0, conv.znear, conv.zfar, 1.0, 12), scene=s)
    gl, depth, cloud = cam.render_to_cloud(cpos[None], cmat.reshape(1, 9))
    hit = np.isfinite(cloud[0][..., 2]) & (depth[0] < 10)
    pts = cloud[0][hit]
    # the blue box: 2 m ahead (its near face at z = 1.95), centred 0.3 m to the RIGHT, on the axis vertically
    near = pts[np.abs(pts[:, 2] - 1.95) < 0.01]
    assert near.shape[0] > 20
    assert 0.12 < near[:, 0].min() and near[:, 0].max() < 0.48        # x = 0.3 +- 0.15 (pixel quantisation)
    assert abs(near[:, 1].mean()) < 0.05
    vs, us = np.nonzero(hit & (np.abs(cloud[0][..., 2] - 1.95) < 0.01))
    assert us.min() >= 32                                               # right of the principal point
    # every cloud point lies on a VISUAL geom at the oracle's depth along the optical axis (the ray
    # direction has z = -1 in the camera frame, so the oracle's distance IS that depth); x / y follow
    # depth_image_proc's integer pixel coordinates, half a pixel off the ray through the pixel centre
    K = cam.info()[0]
    vv, uu = np.nonzero(hit)
    l = np.stack([(uu + 0.5 - K[2]) / K[0], -(vv + 0.5 - K[5]) / K[4], -np.ones(len(uu))], -1)
    d_or, g_or = o.multi_ray(cpos, l @ cmat.T, geomgroup=[1, 1, 1, 0, 0, 0])
    assert set(g_or.tolist()) <= {1, 3, 5}
    np.testing.assert_allclose(pts[:, 2], d_or, rtol=3e-4)
    np.testing.assert_allclose(pts[:, 0], (uu - K[2]) * pts[:, 2] / K[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pts[:, 1], (vv - K[5]) * pts[:, 2] / K[4], rtol=1e-5, atol=1e-6)
    world = OPT_POS + pts @ quat2mat(OPT_QUAT).T       # optical frame -> world: on the box's near face
    on_box = np.abs(pts[:, 2] - 1.95) < 0.01
    box_c = OPT_POS + quat2mat(OPT_QUAT) @ np.array([0.3, 0.0, 2.0])
    assert np.all(np.linalg.norm(world[on_box] - box_c, axis=1) < 0.25)
    s.close()
