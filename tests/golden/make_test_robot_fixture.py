#!/usr/bin/env python
"""Derives tests/golden/test_robot_scene.json from the reference's own fixture scene.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_test_robot_fixture.py

Reads ref:test/test_resources/test_robot.xml (+ scene.xml for stat.extent, + test_robot.urdf for
the lidar <param>s) through the package's MJCF-subset reader, runs forward kinematics at qpos = 0
and stores the flat arrays the engine consumes, plus the known answer SURVEY.md section 8(c) O4
derives for this scene: range[i] = 1.95 / cos(theta_i), theta_i = -0.3 + 0.025 i, geomid 0 for
i <= 12 and 2 for i >= 13.
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from mujoco_ros2_simulation_b200.mjcf import load_mjcf  # noqa: E402

REF = "/root/reference/test/test_resources"


def main():
    sc = load_mjcf(os.path.join(REF, "test_robot.xml"))
    shell = ET.parse(os.path.join(REF, "scene.xml")).getroot()
    extent = float(shell.find("statistic").get("extent"))
    urdf = ET.parse(os.path.join(REF, "test_robot.urdf")).getroot()
    lidar_params = {}
    for s in urdf.iter("sensor"):
        if s.get("name") == "lidar":
            lidar_params = {p.get("name"): p.text for p in s.findall("param")}
    poses = sc.poses()
    bpos, bmat, _ = sc.body_poses()
    lidar_body = sc.body_id("lidar_sensor_frame_lidar_body")
    # rangefinders ordered by the index parsed from "<name>-<idx>" (ref:src/mujoco_lidar.cpp:30-49)
    rfs = sorted(sc.rangefinders, key=lambda r: int(r.name.rsplit("-", 1)[1]))
    from mujoco_ros2_simulation_b200.model import quat2mat
    local_dir = [quat2mat(sc.sites[r.site].frame.quat)[:, 2].tolist() for r in rfs]
    site_frame = sc.sites[sc.site_id("lidar_sensor_frame")]
    cam = sc.cameras[0]
    m = sc.model
    inc, mn, mx = (float(lidar_params[k]) for k in ("angle_increment", "min_angle", "max_angle"))
    n = int((mx - mn) / inc) + 1
    theta = mn + inc * np.arange(n)
    out = {
        "source": "ref:test/test_resources/test_robot.xml at qpos=0 (+scene.xml, test_robot.urdf:203-211)",
        "model": {
            "geom_type": m.geom_type.tolist(), "geom_bodyid": m.geom_bodyid.tolist(),
            "geom_group": m.geom_group.tolist(), "geom_matid": m.geom_matid.tolist(),
            "geom_dataid": m.geom_dataid.tolist(), "geom_size": m.geom_size.tolist(),
            "geom_rgba": m.geom_rgba.tolist(), "body_weldid": m.body_weldid.tolist(),
        },
        "geom_xpos": poses.geom_xpos[0].tolist(),
        "geom_xmat": poses.geom_xmat[0].tolist(),
        "lidar": {
            "bodyid": lidar_body,
            "body_xpos": bpos[lidar_body].tolist(),
            "body_xmat": bmat[lidar_body].reshape(9).tolist(),
            "local_dir": local_dir,
            "site_quat": site_frame.frame.quat.tolist(),
            "sensor_names": [r.name for r in rfs],
            "min_angle": mn, "max_angle": mx, "angle_increment": inc,
            "range_min": float(lidar_params["range_min"]),
            "range_max": float(lidar_params["range_max"]),
            "frame_name": lidar_params["frame_name"],
            "laserscan_topic": lidar_params["laserscan_topic"],
        },
        "camera": {"name": cam.name, "width": cam.resolution[0], "height": cam.resolution[1],
                   "fovy": cam.fovy},
        "znear": sc.znear, "zfar": sc.zfar, "extent": extent,
        "tree": {
            "geom_body": m.geom_bodyid.tolist(),
            "geom_local_pos": [g.frame.pos.tolist() for g in sc.geoms],
            "forearm_pos": sc.bodies[sc.body_id("forearm")].frame.pos.tolist(),
            "plane_z": -0.5,
        },
        "known_answer": {
            "note": "SURVEY.md 8(c) O4: hand-derived, not MuJoCo output; tolerance 2e-6 relative",
            "theta": theta.tolist(),
            "range": (1.95 / np.cos(theta)).tolist(),
            "geomid": [0 if i <= 12 else 2 for i in range(n)],
            "num_rangefinders": n,
        },
    }
    with open(os.path.join(HERE, "test_robot_scene.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote test_robot_scene.json:", n, "rays,", m.ngeom, "geoms")


if __name__ == "__main__":
    main()
