#!/usr/bin/env python
"""Writes tests/golden/clutter_golden.npz: oracle outputs for fixed, seeded inputs on the clutter500
scene (reduced lidar 32x256) -- a frozen copy of what the oracle said when it was pinned, so a later
change to either the oracle or the CUDA path that shifts a single bit shows up in both tiers.

    python tests/golden/make_clutter_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from mujoco_ros2_simulation_b200 import scenes  # noqa: E402
from oracle import binding as orc  # noqa: E402


def main():
    cs = scenes.clutter500(n_el=32, n_az=256)
    o = orc.OracleScene(cs.model, cs.poses)
    pos, mat = cs.sensor_poses(3)
    out = {"pos": pos, "mat": mat}
    for i in range(3):
        r, g, raw = o.lidar_scan(cs.lidar, pos[i], mat[i], nthread=8)
        out[f"ranges{i}"], out[f"geomid{i}"], out[f"dist{i}"] = r, g, raw
    d, g = o.multi_ray(cs.sensor_pos, cs.lidar.local_dir, geomgroup=[1, 0, 1, 1, 0, 0], flg_static=0,
                       bodyexclude=63, nthread=8)
    out["masked_dist"], out["masked_geomid"] = d, g
    P = scenes.clutter_worlds(cs, 3)
    o2 = orc.OracleScene(cs.model, P, world=2)
    r, g, raw = o2.lidar_scan(cs.lidar, pos[1], mat[1], nthread=8)
    out["world2_dist"], out["world2_geomid"] = raw, g
    gl = scenes.gl_depth(1, 64, 48)
    z = orc.depth_linearise_flip(gl, 0.01, 50.0, 1.0)
    out["depth"] = z
    out["cloud"] = orc.depth_to_cloud(z, orc.camera_intrinsics(58.0, 64, 48), 3)
    np.savez_compressed(os.path.join(HERE, "clutter_golden.npz"), **out)
    print("wrote clutter_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
