#!/usr/bin/env python
"""One single-scan host cast (64x2048 on clutter500) between cudaProfilerStart / Stop, for an ncu
launch list of the plugin's duty cycle:
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/lat.csv \\
      --metrics gpu__time_duration.sum,smsp__inst_executed.sum python tools/latency_probe.py
Without ncu it prints the host / device time per call for a few launch-shape settings."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch

    from mujoco_ros2_simulation_b200 import engine, scenes
    robot = "--robot" in sys.argv     # the reference's fixture: 6 boxes, 360-ray planar lidar (cfg1)
    cs = scenes.robot_scene(n_rays=360) if robot else scenes.clutter500()
    sc = engine.Scene(cs.model, nworld=1, device=0)
    sc.update_poses(cs.poses)
    lid = engine.Lidar(sc, cs.lidar)
    nscan = 1
    for a in sys.argv[1:]:
        if a.startswith("--nscan="):
            nscan = int(a.split("=")[1])       # several scans per call: where the small-launch paths hand over
    if robot:
        pos = np.repeat(cs.lidar_pos[None], 2, 0)
        mat = np.repeat(cs.lidar_mat.reshape(1, 9), 2, 0)
    else:
        pos, mat = cs.sensor_poses(nscan + 1)
    bufs = {"ranges": np.empty((nscan, lid.nray), np.float32)}
    profiling = "--profile" in sys.argv
    for _ in range(5):
        lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
    if profiling:
        os.environ["B2_GRAPH"] = "0"  # plain launches: ncu lists the kernels of the sequence
        lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
        torch.cuda.profiler.start()
        lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
        torch.cuda.profiler.stop()
        return
    envs = ({}, {"B2_GRAPH": "0"}, {"B2_MESH_COOP_RAYS": "0"}, {"B2_MESH_COOP_G": "8"}, {"B2_MESH_COOP_G": "16"},
            {"B2_MESH_COOP_G": "32"}, {"B2_SPLIT": "0"}, {"B2_SMALL_SPLIT": "0"}, {"B2_TWO_MIN_RAYS": "1"})
    if "--group" in sys.argv:
        envs = ({}, {"B2_SMALL_SPLIT": "0", "B2_GROUP_TILES_RT": "2"}, {"B2_SMALL_SPLIT": "0", "B2_GROUP_TILES_RT": "4"})
    if "--coop" in sys.argv:
        envs = ({"B2_MESH_COOP_RAYS": "0"}, {"B2_MESH_COOP_G": "8"}, {"B2_MESH_COOP_G": "16"}, {"B2_MESH_COOP_G": "32"}, {})
    for env in envs:
        for k in ("B2_GRAPH", "B2_SPLIT", "B2_SMALL_SPLIT", "B2_TWO_MIN_RAYS", "B2_MESH_COOP_RAYS", "B2_MESH_COOP_G",
                  "B2_GROUP_TILES_RT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        lid2 = engine.Lidar(sc, cs.lidar)  # a fresh lidar: its graph is captured under this setting
        for _ in range(5):
            lid2.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
        wall, dev = [], []
        for _ in range(200):
            t0 = time.perf_counter()
            lid2.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
            wall.append((time.perf_counter() - t0) * 1e6)
            dev.append(lid2.last_device_us())
        print(f"{str(env):50s} host p50 {np.median(wall):7.1f} us  device p50 {np.median(dev):7.1f} us")
        lid2.close()


if __name__ == "__main__":
    main()
