#!/usr/bin/env python
"""One bench step (pose refresh + refit + one cast of --scans 64x2048 scans on clutter500) between
cudaProfilerStart / Stop, for ncu --profile-from-start off:

  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/traffic.csv \\
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,smsp__inst_executed.sum \\
      python tools/traffic_run.py --scans 256
  python tools/ncu_traffic.py gpurun_out/traffic.csv 256 > profiles/traffic.json
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=256)
    args = ap.parse_args()
    import torch

    from mujoco_ros2_simulation_b200 import engine, scenes
    dev = torch.device("cuda", 0)
    cs = scenes.clutter500()
    stream = torch.cuda.Stream(device=dev)
    scene = engine.Scene(cs.model, nworld=1, device=0, stream=stream.cuda_stream)
    scene.update_poses(cs.poses)
    lidar = engine.Lidar(scene, cs.lidar)
    n = args.scans
    pos, mat = cs.sensor_poses(n)
    with torch.cuda.stream(stream):
        d_pos, d_mat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
        d_world = torch.zeros(n, dtype=torch.int32, device=dev)
        d_xpos = torch.from_numpy(cs.poses.geom_xpos).to(dev)
        d_xmat = torch.from_numpy(cs.poses.geom_xmat).to(dev)
        ranges = torch.empty((n, lidar.nray), dtype=torch.float32, device=dev)
        geomid = torch.empty((n, lidar.nray), dtype=torch.int32, device=dev)

        def step():
            scene.update_poses_device(d_xpos, d_xmat)
            lidar.cast_device(n, d_pos, d_mat, d_world, ranges=ranges, geomid=geomid)

        for _ in range(3):
            step()
        stream.synchronize()
        torch.cuda.profiler.start()
        step()
        stream.synchronize()
        torch.cuda.profiler.stop()
    print("rays", n * lidar.nray)


if __name__ == "__main__":
    main()
