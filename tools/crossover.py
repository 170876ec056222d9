"""Latency of one device-side cast vs launch size, fused kernel vs two-kernel cast (B200)."""
import os
import numpy as np
import torch
from mujoco_ros2_simulation_b200 import engine, scenes

import sys
N_AZ = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cs = scenes.clutter500(n_el=64, n_az=N_AZ)          # 64 x N_AZ rays per scan
s = engine.Scene(cs.model); s.update_poses(cs.poses)
dev = torch.device("cuda", 0)
R = 64 * N_AZ
for nray_scan, nscan in ((R, 1), (R, 2), (R, 4), (R, 8)):
    spec = scenes.LidarSpec(cs.lidar.local_dir[:nray_scan], tile_w=N_AZ, bodyexclude=63)
    lid = engine.Lidar(s, spec)
    pos, mat = cs.sensor_poses(nscan)
    dp, dm = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
    r = torch.empty((nscan, nray_scan), dtype=torch.float32, device=dev)
    g = torch.empty((nscan, nray_scan), dtype=torch.int32, device=dev)
    row = [f"{nscan * nray_scan:8d} rays"]
    for two, split, small in ((0, 1, 0), (0, 1, 1), (1, 1, 1)):
        os.environ["B2_TWO_KERNEL"] = str(two); os.environ["B2_TWO_MIN_RAYS"] = "1"; os.environ["B2_SPLIT"] = str(split); os.environ["B2_SMALL_SPLIT"] = str(small)
        for _ in range(5):
            lid.cast_device(nscan, dp, dm, None, ranges=r, geomid=g)
        s.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream()
        torch.cuda.synchronize(); a.record()
        for _ in range(20):
            lid.cast_device(nscan, dp, dm, None, ranges=r, geomid=g)
        s.sync(); b.record(); torch.cuda.synchronize()
        row.append(f"two={two} small_split={small}: {a.elapsed_time(b) / 20 * 1e3:8.1f} us")
    print("  ".join(row))
