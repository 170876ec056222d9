"""Throughput of coarse lidar tables vs tiles per collect group (B2_GROUP_TILES_RT) on B200."""
import os
import sys
import numpy as np
import torch
from mujoco_ros2_simulation_b200 import engine, scenes

dev = torch.device("cuda", 0)
for n_az in (128, 256, 512, 2048):
    cs = scenes.clutter500(n_el=64, n_az=n_az)
    s = engine.Scene(cs.model); s.update_poses(cs.poses)
    nscan = (1 << 23) // (64 * n_az)
    pos, mat = cs.sensor_poses(nscan)
    dp, dm = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
    row = [f"64x{n_az:<5d} {nscan * 64 * n_az} rays"]
    ref = None
    for gt in (0, 4, 2, 1):
        os.environ["B2_GROUP_TILES_RT"] = str(gt)
        lid = engine.Lidar(s, cs.lidar)
        r = torch.empty((nscan, lid.nray), dtype=torch.float32, device=dev)
        g = torch.empty((nscan, lid.nray), dtype=torch.int32, device=dev)
        for _ in range(3):
            lid.cast_device(nscan, dp, dm, None, ranges=r, geomid=g)
        s.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            lid.cast_device(nscan, dp, dm, None, ranges=r, geomid=g)
        s.sync(); b.record(); torch.cuda.synchronize()
        if ref is None:
            ref = (r.clone(), g.clone())
        else:
            assert torch.equal(ref[0], r) and torch.equal(ref[1], g)
        row.append(f"gt={gt if gt else 'auto'}: {nscan * lid.nray / (a.elapsed_time(b) / 5 * 1e-3) / 1e9:6.2f} Grays/s")
    print("  ".join(row), flush=True)
