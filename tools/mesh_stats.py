#!/usr/bin/env python
"""Per-pair work of mesh_pairs_kernel for ONE 64x2048 scan on clutter500 (debug build of the library:
-DB2_MESH_STATS, loaded through B2_LIB_PATH): pairs, nodes / triangles / cycles per pair (mean, max,
log2 histograms).  What the single-scan latency of the mesh kernel is made of."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    from mujoco_ros2_simulation_b200 import _capi, engine, scenes
    lib = _capi.load()
    cs = scenes.clutter500()
    sc = engine.Scene(cs.model, nworld=1, device=0)
    sc.update_poses(cs.poses)
    lid = engine.Lidar(sc, cs.lidar)
    nscan = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    pos, mat = cs.sensor_poses(nscan + 1)
    bufs = {"ranges": np.empty((nscan, lid.nray), np.float32)}
    os.environ["B2_GRAPH"] = "0"
    lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
    out = (C.c_ulonglong * 64)()
    lib.b2_debug_mesh_stats(None, 1)
    lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
    lib.b2_debug_mesh_stats(out, 0)
    s = np.array(out[:], dtype=np.float64)
    n = max(s[0], 1)
    print(f"pairs {int(s[0])}  with a hit {int(s[7])}")
    print(f"nodes  / pair: mean {s[1]/n:8.1f}  max {int(s[4])}")
    print(f"tris   / pair: mean {s[2]/n:8.1f}  max {int(s[5])}")
    print(f"cycles / pair: mean {s[3]/n:8.0f}  max {int(s[6])}")
    print("log2 histogram of nodes :", [int(v) for v in s[16:32]])
    print("log2 histogram of tris  :", [int(v) for v in s[32:48]])
    print("log2 histogram of cyc/256:", [int(v) for v in s[48:64]])
    if hasattr(lib, "b2_debug_tile_stats"):
        t = (C.c_ulonglong * 32)()
        lib.b2_debug_tile_stats(None, 1)
        lid.cast(pos[1:], mat[1:], want=("ranges",), out=bufs)
        lib.b2_debug_tile_stats(t, 0)
        t = np.array(t[:], dtype=np.float64)
        n = max(t[0], 1)
        print(f"fused kernel: tile groups {int(t[0])}  cycles / group mean {t[1]/n:8.0f}  max {int(t[2])}")
        print("log2 histogram of cyc/256:", [int(v) for v in t[16:32]])


if __name__ == "__main__":
    main()
