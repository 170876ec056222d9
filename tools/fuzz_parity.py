"""Soak test on a B200: random scenes x random sensors, the CUDA path (two-kernel cast and fused kernel)
against the CPU oracle, every ray compared bit for bit.  Not part of the pytest tiers (minutes);
its log is kept under profiles/.

    python tools/fuzz_parity.py [iterations] [seed]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_ros2_simulation_b200 import engine, scenes  # noqa: E402
from mujoco_ros2_simulation_b200.model import (GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID,  # noqa: E402
                                               GEOM_PLANE, GEOM_SPHERE, Model, Poses, random_rotations)
from oracle import binding as orc  # noqa: E402


def random_scene(rng):
    """primitives of every type (some coincident, some huge, some tiny), optional ground plane"""
    n = int(rng.integers(3, 400))
    types = rng.choice([GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX], size=n)
    size = 10 ** rng.uniform(-1.5, 0.3, size=(n, 3))
    spread = 10 ** rng.uniform(0, 1.3)
    pos = rng.normal(size=(n, 3)) * spread
    mat = random_rotations(rng, n).reshape(n, 9)
    dup = rng.random(n) < 0.05                      # coincident duplicates: the lowest id must win
    src = rng.integers(0, n, size=n)
    pos[dup], mat[dup], size[dup], types[dup] = pos[src[dup]], mat[src[dup]], size[src[dup]], types[src[dup]]
    if rng.random() < 0.7:
        types = np.concatenate([[GEOM_PLANE], types])
        size = np.concatenate([[[0.0, 0.0, 1.0]], size])
        pos = np.concatenate([[[0.0, 0.0, -spread]], pos])
        mat = np.concatenate([np.eye(3).reshape(1, 9), mat])
        n += 1
    group = rng.integers(0, 6, size=n)
    rgba = np.ones((n, 4))
    rgba[rng.random(n) < 0.05, 3] = 0.0             # invisible geoms are skipped
    m = Model(types, np.zeros(n, int), group, -np.ones(n, int), -np.ones(n, int), size, rgba,
              np.zeros((0, 4)), [0])
    return m, Poses(pos, mat), spread


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    orc.build()
    total = bad = 0
    t0 = time.time()
    clutter = scenes.clutter500(n_el=32, n_az=512)   # the mesh scene, every third iteration
    for it in range(iters):
        if it % 3 == 2:
            m, P, spread = clutter.model, clutter.poses, 20.0
            origin = clutter.sensor_pos + rng.normal(size=3) * 0.5
            excl = 63
        else:
            m, P, spread = random_scene(rng)
            origin = rng.normal(size=3) * spread * 0.5
            excl = -1
        shape = rng.choice(["grid", "random", "fan"])
        if shape == "grid":
            n_el, n_az = int(rng.choice([16, 32, 64])), int(rng.choice([256, 1024, 2048]))
            dirs, tile_w = scenes.spinning_lidar_dirs(n_el, n_az), n_az
        elif shape == "fan":
            k = int(rng.choice([360, 1440, 4000]))
            dirs, tile_w = scenes.planar_lidar_dirs(-np.pi, 2 * np.pi / k, k), 0
        else:
            dirs, tile_w = rng.normal(size=(int(rng.integers(100, 5000)), 3)), 0
        gg = None if rng.random() < 0.5 else rng.integers(0, 2, 6).astype(np.uint8)
        spec = scenes.LidarSpec(dirs, tile_w=tile_w, bodyexclude=excl, geomgroup=gg,
                                cutoff=float(rng.choice([0.0, 0.0, spread])), range_min=0.05,
                                range_max=float(rng.choice([1000.0, spread * 2])))
        nscan = max(1, int(rng.choice([1, 3, (1 << 20) // dirs.shape[0] + 1])))
        nscan = min(nscan, 600)
        pos = origin[None] + rng.normal(size=(nscan, 3)) * 0.3
        mat = random_rotations(rng, nscan).reshape(nscan, 9) if rng.random() < 0.5 else np.tile(np.eye(3).reshape(1, 9), (nscan, 1))
        s = engine.Scene(m)
        s.update_poses(P)
        lid = engine.Lidar(s, spec)
        out = lid.cast(pos, mat, want=("ranges", "geomid", "dist"))
        o = orc.OracleScene(m, P)
        for sc in sorted(set([0, nscan // 2, nscan - 1])):
            r, g, raw = o.lidar_scan(spec, pos[sc], mat[sc], nthread=os.cpu_count() or 1)
            mis = int((g != out["geomid"][sc]).sum() + (raw != out["dist"][sc]).sum() + (r != out["ranges"][sc]).sum())
            total += r.size
            bad += mis
            if mis:
                print(f"MISMATCH iter {it} scan {sc}: {mis} values, scene ngeom {m.ngeom}, shape {shape}", flush=True)
        lid.close()
        s.close()
        print(f"iter {it:3d}: ngeom {m.ngeom:4d} {shape:6s} rays/scan {dirs.shape[0]:6d} scans {nscan:4d}  "
              f"checked {total} rays, mismatches {bad}", flush=True)
    print(f"fuzz done: {iters} iterations, {total} rays compared bit for bit, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
