"""GPU tier: scan -> world mapping.  Without a scan_world table scan s is cast in world
s % nworld -- also when the blocking host call cuts the batch into pipelined chunks and when the
device path cuts it into scratch-sized chunks."""
import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import engine, scenes

pytestmark = pytest.mark.gpu


def test_implicit_world_mapping_across_chunks(oracle):
    import torch
    W, S = 5, 37                       # 37 scans over 5 worlds: chunk boundaries fall mid-cycle
    ws = scenes.arm_worlds(nworld=W, n_rays=256)
    s = engine.Scene(ws.model, nworld=W)
    s.update_poses(ws.poses)
    lid = engine.Lidar(s, ws.lidar)
    rng = np.random.default_rng(4)
    pos = np.tile(ws.sensor_pos[0], (S, 1)) + rng.uniform(-0.2, 0.2, (S, 3))
    mat = np.tile(ws.sensor_mat[0], (S, 1))
    implicit = lid.cast(pos, mat, want=("geomid", "dist"))                       # host path, chunked
    explicit = lid.cast(pos, mat, scan_world=(np.arange(S) % W).astype(np.int32), want=("geomid", "dist"))
    assert np.array_equal(implicit["dist"], explicit["dist"])
    assert np.array_equal(implicit["geomid"], explicit["geomid"])
    dev = torch.device("cuda", 0)
    g = torch.empty((S, 256), dtype=torch.int32, device=dev)
    d = torch.empty((S, 256), dtype=torch.float64, device=dev)
    lid.cast_device(S, torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev), None, geomid=g, dist=d)
    s.sync()
    assert np.array_equal(d.cpu().numpy(), explicit["dist"])
    for sc in (0, 6, 19, 36):
        o = oracle.OracleScene(ws.model, ws.poses, world=sc % W)
        r, gg, raw = o.lidar_scan(ws.lidar, pos[sc], mat[sc])
        assert np.array_equal(raw, explicit["dist"][sc]) and np.array_equal(gg, explicit["geomid"][sc])
    # the worlds really differ
    assert not np.array_equal(explicit["dist"][0], explicit["dist"][1])
