"""GPU tier: shard.PeerGather between two PROCESSES (CUDA IPC).  The driver's GPU box has one GPU,
so both ranks use cuda:0 -- the mapping, the pointer arithmetic, both ways of filling the root's
buffers (the kernels' own stores; copy-engine block push) and the world sharding are exactly what
the 8-GPU job runs, only the bytes do not leave the device.  gloo carries the handles."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mujoco_ros2_simulation_b200 import engine, scenes, shard
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    W = 6                                                   # worlds of the job, 3 per rank
    cs = scenes.clutter500(n_el=16, n_az=128)
    lo, hi = shard.world_block(W, world, rank)
    n = hi - lo
    P = scenes.clutter_world_block(cs, lo, n)
    sc = engine.Scene(cs.model, nworld=n)
    sc.update_poses(P)
    lid = engine.Lidar(sc, cs.lidar)
    pos, mat = cs.sensor_poses(W, seed=31)
    d_pos, d_mat = torch.from_numpy(pos[lo:hi]).to(dev), torch.from_numpy(mat[lo:hi]).to(dev)
    sw = torch.arange(n, dtype=torch.int32, device=dev)
    nray = lid.nray
    res = {}
    for mode in ("store", "push"):
        pg = shard.PeerGather({"ranges": (nray, "float32"), "geomid": (nray, "int32")}, n, rank, world, 0)
        if mode == "store":
            lid.cast_device(n, d_pos, d_mat, sw, ranges=pg.ptr("ranges"), geomid=pg.ptr("geomid"))
        else:
            r = torch.empty((n, nray), dtype=torch.float32, device=dev)
            g = torch.empty((n, nray), dtype=torch.int32, device=dev)
            for a, b in shard.scan_blocks(n, 2):
                lid.cast_device(b - a, d_pos[a:b], d_mat[a:b], sw[a:b], ranges=r[a:b], geomid=g[a:b])
                pg.push(sc, "ranges", a, r[a:b])
                pg.push(sc, "geomid", a, g[a:b])
            sc.join()
        sc.sync()
        dist.barrier()                                       # every rank's data has landed
        if rank == 0:
            res[mode] = (pg.root_tensor("ranges").cpu().numpy().copy(), pg.root_tensor("geomid").cpu().numpy().copy())
        dist.barrier()
        pg.close()
    if rank == 0:
        out["res"] = res
    lid.close()
    sc.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_processes_gather_into_the_roots_peer_buffer(oracle):
    from mujoco_ros2_simulation_b200 import scenes, shard
    world, W = 2, 6
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = out["res"]
    cs = scenes.clutter500(n_el=16, n_az=128)
    P = scenes.clutter_world_block(cs, 0, W)
    pos, mat = cs.sensor_poses(W, seed=31)
    for mode in ("store", "push"):
        R, G = res[mode]
        assert R.shape == (world, W // world, cs.lidar.nray)
        for w in range(W):
            owner = shard.owner_of(w, W, world)
            lo, _ = shard.world_block(W, world, owner)
            o = oracle.OracleScene(cs.model, P, world=w)
            r, g, _ = o.lidar_scan(cs.lidar, pos[w], mat[w], nthread=4)
            assert np.array_equal(R[owner, w - lo], r) and np.array_equal(G[owner, w - lo], g), (mode, w)
