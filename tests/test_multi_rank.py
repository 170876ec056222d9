"""CPU tier, world_size 2 over gloo: the N>1 host logic of bench.py -- the path shards by
independent sensor-pose batches / worlds with NO data-path collective; what the
ranks do exchange is the timing (max over ranks) and, optionally, gathered results."""
import os
import socket
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from mujoco_ros2_simulation_b200 import scenes

    # every rank owns a replica of the scene and its own, different batch of sensor poses
    cs = scenes.clutter500(n_el=4, n_az=32)
    pos, mat = cs.sensor_poses(4, seed=bench.rank_pose_seed(rank))

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_local = 10.0 * (rank + 1)  # rank 1 is the slow one
    value = bench.job_rays_per_s(world, 1000, 5, ms_local, reduce_max)
    # optional result gather (the only cross-rank traffic the path ever has)
    mine = torch.full((3,), float(rank))
    gathered = [torch.empty(3) for _ in range(world)]
    dist.all_gather(gathered, mine)
    out[rank] = (float(pos[1, 0]), value, [float(g[0]) for g in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    # disjoint pose batches
    assert out[0][0] != out[1][0]
    # whole-job throughput = all ranks' rays / the slowest rank's time, identical on every rank
    expect = world * 1000 * 5 / (20.0e-3)
    assert out[0][1] == expect and out[1][1] == expect
    assert out[0][2] == [0.0, 1.0] and out[1][2] == [0.0, 1.0]


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under torchrun: rank 0 prints the line, other ranks exit 0
    without work (and without needing a GPU)."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_world_block_partition():
    """shard.world_block / owner_of: contiguous blocks, sizes differ by <= 1, one owner per world."""
    from mujoco_ros2_simulation_b200 import shard
    for W, N in ((37, 8), (32768, 8), (5, 2), (3, 8), (0, 4), (4096, 1)):
        owned = np.zeros(W, int)
        sizes = []
        prev_hi = 0
        for rank in range(N):
            lo, hi = shard.world_block(W, N, rank)
            assert lo == prev_hi and hi >= lo
            prev_hi = hi
            owned[lo:hi] += 1
            sizes.append(hi - lo)
            for w in range(lo, min(hi, lo + 3)):
                assert shard.owner_of(w, W, N) == rank
            if hi > lo:
                assert shard.owner_of(hi - 1, W, N) == rank
        assert prev_hi == W and (owned == 1).all()
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == shard.padded_rows(W, N)
    assert shard.scan_blocks(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert shard.scan_blocks(2, 4) == [(0, 1), (1, 2)] and shard.scan_blocks(0, 4) == []


def _gather_worker(rank, world, port, out):
    """cfg5 in small: W worlds sharded with shard.world_block, every rank traces ITS worlds (the
    kernels' traversal compiled for the host: tests/emu), shard.BlockGather all-gathers the real
    (range, geomid) rows block by block over gloo."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mujoco_ros2_simulation_b200 import scenes, shard
    from tests.emu import emu_binding as emu

    W = 5
    cs = scenes.clutter500(n_el=4, n_az=32)
    lo, hi = shard.world_block(W, world, rank)
    rows = shard.padded_rows(W, world)
    P = scenes.clutter_world_block(cs, lo, hi - lo)  # this rank's worlds only
    pos, mat = cs.sensor_poses(W)                      # sensor pose of world w = pose w
    rng = torch.full((rows, cs.lidar.nray), -7.0)
    gid = torch.full((rows, cs.lidar.nray), -7, dtype=torch.int32)
    g = shard.BlockGather({"ranges": rng, "geomid": gid}, rows, nblock=2)

    def produce(k, a, b):
        for row in range(a, min(b, hi - lo)):
            r, i, _, _ = emu.lidar(cs.model, P, cs.lidar, pos[lo + row], mat[lo + row], world=row)
            rng[row] = torch.from_numpy(r)
            gid[row] = torch.from_numpy(i)

    res = g.run(produce)
    out[rank] = (res["ranges"].assemble().numpy(), res["geomid"].assemble().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_world_gather_matches_single_rank(oracle):
    from mujoco_ros2_simulation_b200 import scenes, shard
    world, W = 2, 5
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
    # the single-rank run: all W worlds in one process, through the oracle
    cs = scenes.clutter500(n_el=4, n_az=32)
    P = scenes.clutter_world_block(cs, 0, W)
    pos, mat = cs.sensor_poses(W)
    rows = shard.padded_rows(W, world)
    for rank in range(world):
        r_all, g_all = out[rank]
        assert r_all.shape == (world, rows, cs.lidar.nray)
        assert np.array_equal(r_all, out[0][0]) and np.array_equal(g_all, out[0][1])  # all ranks agree
    r_all, g_all = out[0]
    hits = 0
    for w in range(W):
        owner = shard.owner_of(w, W, world)
        lo, _ = shard.world_block(W, world, owner)
        o = oracle.OracleScene(cs.model, P, world=w)
        r, g, _ = o.lidar_scan(cs.lidar, pos[w], mat[w])
        assert np.array_equal(r_all[owner, w - lo], r) and np.array_equal(g_all[owner, w - lo], g)
        hits += int((g >= 0).sum())
    assert hits > 0
    # the padding row of the short block was never produced
    assert (g_all[1, rows - 1] == -7).all()
