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


def test_world_shards_are_contiguous_blocks():
    """cfg5 shape: W worlds split into contiguous blocks per rank; every world owned exactly once."""
    W, N = 37, 8
    owned = np.zeros(W, int)
    for rank in range(N):
        lo, hi = W * rank // N, W * (rank + 1) // N
        owned[lo:hi] += 1
    assert (owned == 1).all()
