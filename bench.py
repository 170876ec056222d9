#!/usr/bin/env python
"""bench.py -- lidar rays/s (Grays/s) on N B200s next to the mj_multiRay-style oracle on host cores.

Workload (BASELINE.json configs[1]): the clutter500 scene (1 plane + 4 walls +
primitives + mesh instances, ~500 geoms, seed 20261018) scanned by a 64x2048 3-D lidar from
`--scans` jittered sensor poses per GPU per step (default 1024 => 1.34e8 rays and 1.07 GB of
results per launch, far larger than the 126 MB L2).  One step = pose refresh + BVH refit (K1) +
one cast launch (K2).  Weak scaling: every rank casts its own batch of sensor poses against its
own replica of the scene; no collective on the data path.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the restatement of
MuJoCo's mj_multiRay; libmujoco itself is not available in this image) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lidar_rays_per_s"
UNIT = "Grays/s"
N_EL, N_AZ = 64, 2048
BYTES_PER_RAY = 8  # f32 range + i32 geomid written; rays are generated on-chip


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scans", type=int, default=1024, help="sensor poses per GPU per step")
    ap.add_argument("--e2e-scans", type=int, default=1024, help="sensor poses per e2e step")
    ap.add_argument("--ref-scans", type=int, default=24, help="scans per step of the CPU arm (~1-2 s/step)")
    ap.add_argument("--cpu-baseline-scans", type=int, default=192, help="~10-20 s of CPU work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--gather", action="store_true", help="also time an NCCL all-gather of results")
    return ap.parse_args()


def rank_pose_seed(rank: int) -> int:
    """Weak scaling: every rank casts its own batch of sensor poses (disjoint seeds) against its
    own replica of the scene; there is nothing to exchange on the data path."""
    return 20261018 + 1 + 7919 * rank


def job_rays_per_s(world: int, rays_per_rank_per_step: int, steps: int, ms_local: float, reduce_max) -> float:
    """Whole-job throughput: all ranks' rays over the slowest rank's device time.
    `reduce_max(x)` returns the max of x over ranks (identity when world == 1)."""
    ms = reduce_max(ms_local)
    return world * rays_per_rank_per_step * steps / (ms * 1e-3)


def env_rank():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


class NumaNear:
    """Run the enclosed block on the CPUs of the NUMA node the GPU hangs off, so that pinned host
    buffers allocated inside it are first-touched on that node (with 8 ranks on a 2-socket host
    the D2H streams otherwise cross the socket link).  Restores the affinity on exit; does nothing
    when the topology cannot be read."""

    def __init__(self, index: int):
        self.index, self.node, self.old = index, None, None

    def __enter__(self):
        try:
            import torch
            p = torch.cuda.get_device_properties(self.index)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            if node < 0:
                return self
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            self.old = os.sched_getaffinity(0)
            cpus &= self.old
            if cpus:
                os.sched_setaffinity(0, cpus)
                self.node = node
        except Exception:
            pass
        return self

    def __exit__(self, *a):
        if self.old is not None and self.node is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except Exception:
                pass


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self) -> bool:
        """NVML in-process (10 ms period): the timed region is ~0.1 s, too short for nvidia-smi."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id
            dev_id = torch.cuda.get_device_properties(self.index).pci_device_id
            dom = torch.cuda.get_device_properties(self.index).pci_domain_id
            h = nv.nvmlDeviceGetHandleByPciBusId(f"{dom:08X}:{bus:02X}:{dev_id:02X}.0")
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                    "sw_thermal_slowdown": 0x20}
            while not self.stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                row = [str(self.index), str(sm), str(mx), str(pw)]
                for name in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"):
                    row.append("Active" if (rs & bits[name]) else "Not Active")
                self.rows.append(row)
                self.stop.wait(0.01)
            return True
        except Exception:
            return False

    def _run(self):
        if self._run_nvml():
            return
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(scene_np, nscan: int, nthread: int, seed_off: int = 0):
    """Oracle (mj_multiRay restatement) on `nscan` full scans; returns (rays, seconds)."""
    from oracle import binding as orc
    oc = orc.OracleScene(scene_np.model, scene_np.poses)
    pos, mat = scene_np.sensor_poses(nscan, seed=20261018 + 1 + seed_off)
    t0 = time.perf_counter()
    for i in range(nscan):
        oc.lidar_scan(scene_np.lidar, pos[i], mat[i], nthread=nthread)
    return nscan * scene_np.lidar.nray, time.perf_counter() - t0


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    from mujoco_ros2_simulation_b200 import scenes
    from oracle import binding as orc
    orc.build()
    cs = scenes.clutter500(n_el=N_EL, n_az=N_AZ)
    cores = os.cpu_count() or 1
    for w in range(args.warmup):
        cpu_arm(cs, 1, cores, w)
    t0 = time.perf_counter()
    rays = 0
    for k in range(args.steps):
        r, _ = cpu_arm(cs, args.ref_scans, cores, 100 + k)
        rays += r
    dt = time.perf_counter() - t0
    val = rays / dt / 1e9
    sample = f"{args.ref_scans} scans x {cs.lidar.nray} rays per step on the same scene"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"clutter500 (~500 geoms, meshes) 64x2048 lidar, {args.ref_scans} scans/step",
                   "note": "CPU oracle = C restatement of MuJoCo 3.3.4 mj_multiRay (libmujoco absent)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def other_configs(engine, scenes, torch, dev, stream, nworld3=4096, nworld4=1024):
    """BASELINE.json configs[2] and configs[3] as secondary measurements (device-resident):
    cfg3: 4096 replicated arm worlds, one 1024-ray lidar each, all poses refit every step;
    cfg4: 1024 worlds, a 1024-ray lidar + a 640x480 depth camera (linearise + cloud) each."""
    out = {}
    ws = scenes.arm_worlds(nworld=nworld3, n_rays=1024)
    sc = engine.Scene(ws.model, nworld=nworld3, device=dev.index, stream=stream.cuda_stream)
    lid = engine.Lidar(sc, ws.lidar)
    with torch.cuda.stream(stream):
        xpos = torch.from_numpy(ws.poses.geom_xpos).to(dev)
        xmat = torch.from_numpy(ws.poses.geom_xmat).to(dev)
        spos = torch.from_numpy(ws.sensor_pos).to(dev)
        smat = torch.from_numpy(ws.sensor_mat).to(dev)
        sw = torch.from_numpy(ws.scan_world).to(dev)
        rng = torch.empty((nworld3, 1024), dtype=torch.float32, device=dev)
        gid = torch.empty((nworld3, 1024), dtype=torch.int32, device=dev)

        def step3(n):
            sc.update_poses_device(xpos[:n], xmat[:n])
            lid.cast_device(n, spos, smat, sw, ranges=rng, geomid=gid)

        for _ in range(3):
            step3(nworld3)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record(stream)
        for _ in range(reps):
            step3(nworld3)
        b.record(stream)
        stream.synchronize()
        ms = a.elapsed_time(b) / reps
        out["cfg3_arm_worlds"] = {"workload": f"{nworld3} worlds x 1024-ray lidar, poses refit per step",
                                  "grays_per_s": nworld3 * 1024 / (ms * 1e-3) / 1e9, "ms_per_step": ms,
                                  "hit_fraction": float((gid >= 0).float().mean().item())}
        # cfg4: the first nworld4 worlds + one depth camera per world
        cam = engine.DepthCamera(scenes.depth_camera(), scene=sc, device=dev.index)
        gl = torch.rand((nworld4, 480, 640), dtype=torch.float32, device=dev)
        dimg = torch.empty_like(gl)
        cloud = torch.empty((nworld4, 480, 640, 3), dtype=torch.float32, device=dev)

        def step4():
            step3(nworld4)
            cam.depth_to_cloud_device(nworld4, gl, dimg, cloud)

        for _ in range(3):
            step4()
        reps = 5
        a.record(stream)
        for _ in range(reps):
            step4()
        b.record(stream)
        stream.synchronize()
        ms = a.elapsed_time(b) / reps
        npx = nworld4 * 480 * 640
        out["cfg4_depth_plus_lidar"] = {
            "workload": f"{nworld4} worlds x (640x480 depth -> 32FC1 + cloud, 1024-ray lidar)",
            "ms_per_step": ms, "gpx_per_s": npx / (ms * 1e-3) / 1e9,
            "hbm_gbs_algorithmic": (npx * 20 + nworld4 * 1024 * 8) / (ms * 1e-3) / 1e9}
        del gl, dimg, cloud
        cam.close()
    lid.close()
    sc.close()
    out["cfg5_clutter_worlds"] = cfg5(engine, scenes, torch, dev, stream)
    return out


def cfg5(engine, scenes, torch, dev, stream, nworld=256):
    """BASELINE.json configs[4] per-GPU shape at a bounded world count: `nworld` worlds of the
    clutter500 topology (per-world body jitter, one shared BVH topology, refit every step), each with
    the 64x2048 lidar and a 640x480 depth camera whose GL depth buffer is ray-cast from the same
    scene mirror (N1), then linearised + projected (K3).  4096 worlds/GPU is the full size."""
    cs = scenes.clutter500()
    P = scenes.clutter_worlds(cs, nworld)
    sc = engine.Scene(cs.model, nworld=nworld, device=dev.index, stream=stream.cuda_stream)
    lid = engine.Lidar(sc, cs.lidar)
    cam = engine.DepthCamera(scenes.depth_camera(), scene=sc, device=dev.index)
    pos, mat = cs.sensor_poses(nworld)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    with torch.cuda.stream(stream):
        xpos, xmat = torch.from_numpy(P.geom_xpos).to(dev), torch.from_numpy(P.geom_xmat).to(dev)
        spos, smat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
        sw = torch.arange(nworld, dtype=torch.int32, device=dev)
        cpos = torch.from_numpy(np.tile([[3.0, 0.0, 1.5]], (nworld, 1))).to(dev)
        cmat = torch.from_numpy(np.tile(R, (nworld, 1))).to(dev)
        rng = torch.empty((nworld, lid.nray), dtype=torch.float32, device=dev)
        gid = torch.empty((nworld, lid.nray), dtype=torch.int32, device=dev)
        gl = torch.empty((nworld, 480, 640), dtype=torch.float32, device=dev)
        dimg = torch.empty_like(gl)
        cloud = torch.empty((nworld, 480, 640, 3), dtype=torch.float32, device=dev)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))

        def step():
            sc.update_poses_device(xpos, xmat)
            lid.cast_device(nworld, spos, smat, sw, ranges=rng, geomid=gid)
            b.record(stream)
            cam.render_depth_device(nworld, cpos, cmat, gl, d_world=sw)
            cam.depth_to_cloud_device(nworld, gl, dimg, cloud)

        for _ in range(2):
            step()
        reps, ms_l, ms_c = 3, 0.0, 0.0
        for _ in range(reps):
            a.record(stream)
            step()
            c.record(stream)
            stream.synchronize()
            ms_l += a.elapsed_time(b) / reps
            ms_c += b.elapsed_time(c) / reps
        res = {"workload": f"{nworld} worlds (of 4096/GPU) x (64x2048 lidar + 640x480 ray-cast depth -> 32FC1 + "
                           "cloud) on the clutter500 topology, poses refit per step",
               "lidar_grays_per_s": nworld * lid.nray / (ms_l * 1e-3) / 1e9, "lidar_ms": ms_l,
               "camera_grays_per_s": nworld * 480 * 640 / (ms_c * 1e-3) / 1e9, "camera_ms": ms_c,
               "hit_fraction": float((gid >= 0).float().mean().item())}
    cam.close()
    lid.close()
    sc.close()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist

    from mujoco_ros2_simulation_b200 import engine, scenes

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 must print exactly ONE line on stdout, and NCCL writes its version banner there when
        # the communicator is created: point fd 1 at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()

    cs = scenes.clutter500(n_el=N_EL, n_az=N_AZ)
    nray = cs.lidar.nray
    stream = torch.cuda.Stream(device=dev)
    scene = engine.Scene(cs.model, nworld=1, device=local_rank, stream=stream.cuda_stream)
    scene.update_poses(cs.poses)
    lidar = engine.Lidar(scene, cs.lidar)
    nscan = args.scans
    pos, mat = cs.sensor_poses(nscan, seed=rank_pose_seed(rank))
    with torch.cuda.stream(stream):
        d_pos = torch.from_numpy(pos).to(dev)
        d_mat = torch.from_numpy(mat).to(dev)
        d_world = torch.zeros(nscan, dtype=torch.int32, device=dev)
        d_xpos = torch.from_numpy(cs.poses.geom_xpos).to(dev)
        d_xmat = torch.from_numpy(cs.poses.geom_xmat).to(dev)
        ranges = torch.empty((nscan, nray), dtype=torch.float32, device=dev)
        geomid = torch.empty((nscan, nray), dtype=torch.int32, device=dev)
    stream.synchronize()

    def step(ev0=None, ev1=None):
        scene.update_poses_device(d_xpos, d_xmat)  # K1: pose refresh + refit, every step
        if ev0 is not None:
            ev0.record(stream)
        lidar.cast_device(nscan, d_pos, d_mat, d_world, ranges=ranges, geomid=geomid)  # K2
        if ev1 is not None:
            ev1.record(stream)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = engine.launch_count()
        with ClockSampler(local_rank) as clk:
            e0.record(stream)
            for k in range(args.steps):
                step(*kev[k])
            e1.record(stream)
            stream.synchronize()
            torch.cuda.synchronize()
        launches = engine.launch_count() - l0
        barrier()
    ms_total = e0.elapsed_time(e1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    def reduce_max(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total_max = reduce_max(ms_total)
    rays_per_step = nscan * nray
    value = job_rays_per_s(world, rays_per_step, args.steps, ms_total, reduce_max) / 1e9

    # sanity on the result of the last step (not timed): hit fraction of this scene is ~0.99
    hit_frac = float((geomid[0] >= 0).float().mean().item())

    # ---- e2e: host buffers through the blocking C-ABI call, copies inside the timed region ------
    es = min(args.e2e_scans, nscan)
    with NumaNear(local_rank) as near:  # pinned buffers on the GPU's own NUMA node
        h_pos = torch.from_numpy(pos[:es].copy()).pin_memory()
        h_mat = torch.from_numpy(mat[:es].copy()).pin_memory()
        h_world = torch.zeros(es, dtype=torch.int32).pin_memory()
        h_ranges = torch.empty((es, nray), dtype=torch.float32).pin_memory()
        h_geomid = torch.empty((es, nray), dtype=torch.int32).pin_memory()
        h_xpos = torch.from_numpy(cs.poses.geom_xpos.copy()).pin_memory()
        h_xmat = torch.from_numpy(cs.poses.geom_xmat.copy()).pin_memory()
        h_ranges.zero_()
        h_geomid.zero_()
    lib = scene.lib

    def e2e_step():
        engine._capi.check(lib.b2_scene_update_poses(scene._h, 0, 1, h_xpos.data_ptr(), h_xmat.data_ptr()))
        engine._capi.check(lib.b2_lidar_cast(lidar._h, es, h_world.data_ptr(), h_pos.data_ptr(),
                                             h_mat.data_ptr(), h_ranges.data_ptr(), h_geomid.data_ptr(),
                                             None, None))

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * es * nray * args.steps / float(t.item()) / 1e9
    h2d = es * (3 + 9) * 8 + es * 4 + int(h_xpos.numel() + h_xmat.numel()) * 8
    d2h = es * nray * BYTES_PER_RAY
    e2e_ok = bool((h_geomid[0].numpy() == geomid[0].cpu().numpy()).all()) if es > 0 else True

    # the same call returning what MujocoLidar::update publishes -- the LaserScan ranges only, 4 B/ray
    # (ref:src/mujoco_lidar.cpp:256-279; geomid is not part of the message)
    def e2e_scan_step():
        engine._capi.check(lib.b2_scene_update_poses(scene._h, 0, 1, h_xpos.data_ptr(), h_xmat.data_ptr()))
        engine._capi.check(lib.b2_lidar_cast(lidar._h, es, h_world.data_ptr(), h_pos.data_ptr(),
                                             h_mat.data_ptr(), h_ranges.data_ptr(), None, None, None))

    e2e_scan_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_scan_step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_scan_val = world * es * nray * args.steps / float(t.item()) / 1e9

    # ---- optional: gather of the results over NVLink (not on the data path) -------------------------
    gather_ms = None
    if args.gather and world > 1:
        g_r = torch.empty((world,) + tuple(ranges.shape), dtype=ranges.dtype, device=dev)
        g_g = torch.empty((world,) + tuple(geomid.shape), dtype=geomid.dtype, device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(g_r, ranges)
            dist.all_gather_into_tensor(g_g, geomid)
        torch.cuda.synchronize()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_gather_into_tensor(g_r, ranges)
        dist.all_gather_into_tensor(g_g, geomid)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = float(t.item())
        del g_r, g_g

    # ---- secondary: K3 depth -> cloud (the genuinely HBM-bound kernel), cfg4 shape ---------------------
    secondary = None
    if not args.no_secondary and rank == 0:
        ncam = 256
        cam = engine.DepthCamera(scenes.depth_camera(), scene=scene, device=local_rank)
        with torch.cuda.stream(stream):
            gl = torch.rand((ncam, 480, 640), dtype=torch.float32, device=dev)
            dimg = torch.empty_like(gl)
            cloud = torch.empty((ncam, 480, 640, 3), dtype=torch.float32, device=dev)
            for _ in range(3):
                cam.depth_to_cloud_device(ncam, gl, dimg, cloud)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            reps = 10
            for _ in range(reps):
                cam.depth_to_cloud_device(ncam, gl, dimg, cloud)
            b.record(stream)
            stream.synchronize()
        ms = a.elapsed_time(b) / reps
        npx = ncam * 480 * 640
        peak, how = measured_peak()
        gbs = npx * 20 / (ms * 1e-3) / 1e9
        secondary = {"kernel": "depth_to_cloud", "workload": f"{ncam} x 640x480 depth -> 32FC1 + xyz (point_step 12)",
                     "gpx_per_s": npx / (ms * 1e-3) / 1e9, "ms": ms,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                  "frac": gbs / peak, "bytes_per_px": 20, "peak_source": how}}
        del gl, dimg, cloud
        cam.close()
        secondary = {"depth_to_cloud": secondary}
        try:
            secondary.update(other_configs(engine, scenes, torch, dev, stream))
        except Exception as e:  # the headline must not die on a secondary shape
            secondary["other_configs_error"] = repr(e)

    # ---- roofline of the dominant kernel (K2 cast) ------------------------------------------------------
    st = scene.stats()
    # GeomStatic 32 + GeomPose 96 per geom, 32 per BVH child slot, 128 per fanout-4 mesh node, 48 per triangle
    scene_bytes = cs.model.ngeom * (32 + 96) + st["bvh_nodes"] * 32 + st["mesh_nodes"] * 128 + st["mesh_tris"] * 48
    alg_bytes = BYTES_PER_RAY * rays_per_step + scene_bytes
    peak, how = measured_peak()
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("cast_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- CPU baseline (rank 0, N=1): the oracle on a bounded sample of the same workload -------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding as orc
        orc.build()
        cores = os.cpu_count() or 1
        cpu_arm(cs, 1, cores)
        rays, secs = cpu_arm(cs, args.cpu_baseline_scans, cores, 50)
        cpu = {"value": rays / secs / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_baseline_scans} of the {nscan} scans ({rays} rays), oracle "
                         "(C restatement of mj_multiRay, -O3) on all host cores"}
        # the same oracle on one thread (what the reference's physics thread has)
        r1, s1 = cpu_arm(cs, 4, 1, 90)
        cpu["single_thread"] = {"value": r1 / s1 / 1e9, "unit": UNIT, "sample": f"4 scans ({r1} rays)"}
        # parity of THIS run's GPU output against the oracle on sampled scans of the timed batch
        oc = orc.OracleScene(cs.model, cs.poses)
        g_host, r_host = geomid.cpu().numpy(), ranges.cpu().numpy()
        checked = bad_id = bad_rng = 0
        for sc in sorted({0, nscan // 3, nscan - 1}):
            rr, gg, _ = oc.lidar_scan(cs.lidar, pos[sc], mat[sc], nthread=cores)
            checked += rr.size
            bad_id += int((gg != g_host[sc]).sum())
            bad_rng += int((rr != r_host[sc]).sum())
        parity = {"rays_checked": checked, "geomid_mismatch": bad_id, "range_bits_mismatch": bad_rng,
                  "against": "CPU oracle (oracle/), same poses; ranges compared bit for bit"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"clutter500 ({cs.model.ngeom} geoms incl. {int((cs.model.geom_type == 7).sum())} "
                            f"mesh instances, {cs.model.nface} unique triangles) 64x2048 3-D lidar x {nscan} "
                            "sensor poses per GPU per step; pose refresh + BVH refit every step",
                "rays_per_step_per_gpu": rays_per_step, "l2": "outputs 1.07 GB per launch > 126 MB L2",
                "parallelism": f"replica x{world}, sensor-pose batches sharded, no data-path collective",
                "gather_ms": gather_ms, "hit_fraction": hit_frac, "e2e_matches_device_path": e2e_ok,
                "bvh_nodes": st["bvh_nodes"], "mesh_nodes": st["mesh_nodes"],
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "K2 cast = collect_kernel -> drain_kernel -> mesh-pair kernels (one b2_lidar_cast_device; drain_kernel is the largest at ~50 %)",
                         "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": how,
                         "note": "ray casting is traversal/FP64-issue bound, not HBM bound (SURVEY.md M3)"},
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "scans_per_step": es, "pinned_numa_node": near.node,
                    "outputs": "ranges f32 + geomid i32 (8 B/ray over PCIe)",
                    "laserscan_only": {"value": e2e_scan_val, "unit": UNIT, "d2h_bytes_per_step": es * nray * 4,
                                       "outputs": "ranges f32 only, as MujocoLidar::update publishes"}},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
