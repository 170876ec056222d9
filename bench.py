#!/usr/bin/env python
"""bench.py -- lidar rays/s (Grays/s) on N B200s next to the mj_multiRay-style oracle on host cores.

Workload (BASELINE.json configs[1]): the clutter500 scene (1 plane + 4 walls +
primitives + mesh instances, ~500 geoms, seed 20261018) scanned by a 64x2048 3-D lidar from
`--scans` jittered sensor poses per GPU per step (default 1024 => 1.34e8 rays and 1.07 GB of
results per launch, far larger than the 126 MB L2).  One step = pose refresh + BVH refit (K1) +
one cast launch (K2).  Weak scaling: every rank casts its own batch of sensor poses against its
own replica of the scene.

`value`  = device-resident rays/s, results left sharded on the ranks.
`gather` = (N > 1) the same step with the results gathered over NVLink: the cast kernels of every
           rank store straight into the root's peer buffer (shard.PeerGather -- the stores ARE the
           gather), and, as the library baseline, torch.distributed all-gather of scan blocks
           overlapped with the next block's cast (shard.BlockGather).
`e2e`    = the blocking host-pointer C-ABI call, copies inside the timed region; headline = the
           LaserScan payload the reference publishes (f32 ranges), with the (ranges, geomid) and
           (ranges, uint16 geomid) variants and the measured host-path ceiling beside it.
cfg5     = BASELINE.json configs[4] on EVERY rank: 4096 worlds per GPU of the clutter topology,
           64x2048 lidar + 640x480 ray-cast depth camera each, gathered at N > 1.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the restatement of
MuJoCo's mj_ray / mj_multiRay; libmujoco itself is not available in this image) on all host cores.
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

# the CPU arm's variants: what the reference executes per rangefinder, and the
# mj_multiRay-style accelerated forms (oracle/b2_oracle.c)
CPU_VARIANTS = {
    "mj_ray_loop": "brute-force mj_ray per ray over all geoms (what MuJoCo runs per rangefinder in mj_step)",
    "multiray_body_cull": "mj_multiRay: per-body azimuth/elevation rectangle cull (SURVEY A.6) + per-mesh BVH",
    "multiray_geom_cull": "mj_multiRay-style with a per-geom bounding-sphere cull (A.6 with one geom per body)",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scans", type=int, default=1024, help="sensor poses per GPU per step")
    ap.add_argument("--e2e-scans", type=int, default=1024, help="sensor poses per e2e step")
    ap.add_argument("--ref-scans", type=int, default=24, help="scans per step of the CPU arm (~1 s/step)")
    ap.add_argument("--cpu-baseline-scans", type=int, default=96, help="scans of the fastest CPU variant (~5-10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the NVLink gather legs")
    ap.add_argument("--gather-blocks", type=int, default=8, help="scan blocks of the overlapped all-gather")
    ap.add_argument("--cfg5-worlds", type=int, default=4096, help="worlds per GPU of the cfg5 leg (0: skip)")
    ap.add_argument("--no-latency", action="store_true")
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


def cpu_arm(scene_np, nscan: int, nthread: int, seed_off: int = 0, variant: str = "mj_ray_loop", oc=None):
    """Oracle on `nscan` full scans with one of CPU_VARIANTS; returns (rays, seconds)."""
    from oracle import binding as orc
    oc = oc or orc.OracleScene(scene_np.model, scene_np.poses)
    pos, mat = scene_np.sensor_poses(nscan, seed=20261018 + 1 + seed_off)
    cull = {"multiray_body_cull": orc.CULL_BODY_ANGULAR, "multiray_geom_cull": orc.CULL_GEOM_SPHERE}.get(variant)
    t0 = time.perf_counter()
    for i in range(nscan):
        if cull is None:
            oc.lidar_scan(scene_np.lidar, pos[i], mat[i], nthread=nthread)
        else:
            oc.lidar_scan_multiray(scene_np.lidar, pos[i], mat[i], cull=cull, nthread=nthread)
    return nscan * scene_np.lidar.nray, time.perf_counter() - t0


def cpu_variants(cs, cores: int, budget_s: float, oc=None):
    """Time every CPU variant on a bounded sample (~budget_s each): {name: {value, cores, sample}}."""
    from oracle import binding as orc
    oc = oc or orc.OracleScene(cs.model, cs.poses)
    out = {}
    for k, name in enumerate(CPU_VARIANTS):
        cpu_arm(cs, 1, cores, 900 + k, name, oc)                       # warm
        r1, s1 = cpu_arm(cs, 2, cores, 910 + k, name, oc)              # rate estimate
        n = int(max(2, min(512, budget_s / max(s1 / 2, 1e-3))))
        rays, secs = cpu_arm(cs, n, cores, 920 + k, name, oc)
        out[name] = {"value": rays / secs / 1e9, "unit": UNIT, "cores": cores,
                     "sample": f"{n} scans ({rays} rays)", "what": CPU_VARIANTS[name]}
    return out


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    from mujoco_ros2_simulation_b200 import scenes
    from oracle import binding as orc
    orc.build()
    cs = scenes.clutter500(n_el=N_EL, n_az=N_AZ)
    oc = orc.OracleScene(cs.model, cs.poses)
    cores = os.cpu_count() or 1
    # every variant on a bounded sample; the line's value is the FASTEST one, timed over the K steps
    variants = cpu_variants(cs, cores, 4.0, oc)
    best = max(variants, key=lambda k: variants[k]["value"])
    for w in range(args.warmup):
        cpu_arm(cs, 1, cores, w, best, oc)
    t0 = time.perf_counter()
    rays = 0
    for k in range(args.steps):
        r, _ = cpu_arm(cs, args.ref_scans, cores, 100 + k, best, oc)
        rays += r
    dt = time.perf_counter() - t0
    val = rays / dt / 1e9
    sample = f"{args.ref_scans} scans x {cs.lidar.nray} rays per step on the same scene, variant {best}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"clutter500 (~500 geoms, meshes) 64x2048 lidar, {args.ref_scans} scans/step",
                   "note": f"CPU oracle = C restatement of MuJoCo 3.3.4 mj_ray / mj_multiRay (libmujoco absent); "
                           f"the line's value is the fastest variant: {best} = {CPU_VARIANTS[best]}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "variant": best, "variants": variants},
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
    return out


def cfg5(ctx, nworld: int):
    """BASELINE.json configs[4], this rank's share: `nworld` worlds of the clutter500 topology (the job
    has nworld x N worlds; rank r owns shard.world_block's r-th block; per-world body jitter, one
    shared BVH topology, every world refit every step), each with the 64x2048 lidar and a 640x480
    depth camera whose GL depth buffer is ray-cast from the same scene mirror (N1), then linearised +
    projected (K3).  At N > 1 the (range, geomid) rows and the 32FC1 depth images are gathered on
    rank 0 over NVLink by the kernels' own stores (shard.PeerGather); clouds stay sharded."""
    torch, dist, engine, scenes, shard = ctx["torch"], ctx["dist"], ctx["engine"], ctx["scenes"], ctx["shard"]
    dev, stream, rank, world = ctx["dev"], ctx["stream"], ctx["rank"], ctx["world"]
    cs = scenes.clutter500()
    lo, hi = shard.world_block(nworld * world, world, rank)
    assert hi - lo == nworld
    P = scenes.clutter_world_block(cs, lo, nworld)
    sc = engine.Scene(cs.model, nworld=nworld, device=dev.index, stream=stream.cuda_stream)
    lid = engine.Lidar(sc, cs.lidar)
    cam = engine.DepthCamera(scenes.depth_camera(), scene=sc, device=dev.index)
    rng_s = np.random.default_rng([20261018, 55, rank])
    pos = cs.sensor_pos[None] + rng_s.uniform(-0.3, 0.3, size=(nworld, 3))
    mat = np.tile(np.eye(3).reshape(1, 9), (nworld, 1))
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    nray, npx = lid.nray, 480 * 640
    pg = None
    res = {}
    with torch.cuda.stream(stream):
        xpos, xmat = torch.from_numpy(P.geom_xpos).to(dev), torch.from_numpy(P.geom_xmat).to(dev)
        spos, smat = torch.from_numpy(pos).to(dev), torch.from_numpy(mat).to(dev)
        sw = torch.arange(nworld, dtype=torch.int32, device=dev)
        cpos = torch.from_numpy(np.tile([[3.0, 0.0, 1.5]], (nworld, 1))).to(dev)
        cmat = torch.from_numpy(np.tile(R, (nworld, 1))).to(dev)
        rng = torch.empty((nworld, nray), dtype=torch.float32, device=dev)
        gid = torch.empty((nworld, nray), dtype=torch.int32, device=dev)
        gl = torch.empty((nworld, 480, 640), dtype=torch.float32, device=dev)
        dimg = torch.empty_like(gl)
        cloud = torch.empty((nworld, 480, 640, 3), dtype=torch.float32, device=dev)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))

        def step(out_r, out_g, out_d):
            sc.update_poses_device(xpos, xmat)
            lid.cast_device(nworld, spos, smat, sw, ranges=out_r, geomid=out_g)
            b.record(stream)
            cam.render_depth_device(nworld, cpos, cmat, gl, d_world=sw)
            cam.depth_to_cloud_device(nworld, gl, out_d, cloud)

        def timed(out_r, out_g, out_d, reps):
            ms_l = ms_c = 0.0
            for _ in range(reps):
                ctx["barrier"]()
                a.record(stream)
                step(out_r, out_g, out_d)
                c.record(stream)
                stream.synchronize()
                ctx["barrier"]()
                ms_l += a.elapsed_time(b) / reps
                ms_c += b.elapsed_time(c) / reps
            return ctx["reduce_max"](ms_l), ctx["reduce_max"](ms_c)

        step(rng, gid, dimg)
        ms_l, ms_c = timed(rng, gid, dimg, 2)
        rays_l, rays_c = world * nworld * nray, world * nworld * npx
        res = {"workload": f"{nworld} worlds per GPU x {world} GPU(s) = {nworld * world} worlds (rank r owns block r) x "
                           "(64x2048 lidar + 640x480 ray-cast depth -> 32FC1 + cloud) on the clutter500 topology, "
                           "all poses refit per step",
               "worlds_per_gpu": nworld, "n_gpus": world,
               "lidar_grays_per_s": rays_l / (ms_l * 1e-3) / 1e9, "lidar_ms": ms_l,
               "camera_grays_per_s": rays_c / (ms_c * 1e-3) / 1e9, "camera_ms": ms_c,
               "step_ms": ms_l + ms_c,
               "hit_fraction": float((gid[:64] >= 0).float().mean().item())}
        if world > 1 and not ctx["args"].no_gather:
            try:
                pg = shard.PeerGather({"ranges": (nray, "float32"), "geomid": (nray, "int32"),
                                       "depth": (npx, "float32")}, nworld, rank, world, dev.index)
                sums = (int(gid.to(torch.int64).sum().item()), int(rng.view(torch.int32).to(torch.int64).sum().item()),
                        int(dimg.view(torch.int32).to(torch.int64).sum().item()))
                # (i) the kernels store straight into the root's buffers
                step(pg.ptr("ranges"), pg.ptr("geomid"), pg.ptr("depth"))
                g_l, g_c = timed(pg.ptr("ranges"), pg.ptr("geomid"), pg.ptr("depth"), 2)
                stream.synchronize()
                ctx["barrier"]()
                # (ii) world blocks: cast / render locally, the copy engine carries each finished block
                wblocks = shard.scan_blocks(nworld, 8)
                if rank == 0:  # what (i) left there must not be what the checksums see
                    for name in ("ranges", "geomid", "depth"):
                        pg.root_tensor(name).zero_()
                stream.synchronize()
                ctx["barrier"]()

                def step_push():
                    sc.update_poses_device(xpos, xmat)
                    for lo_b, hi_b in wblocks:
                        n_b = hi_b - lo_b
                        lid.cast_device(n_b, spos[lo_b:hi_b], smat[lo_b:hi_b], sw[lo_b:hi_b], ranges=rng[lo_b:hi_b],
                                        geomid=gid[lo_b:hi_b])
                        pg.push(sc, "ranges", lo_b, rng[lo_b:hi_b])
                        pg.push(sc, "geomid", lo_b, gid[lo_b:hi_b])
                    for lo_b, hi_b in wblocks:
                        n_b = hi_b - lo_b
                        cam.render_depth_device(n_b, cpos[lo_b:hi_b], cmat[lo_b:hi_b], gl[lo_b:hi_b], d_world=sw[lo_b:hi_b])
                        cam.depth_to_cloud_device(n_b, gl[lo_b:hi_b], dimg[lo_b:hi_b], cloud[lo_b:hi_b])
                        pg.push(sc, "depth", lo_b, dimg[lo_b:hi_b])
                    sc.join()

                step_push()
                push_ms = 0.0
                for _ in range(2):
                    ctx["barrier"]()
                    a.record(stream)
                    step_push()
                    c.record(stream)
                    stream.synchronize()
                    ctx["barrier"]()
                    push_ms += a.elapsed_time(c) / 2
                push_ms = ctx["reduce_max"](push_ms)
                t = torch.tensor(sums, dtype=torch.int64, device=dev)
                allsums = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(allsums, t)
                ok = None
                if rank == 0:
                    ok = True
                    G, Rr, D = pg.root_tensor("geomid"), pg.root_tensor("ranges"), pg.root_tensor("depth")
                    for r in range(world):
                        want = [int(v) for v in allsums[r].tolist()]
                        got = [int(G[r].to(torch.int64).sum().item()),
                               int(Rr[r].view(torch.int32).to(torch.int64).sum().item()),
                               int(D[r].view(torch.int32).to(torch.int64).sum().item())]
                        ok = ok and got == want
                gbytes = (world - 1) * nworld * (nray * 8 + npx * 4)
                res["gather"] = {"arrays": "ranges f32 + geomid i32 + depth 32FC1 of every world, on rank 0",
                                 "root_ingress_gb_per_step": gbytes / 1e9,
                                 "p2p_store_to_root": {
                                     "lidar_ms": g_l, "camera_ms": g_c, "step_ms": g_l + g_c,
                                     "lidar_grays_per_s": rays_l / (g_l * 1e-3) / 1e9,
                                     "camera_grays_per_s": rays_c / (g_c * 1e-3) / 1e9,
                                     "root_ingress_gbs": gbytes / ((g_l + g_c) * 1e-3) / 1e9},
                                 "p2p_copy_blocks": {
                                     "step_ms": push_ms, "world_blocks": len(wblocks),
                                     "rays_plus_pixels_grays_per_s": (rays_l + rays_c) / (push_ms * 1e-3) / 1e9,
                                     "root_ingress_gbs": gbytes / (push_ms * 1e-3) / 1e9},
                                 "sharded_step_ms": ms_l + ms_c,
                                 "verified_checksums": ok}
            except Exception as e:  # the headline must not die on this leg
                res["gather"] = {"error": repr(e)}
    if pg is not None:
        stream.synchronize()
        ctx["barrier"]()
        pg.close()
    cam.close()
    lid.close()
    sc.close()
    return res


def measure_gather(ctx, step_ms_sharded: float):
    """N > 1: one bench step with the (range, geomid) results gathered over NVLink.
    (a) p2p_store_to_root: every rank's cast kernels store into the root's peer buffer -- the result
        stores are the gather (no staging, no collective); one barrier per step orders the processes;
    (c) p2p_copy_blocks: the same peer buffer filled in bulk -- block k is cast locally and handed to
        the copy engine while block k+1 is cast (the better form once the root's ingress limits);
    (b) nccl_allgather_overlap: torch.distributed all-gather of scan block k on a side stream while
        block k+1 is cast (every rank ends up with everything; the library baseline)."""
    torch, dist, shard = ctx["torch"], ctx["dist"], ctx["shard"]
    dev, stream, rank, world, args = ctx["dev"], ctx["stream"], ctx["rank"], ctx["world"], ctx["args"]
    scene, lidar, nscan, nray = ctx["scene"], ctx["lidar"], ctx["nscan"], ctx["nray"]
    d_pos, d_mat, d_world, d_xpos, d_xmat = ctx["d_pos"], ctx["d_mat"], ctx["d_world"], ctx["d_xpos"], ctx["d_xmat"]
    ranges, geomid = ctx["ranges"], ctx["geomid"]
    rays_job = world * nscan * nray
    out = {}

    def run(step, steps):
        for _ in range(3):
            step()
        stream.synchronize()
        ctx["barrier"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        ctx["barrier"]()
        return ctx["reduce_max"](e0.elapsed_time(e1)) / steps

    with torch.cuda.stream(stream):
        sums = torch.tensor([int(geomid.to(torch.int64).sum().item()),
                             int(ranges.view(torch.int32).to(torch.int64).sum().item())], dtype=torch.int64, device=dev)
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        # ---- (a) the kernels' own stores are the gather ------------------------------------------------
        try:
            pg = shard.PeerGather({"ranges": (nray, "float32"), "geomid": (nray, "int32")}, nscan, rank, world,
                                  dev.index)

            def step_p2p():
                scene.update_poses_device(d_xpos, d_xmat)
                lidar.cast_device(nscan, d_pos, d_mat, d_world, ranges=pg.ptr("ranges"), geomid=pg.ptr("geomid"))
                dist.barrier()  # stream-ordered: the root may read once every rank's stores have landed

            ms = run(step_p2p, args.steps)
            ok = None
            if rank == 0:
                G, Rr = pg.root_tensor("geomid"), pg.root_tensor("ranges")
                ok = all([int(G[r].to(torch.int64).sum().item()),
                          int(Rr[r].view(torch.int32).to(torch.int64).sum().item())] ==
                         [int(v) for v in allsums[r].tolist()] for r in range(world))
                ok = ok and bool(torch.equal(G[0], geomid))
            gbytes = (world - 1) * nscan * nray * BYTES_PER_RAY
            out["p2p_store_to_root"] = {
                "ms_per_step": ms, "value_with_gather": rays_job / (ms * 1e-3) / 1e9,
                "gather_ms": ms - step_ms_sharded, "root_ingress_gb_per_step": gbytes / 1e9,
                "root_ingress_gbs": gbytes / (ms * 1e-3) / 1e9, "verified_checksums": ok,
                "how": "cast kernels store float4 / int4 lines into rank 0's peer buffer (CUDA IPC, P2P over "
                       "NVLink/NVSwitch); one barrier per step"}
            stream.synchronize()
            ctx["barrier"]()
            pg.close()
        except Exception as e:
            out["p2p_store_to_root"] = {"error": repr(e)}
        # ---- (c) block by block: cast into local memory, the copy engine carries the finished block ----------
        try:
            pg = shard.PeerGather({"ranges": (nray, "float32"), "geomid": (nray, "int32")}, nscan, rank, world,
                                  dev.index)
            blocks = shard.scan_blocks(nscan, args.gather_blocks)

            def step_push():
                scene.update_poses_device(d_xpos, d_xmat)
                for lo, hi in blocks:
                    lidar.cast_device(hi - lo, d_pos[lo:hi], d_mat[lo:hi], d_world[lo:hi], ranges=ranges[lo:hi],
                                      geomid=geomid[lo:hi])
                    pg.push(scene, "ranges", lo, ranges[lo:hi])
                    pg.push(scene, "geomid", lo, geomid[lo:hi])
                scene.join()    # the scene's stream (and the timing events on it) wait for the copies
                dist.barrier()

            ms = run(step_push, args.steps)

            def verify():
                if rank != 0:
                    return None
                G, Rr = pg.root_tensor("geomid"), pg.root_tensor("ranges")
                good = all([int(G[r].to(torch.int64).sum().item()),
                            int(Rr[r].view(torch.int32).to(torch.int64).sum().item())] ==
                           [int(v) for v in allsums[r].tolist()] for r in range(world))
                G.zero_()
                Rr.zero_()
                return good

            ok = verify()
            stream.synchronize()
            ctx["barrier"]()
            # the same as a STREAMING loop: results double-buffered on the rank, the tail of step s's
            # transfers overlaps the first block of step s+1 (b2_scene_copy_fence: one-step lag); the
            # last step is joined inside the timed region
            bufs = ((ranges, geomid), (torch.empty_like(ranges), torch.empty_like(geomid)))
            parity = [0]
            sblocks = [blocks]

            def step_stream():
                r_, g_ = bufs[parity[0]]
                parity[0] ^= 1
                scene.update_poses_device(d_xpos, d_xmat)
                for lo, hi in sblocks[0]:
                    lidar.cast_device(hi - lo, d_pos[lo:hi], d_mat[lo:hi], d_world[lo:hi], ranges=r_[lo:hi],
                                      geomid=g_[lo:hi])
                    pg.push(scene, "ranges", lo, r_[lo:hi])
                    pg.push(scene, "geomid", lo, g_[lo:hi])
                scene.copy_fence()

            # with no per-step join the last block's transfer is hidden, so fewer, larger blocks (less
            # launch-tail loss) win as long as the root's ingress keeps up: measure 2, 4 and the default
            stream_ms, ok_stream = {}, True
            for nb in sorted({2, 4, args.gather_blocks}):
                sblocks[0] = shard.scan_blocks(nscan, nb)
                for _ in range(3):
                    step_stream()
                scene.join()
                stream.synchronize()
                ctx["barrier"]()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    step_stream()
                scene.join()
                dist.barrier()
                e1.record(stream)
                stream.synchronize()
                torch.cuda.synchronize()
                ctx["barrier"]()
                stream_ms[nb] = ctx["reduce_max"](e0.elapsed_time(e1)) / args.steps
                v = verify()
                ok_stream = None if v is None else (ok_stream and v)
                stream.synchronize()
                ctx["barrier"]()
            best_nb = min(stream_ms, key=stream_ms.get)
            ms_stream = stream_ms[best_nb]
            del bufs
            gbytes = (world - 1) * nscan * nray * BYTES_PER_RAY
            out["p2p_copy_blocks"] = {
                "ms_per_step": ms, "value_with_gather": rays_job / (ms * 1e-3) / 1e9,
                "gather_ms": ms - step_ms_sharded, "blocks": len(blocks),
                "root_ingress_gb_per_step": gbytes / 1e9, "root_ingress_gbs": gbytes / (ms * 1e-3) / 1e9,
                "verified_checksums": ok,
                "how": "scan block k is cast into local memory, then the copy engine moves it into rank 0's peer "
                       "buffer (b2_scene_copy_async, bulk NVLink transfers, no SM) while block k+1 is cast"}
            out["p2p_copy_stream"] = {
                "ms_per_step": ms_stream, "value_with_gather": rays_job / (ms_stream * 1e-3) / 1e9,
                "gather_ms": ms_stream - step_ms_sharded, "blocks": best_nb,
                "ms_per_step_by_blocks": {str(k): v for k, v in stream_ms.items()},
                "root_ingress_gb_per_step": gbytes / 1e9, "root_ingress_gbs": gbytes / (ms_stream * 1e-3) / 1e9,
                "verified_checksums": ok_stream,
                "how": "p2p_copy_blocks as a streaming loop: results double-buffered on the rank, the transfers "
                       "of step s's last block overlap the first block of step s+1 (b2_scene_copy_fence); the "
                       "final step is joined inside the timed region"}
            stream.synchronize()
            ctx["barrier"]()
            pg.close()
        except Exception as e:
            out["p2p_copy_blocks"] = {"error": repr(e)}
        # ---- (b) all-gather of scan blocks overlapped with the cast ---------------------------------------
        try:
            comm = torch.cuda.Stream(device=dev)
            bg = shard.BlockGather({"ranges": ranges, "geomid": geomid}, nscan, nblock=args.gather_blocks,
                                   comm_stream=comm)

            def produce(k, lo, hi):
                lidar.cast_device(hi - lo, d_pos[lo:hi], d_mat[lo:hi], d_world[lo:hi], ranges=ranges[lo:hi],
                                  geomid=geomid[lo:hi])

            def step_nccl():
                scene.update_poses_device(d_xpos, d_xmat)
                bg.run(produce)

            ms = run(step_nccl, args.steps)
            got = bg.out["geomid"].assemble()
            ok = all(int(got[r].to(torch.int64).sum().item()) == int(allsums[r][0].item()) for r in range(world))
            # the collective alone, nothing overlapping it
            for _ in range(2):
                dist.all_gather_into_tensor(bg.out["ranges"].data, ranges.reshape(-1))
            torch.cuda.synchronize()
            ctx["barrier"]()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dist.all_gather_into_tensor(bg.out["ranges"].data, ranges.reshape(-1))
            dist.all_gather_into_tensor(bg.out["geomid"].data, geomid.reshape(-1))
            b.record(stream)
            stream.synchronize()
            alone = ctx["reduce_max"](a.elapsed_time(b))
            gbytes = (world - 1) * nscan * nray * BYTES_PER_RAY
            out["nccl_allgather_overlap"] = {
                "ms_per_step": ms, "value_with_gather": rays_job / (ms * 1e-3) / 1e9,
                "gather_ms": ms - step_ms_sharded, "blocks": len(bg.blocks),
                "collective_alone_ms": alone, "ingress_gbs_per_rank_alone": gbytes / (alone * 1e-3) / 1e9,
                "verified_checksums": bool(ok),
                "how": "torch.distributed (NCCL) all_gather_into_tensor of block k on a side stream while block "
                       "k+1 is cast; every rank receives all results"}
            del bg, got
        except Exception as e:
            out["nccl_allgather_overlap"] = {"error": repr(e)}
    good = {k: v for k, v in out.items() if "value_with_gather" in v}
    best = max(good, key=lambda k: good[k]["value_with_gather"]) if good else None
    return {"mode": best, "value_with_gather": good[best]["value_with_gather"] if best else None,
            "gather_ms": good[best]["gather_ms"] if best else None, "modes": out}


def host_path_gbs(ctx, nbytes: int = 1 << 30, reps: int = 3):
    """The box's device->host ceiling: every rank copies `nbytes` from its GPU into its own pinned
    buffer with plain cudaMemcpyAsync at the same time; aggregate GB/s over the slowest rank."""
    torch = ctx["torch"]
    dev, world = ctx["dev"], ctx["world"]
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        dst.copy_(src, non_blocking=True)
        st.synchronize()
        ctx["barrier"]()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        st.synchronize()
        dt = time.perf_counter() - t0
    dt = ctx["reduce_max"](dt)
    del src, dst
    return world * nbytes * reps / dt / 1e9


def latency_block(ctx):
    """The plugin's real duty cycle (ref:src/mujoco_lidar.cpp:228-242: ONE scan per publish): wall time
    of the blocking host call and the device time inside it, after the graph of the shape exists.
    cfg1 = the reference's fixture scene with a 360-ray planar lidar; one 64x2048 scan of the bench
    scene; one 640x480 ray-cast depth frame -> 32FC1 + cloud.  Next to each: the oracle on ONE thread
    (what the reference's physics thread pays inside mj_step for the same rays)."""
    engine, scenes = ctx["engine"], ctx["scenes"]
    from oracle import binding as orc
    out = {}

    def lidar_case(name, model, poses, spec, pos, mat, reps=200):
        sc = engine.Scene(model, nworld=1, device=ctx["dev"].index)
        sc.update_poses(poses)
        lid = engine.Lidar(sc, spec)
        bufs = {"ranges": np.empty((1, spec.nray), np.float32)}
        for _ in range(5):
            lid.cast(pos, mat, want=("ranges",), out=bufs)
        wall, devus = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            lid.cast(pos, mat, want=("ranges",), out=bufs)
            wall.append((time.perf_counter() - t0) * 1e6)
            devus.append(lid.last_device_us())
        oc = orc.OracleScene(model, poses)
        t0 = time.perf_counter()
        n_or = 3 if spec.nray > 10000 else 50
        for _ in range(n_or):
            oc.lidar_scan(spec, pos[0], mat[0], nthread=1)
        cpu_us = (time.perf_counter() - t0) / n_or * 1e6
        out[name] = {"rays": spec.nray, "host_call_us_p50": float(np.median(wall)),
                     "host_call_us_p99": float(np.percentile(wall, 99)), "device_us_p50": float(np.median(devus)),
                     "oracle_1thread_us": cpu_us, "output": "LaserScan ranges (f32), pageable destination"}
        lid.close()
        sc.close()

    rs = scenes.robot_scene(n_rays=360)
    lidar_case("cfg1_test_robot_360", rs.model, rs.poses, rs.lidar, rs.lidar_pos[None], rs.lidar_mat.reshape(1, 9))
    cs = ctx["cs"]
    p1, m1 = cs.sensor_poses(2)
    lidar_case("clutter500_64x2048_one_scan", cs.model, cs.poses, cs.lidar, p1[1:2], m1[1:2])
    # one 640x480 depth frame, ray-cast + linearise + cloud
    sc = engine.Scene(cs.model, nworld=1, device=ctx["dev"].index)
    sc.update_poses(cs.poses)
    cam = engine.DepthCamera(scenes.depth_camera(), scene=sc, device=ctx["dev"].index)
    R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
    cp = np.array([[3.0, 0.0, 1.5]])
    for _ in range(3):
        cam.render_to_cloud(cp, R, want=("depth", "cloud"))
    wall = []
    for _ in range(50):
        t0 = time.perf_counter()
        cam.render_to_cloud(cp, R, want=("depth", "cloud"))
        wall.append((time.perf_counter() - t0) * 1e6)
    out["depth_frame_640x480_raycast"] = {"pixels": 640 * 480, "host_call_us_p50": float(np.median(wall)),
                                          "host_call_us_p99": float(np.percentile(wall, 99)),
                                          "output": "32FC1 depth + xyz cloud, pageable destination"}
    cam.close()
    sc.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from mujoco_ros2_simulation_b200 import engine, scenes, shard

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

    def reduce_max(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    ms_total_max = reduce_max(ms_total)
    rays_per_step = nscan * nray
    value = job_rays_per_s(world, rays_per_step, args.steps, ms_total, reduce_max) / 1e9

    # sanity on the result of the last step (not timed): hit fraction of this scene is ~0.99
    hit_frac = float((geomid[0] >= 0).float().mean().item())

    ctx = {"torch": torch, "dist": dist, "engine": engine, "scenes": scenes, "shard": shard, "dev": dev,
           "stream": stream, "rank": rank, "world": world, "args": args, "barrier": barrier,
           "reduce_max": reduce_max, "scene": scene, "lidar": lidar, "nscan": nscan, "nray": nray, "cs": cs,
           "d_pos": d_pos, "d_mat": d_mat, "d_world": d_world, "d_xpos": d_xpos, "d_xmat": d_xmat,
           "ranges": ranges, "geomid": geomid}

    # ---- N > 1: the same step with the results gathered over NVLink ---------------------------------
    gather = None
    if world > 1 and not args.no_gather:
        gather = measure_gather(ctx, ms_total_max / args.steps)

    # ---- e2e: host buffers through the blocking C-ABI call, copies inside the timed region ------
    es = min(args.e2e_scans, nscan)
    with NumaNear(local_rank) as near:  # pinned buffers on the GPU's own NUMA node (when the VM shows one)
        h_pos = torch.from_numpy(pos[:es].copy()).pin_memory()
        h_mat = torch.from_numpy(mat[:es].copy()).pin_memory()
        h_world = torch.zeros(es, dtype=torch.int32).pin_memory()
        h_ranges = torch.empty((es, nray), dtype=torch.float32).pin_memory()
        h_geomid = torch.empty((es, nray), dtype=torch.int32).pin_memory()
        h_geomid16 = torch.empty((es, nray), dtype=torch.uint16).pin_memory()
        h_xpos = torch.from_numpy(cs.poses.geom_xpos.copy()).pin_memory()
        h_xmat = torch.from_numpy(cs.poses.geom_xmat.copy()).pin_memory()
        h_ranges.zero_()
        h_geomid.zero_()
        h_geomid16.zero_()
    lib = scene.lib

    def e2e_run(call):
        def one():
            engine._capi.check(lib.b2_scene_update_poses(scene._h, 0, 1, h_xpos.data_ptr(), h_xmat.data_ptr()))
            call()
        for _ in range(2):
            one()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one()
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        return world * es * nray * args.steps / dt / 1e9

    # what MujocoLidar::update publishes: the LaserScan ranges, 4 B/ray (ref:src/mujoco_lidar.cpp:256-279)
    e2e_scan_val = e2e_run(lambda: engine._capi.check(lib.b2_lidar_cast(
        lidar._h, es, h_world.data_ptr(), h_pos.data_ptr(), h_mat.data_ptr(), h_ranges.data_ptr(), None, None, None)))
    # + the id of the geom hit, as int32 (8 B/ray) and as uint16 (6 B/ray)
    e2e_id_val = e2e_run(lambda: engine._capi.check(lib.b2_lidar_cast(
        lidar._h, es, h_world.data_ptr(), h_pos.data_ptr(), h_mat.data_ptr(), h_ranges.data_ptr(),
        h_geomid.data_ptr(), None, None)))
    e2e_ok = bool((h_geomid[0].numpy() == geomid[0].cpu().numpy()).all()) if es > 0 else True
    e2e_u16_val = e2e_run(lambda: engine._capi.check(lib.b2_lidar_cast_u16(
        lidar._h, es, h_world.data_ptr(), h_pos.data_ptr(), h_mat.data_ptr(), h_ranges.data_ptr(),
        h_geomid16.data_ptr())))
    e2e_ok = e2e_ok and bool((h_geomid16[0].numpy().astype(np.int32) ==
                              np.where(h_geomid[0].numpy() < 0, 0xFFFF, h_geomid[0].numpy())).all())
    h2d = es * (3 + 9) * 8 + es * 4 + int(h_xpos.numel() + h_xmat.numel()) * 8
    host_gbs = host_path_gbs(ctx)
    del h_ranges, h_geomid, h_geomid16

    # ---- single-scan latency: the plugin's duty cycle -----------------------------------------------------
    latency = None
    if rank == 0 and not args.no_latency:
        try:
            latency = latency_block(ctx)
        except Exception as e:
            latency = {"error": repr(e)}

    # ---- secondary: K3 depth -> cloud (the genuinely HBM-bound kernel), cfg3, cfg4 -------------------------
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
    barrier()

    # ---- cfg5 on every rank (BASELINE.json configs[4]) ----------------------------------------------------------
    cfg5_res = None
    if not args.no_secondary and args.cfg5_worlds > 0:
        torch.cuda.empty_cache()
        try:
            cfg5_res = cfg5(ctx, args.cfg5_worlds)
        except Exception as e:
            cfg5_res = {"error": repr(e)}
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (K2 cast) ------------------------------------------------------
    st = scene.stats()
    # GeomStatic 32 + GeomPose 96 per geom, 32 per BVH child slot, 128 per fanout-4 mesh node, 48 per triangle
    scene_bytes = cs.model.ngeom * (32 + 96) + st["bvh_nodes"] * 32 + st["mesh_nodes"] * 128 + st["mesh_tris"] * 48
    alg_bytes = BYTES_PER_RAY * rays_per_step + scene_bytes
    peak, how = measured_peak()
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # DRAM bytes of the WHOLE cast (K2a + K2b + K2c + K2d/e) per ray from the committed ncu capture of
            # this build (tools/ncu_traffic.py regenerates the file), scaled to this launch's rays
            traffic = int(tj["cast_dram_bytes_per_ray"] * rays_per_step)
            traffic_src = tj.get("source")
        except Exception:
            traffic = None

    # ---- CPU baseline (rank 0, N=1): the oracle on a bounded sample of the same workload -------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding as orc
        orc.build()
        cores = os.cpu_count() or 1
        oc = orc.OracleScene(cs.model, cs.poses)
        variants = cpu_variants(cs, cores, 4.0, oc)
        best = max(variants, key=lambda k: variants[k]["value"])
        rays, secs = cpu_arm(cs, args.cpu_baseline_scans, cores, 50, best, oc)
        cpu = {"value": rays / secs / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_baseline_scans} of the {nscan} scans ({rays} rays), fastest CPU variant "
                         f"({best}) of the oracle (C restatement, -O3) on all host cores",
               "variant": best, "variants": variants}
        # the brute-force loop on one thread (what the reference's physics thread has)
        r1, s1 = cpu_arm(cs, 2, 1, 90, "mj_ray_loop", oc)
        cpu["single_thread"] = {"value": r1 / s1 / 1e9, "unit": UNIT, "variant": "mj_ray_loop",
                                "sample": f"2 scans ({r1} rays)"}
        # parity of THIS run's GPU output against the oracle on sampled scans of the timed batch
        g_host, r_host = geomid.cpu().numpy(), ranges.cpu().numpy()
        checked = bad_id = bad_rng = 0
        for sc_i in sorted({0, nscan // 3, nscan - 1}):
            rr, gg, _ = oc.lidar_scan(cs.lidar, pos[sc_i], mat[sc_i], nthread=cores)
            checked += rr.size
            bad_id += int((gg != g_host[sc_i]).sum())
            bad_rng += int((rr != r_host[sc_i]).sum())
        parity = {"rays_checked": checked, "geomid_mismatch": bad_id, "range_bits_mismatch": bad_rng,
                  "against": "CPU oracle (oracle/), same poses; ranges compared bit for bit"}

    if rank == 0:
        if secondary is not None and cfg5_res is not None:
            secondary["cfg5_clutter_worlds"] = cfg5_res
        elif cfg5_res is not None:
            secondary = {"cfg5_clutter_worlds": cfg5_res}
        if secondary is not None and latency is not None and "error" not in latency:
            # the same numbers under the key the review asked for
            secondary["latency_us"] = {k: {"host_call_p50": v["host_call_us_p50"], "device_p50": v.get("device_us_p50"),
                                           "oracle_1thread": v.get("oracle_1thread_us")} for k, v in latency.items()}
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
                "parallelism": f"replica x{world}, sensor-pose batches sharded; value = results left on the ranks, "
                               "gather = results collected over NVLink",
                "gather_ms": gather["gather_ms"] if gather else None,
                "hit_fraction": hit_frac, "e2e_matches_device_path": e2e_ok,
                "bvh_nodes": st["bvh_nodes"], "mesh_nodes": st["mesh_nodes"],
            },
            "value_with_gather": gather["value_with_gather"] if gather else None,
            "gather": gather,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "K2 cast = collect_kernel -> drain_kernel -> mesh-pair kernels (one "
                                   "b2_lidar_cast_device; drain_kernel is the largest at ~50 %)",
                         "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": how,
                         "note": "ray casting is traversal/FP64-issue bound, not HBM bound (SURVEY.md M3)"},
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": {"value": e2e_scan_val, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": es * nray * 4, "scans_per_step": es, "pinned_numa_node": near.node,
                    "outputs": "LaserScan ranges f32 (4 B/ray): what MujocoLidar::update publishes",
                    "host_path_gbs": host_gbs,
                    "frac_of_host_path": e2e_scan_val * 4 / host_gbs if host_gbs else None,
                    "with_geomid_i32": {"value": e2e_id_val, "unit": UNIT, "d2h_bytes_per_step": es * nray * 8,
                                        "frac_of_host_path": e2e_id_val * 8 / host_gbs if host_gbs else None},
                    "with_geomid_u16": {"value": e2e_u16_val, "unit": UNIT, "d2h_bytes_per_step": es * nray * 6,
                                        "frac_of_host_path": e2e_u16_val * 6 / host_gbs if host_gbs else None},
                    "note": "host_path_gbs = every rank's plain pinned D2H at once (aggregate): the ceiling of "
                            "this box for results that must reach host memory"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "latency": latency,
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
