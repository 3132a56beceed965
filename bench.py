#!/usr/bin/env python3
"""bench.py - depth frames/s into a 512^3 TSDF (BASELINE.json metric) on N B200s.

A "step" = one pass of the hot path over 1024 synthetic 640x480 depth frames (16 launches of 64 frames each:
bucket kernel -> ray-march kernels -> commit kernel = all of dense_tsdf.py:162-270), state fully materialised at
the end of every launch.  Every frame of the stream has its own depth image (sphere radius varies with the frame
index) and its own pose.

  value        frames/s with the depth frames already resident in HBM (whole job, all ranks)
  e2e          frames/s through the public Python API (taichislam_b200.mapping.DenseTSDF.recast_depth_to_map,
               one call per frame) from pinned HOST frames in the DEFAULT frame mode (the frame is copied - and the
               copy awaited - inside the call, the reference's semantics); H2D copy of every frame and a D2H read of
               the step's counters inside the timed region.  extras.e2e_modes adds the opt-in borrowing mode (the GPU
               reads the sampled rows of page-locked frames itself) and pageable frames.
  roofline     dominant kernel k_march_blocks: 9 B per voxel update / its CUDA-event duration vs the measured HBM
               peak; process_new_pcl_all_kernels states the same for ALL kernels of the ray march (17 B/ray + 9 B/update)
  cpu_baseline the CPU oracle (restatement of the reference's Taichi kernels; Taichi itself is
               not installable here) on all host cores, bounded sample, rank 0 only

`--impl reference` times that CPU restatement alone (the reference's own implementation cannot
run: it needs the taichi package).
"""
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

METRIC = "depth_frames_per_sec_into_512^3_tsdf"
UNIT = "frames/s"
BATCH = 64              # frames per integrate launch (= TSLAM_MAX_BATCH)
LAUNCHES_PER_STEP = 16  # one step = 1024 frames, ~10 ms of GPU work
STEP_FRAMES = BATCH * LAUNCHES_PER_STEP
POOL_BATCHES = 4        # distinct input batches cycled through: 4*64*614 KB = 157 MB > 126 MB L2
WORKLOAD = ("C2: synthetic 640x480 uint16 depth stream (S2: camera-centred sphere, R = 4 m +- 0.2 m varying with the frame "
            "index, circle poses) -> 512^3 TSDF, voxel 0.05 m, recast_step 2, max_ray 10 m")
CONFIG = {"workload": WORKLOAD, "frames_per_step": STEP_FRAMES, "frames_per_launch": BATCH, "grid": "512^3"}
MAP_SCALE = [25.6, 25.6]  # 512^3 voxels of 0.05 m (C2)


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this same command (profiles/)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for fn in sorted(os.listdir(pdir)):
            if fn.endswith("_traffic.json"):
                try:
                    d = json.load(open(os.path.join(pdir, fn)))
                    for key, val in d.items():  # exact name, or the template instance ("void k_raymarch<0>")
                        if key == kernel or (kernel + "<") in key:
                            best = (float(val), fn)
                except Exception:
                    pass
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class NvmlSampler:
    """SM clock and throttle reasons read through NVML every 2 ms DURING the timed region (nvidia-smi's fastest loop,
    100 ms, is longer than a short timed region)."""

    def __init__(self, cuda_index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
        self.h = h
        self.sm, self.reasons, self.run = [], 0, False
        self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def start(self):
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        time.sleep(0.01)

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self.run = False
        self.t.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if self.reasons & int(bit))
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                "reasons": reasons, "source": "nvml, 2 ms period"}


def make_sampler(cuda_index):
    try:
        return NvmlSampler(cuda_index)
    except Exception:
        return ClockSampler(cuda_index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_SPHERES = {}


def frame_depth(t):
    """Depth image of stream frame t: the S2 sphere with a radius that depends on t (11 distinct images)."""
    from taichislam_b200 import synthetic as syn
    k = ((t % 256) * 7) % 11
    if k not in _SPHERES:
        _SPHERES[k] = syn.scene_sphere(4.0 + 0.04 * (k - 5))
    return _SPHERES[k]


def make_inputs(n_frames, start):
    from taichislam_b200 import synthetic as syn
    depth = np.stack([frame_depth(start + q) for q in range(n_frames)]) if n_frames else np.zeros((0, syn.H, syn.W), np.uint16)
    Rs, Ts = syn.stream_poses(n_frames, start=start)
    return depth, Rs, Ts


def cpu_baseline(target_seconds=12.0):
    """Oracle (CPU restatement of dense_tsdf.py:157-270) on all host cores, bounded sample of the same stream."""
    from oracle.oracle import OracleTSDF, integrate_stream_mt
    from taichislam_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    maps = [OracleTSDF(map_scale=MAP_SCALE, K=syn.K_DEPTH, is_global_map=True) for _ in range(cores)]
    d, Rs, Ts = make_inputs(cores, 0)
    d = np.ascontiguousarray(d)
    t0 = time.perf_counter()
    integrate_stream_mt(maps, d, Rs, Ts)            # calibration (also warms the maps)
    t1 = time.perf_counter() - t0
    reps = int(min(64, max(1, target_seconds / max(t1, 1e-3))))
    n = cores * reps
    d, Rs, Ts = make_inputs(n, cores)
    d = np.ascontiguousarray(d)
    t0 = time.perf_counter()
    integrate_stream_mt(maps, d, Rs, Ts)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} frames of the same stream into {cores} independent 512^3 maps, "
                      f"one host thread each, {dt:.1f} s; oracle = CPU restatement of the reference's Taichi kernels "
                      f"(taichi not installable)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle.oracle import OracleTSDF, integrate_stream_mt
    from taichislam_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    maps = [OracleTSDF(map_scale=MAP_SCALE, K=syn.K_DEPTH, is_global_map=True) for _ in range(cores)]
    # bounded sample per step: one or two frames per host thread, so that K steps stay within a few minutes
    # (~0.3 s per 128-frame step on 128 cores)
    per_step = cores if args.steps > 50 else 2 * cores
    t = 0
    for _ in range(args.warmup):
        d, Rs, Ts = make_inputs(per_step, t)
        integrate_stream_mt(maps, np.ascontiguousarray(d), Rs, Ts)
        t += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d, Rs, Ts = make_inputs(per_step, t)
        integrate_stream_mt(maps, np.ascontiguousarray(d), Rs, Ts)
        t += per_step
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f"{per_step} frames per step (bounded sample of the {STEP_FRAMES}-frame step) of the same stream into {cores} independent "
              f"512^3 maps, {cores} host threads; CPU restatement of the reference (taichi cannot be installed: no wheel, no network)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def frame_depth_mod(t):
    return frame_depth(t % 256)


def c5_extras(rank, local_rank, world, dist, torch):
    """BASELINE config 5 flavour on the bench's ranks: 64 submaps (10 frames each) integrated submap-sharded, fused into
    a 2048^3 x 0.05 m global volume TILED over the ranks (data-driven tiling, foreign blocks in one NCCL all-to-all,
    one-block halo, local marching cubes, mesh-vertex all-gather).  Everything is run twice; the second (warm) pass is
    timed on the device (CUDA events, max over ranks)."""
    from taichi_slam.mapping import DenseTSDF, MarchingCubeMesher
    from taichislam_b200 import synthetic as syn
    from taichislam_b200.distributed import TiledGlobalMap, submap_owner
    n_sub, frames_per = 64, 10
    sub = DenseTSDF(map_scale=[25.6, 25.6], voxel_scale=0.05, max_submap_num=64, max_disp_particles=1024, max_blocks=40000)
    sub.set_dep_camera_intrinsic(syn.K_DEPTH)
    glo = DenseTSDF(map_scale=[102.4, 102.4], voxel_scale=0.05, is_global_map=True, max_submap_num=64, max_disp_particles=1024,
                    max_blocks=60000)
    mine = [s_ for s_ in range(n_sub) if submap_owner(s_, world) == rank]
    for s_ in range(n_sub):
        gx, gy = s_ % 8, s_ // 8
        a = 0.3 * s_
        Rb = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        glo.set_base_pose_submap(s_, Rb, np.array([6.0 * (gx - 3.5), 6.0 * (gy - 3.5), 0.0]))
    d = syn.scene_sphere(4.0)
    for s_ in mine:
        sub.active_submap_id[None] = s_
        sub.set_base_pose_submap(s_, np.eye(3), np.zeros(3))
        for q in range(frames_per):
            R, T = syn.stream_pose(q * 7)
            sub.recast_depth_to_map(R, T, d, np.array([]))
    sub._flush()
    tiled = TiledGlobalMap(glo, dist, rank, world)
    mesher = MarchingCubeMesher(glo, 3000000, tsdf_surface_thres=0.25)

    def timed(fn):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    tiled.fuse_submaps_tiled(sub)                      # cold: NCCL channels, allocations
    tiled.generate_mesh_all_gather(mesher)
    fuse_ms, _ = timed(lambda: tiled.fuse_submaps_tiled(sub))
    x = dict(tiled.last_exchange)
    mesh_ms, (mv, mn, counts) = timed(lambda: tiled.generate_mesh_all_gather(mesher))
    x.update(tiled.last_exchange)
    tot = torch.tensor([float(glo.count_active()), float(x["fusion_blocks_sent"]), float(x.get("halo_blocks_sent", 0)),
                        float(x["fusion_bytes_sent"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    return {"config": "C5: 64 submaps -> 2048^3 x 0.05 m global volume tiled over the ranks", "tiles": list(tiled.tiles),
            "cuts": [list(map(int, c)) for c in tiled.cuts] if tiled.cuts else None,
            "fuse_ms": fuse_ms, "mesh_ms": mesh_ms, "global_voxels": int(tot[0].item()),
            "fusion_blocks_exchanged": int(tot[1].item()), "exchanged_MB": tot[3].item() / 1e6,
            "halo_blocks_exchanged": int(tot[2].item()), "triangles": int(sum(counts)), "triangles_per_rank": [int(c) for c in counts],
            "timing": "warm second pass, CUDA events, max over ranks; mesh_ms = halo exchange + local marching cubes + mesh all-gather"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from taichislam_b200 import synthetic as syn
    from taichislam_b200.tsdf_handle import TsdfHandle

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, Nz = 512, 512
    t_rank = rank * 102400  # every rank integrates its own stretch of the stream into its own map (submap sharding); multiple of 256
    # ---- device-resident arm ------------------------------------------------------------------
    g = TsdfHandle(N, Nz, K=syn.K_DEPTH, is_global_map=True)
    pool = []
    for b in range(POOL_BATCHES):  # frame t of the stream uses image t mod 256: the pool holds exactly those 256 frames
        d, _, _ = make_inputs(BATCH, t_rank + b * BATCH)
        pool.append(torch.from_numpy(np.ascontiguousarray(d).view(np.int16)).cuda())
    n_launches = (args.warmup + args.steps) * LAUNCHES_PER_STEP
    Rs_all, Ts_all = syn.stream_poses(n_launches * BATCH, start=t_rank)   # host pose stream, prepared up front
    Rs_all = np.ascontiguousarray(Rs_all.astype(np.float32).reshape(n_launches, BATCH, 9))
    Ts_all = np.ascontiguousarray(Ts_all.astype(np.float32).reshape(n_launches, BATCH, 3))
    launch_no = [0]

    def step_dev():
        for _ in range(LAUNCHES_PER_STEP):
            k = launch_no[0]
            launch_no[0] += 1
            g.integrate_depth(pool[k % POOL_BATCHES], Rs_all[k], Ts_all[k], commit=True)

    for i in range(args.warmup):
        step_dev()
    g.stats(clear=True)
    g.set_profiling(True)
    l0 = g.launch_count()
    sampler = make_sampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    st = g.stats()
    kms = g.kernel_ms2(min(args.steps * LAUNCHES_PER_STEP, 512))
    mstats = g.march_stats()
    g.sync()
    tmax = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total_max = float(tmax.item())
    value = world * args.steps * STEP_FRAMES / (ms_total_max * 1e-3)

    # roofline of process_new_pcl (dense_tsdf.py:236-270), SURVEY 8d algorithmic bytes 17*rays + 9*updates per launch,
    # over the CUDA-event time of ALL its kernels (ray set-up + segment walk, scan, placement, block march); the
    # dominant one (k_march_blocks) is listed beside it
    peak, peak_src = load_peaks()
    km = kms.mean(axis=0) if len(kms) else np.full(7, np.nan)
    bucket_ms, ray_ms, commit_ms = float(km[0]), float(km[1]), float(km[2])
    n_l = max(args.steps * LAUNCHES_PER_STEP, 1)
    rays_per_launch = st["n_rays"] / n_l
    upd_per_launch = st["n_updates"] / n_l
    bytes_launch = 17.0 * rays_per_launch + 9.0 * upd_per_launch
    achieved = bytes_launch / (ray_ms * 1e-3) / 1e9 if ray_ms == ray_ms and ray_ms > 0 else None
    march_only = float(km[6])
    frame_bytes = (2.0 * st["n_px"] + 24.0 * st["n_valid"] + 17.0 * st["n_rays"] + 9.0 * st["n_updates"]) / max(n_l * BATCH, 1)
    march_ach = 9.0 * upd_per_launch / (march_only * 1e-3) / 1e9 if march_only == march_only and march_only > 0 else None
    tr = load_traffic("k_march_blocks") or (None, None)

    # ---- marching cubes of the resulting map (C2 "+ marching cubes"), timed outside the frame metric ----
    mc = {}
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ntri, _, _ = g.marching_cubes(1, 0.25, cap_tri=1 << 22)  # cold: first launch of the three MC kernels, output buffers allocated
        t1 = time.perf_counter()
        ntri, _, _ = g.marching_cubes(1, 0.25, cap_tri=1 << 22)
        mc = {"triangles": int(ntri), "ms_incl_d2h_of_mesh": 1e3 * (time.perf_counter() - t1), "first_call_ms": 1e3 * (t1 - t0)}
    except Exception as ex:  # pragma: no cover
        mc = {"error": str(ex)}
    g.close()

    # ---- end-to-end arm: public API, host frames, H2D + D2H inside the timed region ----------------
    e2e, e2e_modes = None, {}
    if not args.no_e2e:
        from taichislam_b200.mapping import DenseTSDF
        empty_tex = np.array([])
        host_t = torch.from_numpy(np.ascontiguousarray(make_inputs(BATCH, t_rank)[0]).view(np.int16))

        def run_e2e(mode, steps):
            if mode == "dma_pageable":
                os.environ["TSLAM_FRAME_COPY"] = "dma"  # read when the handle is created
            if mode == "ring_pinned":
                os.environ["TSLAM_PINNED_COPY"] = "ring"
            if mode == "fetch_pinned":
                os.environ["TSLAM_PINNED_COPY"] = "fetch"
            try:
                m = DenseTSDF(map_scale=MAP_SCALE, voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True)
            finally:
                os.environ.pop("TSLAM_FRAME_COPY", None)
                os.environ.pop("TSLAM_PINNED_COPY", None)
            return run_e2e_on(m, mode, steps)

        def run_e2e_on(m, mode, steps):
            m.set_dep_camera_intrinsic(syn.K_DEPTH)
            m.set_base_pose_submap(0, np.eye(3), np.zeros(3))  # pose-table rows start at zero (mapping_common.py:106-107)
            if mode == "borrow":
                m.set_frame_borrowing(True)
            if mode == "commit1":
                m.set_commit_granularity(1)
            host = host_t.clone() if mode in ("pageable", "dma_pageable") else host_t.pin_memory()
            host_np = host.numpy().view(np.uint16)
            t = t_rank + 51200
            eRs, eTs = syn.stream_poses((args.warmup + steps) * STEP_FRAMES, start=t)
            ek = [0]

            def step():
                k0 = ek[0]
                ek[0] += STEP_FRAMES
                for q in range(STEP_FRAMES):
                    m.recast_depth_to_map(eRs[k0 + q], eTs[k0 + q], host_np[q % BATCH], empty_tex)
                return m.frame_counters()  # flushes the queue, commits, D2H read of the integrate counters

            for i in range(args.warmup):
                step()
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                res = step()
            barrier()
            dt = time.perf_counter() - t0
            assert res["n_updates"] > 0, "e2e arm integrated nothing"
            tm = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            del m
            return world * steps * STEP_FRAMES / float(tm.item()), int(res["d2h_bytes"])

        try:
            es = max(3, args.steps // 5)
            v, d2h = run_e2e("copy", es)
            e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * syn.H * syn.W * 2), "d2h_bytes_per_step": d2h, "steps": es,
                   "api": "DenseTSDF.recast_depth_to_map per frame; page-locked host uint16 frames, DEFAULT frame mode: every frame is "
                          "DMA-copied and the copy awaited inside the call (the caller may reuse its buffer at once, as with the "
                          "reference); the method's pose arithmetic runs while the copy is in flight"}
            es2 = max(3, args.steps // 10)
            vb, _ = run_e2e("borrow", es2)
            vp, _ = run_e2e("pageable", es2)
            vc, _ = run_e2e("commit1", es2)
            vd, _ = run_e2e("dma_pageable", es2)
            vq, _ = run_e2e("ring_pinned", es2)
            vf, _ = run_e2e("fetch_pinned", es2)
            e2e_modes = {"borrowed_pinned_frames": {"value": vb, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * (syn.H // 2) * syn.W * 2),
                                                    "note": "opt-in set_frame_borrowing(True): no copy, the GPU reads the sampled rows from host "
                                                            "memory; frames must stay untouched until the next flush"},
                         "commit_every_frame": {"value": vc, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * syn.H * syn.W * 2),
                                                "note": "set_commit_granularity(1): one launch sequence + commit per frame (Wmax clamp granule = frame)"},
                         "pageable_frames": {"value": vp, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * (syn.H // 2) * syn.W * 2),
                                             "note": "what np.frombuffer(depth_msg.data) gives the ROS node (taichislam_node.py:381-382): the sampled "
                                                     "rows are copied into the library's page-locked ring (streaming stores) and fetched from there by the GPU"},
                         "pinned_frames_awaited_row_fetch": {"value": vf, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * (syn.H // 2) * syn.W * 2),
                                                             "note": "TSLAM_PINNED_COPY=fetch: the GPU fetches the sampled rows of the caller's page-locked frame inside the call, awaited"},
                         "pinned_frames_through_the_ring": {"value": vq, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * (syn.H // 2) * syn.W * 2),
                                                            "note": "TSLAM_PINNED_COPY=ring: page-locked frames take the pageable frames' path (sampled rows memcpy'd into the ring)"},
                         "pageable_frames_through_cudaMemcpyAsync": {"value": vd, "unit": UNIT, "h2d_bytes_per_step": int(STEP_FRAMES * syn.H * syn.W * 2),
                                                                     "note": "TSLAM_FRAME_COPY=dma: the runtime stages the whole pageable frame (A/B for the ring)"}}
        except Exception as ex:  # pragma: no cover
            e2e = e2e or {"value": None, "unit": UNIT, "error": repr(ex)}
            e2e_modes["error"] = repr(ex)

    c5 = None
    if world > 1 and not args.no_c5:
        try:
            c5 = c5_extras(rank, local_rank, world, dist, torch)
        except Exception as ex:  # pragma: no cover
            c5 = {"error": repr(ex)}

    cfg = dict(CONFIG)
    cfg.update({"parallelism": f"submap-sharded x{world} (no data-path collective in the frame metric; the tiled global-map "
                               f"exchange is measured in extras.c5 for N > 1)",
                "l2": f"inputs cycle through {POOL_BATCHES} batches = {POOL_BATCHES * BATCH * syn.H * syn.W * 2 / 1e6:.0f} MB > 126 MB L2"})
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        # dominant kernel (k_march_blocks: the voxel updates, SURVEY 8d 9 B per update) per the bench contract; the
        # whole process_new_pcl (ray set-up + segment build + march; 17 B per ray + 9 B per update over ALL their time)
        # is stated beside it - that one is what the frame rate follows
        "roofline": {"bound": "hbm", "kernel": "k_march_blocks",
                     "achieved": march_ach, "peak": peak, "unit": "GB/s",
                     "frac": (march_ach / peak) if march_ach else None,
                     "traffic": tr[0], "traffic_source": tr[1], "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": 9.0 * upd_per_launch,
                     "kernel_ms": {"bucket": bucket_ms, "raymarch": ray_ms, "commit": commit_ms,
                                   "setup_walk_class_scan": float(km[3] + km[4]), "seg_place": float(km[5]), "march_blocks": march_only},
                     "process_new_pcl_all_kernels": {"kernels": "k_seg_walk (ray set-up + segment walk) + k_seg_class/scan/place + k_march_blocks",
                                                     "ms": ray_ms, "algorithmic_bytes_per_launch": bytes_launch,
                                                     "achieved": achieved, "frac": (achieved / peak) if achieved else None},
                     "algorithmic_bytes_per_frame_all_kernels": frame_bytes},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "extras": {"marching_cubes": mc, "per_frame": {"rays": st["n_rays"] / max(n_l * BATCH, 1),
                                                         "voxel_updates": st["n_updates"] / max(n_l * BATCH, 1)},
                   "voxel_blocks": st["n_blocks"], "march": {k: v / n_l for k, v in mstats.items()}, "e2e_modes": e2e_modes, "c5": c5},
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline()
        except Exception as ex:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "error": repr(ex)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
