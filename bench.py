#!/usr/bin/env python3
"""bench.py - depth frames/s into a 512^3 TSDF (BASELINE.json metric) on N B200s.

A "step" = one pass of the hot path over one batch of 64 synthetic 640x480 depth frames:
bucket kernel -> ray-march kernel -> commit kernel (all of dense_tsdf.py:162-270 for the
batch), state fully materialised at the end of every step.

  value        frames/s with the depth frames already resident in HBM (whole job, all ranks)
  e2e          frames/s through the public Python API (taichislam_b200.mapping.DenseTSDF) from
               pinned HOST frames: H2D copy of every frame and a D2H read of the step's counters
               inside the timed region
  roofline     ray-march kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak
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
BATCH = 64              # frames per step (= TSLAM_MAX_BATCH: one launch triple per step)
POOL_BATCHES = 4        # distinct input batches cycled through: 4*64*614 KB = 157 MB > 126 MB L2
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
                        if key == kernel or (kernel + "<0>") in key:
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


def make_inputs(n_frames, start):
    from taichislam_b200 import synthetic as syn
    depth = np.broadcast_to(syn.scene_sphere(4.0), (n_frames, syn.H, syn.W))
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
            "sample": f"{n} frames of the S2 stream (sphere R=4 m, circle poses) into {cores} independent 512^3 maps, "
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
    sample = (f"{per_step} frames/step of the S2 stream into {cores} independent 512^3 maps, {cores} host threads; "
              "CPU restatement of the reference (taichi cannot be installed: no wheel, no network)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: synthetic 640x480 depth stream (S2 sphere R=4 m) -> 512^3 TSDF", "frames_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    from oracle.oracle import tsdf_dims  # noqa: F401  (dims helper only; the oracle is not on the measured path)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, Nz = 512, 512
    # ---- device-resident arm ------------------------------------------------------------------
    g = TsdfHandle(N, Nz, K=syn.K_DEPTH, is_global_map=True)
    pool = []
    for b in range(POOL_BATCHES):
        d, _, _ = make_inputs(BATCH, 0)
        pool.append(torch.from_numpy(np.ascontiguousarray(d).view(np.int16)).cuda())
    frame_t = rank * 100000  # every rank integrates its own stretch of the stream into its own map (submap sharding)

    n_steps_total = args.warmup + args.steps
    Rs_all, Ts_all = syn.stream_poses(n_steps_total * BATCH, start=frame_t)   # host pose stream, prepared up front
    Rs_all = np.ascontiguousarray(Rs_all.astype(np.float32).reshape(n_steps_total, BATCH, 9))
    Ts_all = np.ascontiguousarray(Ts_all.astype(np.float32).reshape(n_steps_total, BATCH, 3))
    step_no = [0]

    def step_dev(i, t):
        k = step_no[0]
        step_no[0] += 1
        g.integrate_depth(pool[i % POOL_BATCHES], Rs_all[k], Ts_all[k], commit=True)

    for i in range(args.warmup):
        step_dev(i, frame_t)
        frame_t += BATCH
    g.stats(clear=True)
    g.set_profiling(True)
    l0 = g.launch_count()
    sampler = make_sampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev(i, frame_t)
        frame_t += BATCH
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    st = g.stats()
    kms = g.kernel_ms(min(args.steps, 512))
    g.sync()
    tmax = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total_max = float(tmax.item())
    value = world * args.steps * BATCH / (ms_total_max * 1e-3)

    # roofline of the dominant kernel (ray march): SURVEY 8d algorithmic bytes 17*rays + 9*updates per launch
    peak, peak_src = load_peaks()
    ray_ms = float(kms[:, 1].mean()) if len(kms) else float("nan")
    bucket_ms = float(kms[:, 0].mean()) if len(kms) else float("nan")
    commit_ms = float(kms[:, 2].mean()) if len(kms) else float("nan")
    rays_per_launch = st["n_rays"] / max(args.steps, 1)
    upd_per_launch = st["n_updates"] / max(args.steps, 1)
    bytes_launch = 17.0 * rays_per_launch + 9.0 * upd_per_launch
    achieved = bytes_launch / (ray_ms * 1e-3) / 1e9 if ray_ms == ray_ms and ray_ms > 0 else None
    frame_bytes = (2.0 * st["n_px"] + 24.0 * st["n_valid"] + 17.0 * st["n_rays"] + 9.0 * st["n_updates"]) / max(args.steps * BATCH, 1)

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

    # ---- end-to-end arm: public API, pinned host frames, H2D + D2H inside the timed region --------
    e2e = None
    if not args.no_e2e:
        try:
            from taichislam_b200.mapping import DenseTSDF
            m = DenseTSDF(map_scale=MAP_SCALE, voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True)
            m.set_dep_camera_intrinsic(syn.K_DEPTH)
            m.set_base_pose_submap(0, np.eye(3), np.zeros(3))  # pose-table rows start at zero (mapping_common.py:106-107)
            host = torch.from_numpy(np.ascontiguousarray(make_inputs(BATCH, 0)[0]).view(np.int16)).pin_memory()
            host_np = host.numpy().view(np.uint16)
            empty_tex = np.array([])
            zero_copy = os.environ.get("TSLAM_ZERO_COPY", "1") != "0"
            t = rank * 100000 + 50000
            eRs, eTs = syn.stream_poses(n_steps_total * BATCH, start=t)   # host pose stream prepared up front
            ek = [0]

            def step_e2e():
                k0 = ek[0]
                ek[0] += BATCH
                for q in range(BATCH):
                    m.recast_depth_to_map(eRs[k0 + q], eTs[k0 + q], host_np[q], empty_tex)
                return m.frame_counters()  # flushes the queue, commits, D2H read of the integrate counters

            for i in range(args.warmup):
                step_e2e()
            m.frame_counters()
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                res = step_e2e()
            barrier()
            dt = time.perf_counter() - t0
            assert res["n_updates"] > 0, "e2e arm integrated nothing"
            tm = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e = {"value": world * args.steps * BATCH / float(tm.item()), "unit": UNIT,
                   # page-locked frames: the GPU fetches the sampled rows (every recast_step-th) over PCIe itself
                   "h2d_bytes_per_step": int(BATCH * (syn.H // 2) * syn.W * 2) if zero_copy else int(BATCH * syn.H * syn.W * 2),
                   "d2h_bytes_per_step": int(res["d2h_bytes"]),
                   "api": "DenseTSDF.recast_depth_to_map per frame (pinned host uint16 frames" +
                          (", sampled rows read by the GPU from host memory)" if zero_copy else ", whole-frame DMA copies)")}
        except Exception as ex:  # pragma: no cover
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: synthetic 640x480 uint16 depth stream (S2: camera-centred sphere R=4 m, circle poses) "
                               "-> 512^3 TSDF, voxel 0.05 m, recast_step 2, max_ray 10 m; marching cubes of the result "
                               "reported in extras (per-output, not per-frame)",
                   "frames_per_step": BATCH, "grid": "512^3", "parallelism": f"submap-sharded x{world} (no data-path collective)",
                   "l2": f"inputs cycle through {POOL_BATCHES} batches = {POOL_BATCHES * BATCH * syn.H * syn.W * 2 / 1e6:.0f} MB > 126 MB L2"},
        "roofline": {"bound": "hbm", "kernel": "k_raymarch", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": (load_traffic("k_raymarch") or (None, None))[0], "traffic_source": (load_traffic("k_raymarch") or (None, None))[1],
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": {"bucket": bucket_ms, "raymarch": ray_ms, "commit": commit_ms},
                     "algorithmic_bytes_per_frame_all_kernels": frame_bytes},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "extras": {"marching_cubes": mc, "per_frame": {"rays": st["n_rays"] / max(args.steps * BATCH, 1),
                                                         "voxel_updates": st["n_updates"] / max(args.steps * BATCH, 1)},
                   "voxel_blocks": st["n_blocks"]},
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
