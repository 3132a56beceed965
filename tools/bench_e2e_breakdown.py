"""Where does the end-to-end (host frames -> map) time go?  Run on a GPU box.

  (a) Python + ctypes cost of one DenseTSDF.recast_depth_to_map call (tiny 16x16 frames: copy and kernels negligible)
  (b) H2D rate of 614 KB pinned frame copies issued back to back (cudaMemcpyAsync, what the queue does per frame)
  (c) zero-copy: a kernel reading pinned host memory directly (sampled rows only), as a PCIe read-rate probe
  (d) the e2e step as bench.py runs it, split into enqueue loop / flush+sync
"""
import time

import numpy as np
import torch

from taichislam_b200 import synthetic as syn
from taichislam_b200.mapping import DenseTSDF


def main():
    torch.cuda.init()
    out = {}
    # (a) per-call overhead
    m = DenseTSDF(map_scale=[25.6, 25.6], voxel_scale=0.05, is_global_map=True)
    m.set_dep_camera_intrinsic(syn.K_DEPTH)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    tiny = torch.zeros((16, 16), dtype=torch.int16).pin_memory().numpy().view(np.uint16)
    Rs, Ts = syn.stream_poses(6400)
    e = np.array([])
    for q in range(640):
        m.recast_depth_to_map(Rs[q], Ts[q], tiny, e)
    m.frame_counters()
    t0 = time.perf_counter()
    for q in range(6400):
        m.recast_depth_to_map(Rs[q], Ts[q], tiny, e)
    t1 = time.perf_counter()
    m.frame_counters()
    out["python_call_us"] = (t1 - t0) / 6400 * 1e6
    # (b) H2D of 614 KB pinned frames
    n = 256
    host = torch.zeros((n, 480, 640), dtype=torch.int16).pin_memory()
    dev = torch.empty((64, 480, 640), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for q in range(n):
            dev[q % 64].copy_(host[q], non_blocking=True)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out["h2d_frame_copies_GBps"] = n * 614400 / dt / 1e9
    out["h2d_issue_us_per_copy"] = t_issue / n * 1e6
    big = torch.zeros((64, 480, 640), dtype=torch.int16).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(8):
        dev.copy_(big, non_blocking=True)
    torch.cuda.synchronize()
    out["h2d_39MB_copies_GBps"] = 8 * big.numel() * 2 / (time.perf_counter() - t0) / 1e9
    # (d) e2e step split
    frames = syn.scene_sphere(4.0)
    hostf = torch.from_numpy(np.stack([frames] * 64).view(np.int16)).pin_memory().numpy().view(np.uint16)
    m2 = DenseTSDF(map_scale=[25.6, 25.6], voxel_scale=0.05, is_global_map=True)
    m2.set_dep_camera_intrinsic(syn.K_DEPTH)
    m2.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    enq, fl = [], []
    for s in range(12):
        t0 = time.perf_counter()
        for q in range(64):
            m2.recast_depth_to_map(Rs[64 * s + q], Ts[64 * s + q], hostf[q], e)
        t1 = time.perf_counter()
        m2.frame_counters()
        t2 = time.perf_counter()
        if s >= 2:
            enq.append(t1 - t0)
            fl.append(t2 - t1)
    out["e2e_enqueue_ms_per_64"] = 1e3 * float(np.mean(enq))
    out["e2e_flush_sync_ms"] = 1e3 * float(np.mean(fl))
    out["e2e_fps"] = 64 / (float(np.mean(enq)) + float(np.mean(fl)))
    print(out)


if __name__ == "__main__":
    main()
