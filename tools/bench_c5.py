#!/usr/bin/env python3
"""BASELINE config 5: multi-submap TSDF fusion spatially sharded across the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_c5.py

64 submaps (20 frames of the S2 scene each, base poses on an 8x8 grid, 6 m pitch), integrated submap-sharded
(rank r owns submaps r, r+W, ...), fused into a 2048^3 x 0.05 m global volume tiled over the ranks (2x2x2 for 8):
foreign blocks travel in one NCCL all-to-all, the boundary layer is exchanged as ghost blocks, marching cubes runs
locally and the triangle soup is all-gathered.  Prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from taichi_slam.mapping import DenseTSDF, MarchingCubeMesher
    from taichislam_b200 import synthetic as syn
    from taichislam_b200.distributed import TiledGlobalMap, submap_owner, factor_tiles
    n_sub, frames_per = 64, 20
    sub = DenseTSDF(map_scale=[25.6, 25.6], voxel_scale=0.05, max_submap_num=64, max_disp_particles=1024, max_blocks=60000)
    sub.set_dep_camera_intrinsic(syn.K_DEPTH)
    glo = DenseTSDF(map_scale=[102.4, 102.4], voxel_scale=0.05, is_global_map=True, max_submap_num=64, max_disp_particles=1024,
                    max_blocks=120000)
    d = syn.scene_sphere(4.0)
    pinned = torch.from_numpy(d.view(np.int16)).pin_memory().numpy().view(np.uint16)
    mine = [s for s in range(n_sub) if submap_owner(s, world) == rank]
    for s in range(n_sub):
        gx, gy = s % 8, s // 8
        a = 0.3 * s
        Rb = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        Tb = np.array([6.0 * (gx - 3.5), 6.0 * (gy - 3.5), 0.0])
        glo.set_base_pose_submap(s, Rb, Tb)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for s in mine:
        sub.active_submap_id[None] = s
        sub.set_base_pose_submap(s, np.eye(3), np.zeros(3))
        for q in range(frames_per):
            R, T = syn.stream_pose(q * 7)
            sub.recast_depth_to_map(R, T, pinned, np.array([]))
    sub._flush(); torch.cuda.synchronize(); dist.barrier()
    t_int = time.perf_counter() - t0
    tiled = TiledGlobalMap(glo, dist, rank, world)
    tiled.fuse_submaps_tiled(sub)  # warm (NCCL channels, allocations)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    tiled.fuse_submaps_tiled(sub)
    torch.cuda.synchronize(); dist.barrier()
    t_fuse = time.perf_counter() - t0
    x = dict(tiled.last_exchange)
    mesher = MarchingCubeMesher(glo, 3000000, tsdf_surface_thres=0.25)
    t0 = time.perf_counter()
    mv, mn, counts = tiled.generate_mesh_all_gather(mesher)
    torch.cuda.synchronize(); dist.barrier()
    t_mesh = time.perf_counter() - t0
    x.update(tiled.last_exchange)
    own = torch.tensor([glo.count_active(), sum(sub._h.count_active(s) for s in mine), x["fusion_blocks_sent"], x.get("halo_blocks_sent", 0),
                        glo._h.stats()["n_blocks"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(own)
    if rank == 0:
        print(json.dumps({"config": "C5 multi-submap fusion, 2048^3 effective volume", "n_gpus": world, "tiles": list(factor_tiles(world)),
                          "submaps": n_sub, "frames": n_sub * frames_per, "integrate_s": t_int, "integrate_frames_per_s": n_sub * frames_per / t_int,
                          "fuse_ms": 1e3 * t_fuse, "source_voxels": int(own[1].item()), "global_voxels": int(own[0].item()),
                          "fusion_blocks_exchanged": int(own[2].item()), "fusion_exchange_MB": own[2].item() * (8 + 4096 * 10) / 1e6,
                          "halo_blocks_exchanged": int(own[3].item()), "mesh_ms_incl_halo_and_allgather": 1e3 * t_mesh,
                          "triangles": int(sum(counts)), "triangles_per_rank": counts, "global_blocks": int(own[4].item())}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
