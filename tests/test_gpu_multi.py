"""Multi-GPU (NCCL) test of the tiled global map: fusion with block exchange and mesh all-gather on W ranks
== the single-GPU result.  Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build_inputs():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from taichislam_b200 import synthetic as syn
    from util import rot_xyz
    d = syn.scene_sphere(2.5, 240, 320)
    K = [syn.FX / 2, 0, syn.CX / 2, 0, syn.FY / 2, syn.CY / 2, 0, 0, 1]
    poses = {s: (rot_xyz(0.1 * s + 0.05, 0.2 - 0.07 * s, 0.3 * s + 0.1), np.array([1.5 * (s % 2) - 0.7, 1.2 * (s // 2) - 0.6, 0.3 * s - 0.4]))
             for s in range(4)}
    return d, K, poses


def _make_maps(K):
    from taichi_slam.mapping import DenseTSDF
    sub = DenseTSDF(map_scale=[6.4, 6.4], voxel_scale=0.05, max_ray_length=3.0, max_submap_num=8, max_disp_particles=1024,
                    max_image_pixels=240 * 320)
    sub.set_dep_camera_intrinsic(K)
    glo = DenseTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, is_global_map=True, max_submap_num=8, max_disp_particles=1024)
    return sub, glo


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from taichi_slam.mapping import MarchingCubeMesher
        from taichislam_b200.distributed import TiledGlobalMap, submap_owner, all_gather_ragged
        d, K, poses = _build_inputs()
        sub, glo = _make_maps(K)
        for s, (R, T) in poses.items():
            glo.set_base_pose_submap(s, R, T)
            if submap_owner(s, world) != rank:
                continue
            sub.active_submap_id[None] = s
            sub.set_base_pose_submap(s, np.eye(3), np.zeros(3))
            sub.recast_depth_to_map(np.eye(3), np.zeros(3), d, np.array([]))
        tiled = TiledGlobalMap(glo, dist, rank, world)
        tiled.fuse_submaps_tiled(sub)
        idx, t, w, occ = glo._h.gather_device(0)
        # every voxel this rank holds lies in its own tile
        bi = torch.div(idx, 16, rounding_mode="floor").cpu().numpy()
        owners = {tiled.owner_of_block(*b) for b in np.unique(bi, axis=0)}
        assert owners <= {rank}, owners
        all_idx, _ = all_gather_ragged(dist, idx, world)
        all_t, _ = all_gather_ragged(dist, t, world)
        all_w, _ = all_gather_ragged(dist, w, world)
        mesher = MarchingCubeMesher(glo, 400000, tsdf_surface_thres=0.25)
        mv, mn, counts = tiled.generate_mesh_all_gather(mesher)
        if rank == 0:
            q.put(("ok", all_idx.cpu().numpy(), all_t.cpu().numpy(), all_w.cpu().numpy(), mv.cpu().numpy(), mn.cpu().numpy(), counts,
                   dict(tiled.last_exchange)))
        dist.barrier()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("err", rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_tiled_fusion_and_mesh_match_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = q.get(timeout=600)
    for p in ps:
        p.join(timeout=120)
    assert res[0] == "ok", res
    _, idx, t, w, mv, mn, counts, xstats = res
    assert xstats["fusion_blocks_sent"] > 0 and xstats["halo_blocks_received"] > 0 and min(counts) > 0

    # single-GPU reference: same submaps, one map
    from taichi_slam.mapping import MarchingCubeMesher
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import as_dict_rows
    d, K, poses = _build_inputs()
    sub, glo = _make_maps(K)
    for s, (R, T) in poses.items():
        glo.set_base_pose_submap(s, R, T)
        sub.active_submap_id[None] = s
        sub.set_base_pose_submap(s, np.eye(3), np.zeros(3))
        sub.recast_depth_to_map(np.eye(3), np.zeros(3), d, np.array([]))
    glo.fuse_submaps(sub)
    si, st, sw, _ = as_dict_rows(*glo._h.gather(0))
    gi, gt, gw = as_dict_rows(idx, t, w)
    assert np.array_equal(gi, si), f"{gi.shape} vs {si.shape}"
    fin = np.isfinite(st)
    assert np.array_equal(np.isfinite(gt), fin)
    assert np.abs(gt[fin] - st[fin]).max() <= 1e-4
    assert np.all(np.abs(gw - sw) <= 1e-4 * np.maximum(1.0, sw))
    mesher = MarchingCubeMesher(glo, 400000, tsdf_surface_thres=0.25)
    mesher.generate_mesh(1)
    n = int(mesher.num_facelets[None])
    assert sum(counts) == n or abs(sum(counts) - n) <= max(4, int(1e-4 * n))
    if sum(counts) == n:
        from scipy.spatial import cKDTree
        sv = mesher.mesh_vertices.to_numpy()[:3 * n].reshape(-1, 9)
        dd, _ = cKDTree(sv).query(mv.reshape(-1, 9))
        assert np.quantile(dd, 0.99) <= 1e-4 and dd.max() <= 2.5e-2
