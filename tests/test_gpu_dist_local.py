"""Tiled global map on ONE GPU: `world` virtual ranks in one process (taichislam_b200.distributed.run_local_*), the
all-to-all replaced by slicing.  Drives every pack / unpack kernel of csrc/tslam_dist.cu (k_foreign_count/pack,
k_unpack_add, k_halo_count/pack, k_ghost_unpack, k_dirty_hist) and the data-driven tiling, and compares with the
single-map result of DenseTSDF.fuse_submaps (dense_tsdf.py:282-318) + MarchingCubeMesher - so the multi-GPU data path has
driver-side evidence even on a one-GPU box (the NCCL path differs only in who moves the packed rows:
tests/test_gpu_multi.py, bench.py --gpus N)."""
import numpy as np
import pytest

from taichislam_b200 import synthetic as syn
from util import as_dict_rows, rot_xyz

pytestmark = pytest.mark.gpu


def _inputs():
    d = syn.scene_sphere(2.5, 240, 320)
    K = [syn.FX / 2, 0, syn.CX / 2, 0, syn.FY / 2, syn.CY / 2, 0, 0, 1]
    # submaps spread over a FLAT volume (all z within a block or two): equal slices of the volume would leave the
    # upper half of a 2x2x2 tiling empty
    poses = {s: (rot_xyz(0.05 * s, 0.03 * s, 0.7 * s + 0.1), np.array([2.2 * (s % 3) - 2.0, 2.0 * (s // 3) - 1.5, 0.1 * (s % 2)])) for s in range(6)}
    return d, K, poses


def _maps(K, tex):
    from taichi_slam.mapping import DenseTSDF
    sub = DenseTSDF(map_scale=[6.4, 6.4], voxel_scale=0.05, max_ray_length=3.0, max_submap_num=8, max_disp_particles=1024,
                    max_image_pixels=240 * 320, texture_enabled=tex)
    sub.set_dep_camera_intrinsic(K)
    glo = DenseTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, is_global_map=True, max_submap_num=8, max_disp_particles=1024, texture_enabled=tex)
    return sub, glo


def _integrate(sub, glo, poses, d, tex, only=None):
    img = syn.texture_gradient(4, 240, 320) if tex else np.array([])
    for s, (R, T) in poses.items():
        glo.set_base_pose_submap(s, R, T)
        if only is not None and s not in only:
            continue
        sub.active_submap_id[None] = s
        sub.set_base_pose_submap(s, np.eye(3), np.zeros(3))
        sub.recast_depth_to_map(np.eye(3), np.zeros(3), d, img)


@pytest.mark.parametrize("world,tex", [(4, False), (8, False), (2, True)])
def test_local_tiled_fusion_and_mesh_match_single_map(world, tex):
    import torch
    from taichi_slam.mapping import MarchingCubeMesher
    from taichislam_b200.distributed import LocalGroup, TiledGlobalMap, run_local_fusion, run_local_mesh, submap_owner
    d, K, poses = _inputs()
    grp = LocalGroup(world)
    subs, glos, tiled = [], [], []
    for r in range(world):
        sub, glo = _maps(K, tex)
        _integrate(sub, glo, poses, d, tex, only={s for s in poses if submap_owner(s, world) == r})
        subs.append(sub); glos.append(glo)
        tiled.append(TiledGlobalMap(glo, grp, r, world))
    run_local_fusion(tiled, subs)
    assert tiled[0].cuts is not None and sum(t.last_exchange["fusion_blocks_sent"] for t in tiled) > 0
    parts = []
    for r, (t, glo) in enumerate(zip(tiled, glos)):
        got = glo._h.gather_device(0, color=tex)
        idx, tv, w, occ = got[:4]
        if idx.shape[0]:
            bi = torch.div(idx, 16, rounding_mode="floor").cpu().numpy()
            assert {t.owner_of_block(*b) for b in np.unique(bi, axis=0)} <= {r}   # every rank holds only its own tile
        parts.append((idx.cpu().numpy(), tv.cpu().numpy(), w.cpu().numpy(), got[4].cpu().numpy() if tex else None))
    gi, gt, gw = as_dict_rows(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
    # single map
    sub, glo = _maps(K, tex)
    _integrate(sub, glo, poses, d, tex)
    glo.fuse_submaps(sub)
    si, st, sw, _ = as_dict_rows(*glo._h.gather(0))
    assert np.array_equal(gi, si), f"{gi.shape} vs {si.shape}"
    fin = np.isfinite(st)
    assert np.array_equal(np.isfinite(gt), fin)
    assert np.abs(gt[fin] - st[fin]).max() <= 1e-4 and np.all(np.abs(gw - sw) <= 1e-4 * np.maximum(1.0, sw))
    if tex:  # colours travel with the blocks (dense_tsdf.py:276-277)
        cg = np.concatenate([p[3] for p in parts])
        oi = np.concatenate([p[0] for p in parts])
        order = np.lexsort((oi[:, 2], oi[:, 1], oi[:, 0]))
        cs = glo._h.gather(0, color=True)
        so = np.lexsort((cs[0][:, 2], cs[0][:, 1], cs[0][:, 0]))
        ok = np.isfinite(cs[4][so]).all(axis=1) & np.isfinite(cg[order]).all(axis=1)
        assert np.abs(cg[order][ok] - cs[4][so][ok]).max() <= 1e-4
    # meshes
    meshers = [MarchingCubeMesher(g_, 400000, tsdf_surface_thres=0.25) for g_ in glos]
    mv, mn, counts = run_local_mesh(tiled, meshers)
    assert sum(t.last_exchange["halo_blocks_received"] for t in tiled) > 0
    ref = MarchingCubeMesher(glo, 400000, tsdf_surface_thres=0.25)
    ref.generate_mesh(1)
    n = int(ref.num_facelets[None])
    assert abs(sum(counts) - n) <= max(4, int(1e-4 * n)) and n > 1000
    # data-driven tiling: no rank without surface (equal slices of this flat volume leave the upper tiles empty); the
    # marginal-quantile cuts are exact per axis, not per tile, so a small scene is not perfectly level
    assert min(counts) > 0 and max(counts) < 6 * min(counts), counts
    if sum(counts) == n:
        from scipy.spatial import cKDTree
        sv = ref.mesh_vertices.to_numpy()[:3 * n].reshape(-1, 9)
        gv = mv.cpu().numpy().reshape(-1, 9)
        sv, gv = sv[np.isfinite(sv).all(axis=1)], gv[np.isfinite(gv).all(axis=1)]  # fusion NaNs are reference behaviour (0/0, dense_tsdf.py:275)
        assert abs(len(sv) - len(gv)) <= 4
        dd, _ = cKDTree(sv).query(gv)
        assert np.quantile(dd, 0.99) <= 1e-4 and dd.max() <= 2.5e-2
