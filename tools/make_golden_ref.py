#!/usr/bin/env python3
"""Golden vectors from the REFERENCE'S OWN KERNELS, executed - build container only (/root/reference must exist).

oracle/taichi_emu.py is a stand-in for the `taichi` package that runs the reference's unmodified kernel source in
Python with Taichi's value typing (f16 fields, f32 default) on one legal serial schedule.  This script drives the
reference classes exactly as their callers do (set_dep_camera_intrinsic / set_base_pose_submap / recast_* / fuse /
export / generate_mesh) on small seeded inputs and stores the resulting map state in
tests/golden/ref_exec.npz; tests/test_oracle_vs_reference_exec.py replays the same inputs through the oracle.

Run from anywhere:  python tools/make_golden_ref.py        (takes a few minutes: the kernels run as Python loops)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT] + [ROOT, os.path.join(ROOT, "tests")]
from oracle import taichi_emu as emu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_exec.npz")


def state(m, submap=None):
    """observed voxels of a reference DenseTSDF: idx int16[n,3] (+ submap), TSDF f16, W f16; occupy cells separately."""
    keys = sorted(k for k, v in m.TSDF_observed.d.items() if v > 0 and (submap is None or k[0] == submap))
    idx = np.array([k[1:] for k in keys], np.int16).reshape(-1, 3)
    t = np.array([m.TSDF.d[k] for k in keys], np.float16)
    w = np.array([m.W_TSDF.d[k] for k in keys], np.float16)
    occ = np.array([m.occupy.d.get(k, 0) for k in keys], np.int8)
    occ_cells = np.array(sorted(k[1:] for k, v in m.occupy.d.items() if v != 0 and (submap is None or k[0] == submap)), np.int16).reshape(-1, 3)
    return idx, t, w, occ, occ_cells


def inputs():
    from taichislam_b200 import synthetic as syn
    from util import rot_xyz
    K = [v / 4 if i in (0, 2, 4, 5) else v for i, v in enumerate(syn.K_DEPTH)]
    d1 = np.minimum(syn.scene_room()[::4, ::4], 2800).astype(np.uint16)
    d2 = np.minimum(syn.scene_sphere(2.5)[::4, ::4], 2800).astype(np.uint16)
    P1 = (rot_xyz(0.1, -0.05, 0.3), np.array([0.2, -0.1, 0.05]))
    P2 = (rot_xyz(-0.05, 0.1, 0.25), np.array([0.25, -0.05, 0.1]))
    rng = np.random.default_rng(11)
    dirs = rng.normal(size=(2500, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pcl = (dirs * rng.uniform(0.5, 3.4, (2500, 1))).astype(np.float32)   # some beyond max_ray_length: filtered (:177)
    return K, d1, d2, P1, P2, pcl


def main():
    t00 = time.time()
    ref = emu.load_reference()
    D, O, MC = ref.dense_tsdf.DenseTSDF, ref.taichi_octomap.Octomap, ref.marching_cube_mesher.MarchingCubeMesher
    K, d1, d2, P1, P2, pcl = inputs()
    g = {"K": np.array(K), "d1": d1, "d2": d2, "P1_R": P1[0], "P1_T": P1[1], "P2_R": P2[0], "P2_T": P2[1], "pcl": pcl}
    kw = dict(map_scale=[6.4, 6.4], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=3.0, max_disp_particles=1 << 17,
              max_submap_num=4)
    e = np.array([])

    # A. global map: frame 1, then frame 2 on top (dense_tsdf.py:162-165, :188-270)
    m = D(is_global_map=True, **kw)
    m.set_dep_camera_intrinsic(K)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    m.recast_depth_to_map(P1[0], P1[1], d1, e)
    for n, a in zip(("idx", "T", "W", "occ", "occcells"), state(m)):
        g["A1_" + n] = a
    print(f"A1 {len(g['A1_idx'])} voxels, {time.time() - t00:.0f}s")
    # F. exporters / IO on that state (:339-365, :412-454)
    g["F_count_active"] = np.array(int(m.count_active()))
    n = int(g["F_count_active"])
    di, dt_, dw, do, dc = np.zeros((n, 3), np.int16), np.zeros(n, np.float16), np.zeros(n, np.float16), np.zeros(n, np.int8), np.array([])
    m.to_numpy(di, dt_, dw, do, dc)
    g["F_to_numpy_idx"], g["F_to_numpy_T"], g["F_to_numpy_W"], g["F_to_numpy_occ"] = di, dt_, dw, do
    m.disp_floor, m.disp_ceiling = -3.0, 3.0
    m2 = D(is_global_map=True, disp_floor=-3.0, disp_ceiling=3.0, **kw)   # same map with a display range that keeps everything
    m2.load_numpy(0, di, dt_, dw, do, dc)
    m2.cvt_TSDF_surface_to_voxels()
    ns = int(m2.num_TSDF_particles[None])
    g["F_surface_xyz"] = m2.export_TSDF_xyz.to_numpy()[:ns].astype(np.float32)
    g["F_loaded_count"] = np.array(int(m2.count_active()))
    # slice export (:367-389): k index == int(f16(z) / vs) for dz = 0.5; then dz = 1.5 (three layers)
    for tag, z, dz in (("s1", 0.52, 0.5), ("s3", 0.33, 1.5)):
        m2.cvt_TSDF_to_voxels_slice(z, dz)
        nsl = int(m2.num_TSDF_particles[None])
        g[f"F_{tag}_xyz"] = m2.export_TSDF_xyz.to_numpy()[:nsl].astype(np.float32)
        g[f"F_{tag}_val"] = m2.export_TSDF.to_numpy()[:nsl].astype(np.float32)
        g[f"F_{tag}_z_dz"] = np.array([z, dz])
    # append mode of the surface export (cvt_TSDF_surface_to_voxels_to, :326-328): counter keeps running
    m2.num_TSDF_particles[None] = 7
    m2.cvt_TSDF_surface_to_voxels_to(m2.num_TSDF_particles, m2.max_disp_particles, m2.export_TSDF_xyz, m2.export_color)
    g["F_append_count"] = np.array(int(m2.num_TSDF_particles[None]))
    # H. marching cubes on the loaded map (marching_cube_mesher.py:127-187), step 1
    mesher = MC(m2, max_triangles=200000, tsdf_surface_thres=0.1)
    mesher.generate_mesh(1)
    nt = int(mesher.num_facelets[None])
    g["H_mc_triangles"] = np.array(nt)
    g["H_mc_vertices"] = mesher.mesh_vertices.to_numpy()[:3 * nt].astype(np.float32)
    g["H_mc_normals"] = mesher.mesh_normals.to_numpy()[:3 * nt].astype(np.float32)
    print(f"F/H surface {ns}, triangles {nt}, {time.time() - t00:.0f}s")
    m.recast_depth_to_map(P2[0], P2[1], d2, e)
    for n, a in zip(("idx", "T", "W", "occ", "occcells"), state(m)):
        g["A2_" + n] = a
    print(f"A2 {len(g['A2_idx'])} voxels, {time.time() - t00:.0f}s")

    # B. edge semantics of the depth kernel (:188-203): image size not a multiple of recast_step (range(0, h/step) truncates),
    #    zero pixels, pixels beyond max_ray_length*1000 and below min_ray_length*1000, a different internal_voxels
    dB = d1[:119, :157].copy()
    dB[::7, ::5] = 0
    dB[3::11, 2::13] = 3500        # > max_ray_length * 1000
    dB[5::13, 1::11] = 250         # < min_ray_length * 1000
    mB = D(is_global_map=True, recast_step=3, internal_voxels=5, **kw)
    mB.set_dep_camera_intrinsic(K)
    mB.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    mB.recast_depth_to_map(P2[0], P2[1], dB, e)
    for n, a in zip(("idx", "T", "W", "occ", "occcells"), state(mB)):
        g["B_" + n] = a
    g["B_depth"] = dB
    print(f"B {len(g['B_idx'])} voxels, {time.time() - t00:.0f}s")

    # C. point-cloud variant (:167-186)
    m = D(is_global_map=True, **kw)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    m.recast_pcl_to_map(P1[0], P1[1], pcl, e)
    for n, a in zip(("idx", "T", "W", "occ", "occcells"), state(m)):
        g["C_" + n] = a
    print(f"C {len(g['C_idx'])} voxels, {time.time() - t00:.0f}s")

    # D. texture (:204-213, :233-234, :268-269): uniformly coloured images, so that the racy "last ray wins" overwrite has
    #    one possible outcome - frame 1 in colour A everywhere, frame 2 (left half of the image valid) in colour B
    mt = D(is_global_map=True, texture_enabled=True, **kw)
    mt.set_dep_camera_intrinsic(K)
    mt.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    texA = np.zeros(d1.shape + (3,), np.uint8); texA[:] = (200, 40, 90)
    texB = np.zeros(d1.shape + (3,), np.uint8); texB[:] = (10, 250, 30)
    mt.recast_depth_to_map(P1[0], P1[1], d1, texA)
    d1half = d1.copy(); d1half[:, d1.shape[1] // 2:] = 0
    mt.recast_depth_to_map(P1[0], P1[1], d1half, texB)
    keys = sorted(k for k, v in mt.TSDF_observed.d.items() if v > 0)
    g["D_idx"] = np.array([k[1:] for k in keys], np.int16)
    g["D_W"] = np.array([mt.W_TSDF.d[k] for k in keys], np.float16)
    g["D_color"] = np.array([mt.color.d[k] for k in keys], np.float16)
    g["D_texA"], g["D_texB"], g["D_d1half"] = texA[0, 0], texB[0, 0], d1half
    print(f"D textured {len(keys)} voxels, {time.time() - t00:.0f}s")

    # E. submap collection + fusion into a global map (:272-318)
    from util import rot_xyz
    sub = D(is_global_map=False, **kw)
    sub.set_dep_camera_intrinsic(K)
    glo = D(is_global_map=True, **dict(kw, map_scale=[12.8, 12.8]))
    base = [(rot_xyz(0.1, 0.2, 0.3), np.array([0.5, 0.1, -0.2])), (rot_xyz(-0.3, 0.1, 1.0), np.array([-0.4, 0.6, 0.3]))]
    dsmall = d1[::2, ::2].copy()
    Ks = [v / 2 if i in (0, 2, 4, 5) else v for i, v in enumerate(K)]
    sub.set_dep_camera_intrinsic(Ks)
    for s, (Rb, Tb) in enumerate(base):
        sub.set_base_pose_submap(s, Rb, Tb)
        glo.set_base_pose_submap(s, Rb, Tb)
        Rw, Tw = Rb @ P1[0], Rb @ P1[1] + Tb          # world pose whose submap-relative pose is P1
        sub.recast_depth_to_map(Rw, Tw, dsmall, e)
        for n, a in zip(("idx", "T", "W", "occ", "occcells"), state(sub, s)):
            g[f"E_sub{s}_" + n] = a
        if s == 0:
            sub.switch_to_next_submap()
    glo.fuse_submaps(sub)
    keys = sorted(k for k, v in glo.TSDF_observed.d.items() if v > 0)
    g["E_glo_idx"] = np.array([k[1:] for k in keys], np.int16)
    g["E_glo_T"] = np.array([glo.TSDF.d[k] for k in keys], np.float16)
    g["E_glo_W"] = np.array([glo.W_TSDF.d[k] for k in keys], np.float16)
    g["E_glo_occ"] = np.array([glo.occupy.d.get(k, 0) for k in keys], np.int8)
    g["E_base_R"], g["E_base_T"], g["E_dsmall"], g["E_K"] = np.stack([b[0] for b in base]), np.stack([b[1] for b in base]), dsmall, np.array(Ks)
    print(f"E fused {len(keys)} voxels, {time.time() - t00:.0f}s")

    # G. Octomap: points, depth, fusion, level-1 export (taichi_octomap.py:116-199)
    okw = dict(map_scale=[6.4, 6.4], voxel_scale=0.05, min_occupy_thres=1, K=2, max_ray_length=3.0, max_submap_num=4, max_disp_particles=1 << 17)
    oc = O(**okw)
    oc.set_dep_camera_intrinsic(K)
    oc.set_base_pose_submap(0, base[0][0], base[0][1])
    Rw, Tw = base[0][0] @ P1[0], base[0][0] @ P1[1] + base[0][1]
    oc.recast_pcl_to_map(Rw, Tw, pcl, e, len(pcl))
    oc.recast_depth_to_map(Rw, Tw, d1, e)
    keys = sorted(k for k, v in oc.occupy.d.items() if v > 0)
    g["G_idx"] = np.array([k[1:] for k in keys], np.int16)
    g["G_count"] = np.array([oc.occupy.d[k] for k in keys], np.float32)
    oc.cvt_occupy_to_voxels(1)
    ne = int(oc.num_export_particles[None])
    g["G_export_xyz"] = oc.export_x.to_numpy()[:ne].astype(np.float32)
    # fusion: the reference's Octomap.fuse_submaps reads `submaps.color`, which only exists with texture_enabled=True
    # (taichi_octomap.py:77-79, :198) - an untextured Octomap cannot be fused there (AttributeError).  Textured pair:
    from taichislam_b200 import synthetic as syn
    tex = syn.texture_gradient(9, d1.shape[0], d1.shape[1])
    rgb = np.random.default_rng(3).integers(0, 256, (len(pcl), 3)).astype(np.uint8)
    oct_ = O(**dict(okw, texture_enabled=True))
    oct_.set_dep_camera_intrinsic(K)
    oct_.set_base_pose_submap(0, base[0][0], base[0][1])
    oct_.recast_pcl_to_map(Rw, Tw, pcl, rgb, len(pcl))
    oct_.recast_depth_to_map(Rw, Tw, d1, tex)
    og = O(**dict(okw, map_scale=[12.8, 12.8], is_global_map=True, texture_enabled=True))
    og.set_base_pose_submap(0, base[0][0], base[0][1])
    og.fuse_submaps(oct_)
    keys = sorted(k for k, v in og.occupy.d.items() if v > 0)
    g["G_fused_idx"] = np.array([k[1:] for k in keys], np.int16)
    g["G_fused_count"] = np.array([og.occupy.d[k] for k in keys], np.float32)
    g["G_tex"], g["G_rgb"] = tex, rgb
    print(f"G octomap {len(g['G_idx'])} voxels, fused {len(keys)}, {time.time() - t00:.0f}s")

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def topo():
    """TopoGraphGen (topo_graph.py:128-511) executed on the analytic two-room world of tests/topo_world.py."""
    ref = emu.load_reference()
    from topo_world import two_rooms, START_A
    idx, t, w, occ = two_rooms()
    m = ref.dense_tsdf.DenseTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True,
                                 max_disp_particles=1024, max_submap_num=4)
    m.load_numpy(0, idx, t, w, occ, np.array([]))
    g = {}
    for tag, kw in (("a", dict(coll_det_num=64, max_raycast_dist=2.5)), ("b", dict(coll_det_num=128, max_raycast_dist=2))):
        tg = ref.topo_graph.TopoGraphGen(m, max_facelets=8192, **kw)
        n = tg.generate_topo_graph(START_A, max_nodes=12)
        nf, nfr, ne = tg.num_facelets[None], tg.num_frontiers[None], tg.edge_num[None]
        g[tag + "_counts"] = np.array([n, nf, nfr, ne])
        g[tag + "_node_center"] = np.array([tg.nodes[i].center.a for i in range(n)], np.float32)
        g[tag + "_node_range"] = np.array([[tg.nodes[i].start_facelet_idx, tg.nodes[i].end_facelet_idx, tg.nodes[i].master_idx] for i in range(n)])
        g[tag + "_facelet_normal"] = np.array([tg.facelets[i].normal.a for i in range(nf)], np.float32)
        g[tag + "_facelet_frontier"] = np.array([int(tg.facelets[i].is_frontier) for i in range(nf)], np.int8)
        g[tag + "_tri_vertices"] = tg.tri_vertices.to_numpy()[:3 * nf].astype(np.float32)
        g[tag + "_frontier_valid"] = np.array([int(tg.frontiers[k].is_valid) for k in range(nfr)], np.int8)
        g[tag + "_frontier_proj_center"] = np.array([tg.frontiers[k].projected_center.a for k in range(nfr)], np.float32)
        g[tag + "_frontier_next"] = np.array([tg.frontiers[k].next_node_initial.a for k in range(nfr)], np.float32)
        g[tag + "_edges"] = tg.edges.to_numpy()[:ne].astype(np.float32)
        print(tag, "nodes", n, "facelets", nf, "frontiers", nfr, "edge points", ne)
    out = os.path.join(ROOT, "tests", "golden", "ref_exec_topo.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


def f32_state():
    """The reference's integrate kernels with every `ti.f16` declaration read as `ti.f32` (same algorithm, f32 state): what
    the f32-state modes of the oracle - and with them the CUDA path - should reproduce without the f16 noise."""
    emu.f16.np = np.float32
    ref = emu.load_reference()
    K, d1, d2, P1, P2, pcl = inputs()
    g = {}
    m = ref.dense_tsdf.DenseTSDF(is_global_map=True, map_scale=[6.4, 6.4], voxel_scale=0.05, num_voxel_per_blk_axis=16,
                                 max_ray_length=3.0, max_disp_particles=4096, max_submap_num=4)
    m.set_dep_camera_intrinsic(K)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    e = np.array([])
    for tag, (P, d) in (("A1", (P1, d1)), ("A2", (P2, d2))):
        m.recast_depth_to_map(P[0], P[1], d, e)
        keys = sorted(k for k, v in m.TSDF_observed.d.items() if v > 0)
        g[tag + "_idx"] = np.array([k[1:] for k in keys], np.int16)
        g[tag + "_T"] = np.array([m.TSDF.d[k] for k in keys], np.float32)
        g[tag + "_W"] = np.array([m.W_TSDF.d[k] for k in keys], np.float32)
        print(tag, len(keys), "voxels")
        if tag == "A1":   # marching cubes straight on the f32 map (marching_cube_mesher.py:127-187)
            mesher = ref.marching_cube_mesher.MarchingCubeMesher(m, max_triangles=100000, tsdf_surface_thres=0.1)
            mesher.generate_mesh(1)
            nt = int(mesher.num_facelets[None])
            g["A1_mc_vertices"] = mesher.mesh_vertices.to_numpy()[:3 * nt].astype(np.float32)
            g["A1_mc_normals"] = mesher.mesh_normals.to_numpy()[:3 * nt].astype(np.float32)
            print("   triangles", nt)
    # submaps + fusion with f32 state (dense_tsdf.py:272-318)
    from util import rot_xyz
    kw = dict(map_scale=[6.4, 6.4], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=3.0, max_disp_particles=4096, max_submap_num=4)
    sub = ref.dense_tsdf.DenseTSDF(is_global_map=False, **kw)
    glo = ref.dense_tsdf.DenseTSDF(is_global_map=True, **dict(kw, map_scale=[12.8, 12.8]))
    base = [(rot_xyz(0.1, 0.2, 0.3), np.array([0.5, 0.1, -0.2])), (rot_xyz(-0.3, 0.1, 1.0), np.array([-0.4, 0.6, 0.3]))]
    dsmall = d1[::2, ::2].copy()
    sub.set_dep_camera_intrinsic([v / 2 if i in (0, 2, 4, 5) else v for i, v in enumerate(K)])
    for s_, (Rb, Tb) in enumerate(base):
        sub.set_base_pose_submap(s_, Rb, Tb)
        glo.set_base_pose_submap(s_, Rb, Tb)
        sub.recast_depth_to_map(Rb @ P1[0], Rb @ P1[1] + Tb, dsmall, e)
        if s_ == 0:
            sub.switch_to_next_submap()
    glo.fuse_submaps(sub)
    keys = sorted(k for k, v in glo.TSDF_observed.d.items() if v > 0)
    g["E_glo_idx"] = np.array([k[1:] for k in keys], np.int16)
    g["E_glo_T"] = np.array([glo.TSDF.d[k] for k in keys], np.float32)
    g["E_glo_W"] = np.array([glo.W_TSDF.d[k] for k in keys], np.float32)
    print("fused", len(keys), "voxels")
    out = os.path.join(ROOT, "tests", "golden", "ref_exec_f32.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


def saturation():
    """48 frames of a wall 0.5 m in front of a slowly moving sensor, f32 state: next to the sensor hundreds of rays hit
    the same voxel per frame and W_TSDF sits at Wmax = 1000 (dense_tsdf.py:267) from the first frame on, where the
    reference's per-SAMPLE read-modify-write is an exponential moving average over the most recent samples.  The final
    map goes to tests/golden/ref_exec_sat.npz: tests/test_oracle_vs_reference_exec.py measures how far commits once per
    frame / once per 32 frames (sum of the contributions, then one clamped update) are from it."""
    emu.f16.np = np.float32
    ref = emu.load_reference()
    from taichislam_b200 import synthetic as syn
    K = [v / 10 if i in (0, 2, 4, 5) else v for i, v in enumerate(syn.K_DEPTH)]
    d = np.full((48, 64), 500, np.uint16)
    d[:, 40:] = 650   # a step in the wall
    m = ref.dense_tsdf.DenseTSDF(is_global_map=True, map_scale=[6.4, 6.4], voxel_scale=0.05, num_voxel_per_blk_axis=16,
                                 max_ray_length=3.0, max_disp_particles=4096, max_submap_num=4)
    m.set_dep_camera_intrinsic(K)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    n = 48
    Rs, Ts = syn.stream_poses(n, start=3, period=400, radius=0.08)
    e = np.array([])
    for q in range(n):
        m.recast_depth_to_map(Rs[q], Ts[q], d, e)
    keys = sorted(k for k, v in m.TSDF_observed.d.items() if v > 0)
    g = {"K": np.array(K), "depth": d, "Rs": Rs, "Ts": Ts,
         "idx": np.array([k[1:] for k in keys], np.int16), "T": np.array([m.TSDF.d[k] for k in keys], np.float32),
         "W": np.array([m.W_TSDF.d[k] for k in keys], np.float32)}
    print(len(keys), "voxels,", int((g["W"] >= 999.5).sum()), "at Wmax")
    out = os.path.join(ROOT, "tests", "golden", "ref_exec_sat.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


def texproj():
    """tests/golden/ref_exec_texproj.npz: the colour camera path of the depth kernel (color_same_proj=False,
    dense_tsdf.py:208-210 -> color_ind_from_depth_pt, mapping_common.py:44-59) executed: a colour camera with its own
    intrinsics and a NON-SQUARE image of vertical colour bands, so that (a) the pixel mapping and (b) the swapped bound
    test of :56 (colour x is compared with the image HEIGHT: every depth pixel that projects to x >= h reads texture[0, 0],
    marked white) show in the voxel colours.  Within a band the racy "last ray wins" overwrite (:268-269) has one outcome."""
    t00 = time.time()
    ref = emu.load_reference()
    D = ref.dense_tsdf.DenseTSDF
    K, d1, d2, P1, P2, pcl = inputs()
    kw = dict(map_scale=[6.4, 6.4], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=3.0, max_disp_particles=1 << 17,
              max_submap_num=4)
    th, tw = 90, 200
    Kc = [K[0] * 1.2, 0.0, 100.0, 0.0, K[4] * 0.7, 45.0, 0.0, 0.0, 1.0]
    bands = np.array([(30, 60, 90), (220, 40, 40), (40, 200, 60), (50, 70, 230), (230, 210, 40), (150, 40, 200), (20, 180, 190),
                      (240, 130, 30), (90, 90, 90), (200, 200, 120), (120, 20, 60), (60, 140, 20), (10, 30, 160), (170, 170, 250)], np.uint8)
    tex = np.zeros((th, tw, 3), np.uint8)
    for x in range(tw):
        tex[:, x] = bands[x // 15]
    tex[0, 0] = (255, 255, 255)
    m = D(is_global_map=True, texture_enabled=True, color_same_proj=False, disp_floor=-3.0, disp_ceiling=3.0, **kw)  # (display range that keeps every voxel in the surface export)
    m.set_dep_camera_intrinsic(K)
    m.set_color_camera_intrinsic(Kc)
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    m.recast_depth_to_map(P1[0], P1[1], d1, tex)
    keys = sorted(k for k, v in m.TSDF_observed.d.items() if v > 0)
    g = {"K": np.array(K), "Kc": np.array(Kc), "d1": d1, "P1_R": P1[0], "P1_T": P1[1], "tex": tex, "bands": bands,
         "idx": np.array([k[1:] for k in keys], np.int16), "color": np.array([m.color.d[k] for k in keys], np.float16)}
    # coloured marching cubes on this state (marching_cube_mesher.py:62-82 vertexInterp_color, :104-108, :150-170)
    idx, t, w, occ, _ = state(m)
    g["T"], g["W"], g["occ"] = t, w, occ
    mesher = ref.marching_cube_mesher.MarchingCubeMesher(m, max_triangles=200000, tsdf_surface_thres=0.1)
    mesher.generate_mesh(1)
    nt = int(mesher.num_facelets[None])
    g["mc_triangles"] = np.array(nt)
    g["mc_vertices"] = mesher.mesh_vertices.to_numpy()[:3 * nt].astype(np.float32)
    g["mc_colors"] = mesher.mesh_colors.to_numpy()[:3 * nt].astype(np.float32)
    print(f"texproj: coloured mesh {nt} triangles, {time.time() - t00:.0f}s")
    # surface export of a textured map (:339-362): export_color = the voxel's colour instead of the jet value
    m.cvt_TSDF_surface_to_voxels()
    ns = int(m.num_TSDF_particles[None])
    g["surf_xyz"] = m.export_TSDF_xyz.to_numpy()[:ns].astype(np.float32)
    g["surf_color"] = m.export_color.to_numpy()[:ns].astype(np.float32)
    g["surf_disp"] = np.array([m.disp_floor, m.disp_ceiling])
    # textured point-cloud variant (:178-183): colour per point = one of 8 by the octant of its direction
    pal8 = np.array([(250, 20, 20), (20, 250, 20), (20, 20, 250), (240, 240, 30), (30, 240, 240), (240, 30, 240), (130, 130, 130),
                     (250, 140, 10)], np.uint8)
    octant = (pcl[:, 0] > 0).astype(int) * 4 + (pcl[:, 1] > 0).astype(int) * 2 + (pcl[:, 2] > 0).astype(int)
    rgb = pal8[octant]
    mp = D(is_global_map=True, texture_enabled=True, **kw)
    mp.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    mp.recast_pcl_to_map(np.eye(3), np.zeros(3), pcl, rgb)
    pkeys = sorted(k for k, v in mp.TSDF_observed.d.items() if v > 0)
    g["pcl"], g["pcl_rgb"], g["pal8"] = pcl, rgb, pal8
    g["P_idx"] = np.array([k[1:] for k in pkeys], np.int16)
    g["P_color"] = np.array([mp.color.d[k] for k in pkeys], np.float16)
    print(f"texproj: textured point cloud {len(pkeys)} voxels, {time.time() - t00:.0f}s")
    # textured submaps fused into a global map (fuse_with_interploation :273-280 with the colour line :276-277): submap 0
    # uniformly colour A, submap 1 colour B - the fused colour of a voxel is the weight mix of the two
    from util import rot_xyz
    sub = D(is_global_map=False, texture_enabled=True, **kw)
    glo = D(is_global_map=True, texture_enabled=True, **dict(kw, map_scale=[12.8, 12.8]))
    base = [(rot_xyz(0.1, 0.2, 0.3), np.array([0.5, 0.1, -0.2])), (rot_xyz(-0.3, 0.1, 1.0), np.array([-0.4, 0.6, 0.3]))]
    dsmall = d1[::2, ::2].copy()
    Ks = [v / 2 if i in (0, 2, 4, 5) else v for i, v in enumerate(K)]
    sub.set_dep_camera_intrinsic(Ks)
    colAB = np.array([(200, 40, 90), (10, 250, 30)], np.uint8)
    for s_, (Rb, Tb) in enumerate(base):
        sub.set_base_pose_submap(s_, Rb, Tb)
        glo.set_base_pose_submap(s_, Rb, Tb)
        Rw, Tw = Rb @ P1[0], Rb @ P1[1] + Tb
        texs = np.zeros(dsmall.shape + (3,), np.uint8); texs[:] = colAB[s_]
        sub.recast_depth_to_map(Rw, Tw, dsmall, texs)
        skeys = sorted(k for k, v in sub.TSDF_observed.d.items() if v > 0 and k[0] == s_)
        g[f"S{s_}_idx"] = np.array([k[1:] for k in skeys], np.int16)
        g[f"S{s_}_T"] = np.array([sub.TSDF.d[k] for k in skeys], np.float16)
        g[f"S{s_}_W"] = np.array([sub.W_TSDF.d[k] for k in skeys], np.float16)
        g[f"S{s_}_occ"] = np.array([sub.occupy.d.get(k, 0) for k in skeys], np.int8)
        g[f"S{s_}_color"] = np.array([sub.color.d[k] for k in skeys], np.float16)
        if s_ == 0:
            sub.switch_to_next_submap()
    glo.fuse_submaps(sub)
    gkeys = sorted(k for k, v in glo.TSDF_observed.d.items() if v > 0)
    g["FU_idx"] = np.array([k[1:] for k in gkeys], np.int16)
    g["FU_T"] = np.array([glo.TSDF.d[k] for k in gkeys], np.float16)
    g["FU_W"] = np.array([glo.W_TSDF.d[k] for k in gkeys], np.float16)
    g["FU_color"] = np.array([glo.color.d[k] for k in gkeys], np.float16)
    g["FU_base_R"], g["FU_base_T"], g["colAB"] = np.stack([b[0] for b in base]), np.stack([b[1] for b in base]), colAB
    print(f"texproj: textured fusion {len(gkeys)} voxels, {time.time() - t00:.0f}s")
    out = os.path.join(ROOT, "tests", "golden", "ref_exec_texproj.npz")
    np.savez_compressed(out, **g)
    c = g["color"].astype(np.float32)
    white = (np.abs(c - 1.0).max(1) < 2e-3).sum()
    print(f"texproj: {len(keys)} voxels, {white} white (x >= h -> texture[0,0]), {time.time() - t00:.0f}s -> {out}")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    {"topo": topo, "f32": f32_state, "sat": saturation, "texproj": texproj}.get(mode, main)()
