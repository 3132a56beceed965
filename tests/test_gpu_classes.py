"""GPU tests of the reference-facing Python class surface (taichislam_b200.mapping / taichi_slam.mapping):
same constructor kwargs, methods and field-like attributes as the reference classes, results checked against
the oracle.  These are the calls scripts/taichislam_node.py and SubmapMapping make."""
import os

import numpy as np
import pytest

from taichislam_b200 import synthetic as syn
from util import compare_voxels, as_dict_rows, rot_xyz

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_star_import_namespace():
    ns = {}
    exec("from taichi_slam.mapping import *", ns)
    for name in ("DenseTSDF", "Octomap", "SubmapMapping", "BaseMap", "MarchingCubeMesher", "TopoGraphGen", "ti", "np", "math", "time"):
        assert name in ns
    ns["ti"].init(arch=ns["ti"].cuda, dynamic_index=True, debug=False, device_memory_GB=4)  # taichislam_node.py:34
    with pytest.raises(RuntimeError):
        ns["ti"].init(arch=ns["ti"].cpu)


def test_dense_tsdf_per_frame_api_matches_oracle():
    from taichi_slam.mapping import DenseTSDF, MarchingCubeMesher
    from oracle.oracle import OracleTSDF
    kw = dict(map_scale=[25.6, 25.6], voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True, max_ray_length=5.1)
    m = DenseTSDF(max_disp_particles=1 << 20, **kw)
    assert (m.N, m.Nz, m.block_num_xy) == (512, 512, 32)
    m.set_color_camera_intrinsic(syn.K_DEPTH)
    m.set_dep_camera_intrinsic(syn.K_DEPTH)
    o = OracleTSDF(K=syn.K_DEPTH, **kw)
    base_R, base_T = rot_xyz(0.0, 0.0, 0.3), np.array([1.0, -0.5, 0.2])
    # DenseTSDF always enables the submap pose table (dense_tsdf.py:80 -> mapping_common.py:102-103), whose rows start
    # at ZERO: like the reference, poses are meaningless until the active submap's base pose is set.
    m.set_base_pose_submap(0, base_R, base_T)
    d = syn.scene_room()
    n = 70  # crosses the queue's launch boundaries (a launch every TSLAM_MAX_BATCH/2 = 32 frames)
    for q in range(n):
        R, T = syn.stream_pose(q)
        Rw, Tw = base_R @ R, base_R @ T + base_T
        m.recast_depth_to_map(Rw, Tw, d, np.array([]))
        # set_pose = convert_by_base in f64, then f32 (mapping_common.py:149-156)
        Ri = (base_R.T @ Rw).astype(np.float32)
        Ti = (base_R.T @ (Tw - base_T)).astype(np.float32)
        # the class queues frames and commits once per launch of 32 frames; the oracle commits at the same points
        # (commit granularity only matters for voxels saturated at Wmax=1000, dense_tsdf.py:267 - see DESIGN.md)
        o.integrate_depth(Ri, Ti, d, commit=(q % 32 == 31 or q == n - 1))
    assert m.count_active() == o.count_active()
    num = m.count_active()
    idx = np.zeros((num, 3), np.int16); t = np.zeros(num, np.float16); w = np.zeros(num, np.float16); occ = np.zeros(num, np.int8)
    m.to_numpy(idx, t, w, occ, np.array([]))
    oi, ot, ow, oo = o.gather()
    gi, gt, gw, go = as_dict_rows(idx.astype(np.int32), t, w, occ)
    oi, ot, ow, oo = as_dict_rows(oi, ot, ow, oo)
    assert np.array_equal(gi, oi) and np.array_equal(go.astype(np.int32), oo)
    # exported dtype is f16 (dense_tsdf.py:459-461): within one f16 rounding of the oracle's f32 value
    assert np.all(np.abs(gt.astype(np.float32) - ot) <= np.maximum(np.abs(ot), 6e-5) * 2.0 ** -10 + 1e-4)
    # surface export through the field-like attributes the node reads (taichislam_node.py:346-349)
    m.cvt_TSDF_surface_to_voxels()
    par_count = m.num_TSDF_particles[None]
    no, xo, co = o.surface()
    assert par_count == no
    xyz = m.export_TSDF_xyz.to_numpy()[:par_count]
    assert np.array_equal(xyz[np.lexsort(xyz.T[::-1])], xo[np.lexsort(xo.T[::-1])])
    m.export_TSDF_xyz[0] = [1.0, 2.0, 3.0]  # tests/gen_topo_graph.py:64-66 writes fields directly
    assert np.allclose(m.export_TSDF_xyz[0], [1, 2, 3])
    # mesher with the node's threshold (taichislam_node.py:209)
    mesher = MarchingCubeMesher(m, 300000, tsdf_surface_thres=0.05 * 5)
    mesher.generate_mesh(1)
    ntri = mesher.num_facelets[None]
    no_tri, _, _ = o.marching_cubes(1, 0.25)
    assert abs(ntri - no_tri) <= max(4, int(1e-4 * no_tri)) and ntri > 1000
    assert mesher.vertice_num() == 3 * ntri and mesher.num_vertices[None] == 3 * min(ntri, 300000)
    v = mesher.mesh_vertices.to_numpy()
    assert np.all(v[3 * ntri:] == -1000000.0) and np.all(np.abs(v[:3 * ntri]) < 20.0)


def test_save_load_roundtrip_and_legacy_fixture_key(tmp_path):
    """saveMap / loadMap (dense_tsdf.py:482-498) incl. the shipped fixtures' legacy `voxel_size` key."""
    from taichi_slam.mapping import DenseTSDF
    z = np.load(os.path.join(HERE, "golden", "ri_new_crop.npz"))
    legacy = {"indices": z["indices"], "TSDF": z["TSDF"], "W_TSDF": z["W_TSDF"], "color": np.array([]), "occupy": z["occupy"],
              "map_scale": [float(x) for x in z["map_scale"]], "voxel_size": float(z["voxel_size"]), "texture_enabled": False,
              "num_voxel_per_blk_axis": int(z["num_voxel_per_blk_axis"])}
    fn = str(tmp_path / "legacy.npy")
    np.save(fn, legacy)
    m = DenseTSDF.loadMap(fn)
    assert m.count_active() == z["TSDF"].shape[0]
    obj = m.export_submap()
    assert obj["indices"].dtype == np.int16 and obj["TSDF"].dtype == np.float16 and obj["occupy"].dtype == np.int8
    a = as_dict_rows(obj["indices"].astype(np.int32), obj["TSDF"], obj["W_TSDF"], obj["occupy"])
    b = as_dict_rows(z["indices"].astype(np.int32), z["TSDF"], z["W_TSDF"], z["occupy"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint16), b[1].view(np.uint16))
    assert np.array_equal(a[2].view(np.uint16), b[2].view(np.uint16)) and np.array_equal(a[3], b[3])
    fn2 = str(tmp_path / "again.npy")
    m.saveMap(fn2)
    m2 = DenseTSDF.loadMap(fn2)
    assert m2.count_active() == m.count_active()


def test_submap_mapping_flow():
    """SubmapMapping drives submap creation, PGO correction, export/wire format and local->global fusion."""
    from taichi_slam.mapping import SubmapMapping, DenseTSDF
    from oracle.oracle import OracleTSDF
    sub = dict(map_scale=[12.8, 12.8], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=5.1, max_submap_num=16,
               max_disp_particles=1 << 18)
    glo = dict(map_scale=[25.6, 25.6], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=5.1, max_submap_num=16,
               max_disp_particles=1 << 20, disp_ceiling=4.0)
    sm = SubmapMapping(DenseTSDF, sub_opts=sub, global_opts=glo, keyframe_step=3)
    sent = []
    sm.map_send_handle = sent.append
    sm.traj_send_handle = lambda b: None
    sm.set_dep_camera_intrinsic(syn.K_DEPTH)
    sm.set_color_camera_intrinsic(syn.K_DEPTH)
    so = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, max_ray_length=5.1)
    go = OracleTSDF(map_scale=[25.6, 25.6], is_global_map=True, max_ray_length=5.1)
    d = syn.scene_sphere(3.0)
    ext = (np.eye(3), np.zeros(3))
    sid, bases = -1, {}
    for fid in range(7):
        R, T = rot_xyz(0.02 * fid, -0.03 * fid, 0.1 * fid), np.array([0.2 * fid, 0.1, 0.05 * fid])
        sm.recast_depth_to_map_by_frame(fid, True, (R, T), ext, d, np.array([]))
        if fid % 3 == 0:
            sid += 1
            bases[sid] = (R, T)
            so.set_submap_pose(sid, R, T)
            go.set_submap_pose(sid, R, T)
        bR, bT = bases[sid]
        so.integrate_depth((bR.T @ R).astype(np.float32), (bR.T @ (T - bT)).astype(np.float32), d, submap=sid)
    assert sm.submap_collection.get_active_submap_id() == 2 and len(sent) == 2
    sm.submap_collection.finalization_current_submap()  # flushes the frame queue (frame 6 is still queued)
    for s in range(3):
        compare_voxels(sm.submap_collection._h.gather(s), so.gather(s), 1e-4)
    # wire format round trip of the first exported submap (zlib level-1 of np.save(dict), submap_mapping.py:226-233)
    pkt = SubmapMapping._unpack(sent[0])
    assert pkt["frame_id"] == 0 and pkt["indices"].dtype == np.int16 and pkt["TSDF"].dtype == np.float16
    # local -> global fusion of all three submaps
    sm.local_to_global()
    go.fuse_from(so)
    gi, gt, gw, _ = as_dict_rows(*sm.global_map._h.gather(0))
    oi, ot, ow, _ = as_dict_rows(*go.gather(0))
    assert np.array_equal(gi, oi)
    fin = np.isfinite(ot)
    assert np.abs(gt[fin] - ot[fin]).max() <= 1e-4
    sm.cvt_TSDF_surface_to_voxels()
    assert sm.num_TSDF_particles[None] > 100
    # a remote submap arrives over the wire and lands at the top of the table (dense_tsdf.py:500-515)
    sm.input_remote_submap(sent[1])
    assert sm.submap_collection.remote_submap_num[None] == 1


def test_octomap_class_surface():
    from taichi_slam.mapping import Octomap
    from oracle.oracle import OracleOctomap
    m = Octomap(map_scale=[51.2, 51.2], voxel_scale=0.05, K=2, min_occupy_thres=2, max_disp_particles=1 << 20)
    o = OracleOctomap(map_scale=[51.2, 51.2], voxel_scale=0.05, K=2, min_occupy_thres=2)
    assert (m.Rxy, m.N) == (10, 1024)
    bR, bT = rot_xyz(0.0, 0.0, 0.5), np.array([0.2, 0.1, 0.0])
    m.set_base_pose_submap(0, bR, bT)   # pose-table rows start at zero (mapping_common.py:106-107)
    o.set_submap_pose(0, bR, bT)
    rng = np.random.default_rng(11)
    pts = (rng.normal(size=(60000, 3)) * 0.25 + np.array([0.5, 1.0, 0.2]))
    R, T = rot_xyz(0.1, 0.2, 0.3), np.array([0.3, 0.2, 0.1])
    for _ in range(4):
        m.recast_pcl_to_map(R, T, pts, np.array([]), pts.shape[0])  # f64 cloud like ros_numpy's; computed in f32
        o.integrate_points((bR.T @ R).astype(np.float32), (bR.T @ (T - bT)).astype(np.float32), pts.astype(np.float32))
    gi, gc = as_dict_rows(*m.occupy)
    oi, oc = as_dict_rows(*o.gather())
    assert np.array_equal(gi, oi) and np.array_equal(gc, oc)
    m.cvt_occupy_to_voxels(1)
    n = m.num_export_particles[None]
    no, xo = o.export(1)
    assert n == no > 100
    x = m.export_x.to_numpy()[:n]
    assert np.array_equal(x[np.lexsort(x.T[::-1])], xo[np.lexsort(xo.T[::-1])])
