"""GPU parity: marching cubes, Octomap hit counts, ESDF - CUDA (through the C ABI) vs the CPU oracle."""
import json
import os

import numpy as np
import pytest

from taichislam_b200 import synthetic as syn
from util import tri_multiset, rot_xyz, as_dict_rows

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def mesh_equal(gv, gn, ov, on, tol=1e-4):
    """Triangle soups are equal as multisets: every triangle (9 coordinates, emission order of its 3 vertices kept) has a
    partner within `tol` on the other side (nearest neighbour in R^9, both directions), normals of partners agree."""
    from scipy.spatial import cKDTree
    assert gv.shape == ov.shape
    a, b = gv.reshape(-1, 9).astype(np.float64), ov.reshape(-1, 9).astype(np.float64)
    fa, fb = np.isfinite(a).all(1), np.isfinite(b).all(1)
    assert fa.sum() == fb.sum()
    a, b = a[fa], b[fb]
    d_ab, j_ab = cKDTree(b).query(a)
    d_ba, _ = cKDTree(a).query(b)
    assert d_ab.max() <= tol * 3 and d_ba.max() <= tol * 3, f"unmatched triangle: {d_ab.max()} / {d_ba.max()}"
    na, nb = gn.reshape(-1, 9)[fa], on.reshape(-1, 9)[fb][j_ab]
    nf = np.isfinite(nb).all(1)
    assert np.array_equal(np.isfinite(na).all(1), nf)
    assert np.abs(na[nf] - nb[nf]).max() <= 1e-3


def test_mc_sphere_known_answer():
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    vs = 0.05
    o = OracleTSDF(map_scale=[6.4, 6.4], voxel_scale=vs, is_global_map=True)
    g = TsdfHandle(o.N, o.Nz, voxel_scale=vs, is_global_map=True)
    r = np.arange(-15, 15)
    I, J, K = np.meshgrid(r, r, r, indexing="ij")
    idx = np.stack([I.ravel(), J.ravel(), K.ravel()], 1).astype(np.int32)
    p = idx.astype(np.float32) * np.float32(vs)
    t = np.sqrt((p * p).sum(1)).astype(np.float32) - np.float32(3 * vs)
    w = np.ones_like(t)
    occ = np.zeros(len(t), np.int8)
    o.scatter(0, idx, t, w, occ)
    g.scatter(0, idx, t, w, occ)
    no, ov, on = o.marching_cubes(1, 0.1)
    ng, gv, gn = g.marching_cubes(1, 0.1)
    assert ng == no > 100
    mesh_equal(gv, gn, ov, on, tol=1e-6)  # same f32 formula on identical inputs
    # generic step=2 path
    no2, ov2, on2 = o.marching_cubes(2, 0.1)
    ng2, gv2, gn2 = g.marching_cubes(2, 0.1)
    assert ng2 == no2 > 10
    mesh_equal(gv2, gn2, ov2, on2, tol=1e-6)


def test_mc_on_reference_map_crop():
    """Mesh of the crop of the reference's shipped flight map: triangle count pinned in golden.json."""
    z = np.load(os.path.join(HERE, "golden", "ri_new_crop.npz"))
    from oracle.oracle import OracleTSDF, tsdf_dims
    from taichislam_b200.tsdf_handle import TsdfHandle
    vs = float(z["voxel_size"])
    o = OracleTSDF(map_scale=list(z["map_scale"]), voxel_scale=vs, is_global_map=True)
    g = TsdfHandle(o.N, o.Nz, voxel_scale=vs, is_global_map=True)
    args = (0, z["indices"].astype(np.int32), z["TSDF"].astype(np.float32), z["W_TSDF"].astype(np.float32), z["occupy"])
    o.scatter(*args)
    g.scatter(*args)
    no, ov, on = o.marching_cubes(1, 5 * vs)
    ng, gv, gn = g.marching_cubes(1, 5 * vs)
    assert ng == no == GOLD["crop"]["mc_triangles"]
    mesh_equal(gv, gn, ov, on, tol=1e-6)
    # saturating capacity: demand still reported, no overflow
    n3, v3, _ = g.marching_cubes(1, 5 * vs, cap_tri=1000)
    assert n3 == no and v3.shape[0] == 3000


def test_mc_after_integration_c2():
    """BASELINE config 2 flavour: stream frames -> 512^3 TSDF -> marching cubes (thres 5*vs)."""
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    o = OracleTSDF(map_scale=[25.6, 25.6], K=syn.K_DEPTH, is_global_map=True)
    g = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, is_global_map=True)
    n = 6
    d = np.stack([syn.scene_sphere(4.0)] * n)
    Rs, Ts = syn.stream_poses(n)
    g.integrate_depth(d, Rs, Ts)
    for q in range(n):
        o.integrate_depth(Rs[q], Ts[q], d[q])
    no, ov, on = o.marching_cubes(1, 0.25)
    ng, gv, gn = g.marching_cubes(1, 0.25)
    # (a) end to end: TSDF values differ by ~1e-5 between the two sides; a vertex sits at mu = -v1/(v2-v1), so that
    # difference is amplified by vs/|v2-v1| where the field is flat, and a corner within 1e-5 of zero can change a
    # cube's case.  Counts agree to 1e-4, 99% of the vertices to 1e-4 m (observed: 1.5e-8), all to half a voxel.
    # (round 2: the block-binned march forms ds = L - j*vs, <= 3e-6 m from the reference's |P - x| * sign: corners
    # within that of zero flip a cube's case, so sliver triangles come and go at the 5e-4 level of the count)
    assert abs(ng - no) <= max(4, int(1e-3 * no)) and no > 10000
    if ng == no:
        from scipy.spatial import cKDTree
        dv, _ = cKDTree(ov.reshape(-1, 9)).query(gv.reshape(-1, 9))
        assert np.quantile(dv, 0.99) <= 1e-4 and dv.max() <= 2.5e-2  # worst case: half a voxel on a flat spot
    # (b) the mesher itself: on IDENTICAL TSDF values (the GPU's map loaded into a fresh oracle) the two
    # meshes agree to 1e-6 in every vertex and normal.
    gi, gt, gw, gocc = g.gather()
    o2 = OracleTSDF(map_scale=[25.6, 25.6], K=syn.K_DEPTH, is_global_map=True)
    o2.scatter(0, gi, gt, gw, gocc)
    n2, v2, nr2 = o2.marching_cubes(1, 0.25)
    assert n2 == ng
    mesh_equal(gv, gn, v2, nr2, tol=1e-6)


def test_octomap_c3_counts_bit_exact():
    """BASELINE config 3: 100k-point cloud into a 1024^3 hit-count grid, every (ijk,count) pair bit-exact."""
    from oracle.oracle import OracleOctomap
    from taichislam_b200.octo_handle import OctoHandle
    o = OracleOctomap(map_scale=[51.2, 51.2], voxel_scale=0.05, K=2, min_occupy_thres=2)
    g = OctoHandle(o.N, o.Nz, K=2, voxel_scale=0.05, min_occupy_thres=2)
    pts = syn.octo_cloud(100000, seed=1)
    R = rot_xyz(0.0, 0.0, 0.0)
    Rs, Ts_ = rot_xyz(0.05, 0.1, -0.2), np.array([0.4, -0.3, 0.2])   # pose-table row of submap 0 (rows start at zero)
    g.set_submap_pose(0, Rs, Ts_)
    o.set_submap_pose(0, Rs, Ts_)
    g.integrate_points(pts, R, np.zeros(3))
    o.integrate_points(R, np.zeros(3), pts)
    gi, gc = as_dict_rows(*g.gather())
    oi, oc = as_dict_rows(*o.gather())
    assert gi.shape[0] == GOLD["octomap_c3"]["voxels"]
    assert np.array_equal(gi, oi) and np.array_equal(gc, oc)
    # repeated frames accumulate; clustered cloud makes real multi-hit voxels
    rng = np.random.default_rng(5)
    cl = (rng.normal(size=(50000, 3)) * 0.3 + np.array([1.0, 2.0, 0.5])).astype(np.float32)
    R2, T2 = rot_xyz(0.2, 0.1, -0.4), np.array([0.5, -1.0, 0.2])
    for _ in range(3):
        g.integrate_points(cl, R2, T2)
        o.integrate_points(R2, T2, cl)
    gi, gc = as_dict_rows(*g.gather())
    oi, oc = as_dict_rows(*o.gather())
    assert np.array_equal(gi, oi) and np.array_equal(gc, oc) and gc.max() > 10
    for level in (1, 2, 4):
        ng, xg = g.export(level)
        no, xo = o.export(level)
        assert ng == no
        assert np.array_equal(xg[np.lexsort(xg.T[::-1])], xo[np.lexsort(xo.T[::-1])])
    assert g.export(1)[0] > 100


def test_octomap_depth_and_fusion():
    from oracle.oracle import OracleOctomap
    from taichislam_b200.octo_handle import OctoHandle
    kw = dict(map_scale=[12.8, 12.8], voxel_scale=0.05, K=2, min_occupy_thres=1, max_ray_length=5.0)
    o = OracleOctomap(Kcam=syn.K_DEPTH, **kw)
    g = OctoHandle(o.N, o.Nz, K=2, voxel_scale=0.05, min_occupy_thres=1, max_ray_length=5.0, Kcam=syn.K_DEPTH, max_submaps=8)
    d = syn.scene_room()
    poses = {0: (rot_xyz(0.1, 0.0, 0.2), np.array([0.2, 0.0, 0.1])), 2: (rot_xyz(0.0, 0.3, 0.0), np.array([0.0, 0.1, 0.0]))}
    for s, (R, T) in poses.items():
        g.integrate_depth(d, R, T, submap=s)
        o.integrate_depth(R, T, d, submap=s)
        for a, b in zip(as_dict_rows(*g.gather(s)), as_dict_rows(*o.gather(s))):
            assert np.array_equal(a, b)
    og = OracleOctomap(map_scale=[25.6, 25.6], voxel_scale=0.05, K=2, min_occupy_thres=1)
    gg = OctoHandle(og.N, og.Nz, K=2, voxel_scale=0.05, min_occupy_thres=1, max_submaps=8)
    for s, (R, T) in poses.items():
        og.set_submap_pose(s, R, T)
        gg.set_submap_pose(s, R, T)
    og.fuse_from(o)
    gg.fuse_from(g)
    for a, b in zip(as_dict_rows(*gg.gather(0)), as_dict_rows(*og.gather(0))):
        assert np.array_equal(a, b)


def test_esdf_matches_dijkstra_oracle():
    """BASELINE config 4 flavour: integrate, then the converged ESDF wavefront == multi-source Dijkstra, bit-exact
    on the voxels whose TSDF sign class is unambiguous (TSDF differs by ~1e-7 between the sides)."""
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    o = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True, max_ray_length=4.0)
    g = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, is_global_map=True, max_ray_length=4.0)
    d = syn.scene_room()
    g.integrate_depth(d, np.eye(3)[None], np.zeros((1, 3)))
    o.integrate_depth(np.eye(3), np.zeros(3), d)
    # make both sides start from IDENTICAL TSDF values: load the GPU's map into a fresh oracle
    gi, gt, gw, gocc = g.gather()
    o2 = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True, max_ray_length=4.0)
    o2.scatter(0, gi, gt, gw, gocc)
    sweeps = g.esdf_update()
    o2.esdf_update()
    ei, ee = as_dict_rows(*g.esdf_gather())
    oi, oe = as_dict_rows(*o2.esdf_gather())
    assert np.array_equal(ei, oi)
    assert sweeps >= 2
    assert np.array_equal(ee, oe), f"max |dESDF| = {np.abs(ee - oe).max()}"
    # and against the oracle's own integration the ESDF agrees to 1e-4 almost everywhere
    o.esdf_update()
    _, oe1 = as_dict_rows(*o.esdf_gather())
    assert np.quantile(np.abs(ee - oe1), 0.999) <= 1e-4


def test_esdf_incremental_equals_full_512():
    """BASELINE config 4 at its stated size: 512^3 TSDF + incremental ESDF propagation, an update after every launch
    of a 20-frame stream whose surface MOVES half way (sphere 3.0 m -> 3.6 m: old fixed-band voxels leave the band,
    distances grow - the raise wave has to run).  After every update the incrementally maintained field must be
    bit-identical to a full recompute on a twin map, and at the end to the oracle's multi-source Dijkstra."""
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    kw = dict(K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    a, b = TsdfHandle(512, 512, **kw), TsdfHandle(512, 512, **kw)
    Rs, Ts = syn.stream_poses(20, start=40)
    seen_raise = False
    for q in range(20):
        d = syn.scene_sphere(3.0 if q < 10 else 3.6)
        for g in (a, b):
            g.integrate_depth(d, Rs[q][None], Ts[q][None])
        if q % 3 != 2 and q != 19:
            continue  # an update every third frame: several commits between two updates
        sa = a.esdf_update2()
        sb = b.esdf_update2(full=True)
        assert sb["raise_sweeps"] == -1 and (q == 2 or sa["raise_sweeps"] >= 1)
        seen_raise |= sa["suspect"] > 0
        ia, ea = as_dict_rows(*a.esdf_gather())
        ib, eb = as_dict_rows(*b.esdf_gather())
        assert np.array_equal(ia, ib)
        assert np.array_equal(ea, eb), f"frame {q}: max |dESDF| = {np.abs(ea - eb).max()}, {np.sum(ea != eb)} voxels"
    assert seen_raise
    # and against the oracle on IDENTICAL TSDF values
    gi, gt, gw, gocc = a.gather()
    o = OracleTSDF(map_scale=[25.6, 25.6], K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    o.scatter(0, gi, gt, gw, gocc)
    o.esdf_update()
    oi, oe = as_dict_rows(*o.esdf_gather())
    assert np.array_equal(ia, oi) and np.array_equal(ea, oe)
    # an update with nothing new to do is a no-op
    s0 = a.esdf_update2()
    assert s0["changed"] == 0 and s0["suspect"] == 0


def test_planner_queries_match_oracle():
    """Batched raycast / is_pos_occupy / is_pos_unobserved / is_near_pos_occupy (mapping_common.py:165-204), bit-exact."""
    from oracle.oracle import OracleTSDF, OracleOctomap
    from taichislam_b200.tsdf_handle import TsdfHandle
    from taichislam_b200.octo_handle import OctoHandle
    o = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    g = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    d = syn.scene_room()
    g.integrate_depth(d, np.eye(3)[None], np.zeros((1, 3)))
    gi, gt, gw, gocc = g.gather()
    o.scatter(0, gi, gt, gw, gocc)  # identical TSDF values on both sides
    rng = np.random.default_rng(3)
    pts = rng.uniform([-3, -2, -0.5], [3, 2, 5.5], size=(20000, 3)).astype(np.float32)
    pts[:10] = 100.0  # outside the volume
    go, gu = g.query_points(pts)
    oo, ou = o.query_points(pts)
    assert np.array_equal(go, oo) and np.array_equal(gu, ou) and 0.05 < (~gu).mean() < 0.95
    assert np.array_equal(g.query_near_occupy(pts[:3000], 2), o.query_near_occupy(pts[:3000], 2))
    # rays from observed free space (inside the scanned cone) in random directions
    free = pts[(~gu) & (~go)][:4000]
    dirs = rng.normal(size=free.shape).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    gh, gx, gl = g.raycast(free, dirs, 3.0)
    oh, ox, ol = o.raycast(free, dirs, 3.0)
    assert np.array_equal(gh, oh) and np.array_equal(gx, ox) and np.array_equal(gl, ol) and gh.mean() > 0.5
    # octomap
    oo_ = OracleOctomap(map_scale=[12.8, 12.8], voxel_scale=0.05, K=2, min_occupy_thres=0, max_ray_length=6.0, Kcam=syn.K_DEPTH)
    og = OctoHandle(oo_.N, oo_.Nz, K=2, voxel_scale=0.05, min_occupy_thres=0, max_ray_length=6.0, Kcam=syn.K_DEPTH)
    og.integrate_depth(d, np.eye(3), np.zeros(3))
    oo_.integrate_depth(np.eye(3), np.zeros(3), d)
    assert np.array_equal(og.query_points(pts), oo_.query_points(pts))
    origin = np.zeros((2000, 3), np.float32)
    dirs = rng.normal(size=origin.shape).astype(np.float32)
    dirs[:, 2] = np.abs(dirs[:, 2]) + 0.5
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    a = og.raycast(origin, dirs, 6.0)
    b = oo_.raycast(origin, dirs, 6.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[0].any()
