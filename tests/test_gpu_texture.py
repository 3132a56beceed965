"""GPU parity tests of the texture (colour) path vs the CPU oracle.

The reference overwrites color[xi] with the ray's mean colour at every sample, racing rays, last writer wins
(dense_tsdf.py:268-269).  Both sides implement the same deterministic rule (DESIGN.md "Texture"): the winner is
the sample with the largest word [frame sequence | closeness to the ray's surface point | rgb 3x10 bits], so the
committed colours are compared EXACTLY (they are 10-bit quantised values decoded by the same formula)."""
import numpy as np
import pytest

from taichislam_b200 import synthetic as syn
from util import as_dict_rows, compare_voxels, key_sort, rot_xyz

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_texture(seed, h=480, w=640):
    return syn.texture_gradient(seed, h, w)


def make_pair(map_scale, same_proj=True, Kcolor=None, **kw):
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    gpu_only = {k: kw.pop(k) for k in ("max_submaps", "max_blocks", "max_image_pixels", "max_points") if k in kw}
    o = OracleTSDF(map_scale=map_scale, K=syn.K_DEPTH, **kw)
    o.set_color(True, same_proj, Kcolor)
    g = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, texture_enabled=True, **kw, **gpu_only)
    g.set_color_intrinsics(Kcolor if Kcolor is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1], same_proj)
    return g, o


def compare_colors(g, o, submap=0, exact=True, tol=1e-4):
    gi, gt, gw, go, gc = g.gather(submap, color=True)
    oi, ot, ow, oo = o.gather(submap)
    oc = o.gather_color(submap)
    compare_voxels((gi, gt, gw, go), (oi, ot, ow, oo), TOL)
    gc, oc = gc[key_sort(gi)], oc[key_sort(oi)]
    fin = np.isfinite(oc).all(1)
    assert np.array_equal(np.isfinite(gc).all(1), fin)
    if exact:
        bad = np.nonzero((gc[fin] != oc[fin]).any(1))[0]
        assert bad.size == 0, f"{bad.size} of {len(gc)} voxel colours differ, first: gpu {gc[fin][bad[0]]} oracle {oc[fin][bad[0]]}"
    else:
        assert np.abs(gc[fin] - oc[fin]).max(initial=0.0) <= tol
    assert (oc[fin] > 0).any()
    return gc


def test_textured_frame_same_proj():
    g, o = make_pair([25.6, 25.6], is_global_map=True)
    d, tex = syn.scene_room(), make_texture(1)
    R, T = rot_xyz(0.1, -0.05, 0.3), np.array([0.2, -0.1, 0.05])
    g.integrate_depth(d, R[None], T[None], texture=tex[None])
    o.integrate_depth_tex(R, T, d, tex)
    sg, so = g.stats(), o.stats()
    for k in ("n_px", "n_valid", "n_rays", "n_updates", "n_oob"):
        assert sg[k] == so[k], k
    compare_colors(g, o)


def test_textured_frame_color_camera():
    """color_same_proj=False: colour pixel through color_ind_from_depth_pt (mapping_common.py:43-58) incl. its
    swapped bound test; a smaller colour image with its own intrinsics."""
    th, tw = 240, 320
    Kc = [s * 0.5 for s in syn.K_DEPTH]
    Kc[8] = 1.0
    g, o = make_pair([25.6, 25.6], same_proj=False, Kcolor=Kc, is_global_map=True)
    d, tex = syn.scene_sphere(4.0), make_texture(2, th, tw)
    g.integrate_depth(d, np.eye(3)[None], np.zeros((1, 3)), texture=tex[None])
    o.integrate_depth_tex(np.eye(3), np.zeros(3), d, tex)
    compare_colors(g, o)


def test_textured_stream_later_frames_win():
    """10 frames in ONE batched launch == the oracle integrating them one after the other: colours of voxels seen by
    several frames come from the latest one."""
    n = 10
    g, o = make_pair([25.6, 25.6], is_global_map=True)
    d = np.stack([syn.scene_sphere(4.0)] * n)
    tex = np.stack([make_texture(10 + q) for q in range(n)])
    Rs, Ts = syn.stream_poses(n, start=3)
    g.integrate_depth(d, Rs, Ts, texture=tex)
    for q in range(n):
        o.integrate_depth_tex(Rs[q], Ts[q], d[q], tex[q], commit=(q == n - 1))
    compare_colors(g, o)
    # device-resident inputs take the same path
    import torch
    g2, _ = make_pair([25.6, 25.6], is_global_map=True)
    g2.integrate_depth(torch.from_numpy(d.view(np.int16)).cuda().view(torch.uint16), Rs, Ts, texture=torch.from_numpy(tex).cuda())
    a, b = g.gather(0, color=True), g2.gather(0, color=True)
    assert np.array_equal(a[4][key_sort(a[0])], b[4][key_sort(b[0])])


def test_textured_points():
    rng = np.random.default_rng(5)
    n = 60000
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    xyz = (dirs * rng.uniform(1.0, 6.0, (n, 1))).astype(np.float32)
    rgb = rng.integers(1, 256, (n, 3)).astype(np.uint8)
    R, T = rot_xyz(0.2, 0.1, -0.3), np.array([0.1, 0.2, -0.1])
    g, o = make_pair([25.6, 25.6], is_global_map=True, max_points=n)
    g.integrate_points(xyz, R, T, rgb=rgb)
    o.integrate_points_rgb(R, T, xyz, rgb)
    compare_colors(g, o)


def test_color_roundtrip_surface_and_mesh():
    from scipy.spatial import cKDTree
    g, o = make_pair([12.8, 12.8], is_global_map=True, disp_ceiling=5.0)  # the sphere cap lies at z = 2.4 .. 3 m
    d, tex = syn.scene_sphere(3.0), make_texture(7)
    g.integrate_depth(d, np.eye(3)[None], np.zeros((1, 3)), texture=tex[None])
    o.integrate_depth_tex(np.eye(3), np.zeros(3), d, tex)
    # surface exporter takes the voxel colours (dense_tsdf.py:360-362)
    ng, gx, gc = g.surface(0)
    no, ox, oc = o.surface(0)
    assert ng == no > 1000
    og, oo = np.lexsort(gx.T[::-1]), np.lexsort(ox.T[::-1])
    assert np.allclose(gx[og], ox[oo], atol=1e-6) and np.array_equal(gc[og], oc[oo])
    # coloured marching cubes (vertexInterp_color, marching_cube_mesher.py:62-82)
    ng, gv, gn, gcol = g.marching_cubes(1, 0.1, color=True)
    no, ov, on, ocol = o.marching_cubes_color(1, 0.1)
    assert ng == no > 1000
    a, b = gv.reshape(-1, 9).astype(np.float64), ov.reshape(-1, 9).astype(np.float64)
    dist, j = cKDTree(b).query(a)
    assert dist.max() <= 3e-4
    # TSDF values differ by <=1e-4 between the two sides: interpolation weights differ slightly and a corner value within
    # 1e-6 of zero can flip vertexInterp_color's snap branch (mu = 0) -> a loose statistical check here, the strict one
    # on identical inputs below
    dc = np.abs(gcol.reshape(-1, 9) - ocol.reshape(-1, 9)[j])
    assert np.quantile(dc, 0.999) <= 2e-2 and np.median(dc) <= 1e-4
    # scatter the GPU state into a second pair -> identical inputs -> colours agree to f32 rounding
    gi, gt, gw, gocc, gc = g.gather(0, color=True)
    g2, o2 = make_pair([12.8, 12.8], is_global_map=True)
    g2.scatter(0, gi, gt, gw, gocc, color=gc)
    o2.scatter(0, gi, gt, gw, gocc)
    o2.scatter_color(0, gi, gc)
    compare_colors(g2, o2)
    # A vertex colour depends on the EDGE it was made on, not only on its position (a snapped vertex takes the colour of
    # the edge's first corner, mu = 0), and neighbouring cells emit coincident degenerate triangles: match triangles on
    # positions AND colours (18-D), both directions.
    def same_coloured_mesh(gv, gcol, ov, ocol):
        a = np.concatenate([gv.reshape(-1, 9), gcol.reshape(-1, 9)], 1).astype(np.float64)
        b = np.concatenate([ov.reshape(-1, 9), ocol.reshape(-1, 9)], 1).astype(np.float64)
        assert a.shape == b.shape
        d_ab, _ = cKDTree(b).query(a)
        d_ba, _ = cKDTree(a).query(b)
        assert d_ab.max() <= 1e-5 and d_ba.max() <= 1e-5, (d_ab.max(), d_ba.max())

    ng, gv, gn, gcol = g2.marching_cubes(1, 0.1, color=True)
    no, ov, on, ocol = o2.marching_cubes_color(1, 0.1)
    assert ng == no > 1000
    same_coloured_mesh(gv, gcol, ov, ocol)
    ng2, gv2, gn2, gcol2 = g2.marching_cubes(2, 0.1, color=True)  # generic-step path
    no2, ov2, on2, ocol2 = o2.marching_cubes_color(2, 0.1)
    assert ng2 == no2 > 100
    same_coloured_mesh(gv2, gcol2, ov2, ocol2)


def test_textured_fusion():
    """fuse_submaps with colours: weighted colour average (dense_tsdf.py:276-277)."""
    from oracle.oracle import OracleTSDF
    from taichislam_b200.tsdf_handle import TsdfHandle
    g, o = make_pair([12.8, 12.8], is_global_map=False, max_ray_length=5.0)
    gg, og = make_pair([12.8, 12.8], is_global_map=True, max_ray_length=5.0)
    d = syn.scene_sphere(3.0)
    for s in range(2):
        R, T = rot_xyz(0.02 * s, 0.1 * s, 0.0), np.array([0.13 * s, 0.0, 0.07 * s])
        # generic submap poses: with an axis-aligned pose every voxel lands ON a global grid point, all 7 splatted
        # corners get weight 0 and the fused map is 0/0 = NaN (reference quirk, dense_tsdf.py:300)
        Rs, Ts = rot_xyz(0.1, 0.2 - 0.3 * s, 0.3 + 0.7 * s), np.array([0.5 - 0.9 * s, 0.1 + 0.5 * s, -0.2 + 0.5 * s])
        for m in (g, o, gg, og):
            m.set_submap_pose(s, Rs, Ts)
        tex = make_texture(20 + s)
        g.integrate_depth(d, R[None], T[None], submaps=[s], texture=tex[None])
        o.integrate_depth_tex(R, T, d, tex, submap=s)
    gg.fuse_from(g)
    og.fuse_from(o)
    gi, gt, gw, gocc, gc = gg.gather(0, color=True)
    oi, ot, ow, oocc = og.gather(0)
    oc = og.gather_color(0)
    kg, ko = key_sort(gi), key_sort(oi)
    assert np.array_equal(gi[kg], oi[ko])
    # the reference's sequential RMW poisons a voxel with NaN when a zero-weight corner arrives first (0*NaN,
    # dense_tsdf.py:275-277); the order-free sums only yield NaN when the TOTAL weight is zero
    fin = np.isfinite(ot[ko]) & np.isfinite(oc[ko]).all(1)
    assert fin.mean() > 0.99
    assert np.abs(gt[kg][fin] - ot[ko][fin]).max() <= TOL
    assert np.abs(gc[kg][fin] - oc[ko][fin]).max() <= 2e-4
    assert (oc[ko][fin] > 0).mean() > 0.5


def test_dense_tsdf_class_textured(tmp_path):
    """The reference class surface: texture_enabled=True, recast_depth_to_map(R, T, depth, texture), export/load with
    the colour column, coloured mesh."""
    from oracle.oracle import OracleTSDF
    from taichislam_b200.mapping import DenseTSDF, MarchingCubeMesher
    m = DenseTSDF(map_scale=[12.8, 12.8], texture_enabled=True, is_global_map=True, max_ray_length=5.0)
    m.set_dep_camera_intrinsic(np.array(syn.K_DEPTH).reshape(3, 3))
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    o = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True, max_ray_length=5.0)
    o.set_color(True, True)
    d = syn.scene_sphere(3.0)
    for q in range(3):
        R, T = syn.stream_pose(q)
        tex = make_texture(30 + q)
        m.recast_depth_to_map(R, T, d, tex)
        o.integrate_depth_tex(R, T, d, tex, commit=(q == 2))
    obj = m.export_submap()
    assert obj["color"].shape == (m.count_active(), 3) and obj["color"].dtype == np.float16
    oi, ot, ow, oo = o.gather(0)
    oc = o.gather_color(0)
    gi = obj["indices"].astype(np.int32)
    assert np.array_equal(gi[key_sort(gi)], oi[key_sort(oi)])
    assert np.abs(obj["color"][key_sort(gi)].astype(np.float32) - oc[key_sort(oi)]).max() <= 1e-3  # f16 column
    fn = str(tmp_path / "tex_map.npy")
    m.saveMap(fn)
    m2 = DenseTSDF.loadMap(fn)
    assert m2.enable_texture
    obj2 = m2.export_submap()
    g1, g2 = obj["indices"].astype(np.int32), obj2["indices"].astype(np.int32)
    assert np.array_equal(obj["color"][key_sort(g1)], obj2["color"][key_sort(g2)])
    mesher = MarchingCubeMesher(m, max_triangles=400000, tsdf_surface_thres=0.1)
    mesher.generate_mesh(1)
    n = int(mesher.num_facelets[None])
    assert n > 1000
    col = mesher.mesh_colors.to_numpy()[:3 * n]
    assert (col[:, 0] > 0).mean() > 0.9 and col.max() <= 1.0


def test_untextured_map_rejects_texture():
    from taichislam_b200.tsdf_handle import TsdfHandle
    g = TsdfHandle(256, 256, K=syn.K_DEPTH, is_global_map=True)
    with pytest.raises(Exception):
        g.integrate_depth(syn.scene_plane(3.0), np.eye(3)[None], np.zeros((1, 3)), texture=make_texture(0)[None])


# ---------------------------------------------------------------------------------------------------------------
# Octomap colours (taichi_octomap.py:77-79, :120-124, :160-167, :189): latest integrate call wins, inside a call the
# largest packed RGB (after the BGR->RGB swap).  8-bit channels / 255 on both sides -> exact comparison.
# ---------------------------------------------------------------------------------------------------------------
def octo_pair(map_scale, same_proj=True, Kcolor=None, **kw):
    from oracle.oracle import OracleOctomap
    from taichislam_b200.octo_handle import OctoHandle
    o = OracleOctomap(map_scale=map_scale, voxel_scale=0.05, Kcam=syn.K_DEPTH, **kw)
    o.set_color(True, same_proj, Kcolor)
    g = OctoHandle(o.N, o.Nz, voxel_scale=0.05, Kcam=syn.K_DEPTH, max_submaps=8, texture_enabled=True, **kw)
    g.set_color_intrinsics(Kcolor if Kcolor is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1], same_proj)
    return g, o


def octo_equal(g, o, submap=0):
    gi, gc, gcol = g.gather(submap, color=True)
    oi, oc = o.gather(submap)
    ocol = o.gather_color(submap)
    kg, ko = key_sort(gi), key_sort(oi)
    assert np.array_equal(gi[kg], oi[ko]) and np.array_equal(gc[kg], oc[ko])
    assert np.array_equal(gcol[kg], ocol[ko])
    assert (ocol > 0).any()


def test_octomap_colors_points_depth_fusion():
    rng = np.random.default_rng(11)
    g, o = octo_pair([12.8, 12.8], K=2, min_occupy_thres=1, max_ray_length=5.0)
    # clustered cloud: many points per voxel race for the colour; two calls: the later one wins where it hits
    cl = (rng.normal(size=(40000, 3)) * 0.3 + np.array([1.0, 0.5, 0.5])).astype(np.float32)
    for q in range(2):
        rgb = rng.integers(0, 256, (len(cl), 3)).astype(np.uint8)
        R, T = rot_xyz(0.1 * q, 0.05, -0.2), np.array([0.1, -0.2 * q, 0.05])
        g.set_submap_pose(0, np.eye(3), np.zeros(3))
        o.set_submap_pose(0, np.eye(3), np.zeros(3))
        g.integrate_points(cl, R, T, rgb=rgb)
        o.integrate_points_rgb(R, T, cl, rgb)
    octo_equal(g, o)
    # depth + colour image into submaps 1 and 2, same-projection and colour-camera variants
    d = syn.scene_room()
    for s in (1, 2):
        R, T = rot_xyz(0.0, 0.1 * s, 0.2), np.array([0.1 * s, 0.0, 0.1])
        g.set_submap_pose(s, rot_xyz(0.1, 0.0, 0.1 * s), np.array([0.5, 0.1 * s, 0.0]))  # pose rows start at zero
        o.set_submap_pose(s, rot_xyz(0.1, 0.0, 0.1 * s), np.array([0.5, 0.1 * s, 0.0]))
        tex = make_texture(40 + s)
        g.integrate_depth(d, R, T, submap=s, texture=tex)
        o.integrate_depth_tex(R, T, d, tex, submap=s)
        octo_equal(g, o, s)
    for level in (1, 2):
        ng, xg, cg = g.export(level, submap=1, color=True)
        no, xo, co = o.export_color(level, submap=1)
        assert ng == no > 100
        a, b = np.lexsort(xg.T[::-1]), np.lexsort(xo.T[::-1])
        assert np.array_equal(xg[a], xo[b]) and np.array_equal(cg[a], co[b])
    # fusion into a textured global map: counts add, colour = most recently integrated source voxel
    gg, og = octo_pair([25.6, 25.6], K=2, min_occupy_thres=1)
    for s in range(3):
        Rs, Ts = rot_xyz(0.05 * s, 0.0, 0.3 * s), np.array([0.3 * s, -0.2 * s, 0.0])
        gg.set_submap_pose(s, Rs, Ts)
        og.set_submap_pose(s, Rs, Ts)
    gg.fuse_from(g)
    og.fuse_from(o)
    octo_equal(gg, og)


def test_octomap_color_camera_and_class():
    th, tw = 240, 320
    Kc = [s * 0.5 for s in syn.K_DEPTH]
    Kc[8] = 1.0
    g, o = octo_pair([12.8, 12.8], same_proj=False, Kcolor=Kc, K=2, min_occupy_thres=1, max_ray_length=5.0)
    d, tex = syn.scene_sphere(3.0), make_texture(50, th, tw)
    g.integrate_depth(d, np.eye(3), np.zeros(3), texture=tex)
    o.integrate_depth_tex(np.eye(3), np.zeros(3), d, tex)
    octo_equal(g, o)
    # reference class surface
    from taichislam_b200.mapping import Octomap
    m = Octomap(map_scale=[12.8, 12.8], voxel_scale=0.05, min_occupy_thres=0, texture_enabled=True, max_ray_length=5.0,
                max_disp_particles=200000)
    m.set_dep_camera_intrinsic(np.array(syn.K_DEPTH).reshape(3, 3))
    m.set_base_pose_submap(0, np.eye(3), np.zeros(3))
    tex2 = make_texture(51)
    m.recast_depth_to_map(np.eye(3), np.zeros(3), d, tex2)
    x, c = m.get_occupy_voxels(1)
    n = int(m.num_export_particles[None])
    assert n > 1000
    # exported colours are texture colours with the channels swapped (BGR -> RGB)
    got = set(map(tuple, np.round(c[:n] * 255).astype(np.int32)))
    have = set(map(tuple, tex2.reshape(-1, 3)[:, ::-1].astype(np.int32)))
    assert got <= have
