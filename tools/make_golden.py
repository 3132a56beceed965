#!/usr/bin/env python3
"""Generate the small committed fixtures under tests/golden/ (build container only).

1. `ri_new_crop.npz` - a spatial crop of the reference's shipped map
   /root/reference/data/ri_new_tsdf.npy (an `export_submap()` dict of a real flight,
   dense_tsdf.py:456-480; legacy key `voxel_size`), same per-voxel dtypes
   (int16 idx / f16 TSDF / f16 W / int8 occupy).  The full files (10 MB / 43 MB) are
   too large to commit; the crop keeps every voxel with index in a 96x96x52 window.
2. `golden.json` - values the ORACLE produces on (a) that crop and (b) seeded
   synthetic inputs; pins the oracle against silent drift and records the
   reference-fixture facts (voxel counts of the full files, schema).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleTSDF, OracleOctomap  # noqa: E402
from taichislam_b200 import synthetic as syn  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    g = {}
    new = np.load(os.path.join(REF, "ri_new_tsdf.npy"), allow_pickle=True).item()
    old = np.load(os.path.join(REF, "ri_tsdf.npy"), allow_pickle=True).item()
    g["reference_fixtures"] = {
        "ri_new_tsdf.npy": {"voxels": int(new["TSDF"].shape[0]), "voxel_size": float(new["voxel_size"]),
                            "map_scale": [float(x) for x in new["map_scale"]],
                            "num_voxel_per_blk_axis": int(new["num_voxel_per_blk_axis"]),
                            "dtypes": {k: str(v.dtype) for k, v in new.items() if hasattr(v, "dtype")}},
        "ri_tsdf.npy": {"voxels": int(old["TSDF"].shape[0]), "voxel_size": float(old["voxel_size"]),
                        "dtypes": {k: str(v.dtype) for k, v in old.items() if hasattr(v, "dtype")}},
    }
    idx = new["indices"].astype(np.int32)
    lo = np.array([-100, 0, -17])
    hi = lo + np.array([96, 96, 52])
    keep = np.all((idx >= lo) & (idx < hi), axis=1)
    crop = dict(indices=new["indices"][keep], TSDF=new["TSDF"][keep], W_TSDF=new["W_TSDF"][keep],
                occupy=new["occupy"][keep], map_scale=np.array(new["map_scale"]), voxel_size=np.float64(new["voxel_size"]),
                num_voxel_per_blk_axis=np.int64(new["num_voxel_per_blk_axis"]))
    np.savez_compressed(os.path.join(OUT, "ri_new_crop.npz"), **crop)
    n = int(keep.sum())
    # oracle on the crop: load -> count -> export round trip -> marching cubes -> surface export
    m = OracleTSDF(map_scale=list(new["map_scale"]), voxel_scale=float(new["voxel_size"]), num_voxel_per_blk_axis=16,
                   is_global_map=True)
    m.scatter(0, crop["indices"], crop["TSDF"].astype(np.float32), crop["W_TSDF"].astype(np.float32), crop["occupy"])
    assert m.count_active() == n
    gi, gt, gw, go = m.gather()
    ntri, verts, _ = m.marching_cubes(step=1, thres=5 * float(new["voxel_size"]))
    nsurf, sxyz, _ = m.surface()
    g["crop"] = {"voxels": n, "window_lo": lo.tolist(), "window_hi": hi.tolist(),
                 "mc_triangles": int(ntri), "mc_vertex_nansum": [float(x) for x in np.nansum(verts.astype(np.float64), 0)],
                 "mc_nan_vertices": int(np.isnan(verts).any(1).sum()),
                 "surface_voxels": int(nsurf), "nan_tsdf": int(np.isnan(gt).sum()),
                 "gather_idx_sha256": sha(gi)}
    # oracle on seeded synthetic inputs (canonical mode)
    syn_g = {}
    for name, depth in (("S1_plane3m", syn.scene_plane(3.0)), ("S2_sphere4m", syn.scene_sphere(4.0)), ("room", syn.scene_room())):
        o = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True)
        o.integrate_depth(np.eye(3), np.zeros(3), depth)
        i_, t_, w_, oc_ = o.gather()
        st = o.stats()
        syn_g[name] = {"stats": st, "active": int(i_.shape[0]), "idx_sha256": sha(i_),
                       "tsdf_sum": float(t_.astype(np.float64).sum()), "w_sum": float(w_.astype(np.float64).sum()),
                       "occ_sum": int(oc_.sum())}
    g["integrate_256"] = syn_g
    # texture (canonical colour rule, DESIGN.md "Texture"): two textured frames, colours of the observed voxels in
    # lexicographic voxel order; a coloured mesh; Octomap colours
    o = OracleTSDF(map_scale=[12.8, 12.8], K=syn.K_DEPTH, is_global_map=True, disp_ceiling=5.0)
    o.set_color(True, True)
    for q in range(2):
        R, T = syn.stream_pose(40 * q)
        o.integrate_depth_tex(R, T, syn.scene_room(), syn.texture_gradient(100 + q), commit=True)
    i_, t_, w_, oc_ = o.gather()
    col = o.gather_color(0)
    nt, v, nrm, vc = o.marching_cubes_color(1, 0.1)
    g["texture_256"] = {"active": int(i_.shape[0]), "idx_sha256": sha(i_), "color_sha256": sha(col),
                        "color_sum": [float(x) for x in col.astype(np.float64).sum(0)], "coloured": int((col[:, 0] > 0).sum()),
                        "mc_triangles": int(nt), "mc_color_sum": [float(x) for x in vc.astype(np.float64).sum(0)]}
    oo = OracleOctomap(map_scale=[12.8, 12.8], voxel_scale=0.05, K=2, min_occupy_thres=1, Kcam=syn.K_DEPTH, max_ray_length=5.0)
    oo.set_color(True, True)
    oo.set_submap_pose(0, np.eye(3), np.zeros(3))
    oo.integrate_depth_tex(np.eye(3), np.zeros(3), syn.scene_room(), syn.texture_gradient(102))
    oi, ocnt = oo.gather(0)
    g["octomap_texture"] = {"voxels": int(oi.shape[0]), "idx_sha256": sha(oi), "count_sha256": sha(ocnt), "color_sha256": sha(oo.gather_color(0))}
    oc = OracleOctomap(map_scale=[51.2, 51.2], voxel_scale=0.05, K=2, min_occupy_thres=2)
    oc.integrate_points(np.eye(3), np.zeros(3), syn.octo_cloud(100000, seed=1))
    oi, ocnt = oc.gather()
    g["octomap_c3"] = {"N": oc.N, "voxels": int(oi.shape[0]), "hits": int(ocnt.sum()), "max": int(ocnt.max()),
                       "idx_sha256": sha(oi), "cnt_sha256": sha(ocnt)}
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print(json.dumps(g, indent=1, sort_keys=True)[:2500])
    print("crop bytes", os.path.getsize(os.path.join(OUT, "ri_new_crop.npz")))


if __name__ == "__main__":
    main()
