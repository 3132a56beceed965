import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from taichislam_b200 import synthetic as syn
from taichislam_b200.tsdf_handle import TsdfHandle
from oracle.oracle import OracleTSDF
from util import as_dict_rows, rot_xyz

def worst(g, o, tag):
    gi, gt, gw, go = as_dict_rows(*g.gather())
    oi, ot, ow, oo = as_dict_rows(*o.gather())
    print(tag, 'counts', gi.shape[0], oi.shape[0], 'stats', g.stats(), o.stats())
    if gi.shape != oi.shape or not np.array_equal(gi, oi):
        print('  index sets differ'); return
    d = np.abs(gt - ot); k = np.argsort(-d)[:12]
    print('  max dT', d.max(), ' frac>1e-4', (d > 1e-4).mean(), ' max dW rel', (np.abs(gw-ow)/np.maximum(1,ow)).max())
    for q in k:
        print('   ', gi[q], 'T gpu/orc', gt[q], ot[q], 'W gpu/orc', gw[q], ow[q])

d = syn.scene_room()
for name, R, T in (("identity", np.eye(3), np.zeros(3)), ("generic", rot_xyz(0.15, -0.1, 0.4), np.array([0.31, -0.27, 0.12]))):
    o = OracleTSDF(map_scale=[25.6, 25.6], K=syn.K_DEPTH, is_global_map=True)
    g = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, is_global_map=True)
    g.integrate_depth(d, R[None], T[None]); o.integrate_depth(R, T, d)
    worst(g, o, name)
    # run-to-run determinism of the GPU
    g2 = TsdfHandle(o.N, o.Nz, K=syn.K_DEPTH, is_global_map=True)
    g2.integrate_depth(d, R[None], T[None])
    a = as_dict_rows(*g.gather()); b = as_dict_rows(*g2.gather())
    print('  gpu run-to-run max dT', np.abs(a[1]-b[1]).max(), 'dW', np.abs(a[2]-b[2]).max())

# submap flow
from taichi_slam.mapping import SubmapMapping, DenseTSDF
sub = dict(map_scale=[12.8, 12.8], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=5.1, max_submap_num=16, max_disp_particles=1 << 18)
glo = dict(map_scale=[25.6, 25.6], voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=5.1, max_submap_num=16, max_disp_particles=1 << 20)
sm = SubmapMapping(DenseTSDF, sub_opts=sub, global_opts=glo, keyframe_step=3)
sm.set_dep_camera_intrinsic(syn.K_DEPTH)
dd = syn.scene_sphere(3.0)
c = sm.submap_collection
for fid in range(7):
    R, T = rot_xyz(0.02 * fid, -0.03 * fid, 0.1 * fid), np.array([0.2 * fid, 0.1, 0.05 * fid])
    sm.recast_depth_to_map_by_frame(fid, True, (R, T), (np.eye(3), np.zeros(3)), dd, np.array([]))
    c._flush()
    print('fid', fid, 'active', c.get_active_submap_id(), 'counts', [c._h.count_active(s) for s in range(4)], 'global', sm.global_map._h.count_active(0), c._h.stats())
