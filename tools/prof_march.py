#!/usr/bin/env python3
"""Three 64-frame integrate launches of the bench stream (S2 by default) - the command profiled under ncu.
Usage: python tools/prof_march.py [S1|S2|S3|S4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from taichislam_b200 import synthetic as syn
from taichislam_b200.tsdf_handle import TsdfHandle

scene = sys.argv[1] if len(sys.argv) > 1 else "S2"
d = {"S1": syn.scene_plane(3.0), "S2": syn.scene_sphere(4.0), "S3": syn.scene_sphere(8.0), "S4": syn.scene_noise()}[scene]
g = TsdfHandle(512, 512, K=syn.K_DEPTH, is_global_map=True)
dd = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(d, (64,) + d.shape)).view(np.int16)).cuda()
for w in range(3):
    Rs, Ts = syn.stream_poses(64, start=64 * w)
    g.integrate_depth(dd, Rs, Ts)
torch.cuda.synchronize()
print(g.stats())
