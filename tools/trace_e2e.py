import time, numpy as np, torch
from taichislam_b200 import synthetic as syn
from taichislam_b200.mapping import DenseTSDF
torch.cuda.init()
Rs, Ts = syn.stream_poses(6400); e = np.array([])
frames = syn.scene_sphere(4.0)
hostf = torch.from_numpy(np.stack([frames] * 64).view(np.int16)).pin_memory().numpy().view(np.uint16)
m2 = DenseTSDF(map_scale=[25.6, 25.6], voxel_scale=0.05, is_global_map=True)
m2.set_dep_camera_intrinsic(syn.K_DEPTH); m2.set_base_pose_submap(0, np.eye(3), np.zeros(3))
import sys
for s in range(6):
    t0 = time.perf_counter()
    for q in range(64):
        m2.recast_depth_to_map(Rs[64 * s + q], Ts[64 * s + q], hostf[q], e)
    t1 = time.perf_counter()
    m2.frame_counters()
    t2 = time.perf_counter()
    sys.stderr.write(f"step {s}: enqueue {1e3*(t1-t0):.3f} flush {1e3*(t2-t1):.3f}\n")
    time.sleep(0.01)
    t3=time.perf_counter(); m2.frame_counters(); sys.stderr.write(f"   idle frame_counters {1e3*(time.perf_counter()-t3):.3f}\n")
