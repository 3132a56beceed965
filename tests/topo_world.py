"""A small analytic indoor world for the TopoGraphGen tests: two box rooms joined by a doorway, as a TSDF."""
import numpy as np


def two_rooms(vs=0.05):
    """(idx int32[n,3], tsdf f32[n], w f32[n], occ int8[n]) - observed voxels = free space + a 0.3 m shell of wall."""
    from scipy.ndimage import distance_transform_edt
    nx, ny, nz = 150, 80, 60                      # 7.5 x 4 x 3 m grid centred in x/y: all indices inside a 256^3 map
    ox, oy, oz = 75, 40, 10
    free = np.zeros((nx, ny, nz), bool)
    free[6:66, 6:74, 6:50] = True                 # room A: 3.0 x 3.4 x 2.2 m
    free[84:144, 6:74, 6:50] = True               # room B
    free[66:84, 30:50, 6:46] = True               # doorway / corridor 0.9 m long, 1.0 m wide, 2.0 m high
    din = distance_transform_edt(free).astype(np.float32) * np.float32(vs)
    dout = distance_transform_edt(~free).astype(np.float32) * np.float32(vs)
    tsdf = np.where(free, din - np.float32(0.5 * vs), -(dout - np.float32(0.5 * vs))).astype(np.float32)
    obs = free | (dout <= 0.3)
    I, J, K = np.nonzero(obs)
    idx = np.stack([I - ox, J - oy, K - oz], 1).astype(np.int32)
    t = np.clip(tsdf[I, J, K], -0.3, 1.0).astype(np.float16).astype(np.float32)  # f16-representable: the reference stores f16
    return idx, t, np.ones(len(t), np.float32), np.zeros(len(t), np.int8)


START_A = np.array([(36 - 75) * 0.05, (40 - 40) * 0.05, (28 - 10) * 0.05])  # centre of room A in map coordinates
