#!/usr/bin/env python3
"""Numbers for the round-2 design: what block-binned accumulation of the ray march would
look like on the bench stream.  Pure numpy simulation of the reference's sampling rule (no GPU): for a batch of F
frames of the S2 stream it reports rays, samples, segments (maximal runs of consecutive steps of a ray inside one
16^3 block, steps >= 19 = beyond the shared-memory window), the per-block load distribution and the number of
reductions left when every (block, chunk of C segments) accumulates in shared memory first.

    python tools/sim_binning.py [frames=16]
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from taichislam_b200 import synthetic as syn  # noqa: E402

VS, STEP, B = np.float32(0.05), 2, 16


def frame_rays(depth, R, T):
    K = syn.K_DEPTH
    fx, fy, cx, cy = K[0], K[4], K[2], K[5]
    jj, ii = np.meshgrid(np.arange(0, depth.shape[0], STEP), np.arange(0, depth.shape[1], STEP), indexing="ij")
    d = depth[jj, ii].astype(np.float32) / np.float32(1000.0)
    pt = np.stack([(ii - cx) * d / fx, (jj - cy) * d / fy, d], -1).reshape(-1, 3).astype(np.float32)
    p = (pt @ R.T.astype(np.float32)).astype(np.float32)
    key = np.round(p / VS).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    n = inv.max() + 1
    s = np.zeros((n, 3)); c = np.zeros(n)
    np.add.at(s, inv, p); np.add.at(c, inv, 1)
    m = (s / c[:, None]).astype(np.float32)
    L = np.linalg.norm(m, axis=1).astype(np.float32)
    return m / L[:, None], L, T.astype(np.float32)


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    depth = syn.scene_sphere(4.0)
    Rs, Ts = syn.stream_poses(F)
    seg_per_block = {}
    vox_per_block = {}
    n_rays = n_samples = n_segments = 0
    for f in range(F):
        u, L, T = frame_rays(depth, Rs[f], Ts[f])
        n = np.minimum(L / VS + 10, 200).astype(int)
        n_rays += len(L)
        for r in range(0, len(L), 4096):
            uu, nn = u[r:r + 4096], n[r:r + 4096]
            J = np.arange(19, nn.max() + 1, dtype=np.float32)
            pos = (uu[:, None, :] * J[None, :, None]) * VS + T[None, None, :]
            vox = np.round(pos / VS).astype(np.int32)                      # [rays, steps, 3]
            valid = J[None, :] <= nn[:, None]
            blk = vox >> 4
            bid = ((blk[..., 0] + 64) * 128 + (blk[..., 1] + 64)) * 128 + (blk[..., 2] + 64)
            newseg = np.ones(bid.shape, bool)
            newseg[:, 1:] = bid[:, 1:] != bid[:, :-1]
            newseg &= valid
            n_samples += int(valid.sum())
            n_segments += int(newseg.sum())
            ids, cnt = np.unique(bid[newseg], return_counts=True)
            for i, c in zip(ids, cnt):
                seg_per_block[i] = seg_per_block.get(i, 0) + int(c)
            vkey = (bid.astype(np.int64) << 12) | ((vox[..., 0] & 15) << 8 | (vox[..., 1] & 15) << 4 | (vox[..., 2] & 15))
            for i in np.unique(bid[valid]):
                s = vox_per_block.setdefault(int(i), set())
                s.update(np.unique(vkey[valid & (bid == i)]).tolist())
    segs = np.array(sorted(seg_per_block.values(), reverse=True))
    distinct = sum(len(v) for v in vox_per_block.values())
    print(f"frames {F}: rays {n_rays:,}  far samples (steps >= 19) {n_samples:,}  segments {n_segments:,} "
          f"({n_segments / n_rays:.1f} per ray, {n_samples / n_segments:.1f} samples each)")
    print(f"blocks touched {len(segs)}; segments per block: max {segs[0]:,}  p90 {int(np.percentile(segs, 90)):,}  median {int(np.median(segs)):,}")
    print(f"top 8 blocks hold {100 * segs[:8].sum() / segs.sum():.1f} % of the segments")
    print(f"distinct (block, voxel) pairs touched by the batch: {distinct:,}  -> reductions today {n_samples:,}, "
          f"with one accumulation per block per batch {distinct:,} ({n_samples / distinct:.1f}x fewer)")
    for C in (1024, 4096, 16384):
        chunks = int(np.ceil(segs / C).sum())
        # upper bound on reductions: every chunk touches at most min(4096, its samples) voxels
        red = int(sum(min(4096 * np.ceil(s / C), s * n_samples / n_segments, len(vox_per_block[b]) * np.ceil(s / C))
                      for s, b in zip((seg_per_block[k] for k in seg_per_block), seg_per_block)))
        print(f"chunk = {C:5d} segments: {chunks:,} work items, <= {red:,} reductions ({n_samples / max(red, 1):.1f}x fewer)")


if __name__ == "__main__":
    main()
