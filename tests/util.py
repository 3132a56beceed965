"""Shared helpers for the parity tests."""
import numpy as np


def rot_xyz(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def key_sort(idx):
    """Lexicographic row order of an [n,3] int array."""
    return np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))


def as_dict_rows(idx, *cols):
    o = key_sort(idx)
    return (idx[o],) + tuple(c[o] for c in cols)


def compare_voxels(gpu, orc, tol=1e-4, check_occ=True):
    """gpu/orc = (idx, tsdf, w, occ).  Index set and occupy bit-exact, values within tol."""
    gi, gt, gw, go = as_dict_rows(*gpu)
    oi, ot, ow, oo = as_dict_rows(*orc)
    assert gi.shape == oi.shape, f"voxel count differs: gpu {gi.shape[0]} oracle {oi.shape[0]}"
    assert np.array_equal(gi, oi), "observed voxel index sets differ"
    if check_occ:
        assert np.array_equal(go.astype(np.int32), oo.astype(np.int32)), "occupy differs"
    fin = np.isfinite(ot)
    assert np.array_equal(np.isfinite(gt), fin)
    dt = np.abs(gt[fin] - ot[fin])
    dw = np.abs(gw - ow)
    assert dt.max(initial=0.0) <= tol, f"max |dTSDF| = {dt.max()}"
    # weights: relative tolerance (W grows to 1000)
    assert np.all(dw <= tol * np.maximum(1.0, np.abs(ow))), f"max |dW| = {dw.max()}"
    return float(dt.max(initial=0.0)), float(dw.max(initial=0.0))


def tri_multiset(verts, decimals=5):
    """Canonical sorted array of triangles (each triangle's 3 vertices kept in emission order)."""
    t = np.round(verts.reshape(-1, 9).astype(np.float64), decimals)
    o = np.lexsort(t.T[::-1])
    return t[o]
