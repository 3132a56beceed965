"""Deterministic synthetic inputs for tests and bench.

Depth intrinsics are the reference node's defaults (scripts/taichislam_node.py:69-72).
"""
import numpy as np

FX = FY = 384.2377014160156
CX = 323.4873046875
CY = 235.0628204345703
K_DEPTH = [FX, 0.0, CX, 0.0, FY, CY, 0.0, 0.0, 1.0]
H, W = 480, 640


def _uv(h=H, w=W):
    i = np.arange(w, dtype=np.float64)[None, :]
    j = np.arange(h, dtype=np.float64)[:, None]
    return (i - CX) / FX, (j - CY) / FY


def scene_plane(z_m=3.0, h=H, w=W):
    """S1: fronto-parallel plane at depth z."""
    return np.full((h, w), int(round(z_m * 1000)), np.uint16)


def scene_sphere(radius_m=4.0, h=H, w=W):
    """S2 (R=4) / S3 (R=8): camera-centred sphere, depth = R / sqrt(u^2+v^2+1)."""
    u, v = _uv(h, w)
    z = radius_m / np.sqrt(u * u + v * v + 1.0)
    return np.round(z * 1000.0).astype(np.uint16)


def scene_noise(seed=0, lo=1500, hi=4500, h=H, w=W):
    """S4: uniform random depth per pixel (worst-case bucket fan-out)."""
    return np.random.default_rng(seed).integers(lo, hi, size=(h, w)).astype(np.uint16)


def scene_room(seed=3, h=H, w=W):
    """A box room with a few depth steps and invalid (0) holes - exercises the range filter."""
    rng = np.random.default_rng(seed)
    u, v = _uv(h, w)
    u = np.broadcast_to(u, (h, w))
    v = np.broadcast_to(v, (h, w))
    z = np.minimum(np.minimum(np.full((h, w), 5.0), 2.5 / np.maximum(np.abs(u), 1e-3)), 1.6 / np.maximum(np.abs(v), 1e-3))
    d = np.round(z * 1000.0)
    d[rng.random((h, w)) < 0.02] = 0          # dropouts
    d[40:80, 50:120] = 150                   # closer than min_ray
    d[300:330, 400:460] = 20000              # farther than max_ray
    return d.astype(np.uint16)


def texture_gradient(seed, h=H, w=W):
    """Seeded synthetic colour image uint8 [h,w,3]: channel gradients + noise, channel 0 never 0 (marching cubes treats
    r == 0 as "no colour", marching_cube_mesher.py:73-79)."""
    rng = np.random.default_rng(seed)
    jj, ii = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    t = np.stack([(ii * 255 // max(w - 1, 1)), (jj * 255 // max(h - 1, 1)), ((ii + jj) % 256)], -1).astype(np.int32)
    t = np.clip(t + rng.integers(-20, 21, t.shape), 1, 255)
    return t.astype(np.uint8)


def stream_pose(t, period=1000, radius=0.5):
    """Pose t of the bench stream: R = Rz(2*pi*t/period), T on a circle of `radius` m."""
    a = 2.0 * np.pi * t / period
    c, s = np.cos(a), np.sin(a)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    T = np.array([radius * c, radius * s, 0.0])
    return R, T


def stream_poses(n, start=0, period=1000, radius=0.5):
    Rs = np.zeros((n, 3, 3))
    Ts = np.zeros((n, 3))
    for q in range(n):
        Rs[q], Ts[q] = stream_pose(start + q, period, radius)
    return Rs, Ts


def octo_cloud(n=100000, seed=1, rmin=0.3, rmax=20.0):
    """C3 cloud: points uniform in a 40 m cube, kept when |p| in [rmin, rmax]."""
    rng = np.random.default_rng(seed)
    out = np.zeros((0, 3), np.float32)
    while out.shape[0] < n:
        p = rng.uniform(-20.0, 20.0, size=(2 * n, 3))
        r = np.linalg.norm(p, axis=1)
        out = np.concatenate([out, p[(r >= rmin) & (r <= rmax)].astype(np.float32)])
    return np.ascontiguousarray(out[:n])
