"""ctypes face of the CPU oracle (oracle/tslam_oracle.cpp).

TEST INFRASTRUCTURE.  Imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py - never by taichislam_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtslam_oracle.so")

MODE_CANONICAL, MODE_F32_LITERAL, MODE_F16_FAITHFUL = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "tslam_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtslam_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _TsdfCfg(C.Structure):
    _fields_ = [("voxel_scale", C.c_double), ("N", C.c_int), ("Nz", C.c_int),
                ("max_ray_length", C.c_double), ("min_ray_length", C.c_double),
                ("internal_voxels", C.c_int), ("recast_step", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("mode", C.c_int), ("is_global_map", C.c_int),
                ("disp_floor", C.c_double), ("disp_ceiling", C.c_double)]


class _OctoCfg(C.Structure):
    _fields_ = [("voxel_scale", C.c_double), ("N", C.c_int), ("Nz", C.c_int), ("K", C.c_int),
                ("max_ray_length", C.c_double), ("min_ray_length", C.c_double),
                ("recast_step", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("min_occupy_thres", C.c_int)]


_lib = None


def oracle_colormap():
    """The oracle's jet LUT [1024,3] (mapping_common.py:158-163)."""
    out = np.zeros((1024, 3), np.float32)
    L = lib()
    L.orc_colormap.argtypes = [C.c_void_p]
    L.orc_colormap.restype = None
    L.orc_colormap(_p(out))
    return out


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.orc_tsdf_create.restype = vp
        L.orc_tsdf_create.argtypes = [C.POINTER(_TsdfCfg)]
        L.orc_octo_create.restype = vp
        L.orc_octo_create.argtypes = [C.POINTER(_OctoCfg)]
        for name in ("orc_tsdf_destroy", "orc_tsdf_reset", "orc_tsdf_commit", "orc_tsdf_clear_stats",
                     "orc_octo_destroy", "orc_octo_reset"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = None
        L.orc_tsdf_set_submap_pose.argtypes = [vp, i32, vp, vp]
        L.orc_octo_set_submap_pose.argtypes = [vp, i32, vp, vp]
        L.orc_tsdf_get_stats.argtypes = [vp, vp]
        L.orc_tsdf_integrate_depth.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32]
        L.orc_tsdf_integrate_points.argtypes = [vp, vp, i32, vp, vp, i32, i32]
        L.orc_tsdf_count_active.argtypes = [vp, i32]
        L.orc_tsdf_count_active.restype = i64
        L.orc_tsdf_gather.argtypes = [vp, i32, i64, vp, vp, vp, vp]
        L.orc_tsdf_gather.restype = i64
        L.orc_tsdf_scatter.argtypes = [vp, i32, i64, vp, vp, vp, vp]
        L.orc_tsdf_fuse.argtypes = [vp, vp]
        L.orc_tsdf_surface.argtypes = [vp, i32, i64, vp, vp]
        L.orc_tsdf_surface.restype = i64
        L.orc_tsdf_slice.argtypes = [vp, i32, C.c_float, C.c_float, i64, vp, vp]
        L.orc_tsdf_slice.restype = i64
        L.orc_mc.argtypes = [vp, i32, C.c_float, i64, vp, vp]
        L.orc_mc.restype = i64
        L.orc_esdf_update.argtypes = [vp, i32]
        L.orc_esdf_update.restype = i64
        L.orc_esdf_gather.argtypes = [vp, i32, i64, vp, vp]
        L.orc_esdf_gather.restype = i64
        L.orc_octo_integrate_points.argtypes = [vp, vp, i32, vp, vp, i32]
        L.orc_octo_integrate_depth.argtypes = [vp, vp, i32, i32, vp, vp, i32]
        L.orc_octo_gather.argtypes = [vp, i32, i64, vp, vp]
        L.orc_octo_gather.restype = i64
        L.orc_octo_export.argtypes = [vp, i32, i32, i64, vp]
        L.orc_octo_export.restype = i64
        L.orc_octo_fuse.argtypes = [vp, vp]
        L.orc_octo_set_color.argtypes = [vp, i32, i32, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_octo_integrate_points_rgb.argtypes = [vp, vp, vp, i32, vp, vp, i32]
        L.orc_octo_integrate_depth_tex.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32]
        L.orc_octo_gather_color.argtypes = [vp, i32, i64, vp]
        L.orc_octo_export2.argtypes = [vp, i32, i32, i64, vp, vp]
        L.orc_octo_export2.restype = i64
        L.orc_tsdf_set_color.argtypes = [vp, i32, i32, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_tsdf_integrate_depth_tex.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, i32]
        L.orc_tsdf_integrate_points_rgb.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32]
        L.orc_tsdf_gather_color.argtypes = [vp, i32, i64, vp]
        L.orc_tsdf_gather_color.restype = i64
        L.orc_tsdf_scatter_color.argtypes = [vp, i32, i64, vp, vp]
        L.orc_mc2.argtypes = [vp, i32, C.c_float, i64, vp, vp, vp]
        L.orc_mc2.restype = i64
        L.orc_tsdf_query_points.argtypes = [vp, i32, i64, vp, vp]
        L.orc_tsdf_query_near.argtypes = [vp, i32, i64, vp, i32, vp]
        L.orc_tsdf_raycast.argtypes = [vp, i32, i64, vp, vp, C.c_float, vp, vp, vp]
        L.orc_octo_query_points.argtypes = [vp, i32, i64, vp, vp]
        L.orc_octo_raycast.argtypes = [vp, i32, i64, vp, vp, C.c_float, vp, vp, vp]
        L.orc_tsdf_integrate_stream_mt.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]
        L.orc_tsdf_integrate_stream_mt.restype = i32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def tsdf_dims(map_scale, voxel_scale, blk):
    """N, Nz as derived at dense_tsdf.py:24-25."""
    import math
    N = math.ceil(map_scale[0] / voxel_scale / blk) * blk
    Nz = math.ceil(map_scale[1] / voxel_scale / blk) * blk
    return N, Nz


class OracleTSDF:
    """CPU oracle of DenseTSDF (dense_tsdf.py) - constructor kwargs as dense_tsdf.py:13-16."""

    def __init__(self, map_scale=(10, 10), voxel_scale=0.05, num_voxel_per_blk_axis=16, max_ray_length=10,
                 min_ray_length=0.3, internal_voxels=10, is_global_map=False, disp_ceiling=1.8, disp_floor=-0.3,
                 recast_step=2, K=None, mode=MODE_CANONICAL):
        self.N, self.Nz = tsdf_dims(map_scale, voxel_scale, num_voxel_per_blk_axis)
        K = K if K is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        self.cfg = _TsdfCfg(voxel_scale, self.N, self.Nz, max_ray_length, min_ray_length, internal_voxels,
                            recast_step, K[0], K[4], K[2], K[5], mode, int(is_global_map), disp_floor, disp_ceiling)
        self.voxel_scale = voxel_scale
        self.h = C.c_void_p(lib().orc_tsdf_create(C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tsdf_destroy(self.h)
            self.h = None

    def reset(self):
        lib().orc_tsdf_reset(self.h)

    def set_submap_pose(self, s, R, T):
        R, T = _f32(R), _f32(T)
        lib().orc_tsdf_set_submap_pose(self.h, s, _p(R), _p(T))

    def integrate_depth(self, R, T, depth, submap=0, commit=True):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        R, T = _f32(R), _f32(T)
        lib().orc_tsdf_integrate_depth(self.h, _p(depth), depth.shape[0], depth.shape[1], _p(R), _p(T), submap, int(commit))

    def integrate_points(self, R, T, xyz, submap=0, commit=True):
        xyz = _f32(xyz)
        R, T = _f32(R), _f32(T)
        lib().orc_tsdf_integrate_points(self.h, _p(xyz), xyz.shape[0], _p(R), _p(T), submap, int(commit))

    def set_color(self, enabled=True, same_proj=True, Kcolor=None):
        """texture_enabled / color_same_proj / set_color_camera_intrinsic (dense_tsdf.py:13-16, mapping_common.py:28-29)."""
        K = Kcolor if Kcolor is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        lib().orc_tsdf_set_color(self.h, int(enabled), int(same_proj), K[0], K[4], K[2], K[5])

    def integrate_depth_tex(self, R, T, depth, texture, submap=0, commit=True):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        texture = np.ascontiguousarray(texture, dtype=np.uint8)
        R, T = _f32(R), _f32(T)
        lib().orc_tsdf_integrate_depth_tex(self.h, _p(depth), _p(texture), texture.shape[0], texture.shape[1], depth.shape[0],
                                           depth.shape[1], _p(R), _p(T), submap, int(commit))

    def integrate_points_rgb(self, R, T, xyz, rgb, submap=0, commit=True):
        xyz = _f32(xyz)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        R, T = _f32(R), _f32(T)
        lib().orc_tsdf_integrate_points_rgb(self.h, _p(xyz), _p(rgb), xyz.shape[0], _p(R), _p(T), submap, int(commit))

    def gather_color(self, submap=0):
        n = self.count_active(submap)
        c = np.zeros((n, 3), np.float32)
        lib().orc_tsdf_gather_color(self.h, submap, n, _p(c))
        return c

    def scatter_color(self, submap, idx, col):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        col = _f32(col)
        lib().orc_tsdf_scatter_color(self.h, submap, idx.shape[0], _p(idx), _p(col))

    def marching_cubes_color(self, step=1, thres=0.1, cap_tri=1 << 21):
        v = np.zeros((cap_tri * 3, 3), np.float32); nrm = np.zeros((cap_tri * 3, 3), np.float32); col = np.zeros((cap_tri * 3, 3), np.float32)
        n = int(lib().orc_mc2(self.h, step, thres, cap_tri, _p(v), _p(nrm), _p(col)))
        k = min(n, cap_tri)
        return n, v[:3 * k], nrm[:3 * k], col[:3 * k]

    def commit(self):
        lib().orc_tsdf_commit(self.h)

    def stats(self):
        o = np.zeros(5, np.int64)
        lib().orc_tsdf_get_stats(self.h, _p(o))
        return dict(n_px=int(o[0]), n_valid=int(o[1]), n_rays=int(o[2]), n_updates=int(o[3]), n_oob=int(o[4]))

    def clear_stats(self):
        lib().orc_tsdf_clear_stats(self.h)

    def count_active(self, submap=0):
        return int(lib().orc_tsdf_count_active(self.h, submap))

    def gather(self, submap=0):
        n = self.count_active(submap)
        idx = np.zeros((n, 3), np.int32)
        t = np.zeros(n, np.float32)
        w = np.zeros(n, np.float32)
        occ = np.zeros(n, np.int32)
        lib().orc_tsdf_gather(self.h, submap, n, _p(idx), _p(t), _p(w), _p(occ))
        return idx, t, w, occ

    def scatter(self, submap, idx, tsdf, w, occ):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        tsdf, w = _f32(tsdf), _f32(w)
        occ = np.ascontiguousarray(occ, dtype=np.int32)
        lib().orc_tsdf_scatter(self.h, submap, idx.shape[0], _p(idx), _p(tsdf), _p(w), _p(occ))

    def fuse_from(self, src):
        lib().orc_tsdf_fuse(self.h, src.h)

    def surface(self, submap=0, cap=1 << 22):
        xyz = np.zeros((cap, 3), np.float32)
        rgb = np.zeros((cap, 3), np.float32)
        n = int(lib().orc_tsdf_surface(self.h, submap, cap, _p(xyz), _p(rgb)))
        return n, xyz[:min(n, cap)], rgb[:min(n, cap)]

    def slice(self, z, dz=0.5, submap=0, cap=1 << 22):
        xyz = np.zeros((cap, 3), np.float32)
        val = np.zeros(cap, np.float32)
        n = int(lib().orc_tsdf_slice(self.h, submap, z, dz, cap, _p(xyz), _p(val)))
        return n, xyz[:min(n, cap)], val[:min(n, cap)]

    def marching_cubes(self, step=1, thres=0.1, cap_tri=1 << 21):
        v = np.zeros((cap_tri * 3, 3), np.float32)
        nrm = np.zeros((cap_tri * 3, 3), np.float32)
        n = int(lib().orc_mc(self.h, step, thres, cap_tri, _p(v), _p(nrm)))
        m = min(n, cap_tri)
        return n, v[:3 * m], nrm[:3 * m]

    def query_points(self, xyz, submap=0):
        xyz = _f32(xyz)
        f = np.zeros(xyz.shape[0], np.uint8)
        lib().orc_tsdf_query_points(self.h, submap, xyz.shape[0], _p(xyz), _p(f))
        return (f & 1).astype(bool), (f & 2).astype(bool)

    def query_near_occupy(self, xyz, voxel, submap=0):
        xyz = _f32(xyz)
        f = np.zeros(xyz.shape[0], np.uint8)
        lib().orc_tsdf_query_near(self.h, submap, xyz.shape[0], _p(xyz), int(voxel), _p(f))
        return f.astype(bool)

    def raycast(self, pos, direction, max_dist, submap=0):
        pos, direction = _f32(pos), _f32(direction)
        n = pos.shape[0]
        hit = np.zeros(n, np.uint8); xyz = np.zeros((n, 3), np.float32); ln = np.zeros(n, np.float32)
        lib().orc_tsdf_raycast(self.h, submap, n, _p(pos), _p(direction), float(max_dist), _p(hit), _p(xyz), _p(ln))
        return hit.astype(bool), xyz, ln

    def esdf_update(self, submap=0):
        return int(lib().orc_esdf_update(self.h, submap))

    def esdf_gather(self, submap=0):
        n = int(lib().orc_esdf_gather(self.h, submap, 0, None, None))
        idx = np.zeros((n, 3), np.int32)
        e = np.zeros(n, np.float32)
        lib().orc_esdf_gather(self.h, submap, n, _p(idx), _p(e))
        return idx, e


class OracleOctomap:
    """CPU oracle of Octomap (taichi_octomap.py) - N/Nz as taichi_octomap.py:19-27."""

    def __init__(self, map_scale=(10, 10), voxel_scale=0.05, min_occupy_thres=3, min_ray_length=0.3,
                 max_ray_length=3.0, K=2, recast_step=2, Kcam=None):
        import math
        self.Rxy = math.ceil(math.log2(map_scale[0] / voxel_scale) / math.log2(K))
        self.Rz = math.ceil(math.log2(map_scale[1] / voxel_scale) / math.log2(K))
        self.N, self.Nz = K ** self.Rxy, K ** self.Rz
        Kc = Kcam if Kcam is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        self.cfg = _OctoCfg(voxel_scale, self.N, self.Nz, K, max_ray_length, min_ray_length, recast_step,
                            Kc[0], Kc[4], Kc[2], Kc[5], min_occupy_thres)
        self.h = C.c_void_p(lib().orc_octo_create(C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_octo_destroy(self.h)
            self.h = None

    def reset(self):
        lib().orc_octo_reset(self.h)

    def set_submap_pose(self, s, R, T):
        R, T = _f32(R), _f32(T)
        lib().orc_octo_set_submap_pose(self.h, s, _p(R), _p(T))

    def integrate_points(self, R, T, xyz, submap=0):
        xyz = _f32(xyz)
        R, T = _f32(R), _f32(T)
        lib().orc_octo_integrate_points(self.h, _p(xyz), xyz.shape[0], _p(R), _p(T), submap)

    def integrate_depth(self, R, T, depth, submap=0):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        R, T = _f32(R), _f32(T)
        lib().orc_octo_integrate_depth(self.h, _p(depth), depth.shape[0], depth.shape[1], _p(R), _p(T), submap)

    def gather(self, submap=0):
        n = int(lib().orc_octo_gather(self.h, submap, 0, None, None))
        idx = np.zeros((n, 3), np.int32)
        cnt = np.zeros(n, np.uint32)
        lib().orc_octo_gather(self.h, submap, n, _p(idx), _p(cnt))
        return idx, cnt

    def export(self, level=1, submap=0, cap=1 << 22):
        xyz = np.zeros((cap, 3), np.float32)
        n = int(lib().orc_octo_export(self.h, submap, level, cap, _p(xyz)))
        return n, xyz[:min(n, cap)]

    # -- texture (taichi_octomap.py:77-79, :120-124, :160-167, :189) --
    def set_color(self, enabled=True, same_proj=True, Kcolor=None):
        K = Kcolor if Kcolor is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        lib().orc_octo_set_color(self.h, int(enabled), int(same_proj), K[0], K[4], K[2], K[5])

    def integrate_points_rgb(self, R, T, xyz, rgb, submap=0):
        xyz = _f32(xyz)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        R, T = _f32(R), _f32(T)
        lib().orc_octo_integrate_points_rgb(self.h, _p(xyz), _p(rgb), xyz.shape[0], _p(R), _p(T), submap)

    def integrate_depth_tex(self, R, T, depth, texture, submap=0):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        texture = np.ascontiguousarray(texture, dtype=np.uint8)
        R, T = _f32(R), _f32(T)
        lib().orc_octo_integrate_depth_tex(self.h, _p(depth), _p(texture), texture.shape[0], texture.shape[1], depth.shape[0],
                                           depth.shape[1], _p(R), _p(T), submap)

    def gather_color(self, submap=0):
        n = int(lib().orc_octo_gather(self.h, submap, 0, None, None))
        c = np.zeros((n, 3), np.float32)
        lib().orc_octo_gather_color(self.h, submap, n, _p(c))
        return c

    def export_color(self, level=1, submap=0, cap=1 << 22):
        xyz = np.zeros((cap, 3), np.float32)
        rgb = np.zeros((cap, 3), np.float32)
        n = int(lib().orc_octo_export2(self.h, submap, level, cap, _p(xyz), _p(rgb)))
        return n, xyz[:min(n, cap)], rgb[:min(n, cap)]

    def fuse_from(self, src):
        lib().orc_octo_fuse(self.h, src.h)

    def query_points(self, xyz, submap=0):
        xyz = _f32(xyz)
        f = np.zeros(xyz.shape[0], np.uint8)
        lib().orc_octo_query_points(self.h, submap, xyz.shape[0], _p(xyz), _p(f))
        return f.astype(bool)

    def raycast(self, pos, direction, max_dist, submap=0):
        pos, direction = _f32(pos), _f32(direction)
        n = pos.shape[0]
        hit = np.zeros(n, np.uint8); xyz = np.zeros((n, 3), np.float32); ln = np.zeros(n, np.float32)
        lib().orc_octo_raycast(self.h, submap, n, _p(pos), _p(direction), float(max_dist), _p(hit), _p(xyz), _p(ln))
        return hit.astype(bool), xyz, ln


def integrate_stream_mt(maps, depth_frames, Rs, Ts):
    """Throughput helper for the CPU baseline: frames handed to len(maps) host threads."""
    hs = (C.c_void_p * len(maps))(*[m.h for m in maps])
    depth_frames = np.ascontiguousarray(depth_frames, dtype=np.uint16)
    Rs, Ts = _f32(Rs), _f32(Ts)
    n, H, W = depth_frames.shape
    return lib().orc_tsdf_integrate_stream_mt(hs, len(maps), _p(depth_frames), n, H, W, _p(Rs), _p(Ts))
