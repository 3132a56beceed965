"""BaseMap - host-side pose state shared by the map classes.

Mirrors the reference's BaseMap (taichi_slam/mapping/mapping_common.py) for everything the
callers use from Python: pose plumbing (set_pose / set_base_pose / convert_by_base, :91-100,
:141-156), the submap pose tables (:102-131) and the camera-intrinsic setters (:25-29).  The
device-side helpers of the reference (coordinate transforms, unprojection, :31-58, :221-266) live
inside the CUDA kernels of libtslam.so.
"""
import numpy as np

from .field import HostScalar


def sign(val):  # mapping_common.py:5-7
    return (0 < val) - (val < 0)


class _ZeroD:
    """`field[None]` face of a 0-d Taichi field over a host array: TaichiSLAM_demo.py:44-48 writes the pose straight into
    `mapping.input_T[None][i]` / `mapping.input_R[None][i, j]` - the writes land in the buffers handed to the C ABI."""

    def __init__(self, a):
        self._a = a

    def __getitem__(self, idx):
        return self._a

    def __setitem__(self, idx, v):
        self._a[...] = v

    def to_numpy(self):
        return self._a.copy()


class BaseMap:
    def __init__(self, voxel_scale):
        self.base_T_np = np.zeros(3)          # :16-17
        self.base_R_np = np.eye(3)
        self.input_R_np = np.eye(3, dtype=np.float32)   # input_R / input_T 0-d fields (:12-13), kept on the host
        self.input_T_np = np.zeros(3, dtype=np.float32)
        self._pose_tmp_R, self._pose_tmp_d, self._pose_tmp_T = np.zeros((3, 3)), np.zeros(3), np.zeros(3)
        self.input_R, self.input_T = _ZeroD(self.input_R_np), _ZeroD(self.input_T_np)   # mapping_common.py:12-13
        self.frame_id = 0
        self.submap_enabled = False
        self.voxel_scale = voxel_scale
        self.voxel_scale_ = np.array([voxel_scale] * 3, dtype=np.float32)  # :23
        self.K_cam_dep = None
        self.K_cam_color = None

    # :25-29
    def set_dep_camera_intrinsic(self, K):
        self.K_cam_dep = [float(k) for k in np.asarray(K).reshape(-1)]
        self._on_intrinsics()

    def set_color_camera_intrinsic(self, K):
        self.K_cam_color = [float(k) for k in np.asarray(K).reshape(-1)]

    def _on_intrinsics(self):
        pass

    # :91-100
    def convert_by_base(self, R, T):
        R = np.asarray(R, dtype=np.float64)
        T = np.asarray(T, dtype=np.float64)
        if self.submap_enabled:
            sid = self.active_submap_id[None]
            base_R_inv = self.submaps_base_R_np[sid].T
            R_ = base_R_inv @ R
            T_ = base_R_inv @ (T - self.submaps_base_T_np[sid])
        else:
            base_R_inv = self.base_R_np.T
            R_ = base_R_inv @ R
            T_ = base_R_inv @ (T - self.base_T_np)
        return R_, T_

    # :102-111
    def initialize_submap_fields(self, max_submap_num):
        self.submap_enabled = True
        self.submaps_base_R_np = np.zeros((max_submap_num, 3, 3))
        self.submaps_base_T_np = np.zeros((max_submap_num, 3))
        self.active_submap_id = HostScalar(0)
        self.remote_submap_num = HostScalar(0)

    def get_active_submap_id(self):  # :113-114
        return self.active_submap_id[None]

    def switch_to_next_submap(self):  # :116-119
        self.finalization_current_submap()
        self.active_submap_id[None] += 1
        return self.active_submap_id[None]

    def finalization_current_submap(self):
        pass

    def set_base_pose_submap(self, submap_id, _R, _T):  # :121-131
        self.submaps_base_T_np[submap_id] = _T
        self.submaps_base_R_np[submap_id] = _R
        self._upload_submap_pose(int(submap_id), np.asarray(_R, dtype=np.float64), np.asarray(_T, dtype=np.float64))

    def _upload_submap_pose(self, submap_id, R, T):
        raise NotImplementedError

    def set_base_pose(self, _R, _T):  # :141-147
        self.base_T_np = np.asarray(_T, dtype=np.float64)
        self.base_R_np = np.asarray(_R, dtype=np.float64)

    def set_pose(self, _R, _T):  # :149-156
        # convert_by_base (:91-100) without temporaries - this runs once per frame: same f64 operations in the same
        # order (R' = Rb^T R, T' = Rb^T (T - Tb)), results cast f64 -> f32 into the persistent buffers whose addresses are
        # handed to the C ABI every frame
        R = _R if (type(_R) is np.ndarray and _R.dtype == np.float64) else np.asarray(_R, dtype=np.float64)
        T = _T if (type(_T) is np.ndarray and _T.dtype == np.float64) else np.asarray(_T, dtype=np.float64)
        if self.submap_enabled:
            sid = self.active_submap_id.v
            Rb, Tb = self.submaps_base_R_np[sid], self.submaps_base_T_np[sid]
        else:
            Rb, Tb = self.base_R_np, self.base_T_np
        Rbi = Rb.T
        np.matmul(Rbi, R, out=self._pose_tmp_R)
        np.subtract(T.reshape(3), Tb, out=self._pose_tmp_d)
        np.matmul(Rbi, self._pose_tmp_d, out=self._pose_tmp_T)
        self.input_R_np[...] = self._pose_tmp_R
        self.input_T_np[...] = self._pose_tmp_T

    # ------------------------------------------------------------------ planner queries (:165-204)
    # The reference exposes these as @ti.func helpers that other Taichi kernels (TopoGraphGen) call per ray / per
    # point.  Here they are host-callable and BATCHED: xyz / pos / dir are [n,3] (or a single [3]) arrays in the frame
    # of the active submap, every call is one kernel over the whole batch.  Subclasses implement the _query_* hooks.
    @staticmethod
    def _as_batch(a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        return (a.reshape(1, 3), True) if a.ndim == 1 else (a.reshape(-1, 3), False)

    def raycast(self, pos, dir, max_dist):
        """(succ, last position, length) of BaseMap.raycast (:165-178) for every ray."""
        p, single = self._as_batch(pos)
        d, _ = self._as_batch(dir)
        succ, x, ln = self._query_raycast(p, d, float(max_dist))
        return (bool(succ[0]), x[0], float(ln[0])) if single else (succ, x, ln)

    def is_pos_occupy(self, xyz):  # :188-192
        p, single = self._as_batch(xyz)
        r = self._query_points(p)[0]
        return bool(r[0]) if single else r

    def is_pos_unobserved(self, xyz):  # :181-185
        p, single = self._as_batch(xyz)
        r = self._query_points(p)[1]
        return bool(r[0]) if single else r

    def is_near_pos_occupy(self, xyz, voxel):  # :194-204 (range(-voxel, voxel): voxel = 0 tests nothing)
        p, single = self._as_batch(xyz)
        r = self._query_near(p, int(voxel))
        return bool(r[0]) if single else r

    def _query_raycast(self, pos, dir, max_dist):
        raise NotImplementedError("Not implemented")  # mapping_common.py:207-214

    def _query_points(self, xyz):
        raise NotImplementedError("Not implemented")

    def _query_near(self, xyz, voxel):
        raise NotImplementedError("Not implemented")
