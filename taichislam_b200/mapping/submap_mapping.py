"""SubmapMapping - keyframe -> submap policy, PGO pose correction and local -> global fusion on top of
the B200 map classes.

Host-side orchestration only (no kernels): it mirrors the public behaviour of the reference's
taichi_slam/mapping/submap_mapping.py (same constructor, attributes, method names and wire format) so
scripts/taichislam_node.py can drive it unchanged.  Differences, all fixes of reference quirks
(SURVEY.md Appendix B): no hard-coded save path (submap_mapping.py:144-145 writes to /home/xuhao/...),
`recast_depth_to_map(R, T, ...)` works (the reference calls need_create_new_submap with the wrong arity,
:195-196), send handles are optional.
"""
import io
import time
import zlib

import numpy as np

from .dense_tsdf import DenseTSDF
from .taichi_octomap import Octomap


def _merged(defaults, overrides):
    out = dict(defaults)
    out.update(overrides or {})
    return out


_COMMON = dict(voxel_scale=0.05, texture_enabled=False, min_ray_length=0.3, max_ray_length=3.0,
               max_disp_particles=1024 * 1024)


class SubmapMapping:
    def __init__(self, submap_type=DenseTSDF, keyframe_step=20, sub_opts={}, global_opts={}, autosave_path=None):
        self.submap_type = submap_type
        self.keyframe_step = keyframe_step
        self.autosave_path = autosave_path
        kind = dict(num_voxel_per_blk_axis=10) if submap_type is DenseTSDF else dict(K=2)  # submap_mapping.py:14-34
        self.sub_opts = _merged(_merged(_COMMON, dict(map_scale=[10, 10], max_submap_num=1000, **kind)), sub_opts)
        self.submaps = {}            # frame_id -> submap id
        self.frame_count = 0
        self.exporting_global = False
        self.export_TSDF_xyz = self.export_color = self.export_x = None
        self.submap_collection = submap_type(**self.sub_opts)
        self.global_map = self.create_globalmap(global_opts)
        self.first_init = True
        self.set_exporting_global()
        self.ego_motion_poses = {}
        self.pgo_poses = {}
        self.last_frame_id = None
        self.active_submap_frame_id = 0
        self.enable_texture = self.global_map.enable_texture
        self.post_local_to_global_callback = None
        self.map_send_handle = None
        self.traj_send_handle = None

    def create_globalmap(self, global_opts={}):  # submap_mapping.py:58-84
        kind = dict(num_voxel_per_blk_axis=10, max_submap_num=1024) if self.submap_type is DenseTSDF else dict(K=2, max_submap_num=1000)
        opts = _merged(_merged(_COMMON, dict(map_scale=[100, 100], is_global_map=True, **kind)), global_opts)
        return self.submap_type(**opts)

    # -- intrinsics ----------------------------------------------------------------------------
    def set_dep_camera_intrinsic(self, K):
        self.submap_collection.set_dep_camera_intrinsic(K)

    def set_color_camera_intrinsic(self, K):
        self.submap_collection.set_color_camera_intrinsic(K)

    # -- which map feeds the export buffers (:92-107) -------------------------------------------
    def set_exporting_global(self):
        self.exporting_global = True
        self.set_export_submap(self.global_map)

    def set_exporting_local(self):
        self.exporting_global = False
        self.set_export_submap(self.submap_collection)

    def set_export_submap(self, src):
        self.export_color = src.export_color
        if self.submap_type is DenseTSDF:
            self.export_TSDF_xyz = src.export_TSDF_xyz
            self.num_TSDF_particles = src.num_TSDF_particles
        else:
            self.export_x = src.export_x
            self.num_export_particles = src.num_export_particles

    # -- pose graph updates (:109-124) -----------------------------------------------------------
    def set_frame_poses(self, frame_poses, from_remote=False):
        self.pgo_poses.update(frame_poses)
        used = {}
        for fid, (R, T) in frame_poses.items():
            newer = self.last_frame_id is None or fid > self.last_frame_id
            if newer and fid in self.ego_motion_poses:
                self.last_frame_id = fid
            if fid in self.submaps:
                self.global_map.set_base_pose_submap(self.submaps[fid], R, T)
                used[fid] = (R, T)
        if not from_remote:
            self.send_traj(used)

    def convert_by_pgo(self, frame_id, R, T):  # :162-169
        self.ego_motion_poses[frame_id] = (R, T)
        if self.last_frame_id is None:
            return R, T
        ego_R, ego_T = self.ego_motion_poses[self.last_frame_id]
        pgo_R, pgo_T = self.pgo_poses[self.last_frame_id]
        dR = pgo_R @ ego_R.T
        return dR @ R, dR @ (T - ego_T) + pgo_T

    # -- submap life cycle (:126-160) ------------------------------------------------------------
    def need_create_new_submap(self, is_keyframe, R=None, T=None):
        if self.frame_count == 0:
            return True
        return bool(is_keyframe) and self.frame_count % self.keyframe_step == 0

    def create_new_submap(self, frame_id, R, T):
        print("[SubmapMapping] Create new submap ", frame_id)
        if self.first_init:
            self.first_init = False
        else:
            self.send_submap(self.submap_collection.export_submap())
            self.submap_collection.switch_to_next_submap()
            self.submap_collection.clear_last_TSDF_exporting = True
            self.local_to_global()
        sid = self.submap_collection.get_active_submap_id()
        self.global_map.set_base_pose_submap(sid, R, T)
        self.submap_collection.set_base_pose_submap(sid, R, T)
        self.submaps[frame_id] = sid
        self.pgo_poses[frame_id] = (R, T)
        self.active_submap_frame_id = frame_id
        print(f"[SubmapMapping] Created new submap on frame {frame_id}, now have {sid + 1} submaps")
        if self.autosave_path and sid % 2 == 0:
            self.saveMap(self.autosave_path)
        return self.submap_collection

    def local_to_global(self):
        self.global_map.fuse_submaps(self.submap_collection)
        if self.post_local_to_global_callback is not None:
            self.post_local_to_global_callback(self.global_map)

    # -- integrate (:171-200) --------------------------------------------------------------------
    def _enter_frame(self, frame_id, is_keyframe, pose, ext):
        R, T = self.convert_by_pgo(frame_id, pose[0], pose[1])
        if self.need_create_new_submap(is_keyframe, R, T):
            self.create_new_submap(frame_id, R, T)
        R_ext, T_ext = ext
        return R @ R_ext, T + R @ T_ext

    def recast_depth_to_map_by_frame(self, frame_id, is_keyframe, pose, ext, depthmap, texture):
        Rcam, Tcam = self._enter_frame(frame_id, is_keyframe, pose, ext)
        self.submap_collection.recast_depth_to_map(Rcam, Tcam, depthmap, texture)
        self.frame_count += 1

    def recast_pcl_to_map_by_frame(self, frame_id, is_keyframe, pose, ext, pcl, rgb_array):
        Rcam, Tcam = self._enter_frame(frame_id, is_keyframe, pose, ext)
        self.submap_collection.recast_pcl_to_map(Rcam, Tcam, pcl, rgb_array)
        self.frame_count += 1

    def recast_depth_to_map(self, R, T, depthmap, texture):
        if self.need_create_new_submap(True, R, T):
            self.create_new_submap(self.frame_count, R, T)
        self.submap_collection.recast_depth_to_map(R, T, depthmap, texture)
        self.frame_count += 1

    # -- exporters (:202-224) ----------------------------------------------------------------------
    def _export_src(self):
        return self.global_map if self.exporting_global else self.submap_collection

    def cvt_TSDF_to_voxels_slice(self, z):
        self._export_src().cvt_TSDF_to_voxels_slice(z)

    def cvt_TSDF_surface_to_voxels(self):
        if not self.submaps:
            return
        if self.exporting_global:
            g = self.global_map
            g.cvt_TSDF_surface_to_voxels()
            self.submap_collection.cvt_TSDF_surface_to_voxels_to(g.num_TSDF_particles, g.max_disp_particles,
                                                                 self.export_TSDF_xyz, self.export_color)
        else:
            self.submap_collection.cvt_TSDF_surface_to_voxels()

    def cvt_occupy_to_voxels(self, level):
        if self.exporting_global:
            g = self.global_map
            g.cvt_occupy_to_voxels(level)
            self.submap_collection.cvt_occupy_voxels_to(level, g.num_export_particles, g.max_disp_particles,
                                                        self.export_x, self.export_color)
        else:
            self.submap_collection.cvt_occupy_to_voxels(level)

    # -- wire format: zlib(level 1) of np.save(dict)  (:226-261) -----------------------------------
    @staticmethod
    def _pack(obj):
        f = io.BytesIO()
        np.save(f, obj)
        return f.getbuffer(), zlib.compress(f.getbuffer(), level=1)

    @staticmethod
    def _unpack(buf):
        return np.load(io.BytesIO(zlib.decompress(buf)), allow_pickle=True).item()

    def send_submap(self, submap):
        submap["frame_id"] = self.active_submap_frame_id
        submap["pose"] = self.pgo_poses[self.active_submap_frame_id]
        t0 = time.time()
        raw, packed = self._pack(submap)
        if self.map_send_handle is not None:
            self.map_send_handle(packed)
        print(f"[SubmapMapping] Send submap with {len(raw) / 1024.0:.1f} kB, compressed {len(packed) / 1024:.1f}kB "
              f"compress cost {(time.time() - t0) * 1000:.1f}ms")

    def send_traj(self, traj):
        raw, packed = self._pack(traj)
        if self.traj_send_handle is not None:
            self.traj_send_handle(packed)

    def input_remote_submap(self, buf):
        submap = self._unpack(buf)
        idx = self.submap_collection.input_remote_submap(submap)
        self.global_map.set_base_pose_submap(idx, submap["pose"][0], submap["pose"][1])
        self.local_to_global()
        self.submaps[submap["frame_id"]] = idx

    def input_remote_traj(self, buf):
        self.set_frame_poses(self._unpack(buf), True)

    def saveMap(self, filename):
        self.global_map.saveMap(filename)

    def export_submap(self):
        return self.submap_collection.export_submap()
