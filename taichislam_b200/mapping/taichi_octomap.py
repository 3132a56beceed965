"""Octomap - the reference's class surface (taichi_slam/mapping/taichi_octomap.py) on the hit-count
hash grid of libtslam.so.  The reference "Octomap" is a per-voxel hit counter (occupy += 1 per point,
:116-119) with a threshold test (:86-88): no ray casting, no log-odds."""
import math
import time

import numpy as np

from ..octo_handle import OctoHandle
from .field import Field
from .mapping_common import BaseMap


class Octomap(BaseMap):
    def __init__(self, map_scale=[10, 10], voxel_scale=0.05, min_occupy_thres=3, texture_enabled=False,
                 min_ray_length=0.3, max_ray_length=3.0, max_disp_particles=1000000, K=2,
                 max_submap_num=1024, disp_ceiling=10.0, disp_floor=-10.0,
                 is_global_map=False, recast_step=2, color_same_proj=True, max_blocks=0):
        super(Octomap, self).__init__(voxel_scale)
        import torch
        self._torch = torch
        # :19-28
        self.Rxy = math.ceil(math.log2(map_scale[0] / voxel_scale) / math.log2(K))
        self.Rz = math.ceil(math.log2(map_scale[1] / voxel_scale) / math.log2(K))
        self.map_size_xy = map_scale[0]
        self.map_size_z = map_scale[1]
        self.K = K
        self.N = self.K ** self.Rxy
        self.Nz = self.K ** self.Rz
        ctor_voxel_scale = voxel_scale              # the kernels keep using the constructor's value (mapping_common.py:22-23)
        self.voxel_scale = self.map_size_xy / self.N  # :28
        self.max_disp_particles = max_disp_particles
        self.min_occupy_thres = min_occupy_thres
        self.max_ray_length = max_ray_length
        self.min_ray_length = min_ray_length
        self.enable_texture = texture_enabled
        self.max_submap_num = max_submap_num
        self.disp_ceiling = disp_ceiling
        self.disp_floor = disp_floor
        self.is_global_map = is_global_map
        self.recast_step = recast_step
        self.color_same_proj = color_same_proj
        self.color = None  # the colour plane lives inside the library handle (:77-79); read it with _h.gather(color=True)
        self._h = OctoHandle(self.N, self.Nz, K=K, voxel_scale=ctor_voxel_scale, min_occupy_thres=min_occupy_thres,
                             min_ray_length=min_ray_length, max_ray_length=max_ray_length, recast_step=recast_step,
                             max_submaps=min(max_submap_num, 1024), max_blocks=max_blocks, texture_enabled=texture_enabled)
        if texture_enabled:
            self._h.set_color_intrinsics([1, 0, 0, 0, 1, 0, 0, 0, 1], color_same_proj)
        dev = torch.device("cuda", torch.cuda.current_device())
        n = max_disp_particles
        self.num_export_particles = Field(torch.zeros(1, dtype=torch.int32, device=dev))
        self.export_x = Field(torch.full((n, 3), -100000.0, dtype=torch.float32, device=dev))       # :56-60
        self.export_color = Field(torch.full((n, 3), 0.5, dtype=torch.float32, device=dev))
        self.initialize_submap_fields(self.max_submap_num)
        print(f'The map voxel is:[{self.max_submap_num}x{self.N}x{self.N}x{self.Nz}] voxel scale {self.voxel_scale:3.3f}^3 '
              f'map scale:[{self.map_size_xy}mx{self.map_size_xy}mx{self.map_size_z}m] tree depth [{self.Rxy}, {self.Rz}]')

    @property
    def occupy(self):
        """(idx int32[n,3], count uint32[n]) of the active submap - the leaf field of the reference's tree (:75-76)."""
        return self._h.gather(self._active())

    def _flush(self):
        pass

    def _active(self):
        return self.active_submap_id[None]

    def _on_intrinsics(self):
        self._h.set_intrinsics(self.K_cam_dep)

    def set_color_camera_intrinsic(self, K):
        super(Octomap, self).set_color_camera_intrinsic(K)
        self._h.set_color_intrinsics(self.K_cam_color, self.color_same_proj)

    def _upload_submap_pose(self, submap_id, R, T):
        self._h.set_submap_pose(submap_id, R, T)

    # :126-132
    def recast_pcl_to_map(self, R, T, xyz_array, rgb_array, n=None):
        self.set_pose(R, T)
        xyz = np.ascontiguousarray(np.asarray(xyz_array, dtype=np.float32).reshape(-1, 3))
        rgb = None
        if self.enable_texture and rgb_array is not None and np.size(rgb_array) > 0:  # :142-143; BGR -> RGB inside (:120-124)
            rgb = np.ascontiguousarray(np.asarray(rgb_array, dtype=np.uint8).reshape(-1, 3))
        if n is not None:
            xyz = xyz[:n]
            rgb = rgb[:n] if rgb is not None else None
        self._h.integrate_points(xyz, self.input_R_np, self.input_T_np, submap=self._active(), rgb=rgb)

    def recast_depth_to_map(self, R, T, depthmap, texture):
        self.set_pose(R, T)
        tex = None
        if self.enable_texture and texture is not None and np.size(texture) > 0:  # :160-167
            tex = np.ascontiguousarray(texture, dtype=np.uint8)
        self._h.integrate_depth(np.asarray(depthmap), self.input_R_np, self.input_T_np, submap=self._active(), texture=tex)

    # :90-114
    @staticmethod
    def _lod(level):
        """`occupy.parent(level)` (:95): level 0 is the field's own node, level 1 the innermost pointer node - one cell per
        voxel either way (the node runs with disp_level = 0, taichislam_node.py:37); level L > 1 = K^(L-1)-voxel groups."""
        return max(int(level), 1)

    def cvt_occupy_to_voxels(self, level):
        level = self._lod(level)
        self.num_export_particles.t.zero_()
        self._h.extract(self._active(), level, self.export_x.t, self.num_export_particles.t,
                        self.export_color.t if self.enable_texture else None)  # :101-102

    def cvt_occupy_voxels_to(self, level, cur_num, max_disp_particles, x, color):
        level = self._lod(level)
        self._h.extract(self._active(), level, x.t[:max_disp_particles], cur_num.t,
                        color.t[:max_disp_particles] if self.enable_texture else None)  # :113-114

    def get_occupy_voxels(self, l):
        self.cvt_occupy_to_voxels(l)
        return self.export_x.to_numpy(), self.export_color.to_numpy()

    def fuse_submaps(self, submaps):  # :195-199
        t = time.time()
        self._h.fuse_from(submaps._h)
        print(f"[OctoMap] Fuse submaps {(time.time() - t) * 1000:.1f}ms, active local: {submaps.active_submap_id[None]} "
              f"remote: {submaps.remote_submap_num[None]}")

    # planner queries (mapping_common.py:165-204; is_occupy :86-88; the reference's Octomap has no is_unobserved)
    def _query_raycast(self, pos, dir, max_dist):
        return self._h.raycast(pos, dir, max_dist, submap=self._active())

    def _query_points(self, xyz):
        occ = self._h.query_points(xyz, submap=self._active())
        return occ, np.zeros(len(occ), bool)  # BaseMap.is_unobserved prints "Not implemented" and returns False (:207-209)

    def _query_near(self, xyz, voxel):
        if int(voxel) <= 0:  # range(-voxel, voxel) is empty (mapping_common.py:198): what TopoGraphGen asks for (:328)
            return np.zeros(len(xyz), bool)
        raise NotImplementedError("Octomap.is_near_pos_occupy(voxel > 0) has no batched kernel yet")

    def saveMap(self, path):  # :201-202 (stub in the reference)
        pass

    def export_submap(self):  # :204-205
        return {}

    def finalization_current_submap(self):
        pass

    def reset(self):  # :210-211
        self._h.reset()
