"""DenseTSDF - the reference's class surface (taichi_slam/mapping/dense_tsdf.py) on libtslam.so.

Same constructor kwargs, attributes and methods as the reference class;
every method that launched a Taichi kernel now enqueues hand-written sm_100a kernels through
the C ABI (include/tslam.h).  Like Taichi, calls are asynchronous: `recast_depth_to_map`
hands the frame and its pose to the library's queue (pageable arrays are copied at once, page-locked
ones are fetched by the GPU itself a few calls later) and the queue integrates every 32 frames with
one launch triple; any reader (`count_active`, `to_numpy`, `cvt_*`, field reads, the mesher, the
planner queries) flushes the queue first, so results are indistinguishable from per-frame execution
(up to the commit-granularity note in DESIGN.md for voxels saturated at Wmax).
"""
import collections
import math
import time

import numpy as np

from .. import _capi as capi
from ..tsdf_handle import TsdfHandle
from .field import Field
from .mapping_common import BaseMap

Wmax = 1000  # dense_tsdf.py:8


class DenseTSDF(BaseMap):
    def __init__(self, map_scale=[10, 10], voxel_scale=0.05, texture_enabled=False,
                 max_disp_particles=1024 * 1024, num_voxel_per_blk_axis=16, max_ray_length=10, min_ray_length=0.3,
                 internal_voxels=10, max_submap_num=1024, is_global_map=False,
                 disp_ceiling=1.8, disp_floor=-0.3, recast_step=2, color_same_proj=True,
                 max_blocks=0, max_image_pixels=640 * 480):
        super(DenseTSDF, self).__init__(voxel_scale)
        import torch
        self._torch = torch
        # derived sizes exactly as dense_tsdf.py:18-31
        self.num_voxel_per_blk_axis = num_voxel_per_blk_axis
        self.voxel_scale = voxel_scale
        self.N = math.ceil(map_scale[0] / voxel_scale / num_voxel_per_blk_axis) * num_voxel_per_blk_axis
        self.Nz = math.ceil(map_scale[1] / voxel_scale / num_voxel_per_blk_axis) * num_voxel_per_blk_axis
        self.block_num_xy = math.ceil(map_scale[0] / voxel_scale / num_voxel_per_blk_axis)
        self.block_num_z = math.ceil(map_scale[1] / voxel_scale / num_voxel_per_blk_axis)
        self.map_size_xy = voxel_scale * self.N
        self.map_size_z = voxel_scale * self.Nz
        self.max_disp_particles = max_disp_particles
        self.enable_texture = texture_enabled
        self.max_ray_length = max_ray_length
        self.min_ray_length = min_ray_length
        self.tsdf_surface_thres = self.voxel_scale * 1.8  # :39
        self.internal_voxels = internal_voxels
        self.max_submap_num = max_submap_num
        self.is_global_map = is_global_map
        self.disp_ceiling = disp_ceiling
        self.disp_floor = disp_floor
        self.recast_step = recast_step
        self.color_same_proj = color_same_proj
        self.clear_last_TSDF_exporting = False  # assigned by SubmapMapping (submap_mapping.py:134)
        self.color = None  # the colour plane lives inside the library handle (dense_tsdf.py:48-50); read it with to_numpy

        self._h = TsdfHandle(self.N, self.Nz, voxel_scale=voxel_scale, max_ray_length=max_ray_length,
                             min_ray_length=min_ray_length, internal_voxels=internal_voxels, recast_step=recast_step,
                             K=None, is_global_map=is_global_map, disp_floor=disp_floor, disp_ceiling=disp_ceiling,
                             max_submaps=min(max_submap_num, 1024), max_blocks=max_blocks,
                             max_image_pixels=max_image_pixels, texture_enabled=texture_enabled)
        self.initialize_submap_fields(self.max_submap_num)
        self._init_export_fields()
        self._queue = self._h.L.tslam_tsdf_queue_depth
        self._queue_tex = self._h.L.tslam_tsdf_queue_depth_tex
        self._queue_begin = self._h.L.tslam_tsdf_queue_depth_begin
        self._queue_end = self._h.L.tslam_tsdf_queue_depth_end
        # page-locked frames are read by the GPU straight from host memory a few calls later (include/tslam.h): keep
        # the arrays alive for as long as the library may still read them (it never runs more than 2 batches ahead)
        self._frame_refs = collections.deque(maxlen=6 * capi.MAX_BATCH)  # depth + colour image per frame
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        self._dev_index = torch.cuda.current_device()
        if texture_enabled:
            self._h.set_color_intrinsics([1, 0, 0, 0, 1, 0, 0, 0, 1], color_same_proj)
        self._pR, self._pT = self.input_R_np.ctypes.data, self.input_T_np.ctypes.data  # persistent pose buffers (BaseMap.set_pose)
        print(f"TSDF map initialized blocks {self.block_num_xy}x{self.block_num_xy}x{self.block_num_z}")

    # ------------------------------------------------------------------ fields (dense_tsdf.py:52-60, :129-134)
    def _init_export_fields(self):
        torch = self._torch
        dev = torch.device("cuda", torch.cuda.current_device())
        n = self.max_disp_particles
        fl = self._flush
        self.num_export_particles = Field(torch.zeros(1, dtype=torch.int32, device=dev), fl)
        self.num_TSDF_particles = Field(torch.zeros(1, dtype=torch.int32, device=dev), fl)
        self.num_export_ESDF_particles = Field(torch.zeros(1, dtype=torch.int32, device=dev), fl)
        self.export_x = Field(torch.full((n, 3), -100000.0, dtype=torch.float32, device=dev), fl)
        self.export_color = Field(torch.full((n, 3), 0.5, dtype=torch.float32, device=dev), fl)
        self.export_TSDF = Field(torch.zeros(n, dtype=torch.float32, device=dev), fl)
        self.export_TSDF_xyz = Field(torch.full((n, 3), -100000.0, dtype=torch.float32, device=dev), fl)

    def _on_intrinsics(self):
        self._flush()
        self._h.set_intrinsics(self.K_cam_dep)

    def set_color_camera_intrinsic(self, K):
        super(DenseTSDF, self).set_color_camera_intrinsic(K)
        self._flush()
        self._h.set_color_intrinsics(self.K_cam_color, self.color_same_proj)

    def _upload_submap_pose(self, submap_id, R, T):
        self._h.set_submap_pose(submap_id, R, T)

    # ------------------------------------------------------------------ integrate (dense_tsdf.py:157-165)
    def recast_depth_to_map(self, R, T, depthmap, texture):
        """Queue one depth frame (uint16 mm [h,w]); the library integrates queued frames in batches (:162-165).
        The frame is copied to the device right away (asynchronously when `depthmap` is pinned host memory)."""
        if self.K_cam_dep is None:
            raise RuntimeError("set_dep_camera_intrinsic() must be called before recast_depth_to_map")
        if depthmap.dtype != np.uint16 or not depthmap.flags.c_contiguous:
            depthmap = np.ascontiguousarray(depthmap, dtype=np.uint16)
        h, w = depthmap.shape
        sid = 0 if self.is_global_map else self.active_submap_id.v
        self._frame_refs.append(depthmap)
        if self.enable_texture and texture is not None and getattr(texture, "size", 0) > 0:
            # ti.static(self.enable_texture) (:205): the colour image rides along (uint8 [th,tw,3]); a caller that has no
            # image for this frame (empty array) integrates geometry only
            self.set_pose(R, T)
            if not self.color_same_proj and self.K_cam_color is None:
                raise RuntimeError("set_color_camera_intrinsic() must be called before recast_depth_to_map (color_same_proj=False)")
            if texture.dtype != np.uint8 or not texture.flags.c_contiguous:
                texture = np.ascontiguousarray(texture, dtype=np.uint8)
            th, tw = texture.shape[0], texture.shape[1]
            self._frame_refs.append(texture)
            if self.color_same_proj and (th < h or tw < w):
                raise ValueError("color_same_proj=True needs a colour image at least as large as the depth image")
            rc = self._queue_tex(self._h.h, depthmap.__array_interface__["data"][0], texture.__array_interface__["data"][0], h, w,
                                 th, tw, self._pR, self._pT, sid, self._stream_ptr())
        else:
            # the frame's copy is started first, the pose arithmetic (set_pose: convert_by_base in f64) runs while the
            # DMA is in flight, then the copy is awaited: the array has been consumed when this method returns
            st = self._stream_ptr()
            rc = self._queue_begin(self._h.h, depthmap.__array_interface__["data"][0], h, w, st)
            if rc:
                capi.check(rc)
            self.set_pose(R, T)
            rc = self._queue_end(self._h.h, self._pR, self._pT, sid, st)
        if rc:
            capi.check(rc)

    def set_frame_borrowing(self, on=True):
        """Opt in to zero-copy hand-over of PAGE-LOCKED depth frames: they are not copied, the GPU reads their sampled rows
        from host memory a few calls later, so the caller must leave them untouched until the next reader / flush.
        Default (off): recast_depth_to_map has consumed the array when it returns, like the reference."""
        self._flush()
        capi.check(self._h.L.tslam_tsdf_set_frame_mode(self._h.h, int(bool(on))))

    def set_commit_granularity(self, frames):
        """Frames per internal launch + commit (default 32).  1 = one commit per recast_depth_to_map call: the clamp at
        Wmax (dense_tsdf.py:267) then sees the same granule as a frame-by-frame run of the reference kernels."""
        self._flush()
        capi.check(self._h.L.tslam_tsdf_set_queue_launch(self._h.h, int(frames), int(frames)))

    def _stream_ptr(self):
        # raw handle of torch's current stream; the private accessor skips building a Stream object (once per frame)
        if self._raw_stream is not None:
            return self._raw_stream(self._dev_index)
        return self._torch.cuda.current_stream().cuda_stream

    def _flush(self):
        capi.check(self._h.L.tslam_tsdf_flush(self._h.h, self._stream_ptr()))

    def recast_pcl_to_map(self, R, T, xyz_array, rgb_array):
        """:157-160.  xyz_array [n,3] (any float dtype; computed in f32 like the kernel's ti.f32 cast, :171-174)."""
        self._flush()
        self.set_pose(R, T)
        xyz = np.ascontiguousarray(np.asarray(xyz_array, dtype=np.float32).reshape(-1, 3))
        s = 0 if self.is_global_map else self.active_submap_id[None]
        rgb = None
        if self.enable_texture and rgb_array is not None and np.size(rgb_array) > 0:  # :179-183
            rgb = np.ascontiguousarray(np.asarray(rgb_array, dtype=np.uint8).reshape(-1, 3))
        self._h.integrate_points(xyz, self.input_R_np, self.input_T_np, submap=s, commit=True, rgb=rgb)
        self._torch.cuda.current_stream().synchronize()  # pageable source

    # ------------------------------------------------------------------ submaps / fusion (:272-318)
    def reset(self):
        self._h.reset()

    def fuse_submaps(self, submaps):
        submaps._flush()
        t = time.time()
        self._h.fuse_from(submaps._h)
        print(f"[DenseTSDF] Fuse submaps {(time.time() - t) * 1000:.1f}ms, active local: {submaps.active_submap_id[None]} "
              f"remote: {submaps.remote_submap_num[None]}")

    # ------------------------------------------------------------------ exporters (:320-389)
    def _active(self):
        return 0 if self.is_global_map else self.active_submap_id[None]

    def cvt_occupy_to_voxels(self):
        self.cvt_TSDF_surface_to_voxels()

    def cvt_TSDF_surface_to_voxels(self):
        self._flush()
        self.num_TSDF_particles.t.zero_()  # add_to_cur=False (:342-343)
        self._h.extract_surface(self._active(), self.export_TSDF_xyz.t, self.export_color.t, self.num_TSDF_particles.t)

    def cvt_TSDF_surface_to_voxels_to(self, num_TSDF_particles, max_disp_particles, export_TSDF_xyz, export_color):
        self._flush()
        xyz = export_TSDF_xyz.t[:max_disp_particles]
        rgb = export_color.t[:max_disp_particles]
        self._h.extract_surface(self._active(), xyz, rgb, num_TSDF_particles.t)

    def cvt_TSDF_to_voxels_slice(self, z, dz=0.5, clear_last=True):
        self._flush()
        if clear_last:
            self.num_TSDF_particles.t.zero_()
        self._h.extract_slice(self._active(), z, dz, self.export_TSDF_xyz.t, self.export_TSDF.t, self.export_color.t,
                              self.num_TSDF_particles.t)

    def get_voxels_TSDF_surface(self):
        self.cvt_TSDF_surface_to_voxels()
        return self.export_TSDF_xyz.to_numpy(), self.export_TSDF.to_numpy(), None

    def get_voxels_TSDF_slice(self, z):
        self.cvt_TSDF_to_voxels_slice(z)
        return self.export_TSDF_xyz.to_numpy(), self.export_TSDF.to_numpy()

    def get_voxels_occupy(self):
        self.cvt_occupy_to_voxels()
        return self.export_TSDF_xyz.to_numpy(), self.export_color.to_numpy()

    def finalization_current_submap(self):
        self._flush()

    # ------------------------------------------------------------------ I/O (:412-515)
    def count_active(self):
        self._flush()
        return self._h.count_active(self._active())

    def to_numpy(self, data_indices, data_tsdf, data_wtsdf, data_occ, data_color):
        """Fill caller arrays (dtypes of export_submap, :459-462) with the observed voxels of the active submap."""
        self._flush()
        if self.enable_texture:
            idx, t, w, occ, col = self._h.gather(self._active(), color=True)
        else:
            idx, t, w, occ = self._h.gather(self._active())
        n = idx.shape[0]
        data_indices[:n] = idx
        data_tsdf[:n] = t
        data_wtsdf[:n] = w
        data_occ[:n] = occ
        if self.enable_texture:  # :437-440
            data_color[:n] = col

    def load_numpy(self, submap_id, data_indices, data_tsdf, data_wtsdf, data_occ, data_color):
        self._flush()
        occ = np.clip(np.asarray(data_occ).astype(np.int64), -128, 127).astype(np.int8)
        col = None
        if self.enable_texture:  # :450-453
            col = np.asarray(data_color, dtype=np.float32).reshape(-1, 3)
        self._h.scatter(submap_id, np.asarray(data_indices).astype(np.int32), np.asarray(data_tsdf, dtype=np.float32),
                        np.asarray(data_wtsdf, dtype=np.float32), occ, color=col)

    def export_submap(self):
        s = time.time()
        num = self.count_active()
        indices = np.zeros((num, 3), np.int16)
        tsdf = np.zeros((num), np.float16)
        w_tsdf = np.zeros((num), np.float16)
        occupy = np.zeros((num), np.int8)
        color = np.zeros((num, 3), np.float16) if self.enable_texture else np.array([])  # :463-466
        self.to_numpy(indices, tsdf, w_tsdf, occupy, color)
        obj = {
            'indices': indices,
            'TSDF': tsdf,
            'W_TSDF': w_tsdf,
            'color': color,
            'occupy': occupy,
            "map_scale": [self.map_size_xy, self.map_size_z],
            "voxel_scale": self.voxel_scale,
            "texture_enabled": self.enable_texture,
            "num_voxel_per_blk_axis": self.num_voxel_per_blk_axis,
        }
        print(f"Export submap {self.active_submap_id[None]} to numpy, voxels {num / 1024:.1f}k, time: {1000 * (time.time() - s):.1f}ms")
        return obj

    def saveMap(self, filename):
        np.save(filename, self.export_submap())

    @staticmethod
    def loadMap(filename):
        obj = np.load(filename, allow_pickle=True).item()
        vs = obj['voxel_scale'] if 'voxel_scale' in obj else obj['voxel_size']  # shipped fixtures carry the legacy key
        mapping = DenseTSDF(map_scale=obj['map_scale'], voxel_scale=float(vs), texture_enabled=obj['texture_enabled'],
                            num_voxel_per_blk_axis=obj['num_voxel_per_blk_axis'], is_global_map=True)
        mapping.load_numpy(0, obj['indices'], obj['TSDF'], obj['W_TSDF'], obj['occupy'], obj['color'])
        print(f"[SubmapMapping] Loaded {obj['TSDF'].shape[0]} voxels from {filename}")
        return mapping

    def input_remote_submap(self, submap):
        # remote submaps fill the table from the top (:500-515)
        self.remote_submap_num[None] = self.remote_submap_num[None] + 1
        idx = self.max_submap_num - self.remote_submap_num[None]
        R, T = submap['pose']
        color = submap['color'] if self.enable_texture else np.array([])  # :509-512
        self.load_numpy(idx, submap['indices'], submap['TSDF'], submap['W_TSDF'], submap['occupy'], color)
        self.set_base_pose_submap(idx, R, T)
        return idx

    # ------------------------------------------------------------------ planner queries (mapping_common.py:165-204)
    def _query_raycast(self, pos, dir, max_dist):
        self._flush()
        return self._h.raycast(pos, dir, max_dist, submap=self._active())

    def _query_points(self, xyz):
        self._flush()
        return self._h.query_points(xyz, submap=self._active())   # (is_occupy, is_unobserved), dense_tsdf.py:148-155

    def _query_near(self, xyz, voxel):
        self._flush()
        return self._h.query_near_occupy(xyz, voxel, submap=self._active())

    def init_sphere(self):
        """dense_tsdf.py:136-146 as intended by tests/marching_cube_test.py:20-21: an analytic sphere SDF of radius
        3 voxels in a 30^3 region (the reference indexes its 4-D fields with three indices there and uses uncentred
        coordinates, so the kernel does not run at HEAD; this is the centred, working form: TSDF = |p| - radius)."""
        r = np.arange(-15, 15)
        I, J, K = np.meshgrid(r, r, r, indexing="ij")
        idx = np.stack([I.ravel(), J.ravel(), K.ravel()], 1).astype(np.int32)
        p = idx.astype(np.float32) * np.float32(self.voxel_scale)
        t = np.sqrt((p * p).sum(1)).astype(np.float32) - np.float32(self.voxel_scale * 3)
        self.load_numpy(self._active(), idx, t, np.ones_like(t), np.zeros(len(t), np.int8),
                        np.full((len(t), 3), 0.5, np.float32) if self.enable_texture else np.array([]))

    # ------------------------------------------------------------------ extras (not in the reference)
    def frame_counters(self):
        """Flush, commit and read the integrate counters back (bench.py end-to-end arm)."""
        self._flush()
        st = self._h.stats()
        st["d2h_bytes"] = 8 * 8 + 5 * 8 + 4 * 4
        return st

    def esdf_update(self):
        """Converged ESDF of the active submap (DenseSDF.propogate_esdf semantics, dense_esdf.py:302-333)."""
        self._flush()
        return self._h.esdf_update(self._active())

    def get_voxels_ESDF(self):
        return self._h.esdf_gather(self._active())
