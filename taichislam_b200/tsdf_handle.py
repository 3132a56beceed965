"""TsdfHandle - the thinnest Python face of the TSDF part of the C ABI.

Used by the reference-surface classes in taichislam_b200.mapping, by bench.py and by
the parity tests (which therefore exercise the C ABI, not private Python helpers).
"""
import ctypes as C

import numpy as np

from . import _capi as capi


class TsdfHandle:
    def __init__(self, N, Nz, voxel_scale=0.05, max_ray_length=10.0, min_ray_length=0.3, internal_voxels=10,
                 recast_step=2, K=None, is_global_map=False, disp_floor=-0.3, disp_ceiling=1.8, max_submaps=1024,
                 max_blocks=0, max_image_pixels=0, max_points=0, texture_enabled=False):
        capi.require_gpu()
        import torch  # device buffers / streams only
        self.torch = torch
        L = capi.load()
        K = K if K is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        self.cfg = capi.TsdfConfig(voxel_scale, N, Nz, max_ray_length, min_ray_length, internal_voxels, recast_step,
                                   K[0], K[4], K[2], K[5], int(is_global_map), disp_floor, disp_ceiling,
                                   max_submaps, max_blocks, max_image_pixels, max_points, int(bool(texture_enabled)))
        self.texture_enabled = bool(texture_enabled)
        h = C.c_void_p()
        torch.cuda.init()
        torch.cuda.current_stream()  # make sure the primary context exists and is current
        capi.check(L.tslam_tsdf_create(C.byref(self.cfg), C.byref(h)))
        self.h = h
        self.L = L
        self.N, self.Nz = N, Nz

    def close(self):
        if getattr(self, "h", None):
            self.L.tslam_tsdf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ---------------------------------------------------------------------------
    def reset(self):
        capi.check(self.L.tslam_tsdf_reset(self.h, capi.stream_ptr()))

    def set_intrinsics(self, K9):
        capi.check(self.L.tslam_tsdf_set_intrinsics(self.h, K9[0], K9[4], K9[2], K9[5]))

    def set_color_intrinsics(self, K9, color_same_proj=False):
        """set_color_camera_intrinsic (mapping_common.py:28-29) + color_same_proj (dense_tsdf.py:16)."""
        capi.check(self.L.tslam_tsdf_set_color_intrinsics(self.h, K9[0], K9[4], K9[2], K9[5], int(bool(color_same_proj))))

    def set_submap_pose(self, s, R, T):
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        capi.check(self.L.tslam_tsdf_set_submap_pose(self.h, int(s), capi.np_ptr(R), capi.np_ptr(T)))

    # -- integrate ------------------------------------------------------------------------
    def integrate_depth(self, depth, Rs, Ts, submaps=None, commit=True, texture=None):
        """depth: uint16 [n,h,w] (or [h,w]) numpy (host) or torch CUDA tensor; Rs [n,3,3]; Ts [n,3];
        texture: uint8 [n,th,tw,3] in the same memory space as depth (textured maps)."""
        torch = self.torch
        if isinstance(depth, torch.Tensor):
            assert depth.is_cuda and depth.dtype in (torch.uint16, torch.int16) and depth.is_contiguous()
            mem, ptr, shape = capi.MEM_DEVICE, capi.tptr(depth), tuple(depth.shape)
        else:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            mem, ptr, shape = capi.MEM_HOST, capi.np_ptr(depth), depth.shape
        if len(shape) == 2:
            shape = (1,) + tuple(shape)
        n, h, w = shape
        Rs = capi.f32c(Rs).reshape(n, 9)
        Ts = capi.f32c(Ts).reshape(n, 3)
        sm = None
        if submaps is not None:
            sm = np.ascontiguousarray(np.broadcast_to(np.asarray(submaps, dtype=np.int32), (n,)))
        if texture is not None:
            if isinstance(texture, torch.Tensor):
                assert mem == capi.MEM_DEVICE and texture.is_cuda and texture.dtype == torch.uint8 and texture.is_contiguous()
                tptr, tshape = capi.tptr(texture), tuple(texture.shape)
            else:
                assert mem == capi.MEM_HOST
                texture = np.ascontiguousarray(texture, dtype=np.uint8)
                tptr, tshape = capi.np_ptr(texture), texture.shape
            if len(tshape) == 3:
                tshape = (1,) + tuple(tshape)
            assert tshape[0] == n and tshape[3] == 3
            capi.check(self.L.tslam_tsdf_integrate_depth_tex(self.h, ptr, tptr, mem, n, h, w, tshape[1], tshape[2], capi.np_ptr(Rs),
                                                             capi.np_ptr(Ts), capi.np_ptr(sm) if sm is not None else None,
                                                             capi.F_COMMIT if commit else 0, capi.stream_ptr()))
            return
        capi.check(self.L.tslam_tsdf_integrate_depth(self.h, ptr, mem, n, h, w, capi.np_ptr(Rs), capi.np_ptr(Ts),
                                                     capi.np_ptr(sm) if sm is not None else None,
                                                     capi.F_COMMIT if commit else 0, capi.stream_ptr()))

    def integrate_points(self, xyz, R, T, submap=0, commit=True, rgb=None):
        torch = self.torch
        if isinstance(xyz, torch.Tensor):
            assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()
            mem, ptr, n = capi.MEM_DEVICE, capi.tptr(xyz), xyz.shape[0]
        else:
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            mem, ptr, n = capi.MEM_HOST, capi.np_ptr(xyz), xyz.shape[0]
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        if rgb is not None:
            if isinstance(rgb, torch.Tensor):
                assert mem == capi.MEM_DEVICE and rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
                cptr = capi.tptr(rgb)
            else:
                assert mem == capi.MEM_HOST
                rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
                cptr = capi.np_ptr(rgb)
            assert tuple(rgb.shape) == (n, 3)
            capi.check(self.L.tslam_tsdf_integrate_points_rgb(self.h, ptr, cptr, mem, n, capi.np_ptr(R), capi.np_ptr(T), int(submap),
                                                              capi.F_COMMIT if commit else 0, capi.stream_ptr()))
            return
        capi.check(self.L.tslam_tsdf_integrate_points(self.h, ptr, mem, n, capi.np_ptr(R), capi.np_ptr(T), int(submap),
                                                      capi.F_COMMIT if commit else 0, capi.stream_ptr()))

    def commit(self):
        capi.check(self.L.tslam_tsdf_commit(self.h, capi.stream_ptr()))

    def sync(self):
        capi.check(self.L.tslam_tsdf_sync(self.h, capi.stream_ptr()))

    # -- readers --------------------------------------------------------------------------
    def count_active(self, submap=0):
        self.torch.cuda.current_stream().synchronize()
        n = C.c_int64(0)
        capi.check(self.L.tslam_tsdf_count_active(self.h, int(submap), C.byref(n)))
        return int(n.value)

    def gather_device(self, submap=0, cap=None, color=False):
        """Observed voxels of `submap` as torch CUDA tensors (idx int32[n,3], tsdf, w f32[n], occ int8[n]
        [, color f32[n,3]])."""
        torch = self.torch
        if cap is None:
            cap = self.count_active(submap)
        dev = torch.device("cuda", torch.cuda.current_device())
        idx = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev)
        t = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        w = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        occ = torch.empty(max(cap, 1), dtype=torch.int8, device=dev)
        n = C.c_int64(0)
        if color:
            col = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=dev)
            capi.check(self.L.tslam_tsdf_gather2(self.h, int(submap), cap, capi.tptr(idx), capi.tptr(t), capi.tptr(w),
                                                 capi.tptr(occ), capi.tptr(col), C.byref(n), capi.stream_ptr()))
            k = int(n.value)
            return idx[:k], t[:k], w[:k], occ[:k], col[:k]
        capi.check(self.L.tslam_tsdf_gather(self.h, int(submap), cap, capi.tptr(idx), capi.tptr(t), capi.tptr(w),
                                            capi.tptr(occ), C.byref(n), capi.stream_ptr()))
        k = int(n.value)
        return idx[:k], t[:k], w[:k], occ[:k]

    def gather(self, submap=0, color=False):
        return tuple(a.cpu().numpy() for a in self.gather_device(submap, color=color))

    def scatter(self, submap, idx, tsdf, w, occ, color=None):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())

        def dv(a, dt):
            if isinstance(a, torch.Tensor):
                return a.to(device=dev, dtype=dt).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a)).to(device=dev).to(dt).contiguous()

        idx = dv(idx, torch.int32)
        tsdf, w = dv(tsdf, torch.float32), dv(w, torch.float32)
        occ = dv(occ, torch.int8)
        n = idx.shape[0]
        if color is not None:
            color = dv(color, torch.float32).reshape(n, 3)
            capi.check(self.L.tslam_tsdf_scatter2(self.h, int(submap), n, capi.tptr(idx), capi.tptr(tsdf), capi.tptr(w),
                                                  capi.tptr(occ), capi.tptr(color), capi.stream_ptr()))
        else:
            capi.check(self.L.tslam_tsdf_scatter(self.h, int(submap), n, capi.tptr(idx), capi.tptr(tsdf), capi.tptr(w),
                                                 capi.tptr(occ), capi.stream_ptr()))
        torch.cuda.current_stream().synchronize()  # the temporaries above must outlive the kernel

    def fuse_from(self, src):
        capi.check(self.L.tslam_tsdf_fuse(self.h, src.h, capi.stream_ptr()))

    def extract_surface(self, submap, xyz, rgb, count):
        """Append to torch CUDA buffers xyz/rgb f32[cap,3]; count: int32[1] device counter."""
        capi.check(self.L.tslam_tsdf_extract_surface(self.h, int(submap), xyz.shape[0], capi.tptr(xyz), capi.tptr(rgb),
                                                     capi.tptr(count), capi.stream_ptr()))

    def extract_slice(self, submap, z, dz, xyz, val, rgb, count):
        capi.check(self.L.tslam_tsdf_extract_slice(self.h, int(submap), float(z), float(dz), xyz.shape[0], capi.tptr(xyz),
                                                   capi.tptr(val), capi.tptr(rgb), capi.tptr(count), capi.stream_ptr()))

    def surface(self, submap=0, cap=1 << 22):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        rgb = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.extract_surface(submap, xyz, rgb, cnt)
        n = int(cnt.item())
        k = min(n, cap)
        return n, xyz[:k].cpu().numpy(), rgb[:k].cpu().numpy()

    def slice(self, z, dz=0.5, submap=0, cap=1 << 22):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        val = torch.empty(cap, dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.extract_slice(submap, z, dz, xyz, val, None, cnt)
        n = int(cnt.item())
        k = min(n, cap)
        return n, xyz[:k].cpu().numpy(), val[:k].cpu().numpy()

    def marching_cubes(self, step=1, thres=0.1, cap_tri=1 << 21, color=False):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        v = torch.empty((3 * cap_tri, 3), dtype=torch.float32, device=dev)
        nrm = torch.empty((3 * cap_tri, 3), dtype=torch.float32, device=dev)
        n = C.c_int64(0)
        if color:
            col = torch.empty((3 * cap_tri, 3), dtype=torch.float32, device=dev)
            rc = self.L.tslam_mc_generate2(self.h, int(step), float(thres), cap_tri, capi.tptr(v), capi.tptr(nrm), capi.tptr(col),
                                           C.byref(n), capi.stream_ptr())
        else:
            rc = self.L.tslam_mc_generate(self.h, int(step), float(thres), cap_tri, capi.tptr(v), capi.tptr(nrm), C.byref(n),
                                          capi.stream_ptr())
        if rc != capi.E_CAPACITY:
            capi.check(rc)
        k = min(int(n.value), cap_tri)
        if color:
            return int(n.value), v[:3 * k].cpu().numpy(), nrm[:3 * k].cpu().numpy(), col[:3 * k].cpu().numpy()
        return int(n.value), v[:3 * k].cpu().numpy(), nrm[:3 * k].cpu().numpy()

    # -- planner queries (batched @ti.func helpers of BaseMap, mapping_common.py:165-204) ------------------
    def _dev_f32(self, a):
        torch = self.torch
        if isinstance(a, torch.Tensor):
            return a.to(device="cuda", dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()

    def query_points(self, xyz, submap=0):
        """(is_pos_occupy, is_pos_unobserved) per point."""
        torch = self.torch
        x = self._dev_f32(xyz)
        f = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
        capi.check(self.L.tslam_tsdf_query_points(self.h, int(submap), x.shape[0], capi.tptr(x), capi.tptr(f), capi.stream_ptr()))
        f = f.cpu().numpy()
        return (f & 1).astype(bool), (f & 2).astype(bool)

    def query_near_occupy(self, xyz, voxel, submap=0):
        torch = self.torch
        x = self._dev_f32(xyz)
        f = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
        capi.check(self.L.tslam_tsdf_query_near_occupy(self.h, int(submap), x.shape[0], capi.tptr(x), int(voxel), capi.tptr(f), capi.stream_ptr()))
        return f.cpu().numpy().astype(bool)

    def raycast(self, pos, direction, max_dist, submap=0):
        """BaseMap.raycast for a batch of rays: (hit, last position, length)."""
        torch = self.torch
        p, d = self._dev_f32(pos), self._dev_f32(direction)
        n = p.shape[0]
        hit = torch.empty(n, dtype=torch.uint8, device=p.device)
        xyz = torch.empty((n, 3), dtype=torch.float32, device=p.device)
        ln = torch.empty(n, dtype=torch.float32, device=p.device)
        capi.check(self.L.tslam_tsdf_raycast(self.h, int(submap), n, capi.tptr(p), capi.tptr(d), float(max_dist), capi.tptr(hit), capi.tptr(xyz),
                                             capi.tptr(ln), capi.stream_ptr()))
        return hit.cpu().numpy().astype(bool), xyz.cpu().numpy(), ln.cpu().numpy()

    def esdf_update(self, submap=0):
        sw = C.c_int32(0)
        capi.check(self.L.tslam_esdf_update(self.h, int(submap), C.byref(sw), capi.stream_ptr()))
        return int(sw.value)

    def esdf_update2(self, submap=0, full=False):
        """-> dict(lower_sweeps, raise_sweeps (-1: full recompute), changed, suspect)."""
        st = np.zeros(4, np.int32)
        capi.check(self.L.tslam_esdf_update2(self.h, int(submap), 1 if full else 0, capi.np_ptr(st), capi.stream_ptr()))
        return dict(lower_sweeps=int(st[0]), raise_sweeps=int(st[1]), changed=int(st[2]), suspect=int(st[3]))

    def esdf_gather(self, submap=0):
        torch = self.torch
        cap = self.count_active(submap)
        dev = torch.device("cuda", torch.cuda.current_device())
        idx = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev)
        e = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        n = C.c_int64(0)
        capi.check(self.L.tslam_esdf_gather(self.h, int(submap), cap, capi.tptr(idx), capi.tptr(e), C.byref(n),
                                            capi.stream_ptr()))
        k = int(n.value)
        return idx[:k].cpu().numpy(), e[:k].cpu().numpy()

    # -- introspection --------------------------------------------------------------------
    def stats(self, clear=False):
        o = np.zeros(8, np.int64)
        capi.check(self.L.tslam_tsdf_get_stats(self.h, capi.np_ptr(o), int(clear)))
        return dict(n_px=int(o[0]), n_valid=int(o[1]), n_rays=int(o[2]), n_updates=int(o[3]), n_oob=int(o[4]),
                    n_blocks=int(o[5]), err_flags=int(o[6]), launches=int(o[7]))

    def launch_count(self):
        return int(self.L.tslam_tsdf_launch_count(self.h))

    def set_profiling(self, on=True):
        capi.check(self.L.tslam_tsdf_set_profiling(self.h, int(on)))

    def kernel_ms(self, n=512):
        """[rows,3] ms of (bucket, ray-march, commit) for the most recent profiled integrate launches."""
        o = np.zeros((n, 3), np.float32)
        k = C.c_int32(0)
        capi.check(self.L.tslam_tsdf_kernel_ms(self.h, n, capi.np_ptr(o), C.byref(k)))
        return o[:k.value]

    def kernel_ms2(self, n=512):
        """[rows,7] ms: bucket, ray march (total), commit, then ray set-up+segment count, scan, segment fill, block march."""
        o = np.zeros((n, 7), np.float32)
        k = C.c_int32(0)
        capi.check(self.L.tslam_tsdf_kernel_ms2(self.h, n, capi.np_ptr(o), C.byref(k)))
        return o[:k.value]

    def march_stats(self):
        """Diagnostics of the block-binned ray march since the last stats(clear=True)."""
        o = np.zeros(6, np.int64)
        capi.check(self.L.tslam_tsdf_get_march_stats(self.h, capi.np_ptr(o)))
        return dict(n_segs=int(o[0]), n_items=int(o[1]), n_slow=int(o[2]), n_fallback=int(o[3]), n_generic=int(o[4]),
                    n_verify_bad=int(o[5]))
