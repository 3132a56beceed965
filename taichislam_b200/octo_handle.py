"""OctoHandle - thin Python face of the Octomap part of the C ABI (include/tslam.h)."""
import ctypes as C

import numpy as np

from . import _capi as capi


class OctoHandle:
    def __init__(self, N, Nz, K=2, voxel_scale=0.05, min_occupy_thres=3, min_ray_length=0.3, max_ray_length=3.0,
                 recast_step=2, Kcam=None, max_submaps=1024, max_blocks=0, max_image_pixels=0, max_points=0,
                 texture_enabled=False):
        capi.require_gpu()
        import torch
        self.torch = torch
        L = capi.load()
        Kc = Kcam if Kcam is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        self.cfg = capi.OctoConfig(voxel_scale, N, Nz, K, max_ray_length, min_ray_length, recast_step, Kc[0], Kc[4], Kc[2],
                                   Kc[5], min_occupy_thres, max_submaps, max_blocks, max_image_pixels, max_points,
                                   int(bool(texture_enabled)))
        self.texture_enabled = bool(texture_enabled)
        torch.cuda.init()
        torch.cuda.current_stream()
        h = C.c_void_p()
        capi.check(L.tslam_octo_create(C.byref(self.cfg), C.byref(h)))
        self.h, self.L = h, L
        self.N, self.Nz, self.K = N, Nz, K

    def close(self):
        if getattr(self, "h", None):
            self.L.tslam_octo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.L.tslam_octo_reset(self.h, capi.stream_ptr()))

    def set_submap_pose(self, s, R, T):
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        capi.check(self.L.tslam_octo_set_submap_pose(self.h, int(s), capi.np_ptr(R), capi.np_ptr(T)))

    def set_intrinsics(self, K9):
        capi.check(self.L.tslam_octo_set_intrinsics(self.h, K9[0], K9[4], K9[2], K9[5]))

    def set_color_intrinsics(self, K9, color_same_proj=False):
        capi.check(self.L.tslam_octo_set_color_intrinsics(self.h, K9[0], K9[4], K9[2], K9[5], int(bool(color_same_proj))))

    def integrate_points(self, xyz, R, T, submap=0, rgb=None):
        torch = self.torch
        if isinstance(xyz, torch.Tensor):
            assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()
            mem, ptr, n = capi.MEM_DEVICE, capi.tptr(xyz), xyz.shape[0]
        else:
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            mem, ptr, n = capi.MEM_HOST, capi.np_ptr(xyz), xyz.shape[0]
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        cptr = None
        if rgb is not None:
            if isinstance(rgb, torch.Tensor):
                assert mem == capi.MEM_DEVICE and rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous()
                cptr = capi.tptr(rgb)
            else:
                assert mem == capi.MEM_HOST
                rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
                cptr = capi.np_ptr(rgb)
            assert tuple(rgb.shape) == (n, 3)
        capi.check(self.L.tslam_octo_integrate_points_rgb(self.h, ptr, cptr, mem, n, capi.np_ptr(R), capi.np_ptr(T), int(submap),
                                                          capi.stream_ptr()))
        if mem == capi.MEM_HOST:
            torch.cuda.current_stream().synchronize()  # pageable host source must outlive the copy

    def integrate_depth(self, depth, R, T, submap=0, texture=None):
        torch = self.torch
        if isinstance(depth, torch.Tensor):
            mem, ptr, (h, w) = capi.MEM_DEVICE, capi.tptr(depth), depth.shape
        else:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            mem, ptr, (h, w) = capi.MEM_HOST, capi.np_ptr(depth), depth.shape
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        tptr, th, tw = None, 0, 0
        if texture is not None:
            if isinstance(texture, torch.Tensor):
                assert mem == capi.MEM_DEVICE and texture.is_cuda and texture.dtype == torch.uint8 and texture.is_contiguous()
                tptr = capi.tptr(texture)
            else:
                assert mem == capi.MEM_HOST
                texture = np.ascontiguousarray(texture, dtype=np.uint8)
                tptr = capi.np_ptr(texture)
            th, tw = int(texture.shape[0]), int(texture.shape[1])
        capi.check(self.L.tslam_octo_integrate_depth_tex(self.h, ptr, tptr, mem, h, w, th, tw, capi.np_ptr(R), capi.np_ptr(T),
                                                         int(submap), capi.stream_ptr()))
        if mem == capi.MEM_HOST:
            torch.cuda.current_stream().synchronize()

    def gather_device(self, submap=0, cap=None, color=False):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        n = C.c_int64(0)
        if cap is None:
            rc = self.L.tslam_octo_gather(self.h, int(submap), 0, None, None, C.byref(n), capi.stream_ptr())
            if rc not in (capi.TSLAM_OK, capi.E_CAPACITY):
                capi.check(rc)
            cap = int(n.value)
        idx = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev)
        cnt = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        col = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=dev) if color else None
        capi.check(self.L.tslam_octo_gather2(self.h, int(submap), cap, capi.tptr(idx), capi.tptr(cnt),
                                             capi.tptr(col) if color else None, C.byref(n), capi.stream_ptr()))
        k = int(n.value)
        return (idx[:k], cnt[:k], col[:k]) if color else (idx[:k], cnt[:k])

    def gather(self, submap=0, color=False):
        r = self.gather_device(submap, color=color)
        out = (r[0].cpu().numpy(), r[1].cpu().numpy().view(np.uint32))
        return out + (r[2].cpu().numpy(),) if color else out

    def extract(self, submap, level, xyz, count, rgb=None):
        capi.check(self.L.tslam_octo_extract2(self.h, int(submap), int(level), xyz.shape[0], capi.tptr(xyz),
                                              capi.tptr(rgb) if rgb is not None else None, capi.tptr(count), capi.stream_ptr()))

    def export(self, level=1, submap=0, cap=1 << 22, color=False):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        rgb = torch.zeros((cap, 3), dtype=torch.float32, device=dev) if color else None
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.extract(submap, level, xyz, cnt, rgb)
        n = int(cnt.item())
        if color:
            return n, xyz[:min(n, cap)].cpu().numpy(), rgb[:min(n, cap)].cpu().numpy()
        return n, xyz[:min(n, cap)].cpu().numpy()

    def query_points(self, xyz, submap=0):
        torch = self.torch
        x = torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.float32)).cuda()
        f = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
        capi.check(self.L.tslam_octo_query_points(self.h, int(submap), x.shape[0], capi.tptr(x), capi.tptr(f), capi.stream_ptr()))
        return f.cpu().numpy().astype(bool)

    def raycast(self, pos, direction, max_dist, submap=0):
        torch = self.torch
        p = torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float32)).cuda()
        d = torch.from_numpy(np.ascontiguousarray(direction, dtype=np.float32)).cuda()
        n = p.shape[0]
        hit = torch.empty(n, dtype=torch.uint8, device=p.device)
        xyz = torch.empty((n, 3), dtype=torch.float32, device=p.device)
        ln = torch.empty(n, dtype=torch.float32, device=p.device)
        capi.check(self.L.tslam_octo_raycast(self.h, int(submap), n, capi.tptr(p), capi.tptr(d), float(max_dist), capi.tptr(hit), capi.tptr(xyz),
                                             capi.tptr(ln), capi.stream_ptr()))
        return hit.cpu().numpy().astype(bool), xyz.cpu().numpy(), ln.cpu().numpy()

    def fuse_from(self, src):
        capi.check(self.L.tslam_octo_fuse(self.h, src.h, capi.stream_ptr()))

    def sync(self):
        capi.check(self.L.tslam_octo_sync(self.h, capi.stream_ptr()))

    def launch_count(self):
        return int(self.L.tslam_octo_launch_count(self.h))
