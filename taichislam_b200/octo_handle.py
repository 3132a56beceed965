"""OctoHandle - thin Python face of the Octomap part of the C ABI (include/tslam.h)."""
import ctypes as C

import numpy as np

from . import _capi as capi


class OctoHandle:
    def __init__(self, N, Nz, K=2, voxel_scale=0.05, min_occupy_thres=3, min_ray_length=0.3, max_ray_length=3.0,
                 recast_step=2, Kcam=None, max_submaps=1024, max_blocks=0, max_image_pixels=0, max_points=0):
        capi.require_gpu()
        import torch
        self.torch = torch
        L = capi.load()
        Kc = Kcam if Kcam is not None else [1, 0, 0, 0, 1, 0, 0, 0, 1]
        self.cfg = capi.OctoConfig(voxel_scale, N, Nz, K, max_ray_length, min_ray_length, recast_step, Kc[0], Kc[4], Kc[2],
                                   Kc[5], min_occupy_thres, max_submaps, max_blocks, max_image_pixels, max_points)
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

    def integrate_points(self, xyz, R, T, submap=0):
        torch = self.torch
        if isinstance(xyz, torch.Tensor):
            assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()
            mem, ptr, n = capi.MEM_DEVICE, capi.tptr(xyz), xyz.shape[0]
        else:
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            mem, ptr, n = capi.MEM_HOST, capi.np_ptr(xyz), xyz.shape[0]
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        capi.check(self.L.tslam_octo_integrate_points(self.h, ptr, mem, n, capi.np_ptr(R), capi.np_ptr(T), int(submap),
                                                      capi.stream_ptr()))
        if mem == capi.MEM_HOST:
            torch.cuda.current_stream().synchronize()  # pageable host source must outlive the copy

    def integrate_depth(self, depth, R, T, submap=0):
        torch = self.torch
        if isinstance(depth, torch.Tensor):
            mem, ptr, (h, w) = capi.MEM_DEVICE, capi.tptr(depth), depth.shape
        else:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            mem, ptr, (h, w) = capi.MEM_HOST, capi.np_ptr(depth), depth.shape
        R, T = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        capi.check(self.L.tslam_octo_integrate_depth(self.h, ptr, mem, h, w, capi.np_ptr(R), capi.np_ptr(T), int(submap),
                                                     capi.stream_ptr()))
        if mem == capi.MEM_HOST:
            torch.cuda.current_stream().synchronize()

    def gather_device(self, submap=0, cap=None):
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
        capi.check(self.L.tslam_octo_gather(self.h, int(submap), cap, capi.tptr(idx), capi.tptr(cnt), C.byref(n),
                                            capi.stream_ptr()))
        k = int(n.value)
        return idx[:k], cnt[:k]

    def gather(self, submap=0):
        idx, cnt = self.gather_device(submap)
        return idx.cpu().numpy(), cnt.cpu().numpy().view(np.uint32)

    def extract(self, submap, level, xyz, count):
        capi.check(self.L.tslam_octo_extract(self.h, int(submap), int(level), xyz.shape[0], capi.tptr(xyz), capi.tptr(count),
                                             capi.stream_ptr()))

    def export(self, level=1, submap=0, cap=1 << 22):
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.extract(submap, level, xyz, cnt)
        n = int(cnt.item())
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
