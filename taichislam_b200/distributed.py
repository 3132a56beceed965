"""One process per GPU: submap-sharded integration and a spatially tiled global map.

* Integration shards by SUBMAP - independent volumes (`ti.root.pointer(ti.i, submap_num)`,
  dense_tsdf.py:116): rank r integrates the frames of the submaps it owns, no data-path collective.
* The GLOBAL map is cut into tiles of whole 16^3 blocks, tile t owned by rank t.  `fuse_submaps_tiled`
  = every rank splats its own submaps, blocks that land in a foreign tile travel to their owner in ONE
  all-to-all (NCCL over NVLink on GPUs) and are reduced there; `exchange_halo` copies the one-block
  boundary layer to the neighbouring tiles as ghost blocks so marching cubes is local; `all_gather_mesh`
  is the final mesh-vertex all-gather.

The collectives are torch.distributed calls on flat tensors, so the bookkeeping below (split sizes, padding,
ownership) is exercised on CPU with the gloo backend (tests/test_distributed_cpu.py) and on GPUs with NCCL
(tests/test_gpu_multi.py, `gpurun --gpus 2`).
"""
import ctypes as C

import numpy as np


def factor_tiles(world):
    """tiles (tx,ty,tz) with tx*ty*tz == world, as cubic as possible, x >= y >= z (8 -> 2x2x2, 4 -> 2x2x1)."""
    best = (world, 1, 1)
    for tx in range(1, world + 1):
        if world % tx:
            continue
        for ty in range(1, world // tx + 1):
            if (world // tx) % ty:
                continue
            tz = world // tx // ty
            cand = tuple(sorted((tx, ty, tz), reverse=True))
            if max(cand) - min(cand) < max(best) - min(best):
                best = cand
    return best


def submap_owner(submap_id, world):
    return int(submap_id) % int(world)


def shard_frames(frame_submaps, rank, world):
    """Indices of the frames whose submap this rank owns."""
    fs = np.asarray(frame_submaps)
    return np.nonzero(fs % world == rank)[0]


def split_offsets(counts):
    counts = np.asarray(counts, dtype=np.int64)
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return off


def exchange_counts(dist, send_counts, device):
    """all_to_all of one int64 per peer -> how many rows every peer sends to me."""
    import torch
    s = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    return r.cpu().numpy()


def exchange_rows(dist, tensor, send_counts, recv_counts):
    """all_to_all_single of variable row counts (rows grouped by destination in `tensor`)."""
    import torch
    row = int(np.prod(tensor.shape[1:])) if tensor.dim() > 1 else 1
    flat = tensor.reshape(-1)
    out = torch.empty((int(np.sum(recv_counts)),) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    dist.all_to_all_single(out.reshape(-1), flat, output_split_sizes=[int(c) * row for c in recv_counts],
                           input_split_sizes=[int(c) * row for c in send_counts])
    return out


def all_gather_ragged(dist, tensor, world):
    """All-gather of per-rank row blocks of different length: counts first, then padded payload (mesh all-gather)."""
    import torch
    n = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    pad = torch.zeros((m,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    pad[:tensor.shape[0]] = tensor
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0), counts


class TiledGlobalMap:
    """A DenseTSDF global map whose volume is tiled over the ranks of a torch.distributed group."""

    def __init__(self, global_map, dist, rank, world, tiles=None):
        if getattr(global_map, "enable_texture", False):
            raise NotImplementedError("TiledGlobalMap exchanges geometry planes only; textured global maps are single-GPU")
        self.m = global_map
        self.dist = dist
        self.rank, self.world = int(rank), int(world)
        self.tiles = tuple(tiles) if tiles is not None else factor_tiles(world)
        assert self.tiles[0] * self.tiles[1] * self.tiles[2] == self.world
        self._tiles_c = (C.c_int32 * 3)(*self.tiles)
        self.last_exchange = {}

    # ---- helpers ----------------------------------------------------------------------------------------
    def _L(self):
        return self.m._h.L

    def _h(self):
        return self.m._h.h

    def owner_of_block(self, bx, by, bz):
        from . import _capi as capi
        o = C.c_int32(0)
        capi.check(self._L().tslam_tiling_owner(self._h(), self._tiles_c, self.world, int(bx), int(by), int(bz), C.byref(o)))
        return int(o.value)

    # ---- fusion -----------------------------------------------------------------------------------------
    def fuse_submaps_tiled(self, submaps):
        """fuse_submaps (dense_tsdf.py:312-318) with the result tiled over the ranks.  `submaps` = this rank's
        submap collection; the pose table of the global map must hold the poses of the submaps it contains."""
        import torch
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        dev = torch.device("cuda", torch.cuda.current_device())
        submaps._flush()
        capi.check(L.tslam_tsdf_fuse_pending(h, submaps._h.h, st))
        send = np.zeros(self.world, np.int32)
        capi.check(L.tslam_tsdf_foreign_count(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), st))
        n_send = int(send.sum())
        keys = torch.empty(max(n_send, 1), dtype=torch.int64, device=dev)
        acc = torch.empty((max(n_send, 1), 4096, 2), dtype=torch.float32, device=dev)
        obs = torch.empty((max(n_send, 1), 4096), dtype=torch.uint8, device=dev)
        occ = torch.empty((max(n_send, 1), 4096), dtype=torch.int8, device=dev)
        capi.check(L.tslam_tsdf_foreign_pack(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), n_send, capi.tptr(keys),
                                             capi.tptr(acc), capi.tptr(obs), capi.tptr(occ), st))
        recv = exchange_counts(self.dist, send, dev)
        rk = exchange_rows(self.dist, keys[:n_send], send, recv)
        ra = exchange_rows(self.dist, acc[:n_send], send, recv)
        ro = exchange_rows(self.dist, obs[:n_send], send, recv)
        rc = exchange_rows(self.dist, occ[:n_send], send, recv)
        n_recv = int(recv.sum())
        capi.check(L.tslam_tsdf_unpack_add(h, n_recv, capi.tptr(rk), capi.tptr(ra), capi.tptr(ro), capi.tptr(rc), st))
        capi.check(L.tslam_tsdf_commit_fused(h, st))
        torch.cuda.current_stream().synchronize()  # received buffers must outlive the kernels
        self.last_exchange = {"fusion_blocks_sent": n_send, "fusion_blocks_received": n_recv,
                              "fusion_bytes_sent": n_send * (8 + 4096 * 10)}

    # ---- halo + mesh --------------------------------------------------------------------------------------
    def exchange_halo(self):
        import torch
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        dev = torch.device("cuda", torch.cuda.current_device())
        send = np.zeros(self.world, np.int32)
        capi.check(L.tslam_tsdf_halo_count(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), st))
        n_send = int(send.sum())
        keys = torch.empty(max(n_send, 1), dtype=torch.int64, device=dev)
        tw = torch.empty((max(n_send, 1), 4096, 2), dtype=torch.float32, device=dev)
        obs = torch.empty((max(n_send, 1), 4096), dtype=torch.uint8, device=dev)
        capi.check(L.tslam_tsdf_halo_pack(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), n_send, capi.tptr(keys),
                                          capi.tptr(tw), capi.tptr(obs), st))
        recv = exchange_counts(self.dist, send, dev)
        rk = exchange_rows(self.dist, keys[:n_send], send, recv)
        rt = exchange_rows(self.dist, tw[:n_send], send, recv)
        ro = exchange_rows(self.dist, obs[:n_send], send, recv)
        n_recv = int(recv.sum())
        capi.check(L.tslam_tsdf_ghost_unpack(h, n_recv, capi.tptr(rk), capi.tptr(rt), capi.tptr(ro), st))
        torch.cuda.current_stream().synchronize()
        self.last_exchange.update({"halo_blocks_sent": n_send, "halo_blocks_received": n_recv})

    def generate_mesh_all_gather(self, mesher, step=1):
        """Local two-pass marching cubes on this rank's tile (+ghosts), then the mesh-vertex all-gather."""
        self.exchange_halo()
        mesher.generate_mesh(step)
        n = min(int(mesher.num_facelets[None]), mesher.max_triangles)
        v = mesher.mesh_vertices.t[:3 * n]
        nrm = mesher.mesh_normals.t[:3 * n]
        allv, counts = all_gather_ragged(self.dist, v, self.world)
        alln, _ = all_gather_ragged(self.dist, nrm, self.world)
        return allv, alln, [c // 3 for c in counts]
