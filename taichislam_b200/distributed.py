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


def balanced_tiling(hist, world):
    """Data-driven tiling: `hist` = int[3,1024] marginal histograms of the occupied block coordinates (+512), summed over
    the ranks.  Tries every factorisation tx*ty*tz == world, puts the cuts of each axis at the quantiles of its marginal
    and keeps the factorisation whose heaviest tile (product of the marginal shares) is lightest; ties go to the more
    cubic one.  Returns (tiles, (cuts_x, cuts_y, cuts_z)) with block-coordinate cuts, or (factor_tiles(world), None)
    when nothing is occupied."""
    hist = np.asarray(hist, dtype=np.int64).reshape(3, 1024)
    if hist.sum() == 0:
        return factor_tiles(world), None

    def cuts_for(h, k):
        nz = np.nonzero(h)[0]
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        cs = np.cumsum(h)
        tot = cs[-1]
        cuts = [lo]
        for q in range(1, k):
            c = int(np.searchsorted(cs, tot * q / k, side="left")) + 1  # first coordinate past the q/k quantile
            c = max(c, cuts[-1] + 1)
            cuts.append(c)
        cuts.append(max(hi, cuts[-1] + 1))
        shares = [h[cuts[i]:cuts[i + 1]].sum() / tot for i in range(k)]
        return cuts, max(shares)

    best = None
    for tx in range(1, world + 1):
        if world % tx:
            continue
        for ty in range(1, world // tx + 1):
            if (world // tx) % ty:
                continue
            tz = world // tx // ty
            if max(tx, ty, tz) > 16:
                continue
            cx, sx = cuts_for(hist[0], tx)
            cy, sy = cuts_for(hist[1], ty)
            cz, sz = cuts_for(hist[2], tz)
            score = (round(sx * sy * sz * world, 3), max(tx, ty, tz) - min(tx, ty, tz))
            if best is None or score < best[0]:
                best = (score, (tx, ty, tz), ([c - 512 for c in cx], [c - 512 for c in cy], [c - 512 for c in cz]))
    return best[1], best[2]


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


class LocalGroup:
    """Stand-in for torch.distributed inside ONE process: `world` virtual ranks whose TiledGlobalMap objects are driven
    phase by phase (pack on every rank -> route -> unpack on every rank) by run_local_*.  Exercises every pack / unpack
    kernel of csrc/tslam_dist.cu and the whole ownership / split bookkeeping on a single GPU (tests/test_gpu_dist_local.py);
    the NCCL path differs only in who moves the packed rows."""

    def __init__(self, world):
        self.world = int(world)


def _route(send_bufs, send_counts, world):
    """send_bufs[r] = tuple of tensors with rows grouped by destination, send_counts[r][d] rows for d.  Returns per
    destination the tuple of concatenated tensors (source-major, like all_to_all_single) and the receive counts."""
    import torch
    offs = [split_offsets(c) for c in send_counts]
    nbuf = len(send_bufs[0])
    out, rcounts = [], []
    for d in range(world):
        out.append(tuple(torch.cat([send_bufs[r][b][offs[r][d]:offs[r][d + 1]] for r in range(world)], 0) for b in range(nbuf)))
        rcounts.append([int(send_counts[r][d]) for r in range(world)])
    return out, rcounts


class TiledGlobalMap:
    """A DenseTSDF global map whose volume is tiled over the ranks of a torch.distributed group.

    Textured maps (reference: colour is fused with geometry, dense_tsdf.py:276-277) exchange the colour planes too."""

    def __init__(self, global_map, dist, rank, world, tiles=None, balance=True):
        self.m = global_map
        self.dist = dist
        self.rank, self.world = int(rank), int(world)
        self.tiles = tuple(tiles) if tiles is not None else factor_tiles(world)
        self.balance = bool(balance) and tiles is None
        assert self.tiles[0] * self.tiles[1] * self.tiles[2] == self.world
        self._tiles_c = (C.c_int32 * 3)(*self.tiles)
        self.cuts = None
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

    def set_tiling(self, tiles, cuts):
        """Install a tiling on this rank (every rank must install the same one)."""
        from . import _capi as capi
        self.tiles = tuple(int(t) for t in tiles)
        self._tiles_c = (C.c_int32 * 3)(*self.tiles)
        self.cuts = cuts
        if cuts is None:
            capi.check(self._L().tslam_tiling_set_cuts(self._h(), None, None, None, None))
        else:
            arrs = [np.ascontiguousarray(np.asarray(c, dtype=np.int32)) for c in cuts]
            capi.check(self._L().tslam_tiling_set_cuts(self._h(), self._tiles_c, capi.np_ptr(arrs[0]), capi.np_ptr(arrs[1]), capi.np_ptr(arrs[2])))

    # ---- fusion, phase by phase -------------------------------------------------------------------------
    def fusion_splat(self, submaps):
        """Phase 1: splat this rank's submaps (sums pending) and return the marginal histograms of the touched blocks."""
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        submaps._flush()
        capi.check(L.tslam_tsdf_fuse_pending(h, submaps._h.h, st))  # (resets the global map first, dense_tsdf.py:313)
        hist = np.zeros((3, 1024), np.int32)
        capi.check(L.tslam_tsdf_dirty_hist(h, capi.np_ptr(hist), st))
        return hist

    def fusion_pack(self):
        """Phase 2: blocks that fell into a foreign tile -> (send counts, (keys, acc, obs, occ[, col])) grouped by owner."""
        import torch
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        dev = torch.device("cuda", torch.cuda.current_device())
        send = np.zeros(self.world, np.int32)
        capi.check(L.tslam_tsdf_foreign_count(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), st))
        n_send = int(send.sum())
        keys = torch.empty(max(n_send, 1), dtype=torch.int64, device=dev)
        acc = torch.empty((max(n_send, 1), 4096, 2), dtype=torch.float32, device=dev)
        obs = torch.empty((max(n_send, 1), 4096), dtype=torch.uint8, device=dev)
        occ = torch.empty((max(n_send, 1), 4096), dtype=torch.int8, device=dev)
        tex = bool(getattr(self.m, "enable_texture", False))
        col = torch.empty((max(n_send, 1), 4096, 4), dtype=torch.float32, device=dev) if tex else None
        capi.check(L.tslam_tsdf_foreign_pack2(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), n_send, capi.tptr(keys),
                                              capi.tptr(acc), capi.tptr(obs), capi.tptr(occ), capi.tptr(col), st))
        bufs = (keys[:n_send], acc[:n_send], obs[:n_send], occ[:n_send]) + ((col[:n_send],) if tex else ())
        self.last_exchange = {"fusion_blocks_sent": n_send, "fusion_bytes_sent": n_send * (8 + 4096 * (10 + (16 if tex else 0)))}
        return send, bufs

    def fusion_unpack(self, bufs):
        """Phase 3: add the received blocks, commit this rank's tile."""
        import torch
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        n_recv = int(bufs[0].shape[0])
        col = bufs[4] if len(bufs) > 4 else None
        capi.check(L.tslam_tsdf_unpack_add2(h, n_recv, capi.tptr(bufs[0]), capi.tptr(bufs[1]), capi.tptr(bufs[2]), capi.tptr(bufs[3]),
                                            capi.tptr(col), st))
        capi.check(L.tslam_tsdf_commit_fused(h, st))
        torch.cuda.current_stream().synchronize()  # received buffers must outlive the kernels
        self.last_exchange["fusion_blocks_received"] = n_recv

    def fuse_submaps_tiled(self, submaps):
        """fuse_submaps (dense_tsdf.py:312-318) with the result tiled over the ranks.  `submaps` = this rank's
        submap collection; the pose table of the global map must hold the poses of the submaps it contains."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        hist = self.fusion_splat(submaps)
        if self.balance:
            ht = torch.as_tensor(hist.astype(np.int64), device=dev)
            self.dist.all_reduce(ht)
            tiles, cuts = balanced_tiling(ht.cpu().numpy(), self.world)
            self.set_tiling(tiles, cuts)
        send, bufs = self.fusion_pack()
        recv = exchange_counts(self.dist, send, dev)
        rbufs = tuple(exchange_rows(self.dist, b, send, recv) for b in bufs)
        self.fusion_unpack(rbufs)

    # ---- halo + mesh --------------------------------------------------------------------------------------
    def halo_pack(self):
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
        tex = bool(getattr(self.m, "enable_texture", False))
        col = torch.empty((max(n_send, 1), 4096, 4), dtype=torch.float32, device=dev) if tex else None
        capi.check(L.tslam_tsdf_halo_pack2(h, self._tiles_c, self.rank, self.world, capi.np_ptr(send), n_send, capi.tptr(keys),
                                           capi.tptr(tw), capi.tptr(obs), capi.tptr(col), st))
        self.last_exchange["halo_blocks_sent"] = n_send
        return send, (keys[:n_send], tw[:n_send], obs[:n_send]) + ((col[:n_send],) if tex else ())

    def halo_unpack(self, bufs):
        import torch
        from . import _capi as capi
        L, h, st = self._L(), self._h(), capi.stream_ptr()
        n_recv = int(bufs[0].shape[0])
        col = bufs[3] if len(bufs) > 3 else None
        capi.check(L.tslam_tsdf_ghost_unpack2(h, n_recv, capi.tptr(bufs[0]), capi.tptr(bufs[1]), capi.tptr(bufs[2]), capi.tptr(col), st))
        torch.cuda.current_stream().synchronize()
        self.last_exchange["halo_blocks_received"] = n_recv

    def exchange_halo(self):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        send, bufs = self.halo_pack()
        recv = exchange_counts(self.dist, send, dev)
        self.halo_unpack(tuple(exchange_rows(self.dist, b, send, recv) for b in bufs))

    def local_mesh(self, mesher, step=1):
        """Two-pass marching cubes on this rank's tile (+ghosts): triangles of OWNED blocks only."""
        mesher.generate_mesh(step)
        n = min(int(mesher.num_facelets[None]), mesher.max_triangles)
        return mesher.mesh_vertices.t[:3 * n], mesher.mesh_normals.t[:3 * n]

    def generate_mesh_all_gather(self, mesher, step=1):
        """Local two-pass marching cubes on this rank's tile (+ghosts), then the mesh-vertex all-gather."""
        self.exchange_halo()
        v, nrm = self.local_mesh(mesher, step)
        allv, counts = all_gather_ragged(self.dist, v, self.world)
        alln, _ = all_gather_ragged(self.dist, nrm, self.world)
        return allv, alln, [c // 3 for c in counts]


def run_local_fusion(tiled, submaps):
    """Tiled fusion of `world` virtual ranks living in ONE process: tiled[r] / submaps[r] = rank r's global map wrapper
    and submap collection.  Same phases as TiledGlobalMap.fuse_submaps_tiled, the all-to-all replaced by slicing."""
    world = len(tiled)
    hists = [t.fusion_splat(s) for t, s in zip(tiled, submaps)]
    if tiled[0].balance:
        tiles, cuts = balanced_tiling(np.sum([h.astype(np.int64) for h in hists], axis=0), world)
        for t in tiled:
            t.set_tiling(tiles, cuts)
    packed = [t.fusion_pack() for t in tiled]
    routed, _ = _route([p[1] for p in packed], [p[0] for p in packed], world)
    for t, bufs in zip(tiled, routed):
        t.fusion_unpack(bufs)


def run_local_mesh(tiled, meshers, step=1):
    world = len(tiled)
    packed = [t.halo_pack() for t in tiled]
    routed, _ = _route([p[1] for p in packed], [p[0] for p in packed], world)
    for t, bufs in zip(tiled, routed):
        t.halo_unpack(bufs)
    import torch
    vs, ns = zip(*[t.local_mesh(m, step) for t, m in zip(tiled, meshers)])
    return torch.cat(vs, 0), torch.cat(ns, 0), [int(v.shape[0]) // 3 for v in vs]
