"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic in taichislam_b200/distributed.py:
ownership, frame sharding, variable-size all-to-all bookkeeping and the ragged mesh all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichislam_b200 import distributed as D


def test_factor_tiles():
    assert D.factor_tiles(1) == (1, 1, 1)
    assert D.factor_tiles(2) == (2, 1, 1)
    assert D.factor_tiles(4) == (2, 2, 1)
    assert D.factor_tiles(8) == (2, 2, 2)
    for w in range(1, 17):
        t = D.factor_tiles(w)
        assert t[0] * t[1] * t[2] == w


def test_shard_frames_partitions_the_stream():
    frame_submaps = np.repeat(np.arange(13), 7)
    seen = np.zeros(len(frame_submaps), int)
    for world in (1, 2, 4, 8):
        seen[:] = 0
        for r in range(world):
            idx = D.shard_frames(frame_submaps, r, world)
            assert np.all(frame_submaps[idx] % world == r)
            seen[idx] += 1
        assert np.all(seen == 1)  # every frame integrated exactly once
    assert D.submap_owner(11, 8) == 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        # rank r sends (r+1)*(d+1) rows to rank d; row payload identifies (src, dst, i)
        send = np.array([(rank + 1) * (d + 1) for d in range(world)], np.int64)
        rows = []
        for d in range(world):
            for i in range(send[d]):
                rows.append([rank, d, i])
        t = torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)
        planes = t.to(torch.float32).unsqueeze(-1).repeat(1, 1, 2)  # [n,3,2] like (key, acc-plane) pairs
        recv = D.exchange_counts(dist, send, dev)
        assert list(recv) == [(s + 1) * (rank + 1) for s in range(world)]
        got = D.exchange_rows(dist, t, send, recv)
        gotp = D.exchange_rows(dist, planes, send, recv)
        off = D.split_offsets(recv)
        for s in range(world):
            blk = got[off[s]:off[s + 1]]
            assert torch.all(blk[:, 0] == s) and torch.all(blk[:, 1] == rank)
            assert blk[:, 2].tolist() == list(range(int(recv[s])))
        assert torch.equal(gotp[..., 0], got.to(torch.float32))
        # ragged mesh all-gather: rank r contributes 3*(r+2) vertices
        v = torch.full((3 * (rank + 2), 3), float(rank))
        allv, counts = D.all_gather_ragged(dist, v, world)
        assert counts == [3 * (r + 2) for r in range(world)]
        o = 0
        for r in range(world):
            assert torch.all(allv[o:o + counts[r]] == float(r))
            o += counts[r]
        # zero-row exchange does not hang
        z = D.exchange_rows(dist, t[:0], np.zeros(world, np.int64), np.zeros(world, np.int64))
        assert z.shape[0] == 0
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_balanced_tiling_follows_the_data():
    """Flat flight volume: all blocks within 3 z-layers.  Equal slices (2x2x2) would leave 4 of 8 tiles empty; the
    data-driven tiling puts no cut through empty space and balances the marginals."""
    from taichislam_b200.distributed import balanced_tiling
    hist = np.zeros((3, 1024), np.int64)
    hist[0, 512 - 40:512 + 40] = 100
    hist[1, 512 - 30:512 + 50] = 100
    hist[2, 512 - 1:512 + 2] = [2000, 4000, 2000]
    tiles, cuts = balanced_tiling(hist, 8)
    assert tiles[0] * tiles[1] * tiles[2] == 8 and tiles[2] <= 2
    for a in range(3):
        c = np.asarray(cuts[a]) + 512
        shares = [hist[a][c[i]:c[i + 1]].sum() / hist[a].sum() for i in range(tiles[a])]
        assert len(c) == tiles[a] + 1 and np.all(np.diff(c) > 0) and abs(sum(shares) - 1) < 1e-9
        assert max(shares) <= 1.0 / tiles[a] + 0.26   # (a 3-layer axis cannot be cut finer than its layers)
    assert balanced_tiling(np.zeros((3, 1024)), 4)[1] is None
