#!/usr/bin/env python3
"""Golden vectors from the reference's ESDF code, EXECUTED - build container only.

dense_esdf.py is dead at HEAD: `DenseSDF.__init__` calls `BaseMap.__init__()` without its required argument
(dense_esdf.py:16 vs mapping_common.py:11) and the integrate kernels index 3-D fields with 4 components, so the class
cannot be constructed and nothing of it runs - but the ESDF propagation itself (`propogate_esdf`, `process_raise_queue`,
`process_lower_queue`, `insert_*`, `is_fixed`, dense_esdf.py:228-333) is self-contained.  This script imports the
UNMODIFIED file through oracle/taichi_emu.py, gives the class a constructor that only creates the fields those
functions touch (a subclass - the reference source is not edited), builds a small TSDF state by hand and runs
`propogate_esdf` three times (the surface moves between the rounds, once each way: lower and raise queues).  The resulting ESDF / parent_dir /
observed fields go to tests/golden/ref_exec_esdf.npz; tests/test_oracle_cpu.py::test_esdf_literal_matches_executed_reference
replays the same state through oracle/esdf_literal.py (the literal restatement "for the record").
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT] + [ROOT, os.path.join(ROOT, "tests")]
from oracle import taichi_emu as emu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_exec_esdf.npz")
N = 14


def scene(shift):
    """TSDF on an N^3 grid: a tilted plane (signed distance, truncated like a TSDF band) and a small sphere; a corner
    region stays unobserved."""
    vs = 0.05
    ii, jj, kk = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    p = np.stack([ii, jj, kk], -1).astype(np.float64) * vs
    plane = (p @ np.array([0.15, 0.1, 1.0]) / np.linalg.norm([0.15, 0.1, 1.0])) - (0.33 + shift)
    sphere = np.linalg.norm(p - np.array([0.2, 0.45, 0.2]), axis=-1) - 0.11
    t = np.minimum(plane, sphere).astype(np.float32)
    obs = np.zeros((N, N, N), bool)
    obs[1:-1, 1:-1, 1:-1] = True   # (the reference asserts 0 <= neighbour < N for every queued voxel: keep a one-voxel rim)
    obs[10:, 10:, :5] = False
    return t, obs


def main():
    emu.load_reference()
    mod = importlib.import_module("taichi_slam.mapping.dense_esdf")
    ti = sys.modules["taichi"]
    ti.is_active = lambda node, idx: node.is_active(idx)
    mod.sign = emu.ti_sign

    class Active:
        def __init__(self, field):
            self.f = field

        def is_active(self, idx):
            return emu._key(idx) in self.f.d

    class ESDFOnly(mod.DenseSDF):  # the three lines that fix the constructor: only what propogate_esdf touches
        def __init__(self, voxel_scale=0.05, max_ray_length=3.0):
            self.voxel_scale, self.gamma, self.max_ray_length, self.max_queue_size = voxel_scale, voxel_scale, max_ray_length, 1000000
            self.N_ = ti.Vector([N, N, N], ti.f32)
            self.TSDF, self.ESDF = ti.field(dtype=ti.f32), ti.field(dtype=ti.f32)
            self.observed, self.fixed = ti.field(dtype=ti.i8), ti.field(dtype=ti.i8)
            self.parent_dir = ti.Vector.field(3, dtype=ti.i32)
            self.updated_TSDF = ti.field(dtype=ti.i32)
            self.Broot = Active(self.TSDF)
            self.raise_queue = ti.Vector.field(3, dtype=ti.i32, shape=self.max_queue_size)
            self.lower_queue = ti.Vector.field(3, dtype=ti.i32, shape=self.max_queue_size)
            self.num_raise_queue, self.num_lower_queue = ti.field(dtype=ti.i32, shape=()), ti.field(dtype=ti.i32, shape=())
            self.head_lower_queue, self.head_raise_queue = ti.field(dtype=ti.i32, shape=()), ti.field(dtype=ti.i32, shape=())
            self.neighbors = [ti.Vector([a, b, c], ti.f32) for a in range(-1, 2) for b in range(-1, 2) for c in range(-1, 2) if a or b or c]

    m = ESDFOnly()
    out = {"N": N, "voxel_scale": 0.05, "max_ray_length": 3.0}
    import contextlib, io
    for rnd, shift in enumerate((0.0, 0.06, -0.04)):
        t, obs = scene(shift)
        m.updated_TSDF.d.clear()
        for k in np.ndindex(N, N, N):
            if obs[k]:
                m.TSDF[k] = t[k]
                m.updated_TSDF[k] = 1
        with contextlib.redirect_stdout(io.StringIO()):
            m.propogate_esdf()
        keys = sorted(m.ESDF.d.keys())
        out[f"tsdf{rnd}"], out[f"obs{rnd}"] = t, obs
        out[f"idx{rnd}"] = np.array(keys, np.int32)
        out[f"esdf{rnd}"] = np.array([m.ESDF.d[k] for k in keys], np.float32)
        out[f"observed{rnd}"] = np.array([int(m.observed.d.get(k, 0)) for k in keys], np.int8)
        out[f"parent{rnd}"] = np.array([np.asarray(m.parent_dir.d.get(k, np.zeros(3, np.int32))) for k in keys], np.int32)
        out[f"queues{rnd}"] = np.array([int(m.num_raise_queue[None]), int(m.num_lower_queue[None])])
        print(f"round {rnd}: {len(keys)} ESDF cells, raise queue {out[f'queues{rnd}'][0]}, lower queue {out[f'queues{rnd}'][1]}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
