#!/usr/bin/env python3
"""Golden vectors from the REFERENCE ITSELF for the host-side (non-kernel) parts of the path - build container only.

Taichi cannot be installed here, but everything the reference computes OUTSIDE its @ti.kernel bodies is plain
Python/numpy: constructor arithmetic (grid sizes, tree depths), pose plumbing (convert_by_base), the ray
directions of TopoGraphGen.  This script imports the unmodified reference from /root/reference with a stand-in
`taichi` module (decorators are identities, fields are inert objects), runs those parts and stores inputs and
outputs - floats as exact hex strings - in tests/golden/host_reference.json.  tests/test_host_cpu.py replays them
against taichislam_b200.mapping bit for bit.
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "host_reference.json")


def install_fake_taichi():
    ti = MagicMock(name="taichi")

    def ident(*a, **k):
        return a[0] if (len(a) == 1 and callable(a[0]) and not k) else (lambda f: f)

    ti.kernel = ti.func = ti.data_oriented = ti.dataclass = ident

    import collections

    class _VecField(collections.defaultdict):
        def __init__(self):
            super().__init__(lambda: [0.0, 0.0, 0.0, 0.0])

    class Vector:
        def __new__(cls, x, *a, **k):
            return np.asarray(x)

        @staticmethod
        def field(*a, **k):
            return _VecField()

    ti.Vector = Vector
    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = MagicMock()
    for m in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "lcm", "transformations"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def main():
    install_fake_taichi()
    sys.path.insert(0, "/root/reference")
    from taichi_slam.mapping import mapping_common as mc
    from taichi_slam.mapping.dense_tsdf import DenseTSDF
    from taichi_slam.mapping.taichi_octomap import Octomap
    from taichi_slam.mapping.topo_graph import TopoGraphGen
    g = {"source": "unmodified /root/reference classes executed with a stand-in taichi module (tools/make_golden_host.py)"}

    # 1. constructor arithmetic of DenseTSDF (dense_tsdf.py:13-47)
    cases = [dict(map_scale=[12.8, 12.8], voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True),
             dict(map_scale=[25.6, 25.6], voxel_scale=0.05, num_voxel_per_blk_axis=16, is_global_map=True),
             dict(map_scale=[102.4, 102.4], voxel_scale=0.05, is_global_map=True),
             dict(map_scale=[10, 3], voxel_scale=0.07, num_voxel_per_blk_axis=10),
             dict(map_scale=[50, 10], voxel_scale=0.1, max_ray_length=64)]
    keys = ["N", "Nz", "block_num_xy", "block_num_z", "map_size_xy", "map_size_z", "tsdf_surface_thres", "num_voxel_per_blk_axis"]
    out = []
    for kw in cases:
        m = DenseTSDF(**kw)
        out.append({"kwargs": kw, "attrs": {k: (float(getattr(m, k)).hex() if isinstance(getattr(m, k), float) else int(getattr(m, k))) for k in keys}})
    g["dense_tsdf_ctor"] = out

    # 2. constructor arithmetic of Octomap (taichi_octomap.py:14-45)
    ocases = [dict(map_scale=[51.2, 51.2], voxel_scale=0.05, K=2, min_occupy_thres=2), dict(), dict(map_scale=[30, 7], voxel_scale=0.1, K=4),
              dict(map_scale=[12.8, 12.8], voxel_scale=0.05, K=2)]
    okeys = ["Rxy", "Rz", "N", "Nz", "K", "voxel_scale", "map_size_xy", "map_size_z"]
    out = []
    for kw in ocases:
        m = Octomap(**kw)
        out.append({"kwargs": kw, "attrs": {k: (float(getattr(m, k)).hex() if isinstance(getattr(m, k), float) else int(getattr(m, k))) for k in okeys}})
    g["octomap_ctor"] = out

    # 3. pose plumbing: convert_by_base (mapping_common.py:91-100) with and without the submap table
    rng = np.random.default_rng(2024)

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
                np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))

    class Obj:
        pass

    poses = []
    for q in range(24):
        o = Obj()
        o.submap_enabled = q % 2 == 1
        o.base_R_np, o.base_T_np = rot(*rng.normal(size=3)), rng.normal(size=3) * 3
        o.submaps_base_R_np = np.stack([rot(*rng.normal(size=3)) for _ in range(4)])
        o.submaps_base_T_np = rng.normal(size=(4, 3)) * 2
        sid = int(rng.integers(0, 4))
        o.active_submap_id = {None: sid}
        R, T = rot(*rng.normal(size=3)), rng.normal(size=3) * 5
        R_, T_ = mc.BaseMap.convert_by_base(o, R, T)
        poses.append({"submap_enabled": o.submap_enabled, "sid": sid, "base_R": hx(o.base_R_np), "base_T": hx(o.base_T_np),
                      "sub_R": hx(o.submaps_base_R_np), "sub_T": hx(o.submaps_base_T_np), "R": hx(R), "T": hx(T),
                      "R_out": hx(R_), "T_out": hx(T_)})
    g["convert_by_base"] = poses

    # 4. TopoGraphGen.generate_uniform_sample_points (topo_graph.py:211-224)
    dirs = {}
    for n in (64, 128):
        o = Obj()
        TopoGraphGen.generate_uniform_sample_points(o, n)
        arr = np.stack([np.asarray(o.sample_dirs[i], dtype=np.float32) for i in range(n)])
        dirs[str(n)] = [float(v).hex() for v in arr.reshape(-1)]
    g["topo_sample_dirs_f32"] = dirs

    # 5. SubmapMapping orchestration (submap_mapping.py): the call trace of a scripted session (tests/submap_scenario.py)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import contextlib
    import io as _io
    import submap_scenario as sc
    from taichi_slam.mapping.submap_mapping import SubmapMapping
    rec = sc.Recorder()
    with contextlib.ExitStack() as st:
        for pch in rec.patches(DenseTSDF):
            st.enter_context(pch)
        with contextlib.redirect_stdout(_io.StringIO()):
            sc.run(SubmapMapping, DenseTSDF, rec)
    g["submap_mapping_trace"] = rec.trace
    rec = sc.Recorder()
    with contextlib.ExitStack() as st:
        for pch in rec.patches(Octomap):
            st.enter_context(pch)
        with contextlib.redirect_stdout(_io.StringIO()):
            sc.run(SubmapMapping, Octomap, rec, octomap=True)
    g["submap_mapping_trace_octomap"] = rec.trace
    json.dump(g, open(OUT, "w"), indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
