"""Host-side logic of the class surface that needs no GPU (mapping_common.py mirror, synthetic inputs, tiling maths)."""
import numpy as np

from taichislam_b200 import synthetic as syn
from taichislam_b200.mapping.mapping_common import BaseMap
from util import rot_xyz


class _Map(BaseMap):
    pass


def test_set_pose_equals_convert_by_base_bitwise():
    """set_pose's allocation-free path == convert_by_base (mapping_common.py:91-100) + f32 cast (:149-156), bit for bit,
    with and without the submap pose table, for ndarray / list / f32 / column-vector inputs."""
    rng = np.random.default_rng(0)
    m = _Map(0.05)
    m.set_base_pose(rot_xyz(0.3, -0.2, 0.9), np.array([1.5, -0.25, 0.125]))
    Rs, Ts = syn.stream_poses(200, start=17)
    for q in range(200):
        R, T = Rs[q], Ts[q] + rng.normal(size=3)
        m.set_pose(R, T)
        R_, T_ = m.convert_by_base(R, T)
        assert np.array_equal(m.input_R_np, R_.astype(np.float32)) and np.array_equal(m.input_T_np, T_.astype(np.float32))
    s = _Map(0.05)
    s.initialize_submap_fields(8)
    for sid in range(3):
        s.submaps_base_R_np[sid] = rot_xyz(0.1 * sid, 0.2, -0.3 * sid)
        s.submaps_base_T_np[sid] = [0.3 * sid, -0.2, 0.1]
    for sid in range(3):
        s.active_submap_id[None] = sid
        for variant in (lambda a: a, lambda a: a.tolist(), lambda a: a.astype(np.float32)):
            R, T = variant(Rs[sid]), variant(Ts[sid])
            s.set_pose(R, T)
            R_, T_ = s.convert_by_base(R, T)
            assert np.array_equal(s.input_R_np, R_.astype(np.float32)) and np.array_equal(s.input_T_np, T_.astype(np.float32))
        s.set_pose(Rs[sid], Ts[sid].reshape(3, 1)[:, 0])
    # the buffers handed to the C ABI keep their addresses (DenseTSDF caches the pointers)
    p0 = (s.input_R_np.ctypes.data, s.input_T_np.ctypes.data)
    s.set_pose(Rs[5], Ts[5])
    assert p0 == (s.input_R_np.ctypes.data, s.input_T_np.ctypes.data)
    assert s.input_R_np.dtype == np.float32 and s.input_R_np.flags.c_contiguous


def test_synthetic_inputs_are_stable():
    """The seeded synthetic inputs every golden vector depends on (scene generators, pose stream, texture)."""
    import hashlib
    d = syn.scene_sphere(4.0)
    assert d.shape == (480, 640) and d.dtype == np.uint16 and 3900 < int(d[240, 320]) <= 4000
    assert syn.scene_room().shape == (480, 640) and syn.scene_noise().min() >= 1500
    R, T = syn.stream_pose(250)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.allclose(T, [0.0, 0.5, 0.0], atol=1e-12)
    t = syn.texture_gradient(3)
    assert t.shape == (480, 640, 3) and t.dtype == np.uint8 and t[..., 0].min() >= 1
    assert np.array_equal(t, syn.texture_gradient(3)) and not np.array_equal(t, syn.texture_gradient(4))
    assert hashlib.sha256(syn.texture_gradient(100).tobytes()).hexdigest()[:8] == hashlib.sha256(syn.texture_gradient(100).tobytes()).hexdigest()[:8]


def test_topo_sample_directions_are_the_fibonacci_sphere():
    from taichislam_b200.mapping.topo_graph import TopoGraphGen
    t = TopoGraphGen.__new__(TopoGraphGen)
    t.generate_uniform_sample_points(128)   # topo_graph.py:211-224
    d = t.sample_dirs
    assert d.shape == (128, 3) and d.dtype == np.float32
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
    assert np.allclose(d[0], [0, 1, 0], atol=1e-7) and np.allclose(d[-1], [0, -1, 0], atol=1e-6)
    assert abs(d.mean(0)).max() < 0.02  # evenly spread


# ---------------------------------------------------------------------------------------------------------------
# Replay of golden vectors produced by the REFERENCE'S OWN host-side code (tools/make_golden_host.py: the unmodified
# classes of /root/reference executed with a stand-in taichi module): constructor arithmetic, convert_by_base, the
# TopoGraphGen ray directions.  These rows of the path are therefore pinned against the reference itself.
# ---------------------------------------------------------------------------------------------------------------
import json
import os
from unittest import mock

HOST_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_reference.json")))


def _unhex(v):
    return float.fromhex(v) if isinstance(v, str) else v


def _construct_without_gpu(cls_name, kwargs):
    """Run the class constructor's host arithmetic with the device handle stubbed out (no GPU in the CPU suite)."""
    import taichislam_b200.mapping.dense_tsdf as dt
    import taichislam_b200.mapping.taichi_octomap as oc
    import torch
    with mock.patch.object(dt, "TsdfHandle", mock.MagicMock()), mock.patch.object(oc, "OctoHandle", mock.MagicMock()), \
            mock.patch.object(dt.DenseTSDF, "_init_export_fields", lambda self: None), \
            mock.patch.object(torch.cuda, "current_device", lambda: 0), \
            mock.patch.object(oc, "Field", mock.MagicMock()), mock.patch.object(torch, "zeros", mock.MagicMock()), \
            mock.patch.object(torch, "full", mock.MagicMock()), mock.patch.object(torch, "device", mock.MagicMock()):
        return (dt.DenseTSDF if cls_name == "DenseTSDF" else oc.Octomap)(**kwargs)


def test_constructor_arithmetic_matches_reference():
    for case in HOST_GOLD["dense_tsdf_ctor"]:
        m = _construct_without_gpu("DenseTSDF", case["kwargs"])
        for k, v in case["attrs"].items():
            assert getattr(m, k) == _unhex(v), (case["kwargs"], k, getattr(m, k), _unhex(v))
    for case in HOST_GOLD["octomap_ctor"]:
        m = _construct_without_gpu("Octomap", case["kwargs"])
        for k, v in case["attrs"].items():
            assert getattr(m, k) == _unhex(v), (case["kwargs"], k, getattr(m, k), _unhex(v))


def test_convert_by_base_matches_reference_bitwise():
    for c in HOST_GOLD["convert_by_base"]:
        m = _Map(0.05)
        arr = lambda key, shape: np.array([float.fromhex(x) for x in c[key]]).reshape(shape)  # noqa: E731
        m.base_R_np, m.base_T_np = arr("base_R", (3, 3)), arr("base_T", 3)
        if c["submap_enabled"]:
            m.initialize_submap_fields(4)
            m.submaps_base_R_np[:] = arr("sub_R", (4, 3, 3))
            m.submaps_base_T_np[:] = arr("sub_T", (4, 3))
            m.active_submap_id[None] = c["sid"]
        R, T = arr("R", (3, 3)), arr("T", 3)
        R_, T_ = m.convert_by_base(R, T)
        assert np.array_equal(R_, arr("R_out", (3, 3))) and np.array_equal(T_, arr("T_out", 3))
        m.set_pose(R, T)
        assert np.array_equal(m.input_R_np, arr("R_out", (3, 3)).astype(np.float32))
        assert np.array_equal(m.input_T_np, arr("T_out", 3).astype(np.float32))


def test_topo_sample_dirs_match_reference_bitwise():
    from taichislam_b200.mapping.topo_graph import TopoGraphGen
    for n, vals in HOST_GOLD["topo_sample_dirs_f32"].items():
        t = TopoGraphGen.__new__(TopoGraphGen)
        t.generate_uniform_sample_points(int(n))
        ref = np.array([float.fromhex(v) for v in vals], np.float32).reshape(int(n), 3)
        assert np.array_equal(t.sample_dirs, ref)


def test_submap_mapping_call_trace_matches_reference():
    """The scripted session of tests/submap_scenario.py through taichislam_b200.mapping.SubmapMapping, map methods
    replaced by recorders: the trace (which map receives which call with which bit-exact pose, what goes on the wire,
    final submap table) equals the one recorded from the REFERENCE's SubmapMapping.  Only difference allowed: the
    reference's hard-coded `saveMap("/home/xuhao/output/test_map.npy")` (submap_mapping.py:144-145) is opt-in here."""
    import contextlib
    import io
    import submap_scenario as sc
    import taichislam_b200.mapping.dense_tsdf as dt
    import taichislam_b200.mapping.submap_mapping as smm
    rec = sc.Recorder()

    def light_init(self, map_scale=[10, 10], voxel_scale=0.05, texture_enabled=False, max_disp_particles=1024 * 1024,
                   is_global_map=False, **kw):
        self.is_global_map = is_global_map
        self.enable_texture = texture_enabled
        self.max_disp_particles = max_disp_particles
        self.export_color = self.export_TSDF_xyz = self.num_TSDF_particles = object()

    with contextlib.ExitStack() as st:
        st.enter_context(mock.patch.object(dt.DenseTSDF, "__init__", light_init))
        for pch in rec.patches(dt.DenseTSDF):
            st.enter_context(pch)
        with contextlib.redirect_stdout(io.StringIO()):
            sc.run(smm.SubmapMapping, dt.DenseTSDF, rec)
    ref = [e for e in HOST_GOLD["submap_mapping_trace"] if e[1] != "saveMap"]
    got = json.loads(json.dumps([e for e in rec.trace if e[1] != "saveMap"]))
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a == b, (a[:3], b[:3])


def test_submap_mapping_call_trace_matches_reference_octomap():
    """The same session with Octomap as the map class (SubmapMapping(Octomap, ...), taichislam_node.py:199-201), incl. the
    level-of-detail export calls with display level 0 (what the node passes) and 2."""
    import contextlib
    import io
    import submap_scenario as sc
    import taichislam_b200.mapping.taichi_octomap as oc
    import taichislam_b200.mapping.submap_mapping as smm
    rec = sc.Recorder()

    def light_init(self, map_scale=[10, 10], voxel_scale=0.05, texture_enabled=False, max_disp_particles=1000000, is_global_map=False, **kw):
        self.is_global_map = is_global_map
        self.enable_texture = texture_enabled
        self.max_disp_particles = max_disp_particles
        self.export_color = self.export_x = self.num_export_particles = object()

    with contextlib.ExitStack() as st:
        st.enter_context(mock.patch.object(oc.Octomap, "__init__", light_init))
        for pch in rec.patches(oc.Octomap):
            st.enter_context(pch)
        with contextlib.redirect_stdout(io.StringIO()):
            sc.run(smm.SubmapMapping, oc.Octomap, rec, octomap=True)
    ref = [e for e in HOST_GOLD["submap_mapping_trace_octomap"] if e[1] != "saveMap"]
    got = json.loads(json.dumps([e for e in rec.trace if e[1] != "saveMap"]))
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a == b, (a[:3], b[:3])
