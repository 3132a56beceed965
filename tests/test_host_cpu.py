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
