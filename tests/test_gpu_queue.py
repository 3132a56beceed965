"""The per-frame queue of the C ABI (tslam_tsdf_queue_depth / tslam_tsdf_flush): every way a frame can reach the
device - pageable host memory (DMA copy), page-locked host memory (sampled rows gathered by the GPU, no copy),
widths that do not qualify for the gather (fallback copy), mixed batches, geometry changes mid-queue - must build
the same map as the batched device-resident call."""
import ctypes as C

import numpy as np
import pytest

from taichislam_b200 import synthetic as syn
from util import compare_voxels

pytestmark = pytest.mark.gpu


def queue_frames(g, frames, Rs, Ts, submap=0):
    from taichislam_b200 import _capi as capi
    for d, R, T in zip(frames, Rs, Ts):
        R9, T3 = capi.f32c(R).reshape(9), capi.f32c(T).reshape(3)
        capi.check(g.L.tslam_tsdf_queue_depth(g.h, d.ctypes.data_as(C.c_void_p), d.shape[0], d.shape[1], capi.np_ptr(R9), capi.np_ptr(T3),
                                              submap, capi.stream_ptr()))
    capi.check(g.L.tslam_tsdf_flush(g.h, capi.stream_ptr()))
    g.sync()


def same_map(a, b):
    """Same voxel set, same occupancy (exact); TSDF/W equal up to the order of the f32 reductions (run-to-run noise of
    a few 1e-5 next to the sensor, where thousands of samples are summed)."""
    compare_voxels(a.gather(), b.gather(), 1e-4)


@pytest.mark.parametrize("shape", [(480, 640), (240, 424), (101, 203)])
@pytest.mark.parametrize("step", [1, 2, 3])
def test_queue_paths_equal_batched_call(shape, step):
    import torch
    from taichislam_b200.tsdf_handle import TsdfHandle
    h, w = shape
    n = 40  # crosses the queue's 32-frame launch boundary
    rng = np.random.default_rng(h + step)
    base = syn.scene_room()[:h, :w].astype(np.int32)
    frames = np.clip(base[None] + rng.integers(-40, 41, (n, h, w)), 0, 65535).astype(np.uint16)
    frames[3] = 0  # an empty frame in the stream
    Rs, Ts = syn.stream_poses(n, start=7)
    kw = dict(K=syn.K_DEPTH, is_global_map=True, recast_step=step, max_ray_length=6.0)
    ref = TsdfHandle(256, 256, **kw)
    # the queue commits per launch (32 frames, then the rest): same granularity for the reference map
    ref.integrate_depth(torch.from_numpy(frames[:32].view(np.int16)).cuda().view(torch.uint16), Rs[:32], Ts[:32])
    ref.integrate_depth(frames[32:], Rs[32:], Ts[32:])
    pinned = torch.from_numpy(frames.view(np.int16)).pin_memory().numpy().view(np.uint16)
    a = TsdfHandle(256, 256, **kw)
    queue_frames(a, list(pinned), Rs, Ts)            # page-locked, default mode: copied (and awaited) at call time
    same_map(a, ref)
    a2 = TsdfHandle(256, 256, **kw)
    assert a2.L.tslam_tsdf_set_frame_mode(a2.h, 1) == 0
    queue_frames(a2, list(pinned), Rs, Ts)           # borrowed: row gather when w % 8 == 0 and step >= 2, else copies
    same_map(a2, ref)
    b = TsdfHandle(256, 256, **kw)
    queue_frames(b, [f.copy() for f in frames], Rs, Ts)  # pageable: DMA copies
    same_map(b, ref)
    c = TsdfHandle(256, 256, **kw)
    assert c.L.tslam_tsdf_set_frame_mode(c.h, 1) == 0
    mixed = [pinned[q] if q % 3 else frames[q].copy() for q in range(n)]  # both kinds inside one launch
    queue_frames(c, mixed, Rs, Ts)
    same_map(c, ref)
    assert ref.count_active() > 1000


def test_queue_geometry_change_and_reset():
    import torch
    from taichislam_b200.tsdf_handle import TsdfHandle
    kw = dict(K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    big = torch.from_numpy(np.stack([syn.scene_room()] * 5).view(np.int16)).pin_memory().numpy().view(np.uint16)
    small = torch.from_numpy(np.stack([syn.scene_room()[:240, :320].copy()] * 5).view(np.int16)).pin_memory().numpy().view(np.uint16)
    Rs, Ts = syn.stream_poses(10)
    a, ref = TsdfHandle(256, 256, **kw), TsdfHandle(256, 256, **kw)
    queue_frames(a, list(big) + list(small), Rs, Ts)   # the size change launches the first five frames on their own
    ref.integrate_depth(big, Rs[:5], Ts[:5])
    ref.integrate_depth(small, Rs[5:], Ts[5:])
    same_map(a, ref)
    # frames queued but not flushed are discarded by reset (DenseTSDF.reset, dense_tsdf.py:309-310)
    from taichislam_b200 import _capi as capi
    R9, T3 = capi.f32c(Rs[0]).reshape(9), capi.f32c(Ts[0]).reshape(3)
    capi.check(a.L.tslam_tsdf_queue_depth(a.h, big[0].ctypes.data_as(C.c_void_p), 480, 640, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr()))
    a.reset()
    assert a.count_active() == 0
    queue_frames(a, list(big), Rs[:5], Ts[:5])
    ref.reset()
    ref.integrate_depth(big, Rs[:5], Ts[:5])
    same_map(a, ref)


def test_pinned_frame_may_be_reused_right_after_the_call():
    """Default frame mode = the reference's semantics: recast_depth_to_map has consumed the array when it returns
    (its kernel launch copies it synchronously), so a camera driver may refill ONE page-locked buffer per frame.
    (Borrowing - tslam_tsdf_set_frame_mode(m, 1) - trades that guarantee for zero copies; include/tslam.h.)"""
    import torch
    from taichislam_b200 import _capi as capi
    from taichislam_b200.tsdf_handle import TsdfHandle
    kw = dict(K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    n = 12
    rng = np.random.default_rng(5)
    frames = np.clip(syn.scene_room().astype(np.int32)[None] + rng.integers(-60, 61, (n, 480, 640)), 0, 65535).astype(np.uint16)
    Rs, Ts = syn.stream_poses(n, start=3)
    ref = TsdfHandle(256, 256, **kw)
    ref.integrate_depth(frames, Rs, Ts)
    g = TsdfHandle(256, 256, **kw)
    buf = torch.zeros((480, 640), dtype=torch.int16).pin_memory().numpy().view(np.uint16)  # the one buffer of a "driver"
    for q in range(n):
        buf[:] = frames[q]
        R9, T3 = capi.f32c(Rs[q]).reshape(9), capi.f32c(Ts[q]).reshape(3)
        capi.check(g.L.tslam_tsdf_queue_depth(g.h, buf.ctypes.data_as(C.c_void_p), 480, 640, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr()))
        buf[:] = 0  # scribble over it at once
    capi.check(g.L.tslam_tsdf_flush(g.h, capi.stream_ptr()))
    g.sync()
    same_map(g, ref)


def test_queued_frames_then_host_batch():
    """A host batch issued while frames are queued: the queued frames are integrated first (frame order decides the
    colour / Wmax rules) and the batch's staging copy must not clobber them (shared staging buffer)."""
    from taichislam_b200 import _capi as capi
    from taichislam_b200.tsdf_handle import TsdfHandle
    kw = dict(K=syn.K_DEPTH, is_global_map=True, max_ray_length=6.0)
    rng = np.random.default_rng(11)
    frames = np.clip(syn.scene_room().astype(np.int32)[None] + rng.integers(-60, 61, (9, 480, 640)), 0, 65535).astype(np.uint16)
    Rs, Ts = syn.stream_poses(9, start=20)
    ref = TsdfHandle(256, 256, **kw)
    ref.integrate_depth(frames[:5], Rs[:5], Ts[:5])
    ref.integrate_depth(frames[5:], Rs[5:], Ts[5:])
    g = TsdfHandle(256, 256, **kw)
    for q in range(5):
        R9, T3 = capi.f32c(Rs[q]).reshape(9), capi.f32c(Ts[q]).reshape(3)
        f = frames[q].copy()
        capi.check(g.L.tslam_tsdf_queue_depth(g.h, f.ctypes.data_as(C.c_void_p), 480, 640, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr()))
    g.integrate_depth(frames[5:], Rs[5:], Ts[5:])   # host batch, no flush in between
    for q in range(3):  # and the queue keeps working afterwards
        R9, T3 = capi.f32c(Rs[q]).reshape(9), capi.f32c(Ts[q]).reshape(3)
        f = frames[q].copy()
        capi.check(g.L.tslam_tsdf_queue_depth(g.h, f.ctypes.data_as(C.c_void_p), 480, 640, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr()))
    capi.check(g.L.tslam_tsdf_flush(g.h, capi.stream_ptr()))
    ref.integrate_depth(frames[:3], Rs[:3], Ts[:3])
    same_map(g, ref)


def test_queue_and_batch_argument_errors():
    from taichislam_b200.tsdf_handle import TsdfHandle
    from taichislam_b200 import _capi as capi
    g = TsdfHandle(256, 256, K=syn.K_DEPTH, is_global_map=True, max_image_pixels=320 * 240, max_points=1000, max_submaps=4)
    d = syn.scene_plane(3.0)
    R9, T3 = capi.f32c(np.eye(3)).reshape(9), capi.f32c(np.zeros(3))
    rc = g.L.tslam_tsdf_queue_depth(g.h, d.ctypes.data_as(C.c_void_p), 480, 640, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr())
    assert rc == capi.E_INVALID and b"max_image_pixels" in g.L.tslam_last_error()
    small = np.ascontiguousarray(d[:240, :320])
    assert g.L.tslam_tsdf_queue_depth(g.h, small.ctypes.data_as(C.c_void_p), 240, 320, capi.np_ptr(R9), capi.np_ptr(T3), 7, capi.stream_ptr()) == capi.E_INVALID
    assert g.L.tslam_tsdf_queue_depth(g.h, None, 240, 320, capi.np_ptr(R9), capi.np_ptr(T3), 0, capi.stream_ptr()) == capi.E_INVALID
    with pytest.raises(capi.TslamError):
        g.integrate_points(np.zeros((2000, 3), np.float32), np.eye(3), np.zeros(3))
    g.integrate_points(np.zeros((0, 3), np.float32), np.eye(3), np.zeros(3))  # n = 0 is a no-op
    g.integrate_depth(np.zeros((0, 240, 320), np.uint16), np.zeros((0, 3, 3)), np.zeros((0, 3)))
    assert g.count_active() == 0
