"""A scripted SubmapMapping session used twice: by tools/make_golden_host.py against the REFERENCE's SubmapMapping
(imported with a stand-in taichi, map methods replaced by recorders) and by tests/test_host_cpu.py against
taichislam_b200.mapping.SubmapMapping (map methods replaced by the same recorders).  The recorded call traces - which
map gets which call with which (bit-exact) pose - must be identical: that is the orchestration contract of
submap_mapping.py (keyframe -> submap policy :148-155, PGO correction :162-169, base poses :126-146, fusion :157-160,
pose-graph updates :109-124, remote submaps / trajectories :244-261)."""
import io
import zlib
from unittest import mock

import numpy as np


def hexes(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
            np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))


class Recorder:
    """Replaces the kernel-launching methods of a map class by functions that log the call."""

    NAMES = ["recast_depth_to_map", "recast_pcl_to_map", "fuse_submaps", "export_submap", "saveMap", "set_base_pose_submap",
             "switch_to_next_submap", "get_active_submap_id", "input_remote_submap", "cvt_TSDF_surface_to_voxels",
             "cvt_TSDF_surface_to_voxels_to", "cvt_TSDF_to_voxels_slice", "set_dep_camera_intrinsic", "set_color_camera_intrinsic",
             "cvt_occupy_to_voxels", "cvt_occupy_voxels_to"]

    def __init__(self):
        self.trace = []
        self.remote = 0

    def role(self, m):
        return "global" if getattr(m, "is_global_map", False) else "sub"

    def patches(self, cls):
        rec = self

        def sid(m):
            return getattr(m, "_rec_sid", 0)

        def recast_depth_to_map(m, R, T, depthmap, texture):
            rec.trace.append([rec.role(m), "recast_depth_to_map", sid(m), hexes(R), hexes(T), list(np.shape(depthmap))])

        def recast_pcl_to_map(m, R, T, xyz, rgb):
            rec.trace.append([rec.role(m), "recast_pcl_to_map", sid(m), hexes(R), hexes(T), list(np.shape(xyz))])

        def fuse_submaps(m, submaps):
            rec.trace.append([rec.role(m), "fuse_submaps", rec.role(submaps), sid(submaps)])

        def export_submap(m):
            rec.trace.append([rec.role(m), "export_submap", sid(m)])
            return {"indices": np.zeros((2, 3), np.int16), "TSDF": np.zeros(2, np.float16), "marker": sid(m)}

        def saveMap(m, filename):
            rec.trace.append([rec.role(m), "saveMap"])

        def set_base_pose_submap(m, submap_id, R, T):
            rec.trace.append([rec.role(m), "set_base_pose_submap", int(submap_id), hexes(R), hexes(T)])

        def switch_to_next_submap(m):
            m._rec_sid = sid(m) + 1
            rec.trace.append([rec.role(m), "switch_to_next_submap", m._rec_sid])
            return m._rec_sid

        def get_active_submap_id(m):
            return sid(m)

        def input_remote_submap(m, submap):
            rec.remote += 1
            rec.trace.append([rec.role(m), "input_remote_submap", int(submap["frame_id"]), hexes(submap["pose"][0]), hexes(submap["pose"][1])])
            return 1000 - rec.remote

        def cvt_TSDF_surface_to_voxels(m):
            rec.trace.append([rec.role(m), "cvt_TSDF_surface_to_voxels"])

        def cvt_TSDF_surface_to_voxels_to(m, num, max_disp, xyz, color):
            rec.trace.append([rec.role(m), "cvt_TSDF_surface_to_voxels_to", int(max_disp)])

        def cvt_TSDF_to_voxels_slice(m, z, *a, **k):
            rec.trace.append([rec.role(m), "cvt_TSDF_to_voxels_slice", float(z).hex()])

        def cvt_occupy_to_voxels(m, level=None):
            rec.trace.append([rec.role(m), "cvt_occupy_to_voxels", None if level is None else int(level)])

        def cvt_occupy_voxels_to(m, level, cur_num, max_disp, x, color):
            rec.trace.append([rec.role(m), "cvt_occupy_voxels_to", int(level), int(max_disp)])

        def set_dep_camera_intrinsic(m, K):
            rec.trace.append([rec.role(m), "set_dep_camera_intrinsic", hexes(K)])

        def set_color_camera_intrinsic(m, K):
            rec.trace.append([rec.role(m), "set_color_camera_intrinsic", hexes(K)])

        loc = locals()
        return [mock.patch.object(cls, n, loc[n], create=True) for n in self.NAMES]


def run(SubmapMapping, DenseTSDF, rec, octomap=False):
    """The session.  Returns nothing; everything of interest is in rec.trace.  `DenseTSDF` is the map class handed to
    SubmapMapping (the Octomap class when octomap=True)."""
    rng = np.random.default_rng(7)
    sm = SubmapMapping(DenseTSDF, keyframe_step=4, sub_opts=dict(map_scale=[6.4, 6.4], max_submap_num=16),
                       global_opts=dict(map_scale=[12.8, 12.8]))

    def on_submap(buf):
        d = np.load(io.BytesIO(zlib.decompress(bytes(buf))), allow_pickle=True).item()
        rec.trace.append(["wire", "send_submap", int(d["frame_id"]), hexes(d["pose"][0]), hexes(d["pose"][1]), int(d["marker"])])

    def on_traj(buf):
        d = np.load(io.BytesIO(zlib.decompress(bytes(buf))), allow_pickle=True).item()
        rec.trace.append(["wire", "send_traj", sorted(int(k) for k in d)])

    sm.map_send_handle, sm.traj_send_handle = on_submap, on_traj
    sm.set_dep_camera_intrinsic(np.array([525.0, 0, 319.5, 0, 525.0, 239.5, 0, 0, 1]))
    sm.set_color_camera_intrinsic(np.array([600.0, 0, 320.0, 0, 600.0, 240.0, 0, 0, 1]))
    ext = (rot(0.01, -0.02, 0.03), np.array([0.05, 0.0, -0.02]))
    depth, tex = np.zeros((48, 64), np.uint16), np.zeros((48, 64, 3), np.uint8)
    poses = {}
    for fid in range(14):
        R, T = rot(0.02 * fid, 0.01 * fid, 0.1 * fid), np.array([0.2 * fid, 0.05 * fid, 0.01 * fid]) + rng.normal(size=3) * 0.01
        poses[fid] = (R, T)
        if fid == 9:
            pcl = rng.normal(size=(50, 3))
            sm.recast_pcl_to_map_by_frame(fid, True, (R, T), ext, pcl, np.zeros((50, 3), np.uint8))
        else:
            sm.recast_depth_to_map_by_frame(fid, fid % 2 == 0, (R, T), ext, depth, tex)
        if fid == 6:   # a pose-graph update that moves the submaps created so far and re-anchors the odometry
            upd = {k: (rot(0.01, 0.0, 0.02) @ poses[k][0], poses[k][1] + np.array([0.03, -0.01, 0.0])) for k in (0, 4, 6)}
            sm.set_frame_poses(upd)
        if fid == 11:  # a trajectory received from another robot
            traj = {4: (rot(0.0, 0.01, 0.0) @ poses[4][0], poses[4][1] + 0.02), 8: (poses[8][0], poses[8][1] - 0.01)}
            f = io.BytesIO()
            np.save(f, traj)
            sm.input_remote_traj(zlib.compress(f.getbuffer(), level=1))
    sub = {"indices": np.zeros((1, 3), np.int16), "frame_id": 1234, "pose": (rot(0.3, 0.2, 0.1), np.array([1.0, 2.0, 3.0]))}
    f = io.BytesIO()
    np.save(f, sub)
    sm.input_remote_submap(zlib.compress(f.getbuffer(), level=1))
    if octomap:
        sm.cvt_occupy_to_voxels(0)
        sm.set_exporting_local()
        sm.cvt_occupy_to_voxels(2)
    else:
        sm.cvt_TSDF_surface_to_voxels()
        sm.set_exporting_local()
        sm.cvt_TSDF_surface_to_voxels()
        sm.cvt_TSDF_to_voxels_slice(0.5)
    rec.trace.append(["state", "submaps", sorted((int(k), int(v)) for k, v in sm.submaps.items()), int(sm.frame_count), int(sm.last_frame_id)])
