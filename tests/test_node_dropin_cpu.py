"""Node-level drop-in (north_star: "scripts/taichislam_node.py drops in unchanged").

The reference's ROS node is loaded BY PATH, unmodified, in the build container (it does not exist on the GPU box):
  * `taichi_slam.mapping` and its submodules resolve to this repo's alias package (-> taichislam_b200.mapping), set up the
    way the launcher `python -m taichislam_b200.run_node <node.py>` does it (the node prepends its own checkout to sys.path), set up the
    way the launcher `python -m taichislam_b200.run_node <node.py>` does it (the node prepends its own checkout to sys.path),
  * `taichi_slam.utils.*` (ROS / rendering / LCM glue the backend does not replace) falls through to the reference
    checkout via pkgutil.extend_path,
  * rospy / message_filters / ros_numpy / *_msgs come from tests/fake_ros, taichi (imported by the reference's
    visualization.py only) / lcm / transformations / matplotlib are inert mocks.
The node is then driven like its ROS callbacks would (frames + VIO poses -> process_taichi()) with the map classes'
kernel-launching methods replaced by the recorders of tests/submap_scenario.py - no GPU needed: what is checked is the
whole host path between a ROS message and the map calls (parameter plumbing, SubmapMapping policy, pose conversion,
export + PointCloud2 publication)."""
import contextlib
import importlib.util
import io
import os
import sys
from unittest import mock

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
NODE = os.path.join(REF, "scripts", "taichislam_node.py")

pytestmark = pytest.mark.skipif(not os.path.exists(NODE), reason="the reference checkout lives in the build container only")


@contextlib.contextmanager
def node_environment():
    try:  # (binary packages do not survive being dropped from sys.modules and imported again: load cv2 before the snapshot)
        import cv2  # noqa: F401
    except Exception:
        sys.modules["cv2"] = mock.MagicMock(name="cv2")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for k in [k for k in sys.modules if k == "taichi_slam" or k.startswith("taichi_slam.")]:
        del sys.modules[k]
    sys.path[:0] = [os.path.join(HERE, "fake_ros")]
    for m in ("taichi", "taichi.math", "lcm", "transformations", "matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if m not in sys.modules:
            sys.modules[m] = mock.MagicMock(name=m)
    try:
        # what `python -m taichislam_b200.run_node <checkout>/scripts/taichislam_node.py` does before running the script:
        # the node puts its own checkout FIRST on sys.path (taichislam_node.py:4), so the alias package is imported up front
        from taichislam_b200.run_node import prepare
        prepare(REF)
        yield
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]


def load_node():
    spec = importlib.util.spec_from_file_location("taichislam_node_under_test", NODE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_alias_package_resolves_like_the_reference_layout():
    with node_environment():
        import taichi_slam.mapping as tm
        import taichi_slam.mapping.dense_tsdf as td
        import taichi_slam.mapping.mapping_common as tc
        import taichi_slam.mapping.taichi_octomap as to
        import taichi_slam.mapping.marching_cube_mesher as tmc
        import taichi_slam.mapping.topo_graph as tt
        import taichi_slam.mapping.submap_mapping as ts
        import taichislam_b200.mapping as ours
        assert td.DenseTSDF is ours.DenseTSDF and to.Octomap is ours.Octomap and tc.BaseMap is ours.BaseMap
        assert tmc.MarchingCubeMesher is ours.MarchingCubeMesher and tt.TopoGraphGen is ours.TopoGraphGen and ts.SubmapMapping is ours.SubmapMapping
        assert tm.DenseTSDF is ours.DenseTSDF and hasattr(tm, "ti")
        # what the backend does not replace still comes from the checkout
        import taichi_slam.utils.ros_pcl_transfer as rp
        assert rp.__file__.startswith(REF) and callable(rp.pose_msg_to_numpy)


def _pose_msg(R_quat_xyzw, T):
    from geometry_msgs.msg import Pose, Point, Quaternion
    return Pose(Point(x=T[0], y=T[1], z=T[2]), Quaternion(x=R_quat_xyzw[0], y=R_quat_xyzw[1], z=R_quat_xyzw[2], w=R_quat_xyzw[3]))


@pytest.mark.parametrize("mapping_type", ["tsdf", "octo"])
def test_reference_node_runs_unmodified_on_the_backend(mapping_type):
    import torch
    import submap_scenario as sc
    with node_environment():
        import rospy
        rospy.PARAMS.clear()
        rospy.PUBLISHED.clear()
        rospy.PARAMS.update({"~enable_rendering": False, "~enable_multi": True, "~enable_submap": True, "~texture_enabled": False,
                             "~output_map": True, "~keyframe_step": 3, "~mapping_type": mapping_type, "~map_size_xy": 12.8,
                             "~map_size_z": 6.4, "~enable_mesher": True, "~disp/max_mesh": 1000})
        node_mod = load_node()
        import taichislam_b200.mapping.dense_tsdf as dt
        import taichislam_b200.mapping.taichi_octomap as oc
        import taichislam_b200.mapping.marching_cube_mesher as mcm
        from taichislam_b200.mapping.field import Field
        rec = sc.Recorder()
        created = []

        def light_init(self, map_scale=[10, 10], voxel_scale=0.05, texture_enabled=False, max_disp_particles=1024, is_global_map=False, **kw):
            self.is_global_map, self.enable_texture, self.max_disp_particles = is_global_map, texture_enabled, max_disp_particles
            self.voxel_scale_ctor, self.kwargs = voxel_scale, dict(kw, map_scale=list(map_scale))
            self.export_color = Field(torch.zeros((8, 3)))
            self.export_TSDF_xyz = self.export_x = Field(torch.arange(24, dtype=torch.float32).reshape(8, 3))
            self.num_TSDF_particles = self.num_export_particles = Field(torch.tensor([5], dtype=torch.int32))
            created.append(self)

        def mesher_init(self, mapping, max_triangles, tsdf_surface_thres=0.0):
            self.mapping, self.max_triangles, self.thres = mapping, max_triangles, tsdf_surface_thres

        with contextlib.ExitStack() as st:
            cls = dt.DenseTSDF if mapping_type == "tsdf" else oc.Octomap
            st.enter_context(mock.patch.object(cls, "__init__", light_init))
            st.enter_context(mock.patch.object(mcm.MarchingCubeMesher, "__init__", mesher_init))
            for pch in rec.patches(cls):
                st.enter_context(pch)
            with contextlib.redirect_stdout(io.StringIO()):
                node = node_mod.TaichiSLAMNode()
                from sensor_msgs.msg import Image
                from swarm_msgs.msg import VIOFrame, DroneTraj
                depth = (np.arange(48 * 64, dtype=np.uint16).reshape(48, 64) % 4000 + 500)
                ext = _pose_msg([0.0, 0.0, 0.0, 1.0], [0.05, 0.0, -0.02])
                for fid in range(8):
                    a = 0.1 * fid
                    pose = _pose_msg([0.0, 0.0, np.sin(a / 2), np.cos(a / 2)], [0.2 * fid, 0.1, 0.0])
                    frame = VIOFrame(fid, fid % 2 == 0, pose, [ext])
                    msg = Image(height=48, width=64, data=depth.tobytes())
                    node.ts.callback(msg, frame)       # what message_filters would do with a synchronised pair
                    node.process_taichi()
                node.traj_callback(DroneTraj(1, [0, 2], [_pose_msg([0, 0, 0, 1.0], [0.01, 0, 0]), _pose_msg([0, 0, 0, 1.0], [0.4, 0.1, 0])]))
        t = rec.trace
        # (`ti` in the node's namespace is the one its LAST star import brings - visualization.py's `import taichi as ti`,
        # here a mock; with the real package installed for rendering it is real Taichi, which coexists with the backend)
        assert node_mod.DenseTSDF is dt.DenseTSDF and node_mod.SubmapMapping.__module__ == "taichislam_b200.mapping.submap_mapping"
        # parameter plumbing: map_scale / voxel_scale / ray lengths reach the constructors (taichislam_node.py:157-199)
        assert len(created) == 2 and created[0].kwargs["map_scale"] == [12.8, 6.4] and created[0].kwargs["max_ray_length"] == 5.1
        assert sum(c.is_global_map for c in created) == 1
        names = [e[1] for e in t]
        assert names.count("recast_depth_to_map") == 8 and names.count("set_dep_camera_intrinsic") >= 1
        assert "switch_to_next_submap" in names and "fuse_submaps" in names                     # keyframe policy ran (submap_mapping.py:148-160)
        first = next(e for e in t if e[1] == "recast_depth_to_map")
        assert first[0] == "sub" and first[5] == [48, 64]
        assert any(e[1] in ("cvt_TSDF_surface_to_voxels", "cvt_occupy_to_voxels") for e in t)    # output(): export every processed frame
        # ... and the exported particles went out as PointCloud2 through the reference's own point_cloud()
        assert len(rospy.PUBLISHED) >= 1 and rospy.PUBLISHED[0][0] == "/dense_mapping"
        pc = rospy.PUBLISHED[-1][1]
        assert pc.width == 5 and pc.point_step == 12 and len(pc.data) == 5 * 12
        # submaps and trajectories went on the (mocked) LCM wire through the reference's SLAMComm / Buffer encoding
        sent = [c for c in node.comm.lcm.publish.call_args_list]
        assert {c.args[0] for c in sent} >= {"SUBMAP_CHANNEL"} and all(isinstance(c.args[1], (bytes, bytearray)) for c in sent)
        if mapping_type == "tsdf":
            assert node.mesher.max_triangles == 1000 and abs(node.mesher.thres - 0.25) < 1e-12  # MarchingCubeMesher(global_map, max_mesh, 5 * voxel_scale)
            assert node.mesher.mapping is node.mapping.global_map
