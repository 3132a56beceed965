"""taichi_slam.mapping.dense_tsdf -> taichislam_b200.mapping.dense_tsdf (the reference imports its map classes by submodule,
submap_mapping.py:1-3, topo_graph.py:1-7)."""
from taichislam_b200.mapping.dense_tsdf import *  # noqa: F401,F403
from taichislam_b200.mapping import dense_tsdf as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
