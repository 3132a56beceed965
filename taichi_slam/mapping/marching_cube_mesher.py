"""taichi_slam.mapping.marching_cube_mesher -> taichislam_b200.mapping.marching_cube_mesher (the reference imports its map classes by submodule,
submap_mapping.py:1-3, topo_graph.py:1-7)."""
from taichislam_b200.mapping.marching_cube_mesher import *  # noqa: F401,F403
from taichislam_b200.mapping import marching_cube_mesher as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
