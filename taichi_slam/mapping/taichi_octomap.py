"""taichi_slam.mapping.taichi_octomap -> taichislam_b200.mapping.taichi_octomap (the reference imports its map classes by submodule,
submap_mapping.py:1-3, topo_graph.py:1-7)."""
from taichislam_b200.mapping.taichi_octomap import *  # noqa: F401,F403
from taichislam_b200.mapping import taichi_octomap as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
