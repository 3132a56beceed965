"""taichi_slam.mapping.topo_graph -> taichislam_b200.mapping.topo_graph (the reference imports its map classes by submodule,
submap_mapping.py:1-3, topo_graph.py:1-7)."""
from taichislam_b200.mapping.topo_graph import *  # noqa: F401,F403
from taichislam_b200.mapping import topo_graph as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
