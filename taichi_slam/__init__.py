"""Namesake of the reference's top-level package: put this directory's parent in front of a TaichiSLAM checkout on
sys.path and `from taichi_slam.mapping import *` (scripts/taichislam_node.py:6, TaichiSLAM_demo.py) - as well as the
submodule imports `taichi_slam.mapping.dense_tsdf` etc. (submap_mapping.py:1-3) - resolve to the B200 backend, while
everything this backend does not replace (`taichi_slam.utils.*`: ROS / rendering / LCM glue, scripts/taichislam_node.py:7-9)
still comes from the checkout: `__path__` is extended over every other `taichi_slam` directory on sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
