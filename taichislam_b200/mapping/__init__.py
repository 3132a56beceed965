"""Drop-in namespace of `taichi_slam.mapping` (reference mapping/__init__.py:1-6): the map classes plus the
names callers pick up through its star-imports (`ti`, `np`, `math`, `time`; scripts/taichislam_node.py:6,34-36)."""
import math  # noqa: F401
import time  # noqa: F401

import numpy as np  # noqa: F401

from .ti_shim import ti  # noqa: F401
from .mapping_common import BaseMap, sign  # noqa: F401
from .dense_tsdf import DenseTSDF, Wmax  # noqa: F401
from .taichi_octomap import Octomap  # noqa: F401
from .marching_cube_mesher import MarchingCubeMesher  # noqa: F401
from .submap_mapping import SubmapMapping  # noqa: F401
from .topo_graph import TopoGraphGen  # noqa: F401

__all__ = ["DenseTSDF", "Octomap", "MarchingCubeMesher", "BaseMap", "SubmapMapping", "TopoGraphGen", "ti", "np", "math", "time", "sign", "Wmax"]
