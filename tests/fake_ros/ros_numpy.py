import types

import numpy as np

point_cloud2 = types.SimpleNamespace(pointcloud2_to_xyz_array=lambda msg: np.asarray(msg.xyz, dtype=np.float64),
                                     pointcloud2_to_array=lambda msg: msg.array)
