"""TopoGraphGen on the GPU map backend (batched raycast / occupancy kernels of libtslam.so through DenseTSDF) against
the literal CPU restatement of the reference (oracle/topo_oracle.py) on the oracle's copy of the same map."""
import numpy as np
import pytest

from topo_world import two_rooms, START_A

pytestmark = pytest.mark.gpu


def test_topo_graph_on_dense_tsdf_matches_oracle():
    from oracle.oracle import OracleTSDF
    from oracle.topo_oracle import TopoOracle
    from taichi_slam.mapping import DenseTSDF, TopoGraphGen   # the reference's import path (tests/gen_topo_graph.py:4)
    from test_topo_cpu import ScalarOracleMap, graphs_equal
    idx, t, w, occ = two_rooms()
    m = DenseTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, is_global_map=True)
    m.load_numpy(0, idx, t, w, occ, np.array([]))
    o = OracleTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, is_global_map=True)
    o.scatter(0, idx, t, w, occ)
    # the batched map queries of the class surface == the oracle's
    rng = np.random.default_rng(0)
    pts = (rng.uniform(-0.5, 7.0, (2000, 3)) * np.array([1.0, 0.5, 0.35])).astype(np.float32)
    dirs = rng.normal(size=(2000, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    occ_g, un_g = m.is_pos_occupy(pts), m.is_pos_unobserved(pts)
    occ_o, un_o = o.query_points(pts)
    assert np.array_equal(occ_g, occ_o) and np.array_equal(un_g, un_o)
    sg, xg, lg = m.raycast(pts, dirs, 2.0)
    so, xo, lo = o.raycast(pts, dirs, 2.0)
    assert np.array_equal(sg, so) and np.allclose(lg, lo, atol=1e-6) and np.allclose(xg, xo, atol=1e-5)
    topo = TopoGraphGen(m, coll_det_num=64, max_raycast_dist=2.5)
    ref = TopoOracle(ScalarOracleMap(o), coll_det_num=64, max_raycast_dist=2.5)
    n = topo.generate_topo_graph(START_A, max_nodes=12)
    assert n == ref.generate_topo_graph(START_A, max_nodes=12) >= 2
    graphs_equal(topo, ref)
    # default parameters (128 rays, 2 m) as tests/gen_topo_graph.py uses them, and the benchmark helper runs
    topo2 = TopoGraphGen(m)
    assert topo2.generate_topo_graph(START_A, max_nodes=20) >= 2
    topo2.node_expansion_benchmark(START_A, run_num=2)
