"""TopoGraphGen host logic on CPU: the product class (taichislam_b200.mapping.TopoGraphGen) driven by an oracle-backed
map (batched queries answered by the CPU oracle) against the literal restatement of the reference
(oracle/topo_oracle.py) on the same map.  The GPU-backed run of the same scene is tests/test_gpu_topo.py."""
import numpy as np
import pytest

from oracle.oracle import OracleTSDF
from oracle.topo_oracle import TopoOracle
from taichislam_b200.mapping.mapping_common import BaseMap
from taichislam_b200.mapping.topo_graph import TopoGraphGen
from topo_world import two_rooms, START_A


class OracleBackedMap(BaseMap):
    """BaseMap whose planner queries are answered by the CPU oracle (test wiring; the product never imports oracle/)."""

    def __init__(self, o):
        super().__init__(o.voxel_scale)
        self.o = o

    def _query_raycast(self, pos, dir, max_dist):
        return self.o.raycast(pos, dir, max_dist)

    def _query_points(self, xyz):
        return self.o.query_points(xyz)

    def _query_near(self, xyz, voxel):
        return self.o.query_near_occupy(xyz, voxel)


class ScalarOracleMap:
    """scalar map queries for the literal oracle"""

    def __init__(self, o):
        self.o, self.voxel_scale = o, o.voxel_scale

    def raycast(self, pos, dir, max_dist):
        s, x, ln = self.o.raycast(np.asarray(pos, np.float32)[None], np.asarray(dir, np.float32)[None], max_dist)
        return bool(s[0]), x[0], np.float32(ln[0])

    def is_pos_occupy(self, xyz):
        return bool(self.o.query_points(np.asarray(xyz, np.float32)[None])[0][0])

    def is_pos_unobserved(self, xyz):
        return bool(self.o.query_points(np.asarray(xyz, np.float32)[None])[1][0])

    def is_near_pos_occupy(self, xyz, voxel):
        return bool(self.o.query_near_occupy(np.asarray(xyz, np.float32)[None], voxel)[0])


def make_world():
    o = OracleTSDF(map_scale=[12.8, 12.8], voxel_scale=0.05, is_global_map=True)
    o.scatter(0, *two_rooms())
    return o


def graphs_equal(t, ref, tol=1e-5):
    assert t.num_nodes[None] == len(ref.nodes)
    assert t.num_facelets[None] == len(ref.facelets)
    assert t.num_frontiers[None] == ref.num_frontiers
    for a, b in zip(t.nodes, ref.nodes):
        assert (a["start"], a["end"], a["master_idx"]) == (b["start"], b["end"], b["master_idx"])
        assert np.allclose(a["center"], b["center"], atol=tol)
    nf = t.num_facelets[None]
    assert np.allclose(t.f_normal.a[:nf], np.array([f.normal for f in ref.facelets]), atol=1e-4)
    assert np.array_equal(t.f_is_frontier.a[:nf].astype(bool), np.array([f.is_frontier for f in ref.facelets]))
    e = t.edges.to_numpy()[:t.edge_num[None]].reshape(-1, 6)
    er = np.array([np.concatenate(p) for p in ref.edges]).reshape(-1, 6)
    assert e.shape == er.shape
    if len(e):
        assert np.allclose(e[np.lexsort(e.T[::-1])], er[np.lexsort(er.T[::-1])], atol=tol)
    for k in range(ref.num_frontiers):
        assert t.frontiers[k]["is_valid"] == ref.frontiers[k]["is_valid"]
        assert np.allclose(t.frontiers[k]["projected_center"], ref.frontiers[k]["projected_center"], atol=tol)


def test_topo_graph_two_rooms_matches_literal_oracle():
    o = make_world()
    t = TopoGraphGen(OracleBackedMap(o), coll_det_num=64, max_raycast_dist=2.5)
    ref = TopoOracle(ScalarOracleMap(o), coll_det_num=64, max_raycast_dist=2.5)
    n = t.generate_topo_graph(START_A, max_nodes=12)
    nr = ref.generate_topo_graph(START_A, max_nodes=12)
    assert n == nr >= 2, (n, nr)
    graphs_equal(t, ref)
    # the graph reaches room B through the doorway: some node centre lies beyond the wall between the rooms (x > 0.45 m)
    assert max(nd["center"][0] for nd in t.nodes) > 0.45
    # every polyhedron is a closed triangulated surface: each facelet has three neighbours inside its own node
    nf = t.num_facelets[None]
    assert nf == sum(nd["end"] - nd["start"] for nd in t.nodes) and nf % 2 == 0
    tri = t.tri_vertices.to_numpy()[:3 * nf]
    assert np.isfinite(tri).all()
    assert t.edge_num[None] >= 2 * (n - 1)


def test_topo_start_in_unknown_space_gives_no_node():
    o = make_world()
    t = TopoGraphGen(OracleBackedMap(o), coll_det_num=32, max_raycast_dist=1.5)
    # far outside the observed region every cell reads TSDF 0 = "occupied": all rays hit at length 0, node too small
    assert t.generate_topo_graph(np.array([-5.0, 5.0, 5.0]), max_nodes=4) == 0
    assert t.num_facelets[None] == 0 and t.edge_num[None] == 0


def test_base_map_query_surface():
    o = make_world()
    m = OracleBackedMap(o)
    s, x, ln = m.raycast(START_A, np.array([0.0, 1.0, 0.0]), 3.0)   # side wall of room A at y = 1.7 m
    assert isinstance(s, bool) and s and 1.3 < ln < 2.0
    sb, xb, lb = m.raycast(np.stack([START_A] * 3), np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]), 3.0)
    assert sb.shape == (3,) and list(sb) == [False, True, True] and abs(lb[1] - ln) < 1e-6  # +x leaves through the doorway
    assert m.is_pos_occupy(START_A) is False and m.is_pos_unobserved(START_A) is False
    assert m.is_near_pos_occupy(START_A, 0) is False  # range(-0, 0) is empty (mapping_common.py:198)
    with pytest.raises(NotImplementedError):
        BaseMap(0.05).raycast(START_A, START_A, 1.0)


# ---------------------------------------------------------------------------------------------------------------
# Against the REFERENCE'S TopoGraphGen, executed (tests/golden/ref_exec_topo.npz: topo_graph.py run unmodified through
# oracle/taichi_emu.py on this same world by `tools/make_golden_ref.py topo`): both the literal restatement and the
# product class must reproduce its graph.
# ---------------------------------------------------------------------------------------------------------------
import os

TOPO_GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_exec_topo.npz"))


def matches_executed_reference(tag, n_nodes, n_facelets, n_frontiers, n_edge_pts, node_center, node_range, normals, is_frontier,
                               tri, fr_valid, fr_proj, fr_next, edges):
    g = TOPO_GOLD
    assert [n_nodes, n_facelets, n_frontiers, n_edge_pts] == list(g[tag + "_counts"])
    assert np.allclose(node_center, g[tag + "_node_center"], atol=1e-5)
    assert np.array_equal(np.asarray(node_range), g[tag + "_node_range"])
    assert np.allclose(normals, g[tag + "_facelet_normal"], atol=1e-4)
    assert np.array_equal(np.asarray(is_frontier).astype(np.int8), g[tag + "_facelet_frontier"])
    assert np.allclose(tri, g[tag + "_tri_vertices"], atol=1e-5)
    assert np.array_equal(np.asarray(fr_valid).astype(np.int8), g[tag + "_frontier_valid"])
    assert np.allclose(fr_proj, g[tag + "_frontier_proj_center"], atol=1e-5)
    assert np.allclose(fr_next, g[tag + "_frontier_next"], atol=1e-5)
    e, er = np.asarray(edges).reshape(-1, 6), g[tag + "_edges"].reshape(-1, 6)
    assert np.allclose(e[np.lexsort(e.T[::-1])], er[np.lexsort(er.T[::-1])], atol=1e-5)


@pytest.mark.parametrize("tag,kw", [("a", dict(coll_det_num=64, max_raycast_dist=2.5)), ("b", dict(coll_det_num=128, max_raycast_dist=2))])
def test_topo_graph_matches_executed_reference(tag, kw):
    o = make_world()
    # the literal restatement
    ref = TopoOracle(ScalarOracleMap(o), **kw)
    n = ref.generate_topo_graph(START_A, max_nodes=12)
    matches_executed_reference(
        tag, n, len(ref.facelets), ref.num_frontiers, 2 * len(ref.edges), [nd["center"] for nd in ref.nodes],
        [[nd["start"], nd["end"], nd["master_idx"]] for nd in ref.nodes], [f.normal for f in ref.facelets],
        [f.is_frontier for f in ref.facelets], np.concatenate([[f.v0, f.v1, f.v2] for f in ref.facelets]),
        [ref.frontiers[k]["is_valid"] for k in range(ref.num_frontiers)], [ref.frontiers[k]["projected_center"] for k in range(ref.num_frontiers)],
        [ref.frontiers[k]["next_node_initial"] for k in range(ref.num_frontiers)], [np.concatenate(p) for p in ref.edges])
    # the product class on the oracle-backed map
    t = TopoGraphGen(OracleBackedMap(o), **kw)
    n = t.generate_topo_graph(START_A, max_nodes=12)
    nf, nfr = t.num_facelets[None], t.num_frontiers[None]
    matches_executed_reference(
        tag, n, nf, nfr, t.edge_num[None], [nd["center"] for nd in t.nodes], [[nd["start"], nd["end"], nd["master_idx"]] for nd in t.nodes],
        t.f_normal.a[:nf], t.f_is_frontier.a[:nf], t.tri_vertices.to_numpy()[:3 * nf], [t.frontiers[k]["is_valid"] for k in range(nfr)],
        [t.frontiers[k]["projected_center"] for k in range(nfr)], [t.frontiers[k]["next_node_initial"] for k in range(nfr)],
        t.edges.to_numpy()[:t.edge_num[None]])
