"""CPU oracle of TopoGraphGen - TEST INFRASTRUCTURE ONLY (imported by tests/ only).  The reference
(taichi_slam/mapping/topo_graph.py) needs Taichi, which cannot run here, and ships no test with assertions for this
class (tests/gen_topo_graph.py is a visual smoke script).  Pinned instead by the reference's own source EXECUTED through
oracle/taichi_emu.py: `tools/make_golden_ref.py topo` records the graph topo_graph.py builds on the two-room world,
tests/test_topo_cpu.py::test_topo_graph_matches_executed_reference checks this restatement (and the product class)
against it - nodes, facelets, frontier flags, projections, edges.

A literal, loop-by-loop restatement of the reference in scalar float32 Python: every method cites the lines it
follows.  `mapping` is any object with the scalar map queries of BaseMap (mapping_common.py:165-204):
    raycast(pos, dir, max_dist) -> (succ, pos, len);  is_pos_occupy(xyz);  is_pos_unobserved(xyz);
    is_near_pos_occupy(xyz, voxel);  voxel_scale
(tests wrap oracle.OracleTSDF in such an adapter).  Slow by design (pure Python loops): small cases only.
"""
import numpy as np
from scipy.spatial import ConvexHull

f32 = np.float32


def v3(x):
    return np.asarray(x, dtype=f32).reshape(3).copy()


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], f32)


def dot(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def norm(a):
    return f32(np.sqrt(dot(a, a), dtype=f32))


def normalized(a):
    return (a / norm(a)).astype(f32)


class Facelet:  # topo_graph.py:22-75
    def __init__(self, poly_idx, facelet_idx, v0, v1, v2, naive_norm):  # init :36-50
        self.poly_idx = poly_idx
        self.edge1 = (v1 - v0).astype(f32)
        self.edge2 = (v2 - v0).astype(f32)
        self.center = (((v0 + v1).astype(f32) + v2).astype(f32) / f32(3)).astype(f32)
        self.normal = normalized(cross(self.edge1, self.edge2))
        if dot(self.normal, naive_norm) < 0:
            self.normal = (-self.normal).astype(f32)
        self.v0, self.v1, self.v2 = v0, v1, v2
        self.facelet_idx = facelet_idx
        self.assigned = False
        self.is_frontier = False

    def rayTriangleIntersect(self, P, w):  # :52-71
        q = cross(w, self.edge2)
        a = dot(self.edge1, q)
        succ = False
        t = f32(0.0)
        if abs(a) > f32(0.00001):
            s = ((P - self.v0) / a).astype(f32)
            r = cross(s, self.edge1)
            b0 = dot(s, q)
            b1 = dot(r, w)
            b2 = f32(f32(f32(1.0) - b0) - b1)
            t = dot(self.edge2, r)
            succ = True
            if b0 < 0.0 or b1 < 0.0 or b2 < 0.0:
                succ = False
        return succ, t


class TopoOracle:
    def __init__(self, mapping, coll_det_num=128, max_raycast_dist=2, thres_size=0.5, frontier_creation_threshold=0.5,
                 frontier_verify_threshold=0.5, frontier_backward_check=-0.2, frontier_combine_angle_threshold=40):  # :131-150
        self.mapping = mapping
        self.coll_det_num = coll_det_num
        self.max_raycast_dist = max_raycast_dist
        self.thres_size = thres_size
        self.frontier_creation_threshold = frontier_creation_threshold
        self.frontier_verify_threshold = frontier_verify_threshold
        self.frontier_normal_dot_threshold = np.cos(np.deg2rad(frontier_combine_angle_threshold))
        self.check_frontier_small_distance = 0.1
        self.frontier_backward_check = frontier_backward_check
        # generate_uniform_sample_points :211-224
        phi = np.pi * (3 - np.sqrt(5))
        ret = []
        for i in range(coll_det_num):
            y = 1 - 2 * (i / (coll_det_num - 1))
            radius = np.sqrt(1 - y * y)
            theta = phi * i
            ret.append([np.cos(theta) * radius, y, np.sin(theta) * radius])
        self.sample_dirs = np.array(ret, dtype=f32)
        self.facelets, self.nodes, self.frontiers, self.edges = [], [], {}, []
        self.num_frontiers = 0
        self.search_frontiers_idx = 0
        self.connected = set()
        self.start_point = np.zeros(3, f32)

    # :472-488
    def detect_collision_facelets(self, pos, dir, max_dist, backward_dist=-0.01, skip_idx=-1):
        succ = False
        best_t = f32(max_dist)
        best_poly_ind = -1
        for k in range(len(self.nodes)):
            if k != skip_idx:
                poly = self.nodes[k]
                if norm((pos - poly["center"]).astype(f32)) < f32(max_dist) + f32(self.max_raycast_dist):
                    for i in range(poly["start"], poly["end"]):
                        _succ, t = self.facelets[i].rayTriangleIntersect(pos, dir)
                        if _succ and f32(backward_dist) < t < best_t:
                            best_t = t
                            best_poly_ind = self.facelets[i].poly_idx
                            succ = True
        pos_poly = (pos + dir * best_t).astype(f32)
        return succ, pos_poly, best_t, best_poly_ind

    # :490-507
    def raycast(self, pos, dir, max_dist, skip_idx=-1):
        recast_type = 1
        succ_poly, pos_coll, len_coll, poly_ind = self.detect_collision_facelets(pos, dir, max_dist, -0.01, skip_idx)
        max_dist_recast = max_dist
        if succ_poly:
            max_dist_recast = len_coll
        succ_map, pos_col_map, len_map = self.mapping.raycast(pos, dir, float(max_dist_recast))
        if (not succ_poly) or (succ_map and len_map < len_coll):
            pos_coll = v3(pos_col_map)
            len_coll = f32(len_map)
            recast_type = 0
            succ_poly = bool(succ_map)
        return succ_poly, recast_type, pos_coll, len_coll, poly_ind

    # :444-470
    def detect_collisions(self):
        pos = self.start_point
        ray_len_black = f32(0.0)
        self.black_unit, self.black_len = [], []
        white = 0
        for i in range(self.coll_det_num):
            succ, t, col_pos, _len, node_idx = self.raycast(pos, self.sample_dirs[i], self.max_raycast_dist)
            if succ:
                self.black_unit.append(self.sample_dirs[i])
                self.black_len.append(f32(_len))
                ray_len_black = f32(ray_len_black + f32(_len))
            else:
                white += 1
        nb = len(self.black_len)
        succ = True
        if nb == 0 or (white == 0 and ray_len_black / f32(nb) < self.thres_size):
            succ = False
        return succ

    # :324-342
    def detect_facelet_frontier(self, facelet, neighbor_node_ids):
        is_frontier = True
        m = self.mapping
        if m.is_near_pos_occupy(facelet.center, 0) or m.is_pos_unobserved(facelet.center):
            is_frontier = False
        else:
            start_raycast_pos = (facelet.center + facelet.normal * f32(m.voxel_scale)).astype(f32)
            if m.is_pos_occupy(start_raycast_pos) or m.is_pos_unobserved(facelet.center):
                is_frontier = False
            else:
                succ, t, col_pos, _len, node_idx = self.raycast(start_raycast_pos, facelet.normal, self.frontier_creation_threshold)
                if succ and t == 1:
                    neighbor_node_ids.append(node_idx)
                if succ:
                    is_frontier = False
        return is_frontier

    # :344-378
    def construct_frontier(self, node_idx, idx_start_facelet, queue):
        frontier_idx = self.num_frontiers
        self.num_frontiers += 1
        center = np.zeros(3, f32)
        normal = np.zeros(3, f32)
        for q in queue:
            center = (center + self.facelets[q + idx_start_facelet].center).astype(f32)
            normal = (normal + self.facelets[q + idx_start_facelet].normal).astype(f32)
        center = (center / f32(len(queue))).astype(f32)
        normal = normalized((normal / f32(len(queue))).astype(f32))
        fr = dict(master_idx=node_idx, frontier_idx=frontier_idx, avg_center=center, outwards_unit_normal=normal, is_valid=False,
                  projected_center=np.zeros(3, f32), projected_normal=np.zeros(3, f32), next_node_initial=np.zeros(3, f32))
        self.frontiers[frontier_idx] = fr
        succ, t, projected_normal = False, f32(0.0), np.zeros(3, f32)
        for q in queue:
            fl = self.facelets[q + idx_start_facelet]
            succ, t = fl.rayTriangleIntersect(center, normal)
            projected_normal = fl.normal
            if succ:
                break
        if succ:
            fr["projected_center"] = (center + t * normal).astype(f32)
            fr["projected_normal"] = projected_normal
        else:
            self.num_frontiers -= 1

    # :380-442
    def add_mesh(self, mesh, neighbors, last_node_idx):
        num_facelets = mesh.shape[0]
        facelet_start_idx = len(self.facelets)
        node = len(self.nodes)
        center_pos = np.zeros(3, f32)
        center_count = f32(0.0)
        neighbor_node_ids = []
        for i in range(num_facelets):
            v0, v1, v2 = v3(mesh[i, 0]), v3(mesh[i, 1]), v3(mesh[i, 2])
            vsum = ((v0 + v1).astype(f32) + v2).astype(f32)
            center_pos = (center_pos + vsum).astype(f32)
            center_count = f32(center_count + f32(3.0))
            naive_norm = normalized((vsum - f32(3.0) * self.start_point).astype(f32))
            fl = Facelet(node, i + facelet_start_idx, v0, v1, v2, naive_norm)
            self.facelets.append(fl)
        for i in range(num_facelets):  # detect_facelet_frontier sees only the EXISTING nodes (num_nodes not yet incremented)
            fl = self.facelets[i + facelet_start_idx]
            fl.is_frontier = self.detect_facelet_frontier(fl, neighbor_node_ids)
        new_node_center = (center_pos / center_count).astype(f32)
        self.nodes.append(dict(idx=node, master_idx=last_node_idx, start=facelet_start_idx, end=facelet_start_idx + num_facelets,
                               center=new_node_center))
        # NOTE: the node is appended AFTER the frontier detection above, like num_nodes[None] += 1 at the very end (:442):
        # detect_collision_facelets inside it iterates range(num_nodes) = the nodes that existed before.
        if last_node_idx >= 0:
            self.edges.append((self.nodes[last_node_idx]["center"], new_node_center))
            self.connected.add((node, last_node_idx))
            self.connected.add((last_node_idx, node))
        for neigh_idx in neighbor_node_ids:
            if (node, neigh_idx) not in self.connected:
                self.connected.add((node, neigh_idx))
                self.connected.add((neigh_idx, node))
                self.edges.append((self.nodes[neigh_idx]["center"], new_node_center))
        thr = f32(self.frontier_normal_dot_threshold)
        for i in range(facelet_start_idx, facelet_start_idx + num_facelets):
            idx = i - facelet_start_idx
            if not self.facelets[i].assigned and self.facelets[i].is_frontier:
                queue = [idx]
                search_idx = 0
                normal = self.facelets[i].normal
                while search_idx < len(queue):
                    _idx = queue[search_idx]
                    search_idx += 1
                    self.facelets[_idx + facelet_start_idx].assigned = True
                    for j in range(3):
                        idx_neighbor = int(neighbors[_idx, j]) + facelet_start_idx
                        fn = self.facelets[idx_neighbor]
                        if fn.is_frontier and not fn.assigned and dot(normal, fn.normal) > thr:
                            queue.append(int(neighbors[_idx, j]))
                self.construct_frontier(node, facelet_start_idx, queue)

    # :296-315
    def generate_poly_on_blacks(self, start_pt, last_node_idx=-1):
        black_dirs = np.array(self.black_unit, f32)
        hull = ConvexHull(black_dirs)
        lens = np.array(self.black_len, f32)
        vertices = hull.points * lens[:, None]
        vertices = np.apply_along_axis(lambda x: x + np.asarray(start_pt).reshape(3), 1, vertices)
        self.add_mesh(vertices[hull.simplices], hull.neighbors, last_node_idx)

    # :245-253
    def node_expansion(self, start_pt, last_node_idx=-1):
        self.start_point = v3(start_pt)
        if self.detect_collisions():
            self.generate_poly_on_blacks(start_pt, last_node_idx)

    # :255-282
    def verify_frontier(self, frontier_idx):
        fr = self.frontiers[frontier_idx]
        normal = fr["projected_normal"]
        proj_center = (fr["projected_center"] + normal * f32(self.check_frontier_small_distance)).astype(f32)
        succ, t, col_pos, _len, node_idx = self.raycast(proj_center, normal, self.max_raycast_dist * 2)
        if succ and _len < self.frontier_verify_threshold:
            fr["is_valid"] = False
        else:
            proj_center = (fr["projected_center"] - normal * f32(self.check_frontier_small_distance)).astype(f32)
            succ2, col_pos2, _len2, node_idx2 = self.detect_collision_facelets(proj_center, normal, self.frontier_verify_threshold,
                                                                               self.frontier_backward_check, fr["master_idx"])
            if succ2 and _len2 < self.frontier_verify_threshold:
                fr["is_valid"] = False
            else:
                if not succ or succ2 and _len2 < _len:
                    _len = _len2
                fr["is_valid"] = True
                fr["next_node_initial"] = (fr["projected_center"] + fr["projected_normal"] * f32(_len) / f32(2)).astype(f32)
        return fr["is_valid"]

    # :284-294
    def generate_topo_graph(self, start_pt, max_nodes=100):
        self.node_expansion(start_pt)
        while self.search_frontiers_idx < self.num_frontiers and self.search_frontiers_idx < max_nodes:
            if self.verify_frontier(self.search_frontiers_idx):
                fr = self.frontiers[self.search_frontiers_idx]
                self.node_expansion(fr["next_node_initial"], last_node_idx=fr["master_idx"])
            self.search_frontiers_idx += 1
        return len(self.nodes)
