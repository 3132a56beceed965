"""TopoGraphGen - sparse topological skeleton graph on the B200 map backend.

Reference: taichi_slam/mapping/topo_graph.py (511 lines of Taichi kernels + scipy).  The algorithm grows a graph
of star-convex free-space polyhedra: from a start point it casts `coll_det_num` rays (:444-470), builds the convex
hull of the directions that hit something (:305-315, scipy.spatial.ConvexHull as in the reference), scales the hull
vertices by the hit distances into a polyhedron of facelets (:296-303, :380-442), marks the facelets that face
free space as frontiers (:324-342), clusters neighbouring frontier facelets (:415-441), and expands a new node
behind every verified frontier (:255-294).

What runs where: every map query of the reference (`mapping.raycast`, `is_pos_occupy`, `is_pos_unobserved`,
`is_near_pos_occupy`, mapping_common.py:165-204) is a BATCHED kernel of libtslam.so behind the map class
(`BaseMap.raycast` etc. in this package) - one launch for the 128 rays of a node expansion, one per query kind for
all facelets of a new polyhedron.  The facelet bookkeeping (a few hundred triangles per node, ray-triangle tests
against the nearby polyhedra, the clustering queue) is host numpy in f32, written to follow the reference's
operation order; the reference runs those parts as serialised Taichi loops as well.

Deviations: fields are host arrays that grow on demand (the reference pre-allocates `max_facelets` = 1 M entries of
every field); the parallel loop over a new polyhedron's facelets (:386-402) is executed in index order, so the
order of the skeleton edges is deterministic (the reference's depends on thread timing; the edge SET is the same).
"""
import time

import numpy as np

from .field import HostScalar

F = np.float32


def _normalized(v):
    n = np.sqrt((v * v).sum(-1, dtype=F), dtype=F)
    return (v / n[..., None]).astype(F)


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(F)


def _dot(a, b):
    return ((a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]).astype(F)


class HostField:
    """Host array with the `to_numpy()` / `[i]` face of a Taichi field; grows on demand (capacity doubles)."""

    def __init__(self, width, dtype=F, cap=4096):
        self.a = np.zeros((cap, width) if width else (cap,), dtype)

    def ensure(self, n):
        if n > self.a.shape[0]:
            cap = self.a.shape[0]
            while cap < n:
                cap *= 2
            b = np.zeros((cap,) + self.a.shape[1:], self.a.dtype)
            b[:self.a.shape[0]] = self.a
            self.a = b

    def to_numpy(self):
        return self.a

    def __getitem__(self, i):
        return self.a[i]

    def __setitem__(self, i, v):
        self.a[i] = v


class TopoGraphGen:
    def __init__(self, mapping, coll_det_num=128, max_raycast_dist=2, max_facelets=1024 * 1024,
                 thres_size=0.5, transparent=0.7, transparent_frontier=0.6, frontier_creation_threshold=0.5,
                 frontier_verify_threshold=0.5, frontier_backward_check=-0.2, frontier_combine_angle_threshold=40):
        self.mapping = mapping
        self.coll_det_num = coll_det_num
        self.generate_uniform_sample_points(coll_det_num)
        self.max_raycast_dist = max_raycast_dist
        self.max_facelets = max_facelets
        self.init_fields(max_facelets, coll_det_num)
        self.thres_size = thres_size
        self.init_colormap(max_facelets, transparent)
        self.frontier_creation_threshold = frontier_creation_threshold
        self.frontier_verify_threshold = frontier_verify_threshold
        self.transparent_frontier = transparent_frontier
        self.frontier_normal_dot_threshold = np.cos(np.deg2rad(frontier_combine_angle_threshold))
        self.check_frontier_small_distance = 0.1
        self.frontier_backward_check = frontier_backward_check

    # ------------------------------------------------------------------ set-up (:152-231)
    def init_colormap(self, max_facelets, transparent):
        # one random RGBA per NODE is all that is ever read (:398); the reference draws max_facelets of them
        self._color_rng = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
        self._transparent = F(transparent)
        self._colormap = {}
        self.debug_frontier_color = np.array([1.0, 1.0, 1.0, 0.6], F)

    def _node_color(self, k):
        c = self._colormap.get(k)
        if c is None:
            c = self._color_rng.rand(4).astype(F)
            c[3] = self._transparent
            self._colormap[k] = c
        return c

    def init_fields(self, max_facelets, coll_det_num, facelet_nh_search_queue_size=1024, max_map_node=4 * 1024):
        self.tri_vertices = HostField(3)
        self.tri_colors = HostField(4)
        # Facelet struct (:22-35) as columns
        self.f_normal, self.f_edge1, self.f_edge2 = HostField(3), HostField(3), HostField(3)
        self.f_v0, self.f_center = HostField(3), HostField(3)
        self.f_poly = HostField(0, np.int32)
        self.f_is_frontier = HostField(0, np.int32)
        self.f_assigned = HostField(0, np.int32)
        # Frontier struct (:76-94): dict per frontier index
        self.frontiers = {}
        # MapNode (:112-127)
        self.nodes = []
        self.num_facelets = HostScalar(0)
        self.num_nodes = HostScalar(0)
        self.num_frontiers = HostScalar(0)
        self.start_point = np.zeros(3, F)
        self.search_frontiers_idx = HostScalar(0)
        self.edges = HostField(3)
        self.edge_color = HostField(3)
        self.edge_num = HostScalar(0)
        self.connected = set()
        self.black_num = HostScalar(0)
        self.white_num = HostScalar(0)
        self.black_list = np.zeros((coll_det_num, 3), F)
        self.black_unit_list = np.zeros((coll_det_num, 3), F)
        self.black_len_list = np.zeros(coll_det_num, F)
        self.white_list = np.zeros((coll_det_num, 3), F)
        self.facelet_nh_search_queue_cap = facelet_nh_search_queue_size

    def reset(self):  # :199-209
        self.num_facelets[None] = 0
        self.num_nodes[None] = 0
        self.num_frontiers[None] = 0
        self.search_frontiers_idx[None] = 0
        self.edge_num[None] = 0
        self.black_num[None] = 0
        self.white_num[None] = 0
        self.connected = set()
        self.nodes = []
        self.frontiers = {}

    def generate_uniform_sample_points(self, npoints):  # :211-224 (Fibonacci sphere)
        phi = np.pi * (3 - np.sqrt(5))
        ret = []
        for i in range(npoints):
            y = 1 - 2 * (i / (npoints - 1))
            radius = np.sqrt(1 - y * y)
            theta = phi * i
            ret.append([np.cos(theta) * radius, y, np.sin(theta) * radius])
        self.sample_dirs = np.array(ret, dtype=F)

    def generate_random_sample_points(self, npoints):  # :226-231
        vec = np.random.randn(3, npoints)
        vec /= np.linalg.norm(vec, axis=0)
        self.sample_dirs = np.ascontiguousarray(vec.T.astype(F))

    # ------------------------------------------------------------------ ray casting against polyhedra + map (:472-507)
    def _facelet_range(self, pos, max_dist, skip_idx):
        """indices of the facelets of every node k != skip_idx with |pos - centre_k| < max_dist + max_raycast_dist, in
        the reference's iteration order (:477-481)."""
        idx = []
        lim = F(max_dist) + F(self.max_raycast_dist)
        for k, nd in enumerate(self.nodes):
            if k == skip_idx:
                continue
            dv = pos - nd["center"]
            if np.sqrt(_dot(dv, dv), dtype=F) < lim:
                idx.append(np.arange(nd["start"], nd["end"]))
        return np.concatenate(idx) if idx else np.zeros(0, np.int64)

    def detect_collision_facelets(self, pos, dir, max_dist, backward_dist=-0.01, skip_idx=-1):
        """:472-488 for ONE ray: (succ, position, t, poly index)."""
        pos, dir = np.asarray(pos, F), np.asarray(dir, F)
        best_t, best_poly, succ = F(max_dist), -1, False
        fi = self._facelet_range(pos, max_dist, skip_idx)
        if fi.size:
            e1, e2, v0 = self.f_edge1.a[fi], self.f_edge2.a[fi], self.f_v0.a[fi]
            # Facelet.rayTriangleIntersect (:52-71)
            q = _cross(dir[None], e2)
            a = _dot(e1, q)
            ok = np.abs(a) > F(0.00001)
            with np.errstate(divide="ignore", invalid="ignore"):
                s = ((pos[None] - v0) / a[:, None]).astype(F)
            r = _cross(s, e1)
            b0 = _dot(s, q)
            b1 = _dot(r, dir[None])
            b2 = (F(1.0) - b0 - b1).astype(F)
            t = _dot(e2, r)
            hit = ok & ~((b0 < 0) | (b1 < 0) | (b2 < 0))
            # sequential "if _succ and backward_dist < t < best_t: take it" == first strict minimum (:483-486)
            cand = hit & (t > F(backward_dist)) & (t < best_t)
            if cand.any():
                tt = np.where(cand, t, np.inf)
                j = int(np.argmin(tt))
                best_t, best_poly, succ = F(t[j]), int(self.f_poly.a[fi[j]]), True
        return succ, (pos + dir * best_t).astype(F), best_t, best_poly

    def raycast(self, pos, dir, max_dist, skip_idx=-1):
        """:490-507 for one ray: (succ, type 0=map 1=polyhedron, position, length, poly index)."""
        pos, dir = np.asarray(pos, F), np.asarray(dir, F)
        recast_type = 1
        succ_poly, pos_coll, len_coll, poly_ind = self.detect_collision_facelets(pos, dir, max_dist, -0.01, skip_idx)
        max_dist_recast = len_coll if succ_poly else max_dist
        succ_map, pos_col_map, len_map = self.mapping.raycast(pos, dir, float(max_dist_recast))
        if (not succ_poly) or (succ_map and len_map < len_coll):
            pos_coll, len_coll, recast_type, succ_poly = np.asarray(pos_col_map, F), F(len_map), 0, bool(succ_map)
        return succ_poly, recast_type, pos_coll, len_coll, poly_ind

    def _raycast_batch(self, pos, dirs, max_dist):
        """self.raycast for many rays at once (no skip index): the map part is ONE batched kernel per distinct length
        limit - rays that hit no polyhedron share `max_dist`, the others are cut at their polyhedron hit (:496-498)."""
        n = pos.shape[0]
        succ = np.zeros(n, bool)
        rtype = np.ones(n, np.int32)
        col = np.zeros((n, 3), F)
        ln = np.zeros(n, F)
        poly = np.full(n, -1, np.int32)
        limit = np.full(n, F(max_dist), F)
        for i in range(n):
            s, p, t, k = self.detect_collision_facelets(pos[i], dirs[i], max_dist, -0.01, -1)
            succ[i], col[i], ln[i], poly[i] = s, p, t, k
            if s:
                limit[i] = t
        free = ~succ
        if free.any():
            ms, mp, ml = self.mapping.raycast(pos[free], dirs[free], float(max_dist))
            succ[free], col[free], ln[free], rtype[free] = ms, mp, ml, 0
        for i in np.nonzero(~free)[0]:
            ms, mp, ml = self.mapping.raycast(pos[i], dirs[i], float(limit[i]))
            if ms and ml < ln[i]:
                col[i], ln[i], rtype[i] = mp, ml, 0
        return succ, rtype, col, ln, poly

    # ------------------------------------------------------------------ node expansion (:233-253, :444-470)
    def detect_collisions(self):
        pos = self.start_point
        n = self.coll_det_num
        succ, rtype, col, ln, poly = self._raycast_batch(np.broadcast_to(pos, (n, 3)).astype(F), self.sample_dirs, self.max_raycast_dist)
        nb = int(succ.sum())
        self.black_num[None] = nb
        self.white_num[None] = n - nb
        self.black_list[:nb] = col[succ]
        self.black_unit_list[:nb] = self.sample_dirs[succ]
        self.black_len_list[:nb] = ln[succ]
        self.white_list[:n - nb] = col[~succ]
        ray_len_black = F(0.0)
        for v in ln[succ]:
            ray_len_black = F(ray_len_black + v)
        self._neighbor_node_ids = []
        if nb == 0:
            return False
        node_size = ray_len_black / F(nb)
        if self.white_num[None] == 0 and node_size < self.thres_size:
            return False
        return True

    def node_expansion(self, start_pt, show=False, last_node_idx=-1):
        self.start_point = np.asarray(start_pt, F).reshape(3).copy()  # start_point field is f32 (:168)
        if self.detect_collisions():
            # the hull vertices are offset by the caller's start_pt as given (f64 for a user-supplied point, :296-303)
            self.generate_poly_on_blacks(np.asarray(start_pt).reshape(3), show, last_node_idx)

    def node_expansion_benchmark(self, start_pt, show=False, run_num=100):
        self.start_point = np.asarray(start_pt, F).reshape(3).copy()
        s = time.time()
        for _ in range(run_num):
            self.detect_collisions()
        print(f"avg detect_collisions time {(time.time() - s) * 1000 / run_num:.3f}ms")
        s = time.time()
        for _ in range(run_num):
            self.generate_poly_on_blacks(start_pt, show)
        print(f"avg gen convex cost time {(time.time() - s) * 1000 / run_num:.3f}ms")

    def generate_mesh_from_hull(self, hull, start_pt):  # :296-303
        start_pt = np.asarray(start_pt)
        lens = self.black_len_list[0:self.black_num[None]]
        vertices = hull.points * lens[:, None]
        vertices = vertices + start_pt[None]
        return vertices[hull.simplices], hull.neighbors

    def generate_poly_on_blacks(self, start_pt, show=False, last_node_idx=-1):  # :305-315
        from scipy.spatial import ConvexHull
        black_dirs = self.black_unit_list[0:self.black_num[None]]
        hull = ConvexHull(black_dirs)
        mesh, neighbors = self.generate_mesh_from_hull(hull, start_pt)
        self.add_mesh(mesh, neighbors, last_node_idx)

    # ------------------------------------------------------------------ facelets, frontiers (:317-442)
    def add_edge(self, a, b, color_a, color_b):
        e = self.edge_num[None]
        self.edges.ensure(e + 2)
        self.edge_color.ensure(e + 2)
        self.edges.a[e], self.edges.a[e + 1] = a, b
        self.edge_color.a[e], self.edge_color.a[e + 1] = color_a, color_b
        self.edge_num[None] = e + 2

    def _detect_facelet_frontier(self, lo, hi):
        """detect_facelet_frontier (:324-342) for the facelets [lo, hi) of the node being added, batched per query."""
        m = self.mapping
        center, normal = self.f_center.a[lo:hi], self.f_normal.a[lo:hi]
        is_frontier = np.ones(hi - lo, bool)
        near = m.is_near_pos_occupy(center, 0)
        unobs = m.is_pos_unobserved(center)
        is_frontier &= ~(near | unobs)
        start = (center + normal * F(m.voxel_scale)).astype(F)
        todo = np.nonzero(is_frontier)[0]
        if todo.size:
            occ = m.is_pos_occupy(start[todo])
            is_frontier[todo[occ]] = False
            todo = todo[~occ]
        if todo.size:
            succ, rtype, col, ln, poly = self._raycast_batch(start[todo], normal[todo], self.frontier_creation_threshold)
            for j in np.nonzero(succ & (rtype == 1))[0]:
                self._neighbor_node_ids.append(int(poly[j]))
            is_frontier[todo[succ]] = False
        return is_frontier

    def construct_frontier(self, node_idx, idx_start_facelet, queue):  # :344-378
        frontier_idx = self.num_frontiers[None]
        self.num_frontiers[None] = frontier_idx + 1
        center = np.zeros(3, F)
        normal = np.zeros(3, F)
        for q in queue:
            center = (center + self.f_center.a[q + idx_start_facelet]).astype(F)
            normal = (normal + self.f_normal.a[q + idx_start_facelet]).astype(F)
        center = (center / F(len(queue))).astype(F)
        normal = _normalized((normal / F(len(queue))).astype(F))
        fr = dict(master_idx=node_idx, frontier_idx=frontier_idx, avg_center=center, outwards_unit_normal=normal, is_valid=False,
                  projected_center=np.zeros(3, F), projected_normal=np.zeros(3, F), next_node_initial=np.zeros(3, F))
        self.frontiers[frontier_idx] = fr
        succ, t, projected_normal = False, F(0.0), np.zeros(3, F)
        for q in queue:
            fi = q + idx_start_facelet
            succ, t = self._ray_triangle(fi, center, normal)
            projected_normal = self.f_normal.a[fi].copy()
            if succ:
                break
        if succ:
            fr["projected_center"] = (center + t * normal).astype(F)
            fr["projected_normal"] = projected_normal
        else:
            self.num_frontiers[None] -= 1  # the slot is reused by the next frontier (:378)

    def _ray_triangle(self, fi, P, w):  # Facelet.rayTriangleIntersect (:52-71), scalar form
        e1, e2, v0 = self.f_edge1.a[fi], self.f_edge2.a[fi], self.f_v0.a[fi]
        q = _cross(w, e2)
        a = _dot(e1, q)
        if not (abs(a) > F(0.00001)):
            return False, F(0.0)
        s = ((P - v0) / a).astype(F)
        r = _cross(s, e1)
        b0, b1 = _dot(s, q), _dot(r, w)
        b2 = F(F(1.0) - b0 - b1)
        t = _dot(e2, r)
        return (not (b0 < 0.0 or b1 < 0.0 or b2 < 0.0)), F(t)

    def add_mesh(self, mesh, neighbors, last_node_idx):  # :380-442
        mesh = np.asarray(mesh)
        nf = mesh.shape[0]
        lo = self.num_facelets[None]
        hi = lo + nf
        if hi > self.max_facelets:
            raise RuntimeError(f"TopoGraphGen: {hi} facelets exceed max_facelets={self.max_facelets}")
        self.num_facelets[None] = hi
        for fld in (self.f_normal, self.f_edge1, self.f_edge2, self.f_v0, self.f_center, self.f_poly, self.f_is_frontier, self.f_assigned):
            fld.ensure(hi)
        self.tri_vertices.ensure(3 * hi)
        self.tri_colors.ensure(3 * hi)
        node = self.num_nodes[None]
        tv = mesh.astype(F).reshape(nf * 3, 3)
        self.tri_vertices.a[3 * lo:3 * hi] = tv
        v0, v1, v2 = tv[0::3], tv[1::3], tv[2::3]
        # Facelet.init (:36-50)
        e1, e2 = (v1 - v0).astype(F), (v2 - v0).astype(F)
        vsum = ((v0 + v1) + v2).astype(F)
        center = (vsum / F(3)).astype(F)
        normal = _normalized(_cross(e1, e2))
        naive = _normalized((vsum - F(3.0) * self.start_point[None]).astype(F))
        flip = _dot(normal, naive) < 0
        normal[flip] = -normal[flip]
        self.f_edge1.a[lo:hi], self.f_edge2.a[lo:hi], self.f_v0.a[lo:hi] = e1, e2, v0
        self.f_center.a[lo:hi], self.f_normal.a[lo:hi] = center, normal
        self.f_poly.a[lo:hi] = node
        self.f_assigned.a[lo:hi] = 0
        center_pos = np.zeros(3, F)
        for i in range(nf):  # f32 running sum in index order (:393-394)
            center_pos = (center_pos + vsum[i]).astype(F)
        center_count = F(3.0 * nf)
        self._neighbor_node_ids = []
        is_frontier = self._detect_facelet_frontier(lo, hi)
        self.f_is_frontier.a[lo:hi] = is_frontier
        col = self._node_color(node)
        tc = np.broadcast_to(col, (3 * nf, 4)).copy()
        tc[np.repeat(is_frontier, 3), 3] = F(self.transparent_frontier)
        self.tri_colors.a[3 * lo:3 * hi] = tc
        new_center = (center_pos / center_count).astype(F)
        self.nodes.append(dict(idx=node, master_idx=last_node_idx, start=lo, end=hi, center=new_center))
        black = np.zeros(3, F)
        if last_node_idx >= 0:
            self.add_edge(self.nodes[last_node_idx]["center"], new_center, black, black)
            self.connected.add((node, last_node_idx))
            self.connected.add((last_node_idx, node))
        for neigh in self._neighbor_node_ids:  # :407-413
            if (node, neigh) not in self.connected:
                self.connected.add((node, neigh))
                self.connected.add((neigh, node))
                self.add_edge(self.nodes[neigh]["center"], new_center, black, black)
        # cluster neighbouring frontier facelets (:415-441); queue entries are hull-local facelet indices
        thr = F(self.frontier_normal_dot_threshold)
        for i in range(lo, hi):
            if not self.f_assigned.a[i] and self.f_is_frontier.a[i]:
                queue = [i - lo]
                normal_i = self.f_normal.a[i]
                head = 0
                while head < len(queue):
                    _idx = queue[head]
                    head += 1
                    self.f_assigned.a[_idx + lo] = 1
                    for j in range(3):
                        nb = int(neighbors[_idx, j]) + lo
                        if self.f_is_frontier.a[nb] and not self.f_assigned.a[nb] and _dot(normal_i, self.f_normal.a[nb]) > thr:
                            if len(queue) >= self.facelet_nh_search_queue_cap:
                                raise RuntimeError("TopoGraphGen: facelet_nh_search_queue overflow")
                            queue.append(int(neighbors[_idx, j]))
                self.construct_frontier(node, lo, queue)
        self.num_nodes[None] = node + 1

    # ------------------------------------------------------------------ graph growth (:255-294)
    def verify_frontier(self, frontier_idx):
        fr = self.frontiers[frontier_idx]
        normal = fr["projected_normal"]
        small = F(self.check_frontier_small_distance)
        proj_center = (fr["projected_center"] + normal * small).astype(F)
        succ, t, col_pos, _len, node_idx = self.raycast(proj_center, normal, self.max_raycast_dist * 2)
        if succ and _len < self.frontier_verify_threshold:
            fr["is_valid"] = False
        else:
            proj_center = (fr["projected_center"] - normal * small).astype(F)
            succ2, col_pos2, _len2, node_idx2 = self.detect_collision_facelets(proj_center, normal, self.frontier_verify_threshold,
                                                                               self.frontier_backward_check, fr["master_idx"])
            if succ2 and _len2 < self.frontier_verify_threshold:
                fr["is_valid"] = False
            else:
                if not succ or succ2 and _len2 < _len:
                    _len = _len2
                fr["is_valid"] = True
                fr["next_node_initial"] = (fr["projected_center"] + fr["projected_normal"] * F(_len) / F(2)).astype(F)
        return fr["is_valid"]

    def generate_topo_graph(self, start_pt, max_nodes=100, show=False):
        self.node_expansion(start_pt, show)
        while self.search_frontiers_idx[None] < self.num_frontiers[None] and self.search_frontiers_idx[None] < max_nodes:
            k = self.search_frontiers_idx[None]
            if self.verify_frontier(k):
                fr = self.frontiers[k]
                self.node_expansion(fr["next_node_initial"], show, last_node_idx=fr["master_idx"])
            self.search_frontiers_idx[None] += 1
        return self.num_nodes[None]

    def test_detect_collisions(self, start_pt):
        self.start_point = np.asarray(start_pt, F).reshape(3).copy()
        self.detect_collisions()
