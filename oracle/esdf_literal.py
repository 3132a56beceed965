"""ESDF propagation of the reference, restated LITERALLY (test infrastructure; SURVEY.md Appendix A.6: "a
literal_single_hop oracle switch reproduces the code as written for the record").

Follows DenseSDF.propogate_esdf and its queues line by line (taichi_slam/mapping/dense_esdf.py:228-333) on a dense
N^3 state, f32 values:
  * seeds (:309-330): for every voxel of `updated` in lexicographic order - a FIXED voxel (|TSDF| < gamma, :228-230)
    takes ESDF = TSDF; it goes to the lower queue if its old ESDF was larger or it was unobserved (:316-319), else to
    the raise AND the lower queue (:320-323); a non-fixed voxel seen for the first time takes sign(TSDF)*max_ray_length
    and queues its 26 NEIGHBOURS for lowering (:328-330) (`self.fixed` is never written: the :324 branch is dead);
  * raise queue (:260-272): the head goes back to sign(ESDF)*max_ray_length; a neighbour whose parent_dir equals the
    step direction is raised too, every other neighbour is queued for lowering;
  * lower queue (:274-299): every queued voxel relaxes its 26 neighbours ONE hop - `n > 0 and head + dis < n` or
    `n < 0 and head - dis > n` - and improved neighbours are NOT re-queued (the insert_lower calls at :292/:298 are
    commented out), so the result depends on the queue order and is not a converged distance field.
Only voxels with a TSDF entry are "active" (insert_* :233-251); queue entries are the float index truncated to i32.

Pinned by tests/golden/ref_exec_esdf.npz: the reference's own functions executed through oracle/taichi_emu.py on a
hand-built state (tools/make_golden_esdf.py) - tests/test_oracle_cpu.py::test_esdf_literal_matches_executed_reference.
The product (csrc/tslam_esdf.cu) and the oracle's default ESDF implement the CONVERGED field instead; what
the two share with this literal code - band, seeds, edge costs, sign rules - is what the pin covers.
"""
import numpy as np

NEIGHBORS = [(a, b, c) for a in range(-1, 2) for b in range(-1, 2) for c in range(-1, 2) if a or b or c]  # dense_esdf.py:141-146


def _sign(v):  # mapping_common.py:5-7
    return int(0 < v) - int(v < 0)


class LiteralESDF:
    def __init__(self, n, voxel_scale=0.05, max_ray_length=3.0):
        self.n = n
        self.vs = np.float32(voxel_scale)
        self.gamma = np.float32(voxel_scale)          # :40
        self.far = np.float32(max_ray_length)
        self.tsdf = {}                                 # active cells
        self.esdf, self.observed, self.parent = {}, {}, {}

    def _active(self, k):
        return k in self.tsdf

    def propagate(self, tsdf, updated_mask):
        f32 = np.float32
        upd = sorted(tuple(int(v) for v in k) for k in np.argwhere(updated_mask))
        for k in upd:
            self.tsdf[k] = f32(tsdf[k])
        raise_q, lower_q = [], []

        def ins_lower(k):
            if self._active(k):
                lower_q.append(k)

        def ins_raise(k):
            if self._active(k):
                raise_q.append(k)

        for k in upd:  # :309-330
            t = self.tsdf[k]
            if abs(t) < self.gamma:
                if self.esdf.get(k, f32(0)) > t or self.observed.get(k, 0) == 0:
                    self.observed[k] = 1
                    self.esdf[k] = t
                    ins_lower(k)
                else:
                    self.esdf[k] = t
                    ins_raise(k)
                    ins_lower(k)
            elif self.observed.get(k, 0) == 0:
                self.observed[k] = 1
                self.esdf[k] = f32(_sign(t)) * self.far
                for d in NEIGHBORS:
                    ins_lower((k[0] + d[0], k[1] + d[1], k[2] + d[2]))
        h = 0
        while h < len(raise_q):  # :260-272
            k = raise_q[h]
            h += 1
            self.esdf[k] = f32(_sign(self.esdf.get(k, f32(0)))) * self.far
            for d in NEIGHBORS:
                nk = (k[0] + d[0], k[1] + d[1], k[2] + d[2])
                if tuple(self.parent.get(nk, (0, 0, 0))) == d:
                    ins_raise(nk)
                else:
                    ins_lower(nk)
        h = 0
        while h < len(lower_q):  # :274-299
            k = lower_q[h]
            h += 1
            for d in NEIGHBORS:
                nk = (k[0] + d[0], k[1] + d[1], k[2] + d[2])
                if not self._active(nk):
                    continue
                dis = f32(np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))) * self.vs   # dir.norm()*voxel_scale
                n_esdf = self.esdf.get(nk, f32(0))
                cand = self.esdf.get(k, f32(0)) + dis
                if n_esdf > 0 and cand < n_esdf:
                    self.esdf[nk] = cand
                    self.parent[nk] = (-d[0], -d[1], -d[2])
                else:
                    cand = self.esdf.get(k, f32(0)) - dis
                    if n_esdf < 0 and cand > n_esdf:
                        self.esdf[nk] = cand
                        self.parent[nk] = (-d[0], -d[1], -d[2])
        return len(raise_q), len(lower_q)
