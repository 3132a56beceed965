"""Field - a minimal stand-in for the Taichi fields the reference's callers touch.

Callers of the reference classes read and write a handful of fields directly
(`num_TSDF_particles[None]`, `export_TSDF_xyz.to_numpy()`, `export_color[i] = ...`,
tests/gen_topo_graph.py:64-66; taichislam_node.py:304-346).  A Field wraps a torch CUDA
tensor (bulk, written by the CUDA kernels through its data pointer) or a host scalar.
Reads synchronise lazily, like Taichi's.
"""
import numpy as np


class Field:
    def __init__(self, tensor, on_read=None):
        self.t = tensor          # torch tensor (CUDA) - shape () / (n,) / (n, k)
        self._on_read = on_read  # callable run before any host read (flushes pending GPU work)

    @property
    def shape(self):
        return tuple(self.t.shape[:1]) if self.t.dim() > 1 else tuple(self.t.shape)

    def _sync(self):
        if self._on_read is not None:
            self._on_read()

    def __getitem__(self, i):
        self._sync()
        if i is None:
            v = self.t.reshape(-1)[0].item()
            return v
        v = self.t[i]
        return v.item() if v.dim() == 0 else v.cpu().numpy()

    def __setitem__(self, i, value):
        import torch
        if i is None:
            self.t.reshape(-1)[0] = value
        else:
            self.t[i] = torch.as_tensor(np.asarray(value), dtype=self.t.dtype, device=self.t.device)

    def to_numpy(self):
        self._sync()
        return self.t.cpu().numpy()

    def from_numpy(self, a):
        import torch
        self.t.copy_(torch.as_tensor(np.ascontiguousarray(a), dtype=self.t.dtype).reshape(self.t.shape))

    def to_torch(self):
        self._sync()
        return self.t

    def fill(self, v):
        self.t.fill_(v)


class HostScalar:
    """0-d int field kept on the host (active_submap_id, remote_submap_num, mapping_common.py:108-111)."""

    def __init__(self, v=0):
        self.v = v

    def __getitem__(self, i):
        return self.v

    def __setitem__(self, i, value):
        self.v = int(value)

    def to_numpy(self):
        return np.array(self.v)
