"""MarchingCubeMesher - the reference's class surface (marching_cube_mesher.py:12-193) on the
two-pass CUDA marching cubes of libtslam.so (tslam_mc_generate2)."""
import ctypes as C

from .. import _capi as capi
from .field import Field


class MarchingCubeMesher:
    def __init__(self, mapping, max_triangles=1000000, tsdf_surface_thres=0.1):
        import torch
        self._torch = torch
        dev = torch.device("cuda", torch.cuda.current_device())
        self.max_triangles = max_triangles
        # mesh_vertices pre-filled with -1e6 sentinels (:35-38); colours/normals start at zero like ti fields
        self.mesh_vertices = Field(torch.full((max_triangles * 3, 3), -1000000.0, dtype=torch.float32, device=dev))
        self.mesh_colors = Field(torch.zeros((max_triangles * 3, 3), dtype=torch.float32, device=dev))
        self.mesh_normals = Field(torch.zeros((max_triangles * 3, 3), dtype=torch.float32, device=dev))
        self.mesh_indices = None
        self.num_facelets = Field(torch.zeros(1, dtype=torch.int32, device=dev))
        self.num_vertices = Field(torch.zeros(1, dtype=torch.int32, device=dev))
        self.mapping = mapping
        self.enable_texture = mapping.enable_texture
        self.tsdf_surface_thres = tsdf_surface_thres

    def generate_mesh(self, step=1):
        """:192-193 -> generate_mesh_kernel (:180-187).  num_facelets holds the triangle DEMAND; when it exceeds
        max_triangles the buffers hold the first max_triangles (the reference writes out of bounds there, :175-177)."""
        m = self.mapping
        m._flush()
        n = C.c_int64(0)
        colors = capi.tptr(self.mesh_colors.t) if self.enable_texture else None  # add_triangle_color (:104-108, :120-125)
        rc = m._h.L.tslam_mc_generate2(m._h.h, int(step), float(self.tsdf_surface_thres), self.max_triangles,
                                       capi.tptr(self.mesh_vertices.t), capi.tptr(self.mesh_normals.t), colors, C.byref(n),
                                       capi.stream_ptr())
        if rc != capi.E_CAPACITY:
            capi.check(rc)
        self.num_facelets[None] = int(n.value)
        # the reference never writes num_vertices (:22) although the node reads it (taichislam_node.py:342)
        self.num_vertices[None] = 3 * min(int(n.value), self.max_triangles)
        print("Total triangles", int(n.value))

    def vertice_num(self):  # :189-190
        return self.num_facelets[None] * 3
