"""`ti` - the few taichi entry points callers reach through `from taichi_slam.mapping import *`
(scripts/taichislam_node.py:34-36: ti.init(arch=ti.cuda, dynamic_index=True, debug=False,
device_memory_GB=4) / ti.cpu).  There is no Taichi here: ti.init only checks that a CUDA
device exists (the backend has no CPU path, so arch=ti.cpu raises)."""


class _Arch:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"ti.{self.name}"


class _TiShim:
    cpu = _Arch("cpu")
    cuda = _Arch("cuda")
    gpu = cuda
    f16, f32, f64, i8, i16, i32, i64, u8, u16 = "f16", "f32", "f64", "i8", "i16", "i32", "i64", "u8", "u16"

    def init(self, arch=None, **kwargs):
        from .. import _capi
        if arch is self.cpu:
            raise RuntimeError("taichislam_b200 has no CPU backend: use ti.init(arch=ti.cuda)")
        _capi.require_gpu()
        self.arch = arch or self.cuda
        return self

    def sync(self):
        import torch
        torch.cuda.synchronize()

    def reset(self):
        pass


ti = _TiShim()
