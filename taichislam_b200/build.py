"""Build libtslam.so (hand-written sm_100a CUDA + the C ABI of include/tslam.h) in-tree.

    python -m taichislam_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands next to this file so it
travels to the GPU box with the repo snapshot.  cudart is linked statically (nvcc's
default): the library shares the primary context - hence device pointers and
streams - with PyTorch in the same process.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
DEPS = SRC + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
OUT = os.path.join(HERE, "libtslam.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # index-forming f32 arithmetic must round like the strict-IEEE oracle
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + SRC
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtslam.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
