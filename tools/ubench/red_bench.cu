// Micro-benchmark: throughput of 8-byte vector reductions (RED.ADD.F32x2) on B200 for the access patterns
// of the ray-march kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void red2(float2* a, float x, float y) {
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
// mode 0: random address per lane per step within `n_vox` voxels
// mode 1: ray-like: lane walks a straight line through a dense 3D grid of side S (voxel (x,y,z) -> blocked 16^3 layout)
// mode 2: as 1 but all lanes of a warp are neighbouring rays (coherent)
// mode 3: no memory op (ALU only baseline with the same address math)
__global__ void k(float2* acc, uint32_t n_vox, int S, int steps, int mode, int reps, float* sink, int skip, int n_origins) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  float keep = 0.f;
  for (int r = 0; r < reps; ++r) {
    uint32_t seed = hash32(tid * 9781u + r * 7919u + 1u);
    float dx, dy, dz;
    if (mode == 2) {
      const uint32_t w = tid >> 5, l = tid & 31;
      seed = hash32(w * 9781u + r * 7919u + 1u);
      dx = ((seed & 1023) / 1023.f - 0.5f) + (l & 7) * 0.01f;
      dy = (((seed >> 10) & 1023) / 1023.f - 0.5f) + (l >> 3) * 0.01f;
      dz = 0.8f;
    } else {
      dx = (seed & 1023) / 1023.f - 0.5f; dy = ((seed >> 10) & 1023) / 1023.f - 0.5f; dz = 0.8f;
    }
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    const float c = S * 0.5f;
    const uint32_t og = hash32((tid >> 7) % (uint32_t)n_origins + 17u);
    const float ox = c + (float)(og & 31) - 16.f, oy = c + (float)((og >> 5) & 31) - 16.f;
    for (int s = 1 + skip; s <= steps + skip; ++s) {
      uint32_t off;
      if (mode == 0) {
        off = hash32(seed + s * 2654435761u) % n_vox;
      } else {
        const int x = (int)roundf(ox + dx * s), y = (int)roundf(oy + dy * s), z = (int)roundf(dz * s);
        const int bx = x >> 4, by = y >> 4, bz = z >> 4, nb = S >> 4;
        off = (uint32_t)(((bx * nb + by) * nb + bz) * 4096 + (((x & 15) << 4 | (y & 15)) << 4 | (z & 15)));
      }
      if (mode == 3) keep += (float)off; else red2(&acc[off], 1.0f, 0.5f);
    }
  }
  if (keep == 123.456f) *sink = keep;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  const int S = 256;                       // 256^3 voxels * 8 B = 134 MB dense (ray-like modes touch a cone)
  const size_t nv = (size_t)S * S * S;
  float2* acc; cudaMalloc(&acc, nv * 8); cudaMemset(acc, 0, nv * 8);
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int steps = 90, reps = 3;
  {
    const int tpb = 128, grid = p.multiProcessorCount * 16;
    auto run = [&](int mode, uint32_t foot, int skip, int norg, const char* tag) {
      k<<<grid, tpb>>>(acc, foot, S, steps, mode, 1, sink, skip, norg);
      cudaEventRecord(e0);
      k<<<grid, tpb>>>(acc, foot, S, steps, mode, reps, sink, skip, norg);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)grid * tpb * steps * reps;
      printf("%-28s mode %d skip %2d origins %5d: %.3f ms  %.1f G updates/s\n", tag, mode, skip, norg, ms, ops / ms * 1e-6);
    };
    run(0, 1u << 20, 0, 1, "random 8MB");
    run(3, 0, 0, 1, "alu only");
    for (int skip : {0, 2, 4, 8, 16, 32}) run(1, 0, skip, 1, "ray-like, one origin");
    for (int norg : {1, 8, 64, 512, 2368}) run(1, 0, 0, norg, "ray-like, origins per CTA");
    for (int norg : {64, 2368}) for (int skip : {4, 16}) run(1, 0, skip, norg, "ray-like, origins+skip");
    for (int norg : {1, 64}) run(2, 0, 0, norg, "coherent warps");
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
