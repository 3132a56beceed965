// Random 64-byte reads from a table of S bytes on B200: (a) throughput of INDEPENDENT reads (one per thread, many
// threads in flight) and (b) latency of a dependent chain - as a function of S.  Question behind it (round 2): the
// per-frame bucket tables are 537 MB; k_ray_setup reads one random 64-byte entry per ray and runs at 6 % issue
// utilisation - is that DRAM latency, DRAM random-access throughput or address translation?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ unsigned hash32(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

__global__ void k_indep(const uint4* __restrict__ tab, unsigned mask, int per_thread, unsigned* out, int write_back, uint4* wtab){
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (int i = 0; i < per_thread; i++) {
    unsigned idx = hash32(t * 131u + i * 7919u) & mask;       // 64-byte entry index
    const uint4 a = tab[(size_t)idx * 4], b = tab[(size_t)idx * 4 + 2];
    acc += a.x ^ b.y;
    if (write_back) { wtab[(size_t)idx * 4] = make_uint4(0,0,0,0); wtab[(size_t)idx * 4 + 1] = make_uint4(0,0,0,0); wtab[(size_t)idx * 4 + 2] = make_uint4(0,0,0,0); wtab[(size_t)idx * 4 + 3] = make_uint4(0,0,0,0); }
  }
  if (acc == 0x12345u) out[0] = acc;
}
__global__ void k_chain(const unsigned* __restrict__ next, unsigned mask, int steps, unsigned* out, long long* cyc){
  unsigned p = hash32(blockIdx.x * blockDim.x + threadIdx.x) & mask;
  long long t0 = clock64();
  for (int i = 0; i < steps; i++) p = next[(size_t)p * 16] & mask;
  long long t1 = clock64();
  if (p == 0xFFFFFFFFu) out[0] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fill(unsigned* tab, size_t n_words){ for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) tab[i] = hash32((unsigned)i * 2654435761u); }

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("tools/ubench/rand_bench.cu on %s: random 64-byte entries\n", pr.name);
  unsigned* out; long long* cyc; CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
  for (size_t mb : {32, 128, 512, 2048}) {
    const size_t bytes = mb << 20, entries = bytes / 64;
    unsigned* tab; CK(cudaMalloc(&tab, bytes));
    k_fill<<<1184, 256>>>(tab, bytes / 4); CK(cudaDeviceSynchronize());
    const unsigned mask = (unsigned)(entries - 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wb = 0; wb < 2; wb++) {
      const int threads = 739000, per = 1;  // one entry per "ray", like k_ray_setup
      k_indep<<<(threads + 255) / 256, 256>>>((const uint4*)tab, mask, per, out, wb, (uint4*)tab); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int r = 0; r < 5; r++) k_indep<<<(threads + 255) / 256, 256>>>((const uint4*)tab, mask, per, out, wb, (uint4*)tab);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("  table %5zu MB  independent %s: 739k entries in %.1f us = %.1f G entries/s\n", mb, wb ? "read+zero" : "read     ", ms * 1e3, 739000 / ms * 1e-6);
    }
    k_chain<<<1, 32>>>(tab, mask, 2000, out, cyc); CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("  table %5zu MB  dependent chain (1 warp, idle GPU): %.0f cycles per load\n", mb, (double)h / 2000);
    k_chain<<<148 * 8, 256>>>(tab, mask, 200, out, cyc); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_chain<<<148 * 8, 256>>>(tab, mask, 200, out, cyc); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("  table %5zu MB  dependent chains, 303k threads x 200 loads: %.2f ms = %.1f G loads/s, %.0f ns per load per thread\n", mb, ms, 148.0 * 8 * 256 * 200 / ms * 1e-6, ms * 1e6 / 200);
    cudaFree(tab);
  }
  return 0;
}
