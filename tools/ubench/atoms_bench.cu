// Shared-memory integer atomic throughput on sm_100a: what does ONE ATOMS.ADD cost per warp instruction
// as a function of the address pattern, return-value use and warps per SM?  (round-2 design input for the
// block-binned ray march: can 2-4 shared atomics per sample beat one global RED per sample?)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
#define WORDS 16384  // 64 KB of accumulators per CTA

__device__ __forceinline__ unsigned hash32(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// mode 0: conflict-free (lane-consecutive, rotating base)   1: random words   2: ray-like (stride 256/16/1 mixes)
// 3: single bank (stride 32)   4: random + swizzle-free 4 atomics per sample (lo/hi x A/B: addr, +4096, +8192, +12288)
// ret: 0 = result unused, 1 = result used (carry-style dependency)
template<int MODE, int RET>
__global__ void __launch_bounds__(512) k_atoms(int iters, unsigned* out, long long* cyc){
  extern __shared__ unsigned sm[];
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned acc = 0;
  unsigned r = hash32(blockIdx.x * 1024 + threadIdx.x);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    unsigned a;
    if (MODE == 0) a = (lane + it * 33 + wid * 97) & (WORDS - 1);
    else if (MODE == 1 || MODE == 4) { r = r * 1664525u + 1013904223u; a = (r >> 10) & 4095; }
    else if (MODE == 2) { // a warp = 32 consecutive steps of one ray through a 16^3 block, direction varies per warp/iter
      const unsigned d = (it + wid) % 3; const unsigned st = d == 0 ? 1u : (d == 1 ? 16u : 256u);
      a = ((lane & 15) * st + (it * 37 + wid * 11) * (st == 1 ? 16 : 1)) & 4095; }
    else a = (lane * 32 + it) & (WORDS - 1);
    if (MODE == 4) {
      if (RET) { unsigned o = atomicAdd(&sm[a], r | 1u); if (o + (r | 1u) < o) atomicAdd(&sm[a + 4096], 1u);
                 unsigned o2 = atomicAdd(&sm[a + 8192], lane | 1u); if (o2 + (lane | 1u) < o2) atomicAdd(&sm[a + 12288], 1u); acc += o ^ o2; }
      else { atomicAdd(&sm[a], r); atomicAdd(&sm[a + 4096], r >> 16); atomicAdd(&sm[a + 8192], lane); atomicAdd(&sm[a + 12288], lane >> 3); }
    } else {
      if (RET) acc += atomicAdd(&sm[a], it | 1u); else atomicAdd(&sm[a], it | 1u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  unsigned s = acc;
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s += sm[i];
  if (s == 0x12345678u) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE, int RET>
static void run(const char* name, int threads, int ctas_per_sm, int sms){
  const int iters = 4096;
  unsigned* out; long long* cyc; CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8 * sms * ctas_per_sm));
  CK(cudaFuncSetAttribute(k_atoms<MODE,RET>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_atoms<MODE,RET><<<sms * ctas_per_sm, threads, WORDS * 4>>>(16, out, cyc);
  cudaEventRecord(e0);
  k_atoms<MODE,RET><<<sms * ctas_per_sm, threads, WORDS * 4>>>(iters, out, cyc);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024]; CK(cudaMemcpy(h, cyc, 8 * sms * ctas_per_sm, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < sms * ctas_per_sm; i++) avg += h[i]; avg /= sms * ctas_per_sm;
  const double per = (MODE == 4) ? 4.0 : 1.0;
  const double warps_sm = threads / 32.0 * ctas_per_sm;
  const double winstr_sm = warps_sm * iters * per;   // atomic warp-instructions per SM (mode 4 RET: 2 + rare carries)
  printf("%-34s ret %d  thr %4d x %d CTA/SM: %.3f ms  %.2f cyc per atomic warp-instr per SM   %.1f G lane-atomics/s\n", name, RET, threads, ctas_per_sm, ms,
         avg / winstr_sm, (double)sms * ctas_per_sm * threads * iters * per / ms * 1e-6);
  cudaFree(out); cudaFree(cyc);
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("tools/ubench/atoms_bench.cu on %s (%d SMs)\n", p.name, sms);
  for (int cps = 1; cps <= 2; cps++) for (int thr = 256; thr <= 512; thr *= 2) {
    run<0,0>("conflict-free", thr, cps, sms); run<0,1>("conflict-free", thr, cps, sms);
    run<1,0>("random in 4096 words", thr, cps, sms); run<1,1>("random in 4096 words", thr, cps, sms);
    run<2,0>("ray-like in 16^3", thr, cps, sms); run<2,1>("ray-like in 16^3", thr, cps, sms);
    run<3,0>("single bank", thr, cps, sms);
    run<4,0>("random, 4 no-ret atomics/sample", thr, cps, sms); run<4,1>("random, 2 ret atomics + carry", thr, cps, sms);
  }
  return 0;
}
