// libtslam.so - ESDF: converged 26-neighbour signed-distance wavefront over the
// observed voxels of one submap.
//
// Semantic source: DenseSDF.propogate_esdf and its raise/lower queues
// (dense_esdf.py:228-333).  That code is dead at HEAD (not exported, stale ctor and
// indices) and its lower queue never re-inserts improved voxels (:292,:298 commented
// out), so there is no runnable reference behaviour: PARITY UNPINNED.  We implement
// the converged fixed point the code describes (DESIGN.md "ESDF"):
//   fixed band |TSDF| < gamma = voxel_scale (is_fixed :228-230): ESDF = TSDF (:315-320)
//   other observed voxels start at sign(TSDF)*max_ray_length (:324,:328) and relax
//     positive side: E(v) = min(E(v), E(h) + |dir|*vs)   h in fixed U positive   (:286-291)
//     negative side: E(v) = max(E(v), E(h) - |dir|*vs)   h in fixed U negative   (:294-298)
//   to convergence.  f32 min/+ is monotone, so the least fixed point is unique and
//   equals the oracle's multi-source Dijkstra bit-for-bit.
// Implementation: block-parallel wavefront with a block-level WORK QUEUE.
//   * a CTA stages its block plus a one-voxel halo (18^3) in shared memory, relaxes to
//     LOCAL convergence (ballot-voted iterations), writes back and stamps the block
//     with the sweep number when anything changed;
//   * sweep k only runs the blocks that have a neighbour (26 + self) stamped >= k-1 -
//     the wave front; every block's 27 neighbour indices are resolved once per update;
//   * sweeps are enqueued in groups; a sweep whose predecessor changed nothing returns
//     at once, so the host only synchronises once per group to test for convergence.
#include <cstring>
#include "tslam_internal.cuh"

#define ES_T 18
#define ES_T3 (ES_T * ES_T * ES_T)
#define ES_GROUP 6       // sweeps enqueued per host synchronisation
#define ES_MAX_SWEEPS 4096

enum { ES_UNOBS = 0, ES_FIXED = 1, ES_POS = 2, ES_NEG = 3, ES_INERT = 4 };  // INERT: observed, TSDF is NaN

struct EsAux {
  int* nbr;      // [max_blocks*27] neighbour block indices (-1 = absent)
  int* epoch;    // [max_blocks] last sweep in which the block changed
  int* changed;  // [ES_MAX_SWEEPS+2] changed[k] != 0 <=> sweep k changed something
};

__device__ __forceinline__ float es_sgn(float v) { return (float)((0.0f < v) - (v < 0.0f)); }

__global__ void __launch_bounds__(256) k_esdf_init(TsGrid g, EsAux aux, int submap, float gamma, float far_v) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    int s, bx, by, bz;
    ts_unpack_key(g.block_key[b], s, bx, by, bz);
    if (threadIdx.x == 0) aux.epoch[b] = (s == submap) ? 0 : -1000000;
    if (s != submap) continue;
    if (threadIdx.x < 27) {
      const int dx = threadIdx.x / 9 - 1, dy = (threadIdx.x / 3) % 3 - 1, dz = threadIdx.x % 3 - 1;
      aux.nbr[b * 27 + threadIdx.x] = ts_find(g, ts_pack_key(s, bx + dx, by + dy, bz + dz));
    }
    const size_t base = (size_t)b * TS_B3;
    for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
      float e = 0.0f;
      if (g.obs[base + v]) {
        const float t = g.tw[base + v].x;
        e = (fabsf(t) < gamma) ? t : es_sgn(t) * far_v;
      }
      g.esdf[base + v] = e;
    }
  }
}

struct EsTile {
  float e[ES_T3];
  unsigned char cls[ES_T3];
  int nbr[27];
  int active;
  int changed;
  int any_write;
};

__device__ __forceinline__ int es_tidx(int lx, int ly, int lz) { return ((lx + 1) * ES_T + (ly + 1)) * ES_T + (lz + 1); }

__global__ void __launch_bounds__(256) k_esdf_sweep(TsGrid g, EsAux aux, int submap, int sweep, float gamma, float vs) {
  __shared__ EsTile tile;
  if (sweep > 1 && aux.changed[sweep - 1] == 0) return;  // already converged: nothing to do
  const int nb = min(*g.n_blocks, g.max_blocks);
  const float d1 = vs, d2 = sqrtf(2.0f) * vs, d3 = sqrtf(3.0f) * vs;  // dir.norm()*voxel_scale (dense_esdf.py:285)
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) { tile.active = 0; tile.any_write = 0; }
    __syncthreads();
    if (aux.epoch[b] < -1) continue;  // block of another submap (uniform per CTA)
    if (threadIdx.x < 27) {
      const int n = aux.nbr[b * 27 + threadIdx.x];
      tile.nbr[threadIdx.x] = n;
      if (n >= 0 && *(volatile int*)&aux.epoch[n] >= sweep - 1) tile.active = 1;  // the wave front reaches this block
    }
    __syncthreads();
    if (!tile.active) continue;
    for (int c = threadIdx.x; c < ES_T3; c += blockDim.x) {
      const int lz = c % ES_T - 1, ly = (c / ES_T) % ES_T - 1, lx = c / (ES_T * ES_T) - 1;
      const int nx = (lx + 16) >> 4, ny = (ly + 16) >> 4, nz = (lz + 16) >> 4;
      const int nbk = tile.nbr[(nx * 3 + ny) * 3 + nz];
      float ev = 0.0f;
      unsigned char cl = ES_UNOBS;
      if (nbk >= 0) {
        const size_t off = (size_t)nbk * TS_B3 + ((((lx & 15) << 4) | (ly & 15)) << 4 | (lz & 15));
        if (g.obs[off]) {
          const float t = g.tw[off].x;
          cl = (fabsf(t) < gamma) ? ES_FIXED : (t > 0.0f ? ES_POS : (t < 0.0f ? ES_NEG : ES_INERT));
          ev = *(volatile float*)&g.esdf[off];
        }
      }
      tile.e[c] = ev;
      tile.cls[c] = cl;
    }
    __syncthreads();
    // relax to local convergence (values only move monotonically, so in-place racy updates are safe)
    for (int iter = 0; iter < 64; ++iter) {
      if (threadIdx.x == 0) tile.changed = 0;
      __syncthreads();
      bool mine = false;
      for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
        const int lx = v >> 8, ly = (v >> 4) & 15, lz = v & 15;
        const int c0 = es_tidx(lx, ly, lz);
        const unsigned char cl = tile.cls[c0];
        if (cl != ES_POS && cl != ES_NEG) continue;
        float best = tile.e[c0];
        const float cur = best;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dz = -1; dz <= 1; dz++) {
              const int nz_ = (dx != 0) + (dy != 0) + (dz != 0);
              if (nz_ == 0) continue;
              const int cn = c0 + (dx * ES_T + dy) * ES_T + dz;
              const unsigned char hc = tile.cls[cn];
              if (!(hc == ES_FIXED || hc == cl)) continue;
              const float dis = nz_ == 1 ? d1 : (nz_ == 2 ? d2 : d3);
              const float he = tile.e[cn];
              if (cl == ES_POS) best = fminf(best, he + dis); else best = fmaxf(best, he - dis);
            }
        if (best != cur) {
          tile.e[c0] = best;
          mine = true;
        }
      }
      if (__any_sync(0xffffffffu, mine) && (threadIdx.x & 31) == 0) tile.changed = 1;
      __syncthreads();
      const int ch = tile.changed;
      if (ch && threadIdx.x == 0) tile.any_write = 1;
      __syncthreads();
      if (!ch) break;
    }
    if (tile.any_write) {
      const size_t base = (size_t)b * TS_B3;
      for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) g.esdf[base + v] = tile.e[es_tidx(v >> 8, (v >> 4) & 15, v & 15)];
      if (threadIdx.x == 0) {
        __threadfence();
        aux.epoch[b] = sweep;
        aux.changed[sweep] = 1;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_esdf_gather(TsGrid g, int submap, long long cap, int32_t* idx, float* esdf,
                                                      unsigned long long* counter) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    int s, bx, by, bz;
    ts_unpack_key(g.block_key[b], s, bx, by, bz);
    if (s != submap || g.ghost[b]) continue;
    const size_t base = (size_t)b * TS_B3;
    for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
      const bool want = g.obs[base + v] > 0;
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (!m) continue;
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      unsigned long long rb = 0;
      if (lane == leader) rb = atomicAdd(counter, (unsigned long long)__popc(m));
      rb = __shfl_sync(0xffffffffu, rb, leader);
      const long long row = (long long)(rb + __popc(m & ((1u << lane) - 1)));
      if (want && row < cap) {
        idx[3 * row] = bx * TS_B + (v >> 8);
        idx[3 * row + 1] = by * TS_B + ((v >> 4) & 15);
        idx[3 * row + 2] = bz * TS_B + (v & 15);
        esdf[row] = g.esdf[base + v];
      }
    }
  }
}

extern "C" int tslam_esdf_update(tslam_tsdf_t* m, int32_t submap, int32_t* n_sweeps_out, void* stream) {
  if (!m) return TSLAM_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ts_flush_pending(m, st);
  if (rc) return rc;
  if (!m->g.esdf) {
    TS_CUDA(cudaMalloc(&m->g.esdf, (size_t)m->g.max_blocks * TS_B3 * 4));
    TS_CUDA(cudaMemset(m->g.esdf, 0, (size_t)m->g.max_blocks * TS_B3 * 4));
    TS_CUDA(cudaMalloc(&m->esdf_aux, ((size_t)m->g.max_blocks * 28 + ES_MAX_SWEEPS + 2) * 4));
  }
  EsAux aux;
  aux.nbr = (int*)m->esdf_aux;
  aux.epoch = aux.nbr + (size_t)m->g.max_blocks * 27;
  aux.changed = aux.epoch + m->g.max_blocks;
  const float gamma = (float)m->cfg.voxel_scale;     // dense_esdf.py:40
  const float far_v = (float)m->cfg.max_ray_length;  // dense_esdf.py:324
  TS_CUDA(cudaMemsetAsync(aux.changed, 0, (ES_MAX_SWEEPS + 2) * 4, st));
  k_esdf_init<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, gamma, far_v);
  TS_LAUNCH_CHECK(m);
  int sweeps = 0;
  for (;;) {
    for (int q = 0; q < ES_GROUP && sweeps < ES_MAX_SWEEPS; q++) {
      sweeps++;
      k_esdf_sweep<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, sweeps, gamma, m->in.vs);
      TS_LAUNCH_CHECK(m);
    }
    int changed = 0;
    TS_CUDA(cudaMemcpyAsync(&changed, aux.changed + sweeps, 4, cudaMemcpyDeviceToHost, st));
    TS_CUDA(cudaStreamSynchronize(st));
    if (!changed || sweeps >= ES_MAX_SWEEPS) break;
  }
  if (n_sweeps_out) {  // sweeps the wave needed: the first sweep that changed nothing ends it
    static int h_changed[ES_MAX_SWEEPS + 2];
    TS_CUDA(cudaMemcpy(h_changed, aux.changed, (size_t)(sweeps + 1) * 4, cudaMemcpyDeviceToHost));
    int k = 1;
    while (k <= sweeps && h_changed[k]) k++;
    *n_sweeps_out = k;
  }
  return TSLAM_OK;
}

extern "C" int tslam_esdf_gather(tslam_tsdf_t* m, int32_t submap, int64_t cap, int32_t* idx, float* esdf, int64_t* n_out,
                                 void* stream) {
  if (!m || !n_out) return TSLAM_E_INVALID;
  if (!m->g.esdf) { ts_set_error("tslam_esdf_update has not run"); return TSLAM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* ctr = (unsigned long long*)(m->scratch_i + 8);
  TS_CUDA(cudaMemsetAsync(ctr, 0, 8, st));
  k_esdf_gather<<<m->sm_count * 4, 256, 0, st>>>(m->g, submap, cap, idx, esdf, ctr);
  TS_LAUNCH_CHECK(m);
  unsigned long long v = 0;
  TS_CUDA(cudaMemcpyAsync(&v, ctr, 8, cudaMemcpyDeviceToHost, st));
  TS_CUDA(cudaStreamSynchronize(st));
  *n_out = (int64_t)v;
  if ((int64_t)v > cap) { ts_set_error("esdf gather: capacity"); return TSLAM_E_CAPACITY; }
  return TSLAM_OK;
}
