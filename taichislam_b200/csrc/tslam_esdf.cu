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
//
// INCREMENTAL update (round 2; the reference seeds its wave from the voxels the frame just touched,
// dense_esdf.py:309-330, and has a raise and a lower queue, :260-299).  State kept between updates: the field and a
// class byte per voxel.  An update after new commits:
//   1. detect   (blocks committed since the last update): voxels whose class changed, that became observed, or whose
//               fixed-band value moved are flagged CHANGED;
//   2. raise    wave: a non-fixed voxel whose value is EXACTLY a flagged neighbour's old value +- the edge cost may
//               have been derived from it -> flagged SUSPECT, and so on to closure (f32 equality is exact: at a
//               fixed point every relaxed value IS some neighbour's value +- cost);
//   3. re-init  changed voxels take their new class / seed, suspects go back to +-max_ray_length;
//   4. lower    the sweep kernel above, started from the re-initialised blocks only.
// Unflagged voxels keep values that are still attainable (their support chain avoids every changed voxel), i.e. upper
// bounds; min/+ relaxation from there converges to the SAME least fixed point as a full recompute - tested
// bit-for-bit (tests/test_gpu_mc_octo_esdf.py::test_esdf_incremental_equals_full_512).
#include <cstring>
#include "tslam_internal.cuh"

#define ES_T 18
#define ES_T3 (ES_T * ES_T * ES_T)
#define ES_GROUP 6       // sweeps enqueued per host synchronisation
#define ES_MAX_SWEEPS 4096

enum { ES_UNOBS = 0, ES_FIXED = 1, ES_POS = 2, ES_NEG = 3, ES_INERT = 4 };  // INERT: observed, TSDF is NaN
#define ES_CH 0x40   // class-byte flags of an incremental update: CHANGED (detect) / SUSPECT (raise wave)
#define ES_SU 0x80
#define ES_CLS 0x0F

struct EsAux {
  int* nbr;      // [max_blocks*27] neighbour block indices (-1 = absent)
  int* epoch;    // [max_blocks] last sweep in which the block changed
  int* changed;  // [ES_MAX_SWEEPS+2] changed[k] != 0 <=> sweep k changed something
  unsigned char* cls;  // [max_blocks*4096] class of the voxel at the last update (+ flags during an update)
  int* repoch;   // [max_blocks] raise wave: last raise sweep in which the block gained suspects
  int* rchanged; // [ES_MAX_SWEEPS+2]
  int* counts;   // [4] changed voxels, suspect voxels, re-initialised blocks, (spare)
};
__device__ __forceinline__ unsigned char es_class(bool obs, float t, float gamma) {
  if (!obs) return ES_UNOBS;
  return (fabsf(t) < gamma) ? ES_FIXED : (t > 0.0f ? ES_POS : (t < 0.0f ? ES_NEG : ES_INERT));
}

__device__ __forceinline__ int es_tidx(int lx, int ly, int lz) { return ((lx + 1) * ES_T + (ly + 1)) * ES_T + (lz + 1); }
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
      const bool ob = g.obs[base + v] != 0;
      const float t = g.tw[base + v].x;
      if (ob) e = (fabsf(t) < gamma) ? t : es_sgn(t) * far_v;
      g.esdf[base + v] = e;
      aux.cls[base + v] = es_class(ob, t, gamma);
    }
    if (threadIdx.x == 0) g.esdf_dirty[b] = 0;
  }
}

// ---- incremental update ------------------------------------------------------------------------------------------
// neighbour table of every block of the submap + epochs reset; the voxel state is left alone
__global__ void __launch_bounds__(256) k_esdf_prepare(TsGrid g, EsAux aux, int submap) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * 256) >> 5;
  for (int b = w; b < nb; b += nw) {
    int s, bx, by, bz;
    ts_unpack_key(g.block_key[b], s, bx, by, bz);
    if (lane == 0) { aux.epoch[b] = (s == submap) ? -1 : -1000000; aux.repoch[b] = -1; }
    if (s != submap) continue;
    if (lane < 27) {
      const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
      aux.nbr[b * 27 + lane] = ts_find(g, ts_pack_key(s, bx + dx, by + dy, bz + dz));
    }
  }
}

// 1. detect: blocks committed since the last update
__global__ void __launch_bounds__(256) k_esdf_detect(TsGrid g, EsAux aux, int submap, float gamma) {
  __shared__ int s_any;
  const int nb = min(*g.n_blocks, g.max_blocks);
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    if (!g.esdf_dirty[b] || aux.epoch[b] < -1) continue;  // (uniform per CTA)
    __syncthreads();
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    const size_t base = (size_t)b * TS_B3;
    int mine = 0;
    for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
      const unsigned char oc = aux.cls[base + v] & ES_CLS;
      const float t = g.tw[base + v].x;
      const unsigned char nc = es_class(g.obs[base + v] != 0, t, gamma);
      const bool ch = (nc != oc) || (nc == ES_FIXED && g.esdf[base + v] != t);
      if (ch) { aux.cls[base + v] = oc | ES_CH; mine++; }
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) { atomicAdd(&s_any, mine); }
    __syncthreads();
    if (threadIdx.x == 0) {
      g.esdf_dirty[b] = 0;
      if (s_any) { aux.repoch[b] = 0; atomicAdd(&aux.counts[0], s_any); }
    }
  }
}

struct EsRTile {
  float e[ES_T3];
  unsigned char cls[ES_T3];
  int nbr[27];
  int active, changed, any_write;
};

// 2. raise wave: flag the voxels that may have been derived from a flagged neighbour
__global__ void __launch_bounds__(256) k_esdf_raise(TsGrid g, EsAux aux, int sweep, float vs) {
  __shared__ EsRTile tile;
  if (sweep > 1 && aux.rchanged[sweep - 1] == 0) return;
  const int nb = min(*g.n_blocks, g.max_blocks);
  const float d1 = vs, d2 = sqrtf(2.0f) * vs, d3 = sqrtf(3.0f) * vs;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) { tile.active = 0; tile.any_write = 0; }
    __syncthreads();
    if (aux.epoch[b] < -1) continue;
    if (threadIdx.x < 27) {
      const int n = aux.nbr[b * 27 + threadIdx.x];
      tile.nbr[threadIdx.x] = n;
      if (n >= 0 && *(volatile int*)&aux.repoch[n] >= sweep - 1) tile.active = 1;
    }
    __syncthreads();
    if (!tile.active) continue;
    for (int c = threadIdx.x; c < ES_T3; c += blockDim.x) {
      const int lz = c % ES_T - 1, ly = (c / ES_T) % ES_T - 1, lx = c / (ES_T * ES_T) - 1;
      const int nbk = tile.nbr[((((lx + 16) >> 4) * 3 + ((ly + 16) >> 4)) * 3) + ((lz + 16) >> 4)];
      float ev = 0.0f;
      unsigned char cl = ES_UNOBS;
      if (nbk >= 0) {
        const size_t off = (size_t)nbk * TS_B3 + ((((lx & 15) << 4) | (ly & 15)) << 4 | (lz & 15));
        cl = *(volatile unsigned char*)&aux.cls[off];
        ev = g.esdf[off];  // old values: nothing writes the field during the raise wave
      }
      tile.e[c] = ev;
      tile.cls[c] = cl;
    }
    __syncthreads();
    for (int iter = 0; iter < 64; ++iter) {
      if (threadIdx.x == 0) tile.changed = 0;
      __syncthreads();
      bool mine = false;
      for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
        const int c0 = es_tidx(v >> 8, (v >> 4) & 15, v & 15);
        const unsigned char cl = tile.cls[c0];
        if (cl != ES_POS && cl != ES_NEG) continue;  // flagged, fixed, unobserved: nothing to decide
        const float cur = tile.e[c0];
        bool hit = false;
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dz = -1; dz <= 1; dz++) {
              const int nz_ = (dx != 0) + (dy != 0) + (dz != 0);
              if (nz_ == 0) continue;
              const int cn = c0 + (dx * ES_T + dy) * ES_T + dz;
              if (!(tile.cls[cn] & (ES_CH | ES_SU))) continue;
              const float dis = nz_ == 1 ? d1 : (nz_ == 2 ? d2 : d3);
              const float he = tile.e[cn];
              hit |= (cl == ES_POS) ? (he + dis == cur) : (he - dis == cur);
            }
        if (hit) { tile.cls[c0] = cl | ES_SU; mine = true; }
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
      for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
        const unsigned char cl = tile.cls[es_tidx(v >> 8, (v >> 4) & 15, v & 15)];
        if (cl & ES_SU) aux.cls[base + v] = cl;
      }
      if (threadIdx.x == 0) {
        __threadfence();
        aux.repoch[b] = sweep;
        aux.rchanged[sweep] = 1;
      }
    }
  }
}

// 3. re-initialise flagged voxels; their blocks start the lower wave (epoch 0)
__global__ void __launch_bounds__(256) k_esdf_reinit(TsGrid g, EsAux aux, float gamma, float far_v) {
  __shared__ int s_n[2];
  const int nb = min(*g.n_blocks, g.max_blocks);
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    if (aux.repoch[b] < 0) continue;  // no flagged voxel in this block (uniform per CTA)
    __syncthreads();
    if (threadIdx.x == 0) { s_n[0] = 0; s_n[1] = 0; }
    __syncthreads();
    const size_t base = (size_t)b * TS_B3;
    int nsu = 0;
    for (int v = threadIdx.x; v < TS_B3; v += blockDim.x) {
      const unsigned char c = aux.cls[base + v];
      if (!(c & (ES_CH | ES_SU))) continue;
      const float t = g.tw[base + v].x;
      const bool ob = g.obs[base + v] != 0;
      const unsigned char nc = es_class(ob, t, gamma);
      float e = 0.0f;
      if (ob) e = (nc == ES_FIXED) ? t : es_sgn(t) * far_v;
      g.esdf[base + v] = e;
      aux.cls[base + v] = nc;
      nsu += (c & ES_SU) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) nsu += __shfl_xor_sync(0xffffffffu, nsu, o);
    if ((threadIdx.x & 31) == 0 && nsu) atomicAdd(&s_n[0], nsu);
    __syncthreads();
    if (threadIdx.x == 0) {
      aux.epoch[b] = 0;
      if (s_n[0]) atomicAdd(&aux.counts[1], s_n[0]);
      atomicAdd(&aux.counts[2], 1);
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

static int es_alloc(tslam_tsdf* m) {
  if (m->g.esdf) return TSLAM_OK;
  const size_t nb = (size_t)m->g.max_blocks;
  TS_CUDA(cudaMalloc(&m->g.esdf, nb * TS_B3 * 4));
  TS_CUDA(cudaMemset(m->g.esdf, 0, nb * TS_B3 * 4));
  TS_CUDA(cudaMalloc(&m->esdf_aux, (nb * 29 + 2 * (ES_MAX_SWEEPS + 2) + 8) * 4 + nb * TS_B3));
  TS_CUDA(cudaMemset(m->esdf_aux, 0, (nb * 29 + 2 * (ES_MAX_SWEEPS + 2) + 8) * 4 + nb * TS_B3));
  m->esdf_full_needed = 1;
  return TSLAM_OK;
}
static EsAux es_aux(tslam_tsdf* m) {
  EsAux aux;
  const size_t nb = (size_t)m->g.max_blocks;
  aux.nbr = (int*)m->esdf_aux;
  aux.epoch = aux.nbr + nb * 27;
  aux.repoch = aux.epoch + nb;
  aux.changed = aux.repoch + nb;
  aux.rchanged = aux.changed + (ES_MAX_SWEEPS + 2);
  aux.counts = aux.rchanged + (ES_MAX_SWEEPS + 2);
  aux.cls = (unsigned char*)(aux.counts + 8);
  return aux;
}

// One wave (raise or lower) to convergence: sweeps are enqueued in groups, a sweep whose predecessor changed nothing
// returns at once; the host reads ONE flag per group.  Returns the number of sweeps that did something + 1.
template <class Launch>
static int es_wave(tslam_tsdf* m, cudaStream_t st, int* flags, int group, Launch launch, int* n_out) {
  int sweeps = 0;
  for (;;) {
    for (int q = 0; q < group && sweeps < ES_MAX_SWEEPS; q++) { sweeps++; launch(sweeps); TS_LAUNCH_CHECK(m); }
    int changed = 0;
    TS_CUDA(cudaMemcpyAsync(&changed, flags + sweeps, 4, cudaMemcpyDeviceToHost, st));
    TS_CUDA(cudaStreamSynchronize(st));
    if (!changed || sweeps >= ES_MAX_SWEEPS) break;
  }
  if (n_out) {
    static int h_changed[ES_MAX_SWEEPS + 2];
    TS_CUDA(cudaMemcpy(h_changed, flags, (size_t)(sweeps + 1) * 4, cudaMemcpyDeviceToHost));
    int k = 1;
    while (k <= sweeps && h_changed[k]) k++;
    *n_out = k;
  }
  return TSLAM_OK;
}

// mode 0 = incremental when the state allows it (same submap as the last update, no reset / bulk load in between),
// 1 = full recompute.  stats4 (HOST, may be NULL): lower sweeps, raise sweeps, changed voxels, suspect voxels.
extern "C" int tslam_esdf_update2(tslam_tsdf_t* m, int32_t submap, int32_t mode, int32_t* stats4, void* stream) {
  if (!m) return TSLAM_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ts_flush_pending(m, st);
  if (rc) return rc;
  rc = es_alloc(m);
  if (rc) return rc;
  EsAux aux = es_aux(m);
  const float gamma = (float)m->cfg.voxel_scale;     // dense_esdf.py:40
  const float far_v = (float)m->cfg.max_ray_length;  // dense_esdf.py:324
  const bool full = mode == 1 || m->esdf_full_needed || m->esdf_submap != submap;
  int h_counts[4] = {0, 0, 0, 0}, n_lower = 0, n_raise = 0;
  TS_CUDA(cudaMemsetAsync(aux.changed, 0, (2 * (ES_MAX_SWEEPS + 2) + 8) * 4, st));  // changed, rchanged, counts
  if (full) {
    k_esdf_init<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, gamma, far_v);
    TS_LAUNCH_CHECK(m);
    rc = es_wave(m, st, aux.changed, ES_GROUP, [&](int k) { k_esdf_sweep<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, k, gamma, m->in.vs); }, &n_lower);
    if (rc) return rc;
  } else {
    k_esdf_prepare<<<m->sm_count * 2, 256, 0, st>>>(m->g, aux, submap);
    TS_LAUNCH_CHECK(m);
    k_esdf_detect<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, gamma);
    TS_LAUNCH_CHECK(m);
    rc = es_wave(m, st, aux.rchanged, 4, [&](int k) { k_esdf_raise<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, k, m->in.vs); }, &n_raise);
    if (rc) return rc;
    k_esdf_reinit<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, gamma, far_v);
    TS_LAUNCH_CHECK(m);
    rc = es_wave(m, st, aux.changed, 4, [&](int k) { k_esdf_sweep<<<m->sm_count * 4, 256, 0, st>>>(m->g, aux, submap, k, gamma, m->in.vs); }, &n_lower);
    if (rc) return rc;
    TS_CUDA(cudaMemcpy(h_counts, aux.counts, 16, cudaMemcpyDeviceToHost));
  }
  m->esdf_full_needed = 0;
  m->esdf_submap = submap;
  if (stats4) { stats4[0] = n_lower; stats4[1] = full ? -1 : n_raise; stats4[2] = h_counts[0]; stats4[3] = h_counts[1]; }
  return TSLAM_OK;
}

extern "C" int tslam_esdf_update(tslam_tsdf_t* m, int32_t submap, int32_t* n_sweeps_out, void* stream) {
  int32_t st4[4] = {0, 0, 0, 0};
  int rc = tslam_esdf_update2(m, submap, 0, st4, stream);
  if (n_sweeps_out) *n_sweeps_out = st4[0];
  return rc;
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
