// libtslam.so - block-binned ray march (round 2).  sm_100a.
//
// Replaces the one-thread-per-ray march of round 1 (k_raymarch: one 8-byte global reduction per sample, bound by the
// L2 reduction rate of ~150 G/s whatever the kernel did) for untextured maps.  process_new_pcl
// (dense_tsdf.py:236-270) becomes
//
//   k_ray_setup    one thread per ray (= bucket, read sequentially from the dense per-ray array): exact bucket mean
//                  -> unit direction, length, weight, step count n (:240-251), occupy flag (:248), a 32-byte record;
//   k_seg_walk     one lane per ray: a 3-D DDA cuts the ray into SEGMENTS - runs of consecutive steps that stay in
//                  one 16^3 voxel block - activates the blocks on the way, appends the segments to per-CTA lists and
//                  counts them per (block, length class) with warp-aggregated reductions;
//   k_seg_class    per touched block: class counts -> offsets inside the block's run, list of touched blocks;
//   k_seg_scan     exclusive scan over the touched blocks -> run offsets + work items (<= MR_CHUNK segments of one
//                  block each, full chunks first);
//   k_seg_place    moves every listed segment to its place: a block's run is ordered by segment length;
//   k_march_blocks persistent CTAs take work items.  The block's pending sums live in SHARED MEMORY as exact 64-bit
//                  fixed-point words (no-return ATOMS.ADD: 2.4-6 T lane-atomics/s measured against 0.15-0.19 T/s
//                  for global reductions, tools/ubench/atoms_bench.cu).  A warp walks 32 segments of equal length
//                  in lock step, one per lane, ray parameters in registers.  Each sample's voxel index comes from
//                  ONE fused multiply-add per axis in block-local voxel units and is accepted only when its
//                  fractional part is at least near_eps away from .5 - otherwise (~2e-3 of the samples) the lane
//                  recomputes round((u*j*vs + T)/vs) exactly as the reference states it (:253-254), so voxel
//                  indices stay bit-exact; a sample whose exact voxel is not in the item's block takes a global
//                  reduction.  At the end of the item every touched voxel is flushed with ONE coalesced
//                  REDG.ADD.F32x2 into the block's `acc` plane (what k_commit folds into TSDF/W, :264-267);
//   k_march_generic  the exact one-reduction-per-sample march for what the binned path does not take: segments
//                  next to the volume boundary, rays with more than 65535 steps, workspace overflow.
//
// Sample value: ds = L - j*vs (the reference forms |P - x| * sign((P - x).m), :258-260, which equals it up to the
// f32 rounding of x: <= 3e-6 m on a 25 m map, tolerance 1e-4).
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "tslam_internal.cuh"

#define FULL 0xffffffffu
#define MS_THREADS 256
// shared accumulator layout: word index = x*MB_SX + y*MB_SY + z (skewed strides: consecutive samples of a ray
// spread over the banks whatever its direction - 3.6-way average conflict vs 5.7 for x<<8|y<<4|z)
#define MB_SX 277
#define MB_SY 17
#define MB_WORDS (15 * MB_SX + 15 * MB_SY + 16)  // 4426
#define MB_MIN_WQ 16384                          // a ray whose weight is < 2^14 fixed-point units (relative rounding error > 3e-5)
                                                // is applied with f32 global reductions instead
#define MB_SMEM (4 * MB_WORDS * 4)
#define SEG_CLS 16                               // length classes per block: (count - 1) >> 1
#define SEG_REP 4                                // counter replicas per (block, class), picked by the walk CTA: the hot
                                                // blocks' counters would otherwise serialise in one L2 slice
#define SEG_KEYS (SEG_CLS * SEG_REP)

// ---------------------------------------------------------------------------
// K2a: ray set-up (process_new_pcl :240-251), first stage of the walk kernel: one bucket record -> unit direction,
// length, weight, step count n, occupy flag (:248) and the 32-byte ray record the march kernels read.  The record and
// its index entry are handed back zeroed (PCLroot.deactivate_all() / new_pcl_count = 0, :163, :270).
// ---------------------------------------------------------------------------
struct RaySetup { float ux, uy, uz, tx, ty, tz; int n, s; uint32_t f; bool wide; };
__device__ __forceinline__ RaySetup ray_setup(const TsBatch& batch, const TsIntrin& in, const TsGrid& g, unsigned long long* btab, TsBucket* buckets,
                                              unsigned long long* bidx, uint32_t bucket_shift, uint32_t id, uint32_t r, const TsMarchWs& w,
                                              unsigned int& my_rays, unsigned int& my_fmax) {
  const float vs = in.vs;
  const uint32_t f = id >> bucket_shift;
  uint4* q = reinterpret_cast<uint4*>(&buckets[id]);  // key sx | sy sz | sd cnt cr | cg cb pad
  const uint4 q0 = q[0], q1 = q[1], q2 = q[2];
  const unsigned long long ie = (unsigned long long)q0.x | ((unsigned long long)q0.y << 32);
  if (ie) bidx[ie - 1ull] = 0ull;
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  q[0] = z4; q[1] = z4; q[2] = z4; q[3] = z4;
  const long long sx = (long long)((unsigned long long)q0.z | ((unsigned long long)q0.w << 32));
  const long long sy = (long long)((unsigned long long)q1.x | ((unsigned long long)q1.y << 32));
  const long long sz = (long long)((unsigned long long)q1.z | ((unsigned long long)q1.w << 32));
  const long long sd = (long long)((unsigned long long)q2.x | ((unsigned long long)q2.y << 32));
  const int cnt = (int)q2.z;
  const TsFrame& fr = batch.f[f];
  RaySetup o;
  o.ux = o.uy = o.uz = o.tx = o.ty = o.tz = 0.0f;
  o.n = 0; o.s = fr.submap; o.f = f; o.wide = false;
  float L = 0.0f, wgt = 0.0f;
  if (cnt > 0) {  // :240
    my_rays++;
    const double den = (double)cnt * FIXQ_D;
    const float mx = (float)((double)sx / den);  // pos_s2p = sum/c (:243), exact mean
    const float my = (float)((double)sy / den);
    const float mz = (float)((double)sz / den);
    const float zc = (float)((double)sd / den);  // z = new_pcl_z/c (:247)
    L = sqrtf((mx * mx + my * my) + mz * mz);  // :244
    if (L > 0.0f) {
      o.ux = mx / L; o.uy = my / L; o.uz = mz / L;  // :245
      const float Px = mx + fr.T[0], Py = my + fr.T[1], Pz = mz + fr.T[2];  // :246
      // occupy[sxyz_to_ijk(pos_p)] = 1 (:248)
      const int oi = iroundf(Px / vs), oj = iroundf(Py / vs), ok = iroundf(Pz / vs);
      if (ts_in_bounds(g, oi, oj, ok)) {
        const int bx = oi >> TS_BSHIFT, by = oj >> TS_BSHIFT, bz = ok >> TS_BSHIFT;
        const int blk = rm_lookup(g, btab, ts_pack_key(o.s, bx, by, bz), bx, by, bz);  // touched blocks are listed even when only `occupy` changed
        if (blk >= 0) g.occ[(size_t)blk * TS_B3 + ts_voxel_off(oi, oj, ok)] = 1;
      }
      o.n = (int)fminf(L / vs + (float)in.internal_voxels, in.max_steps);  // :249-251
      wgt = 1.0f / (zc * zc);  // w_x_p(d>=0, z) (:216-225, :262)
      o.tx = (float)((double)fr.T[0] / (double)vs);
      o.ty = (float)((double)fr.T[1] / (double)vs);
      o.tz = (float)((double)fr.T[2] / (double)vs);
      // fixed-point scale of the launch: every |w| and |w*ds| must fit 31 bits (|ds| <= max(L, n*vs - L) + vs)
      const float dmax = fmaxf(L, (float)o.n * vs - L) + vs;
      const unsigned fb = __float_as_uint(fminf(fmaxf(wgt, wgt * dmax), 3.0e38f));  // non-negative floats order like their bits
      if (fb > my_fmax) my_fmax = fb;
      o.wide = o.n > 65535;
    } else {
      L = 0.0f;
    }
  }
  float4* dst = reinterpret_cast<float4*>(&w.rays[r]);
  dst[0] = make_float4(o.ux, o.uy, o.uz, L);
  dst[1] = make_float4(o.tx, o.ty, o.tz, wgt);
  o.n = max(0, min(o.n, 65535));
  w.aux[r] = ((uint32_t)o.n << 16) | (f << 8) | (o.wide ? TS_AUX_WIDE : 0u);
  return o;
}

// ---------------------------------------------------------------------------
// K2b: segment walk: 3-D DDA over the 16^3 blocks in voxel-index space, one lane per ray, warp-uniform loop (one
// block crossing per lane and iteration).  Segments are HINTS - every sample is re-checked exactly by the march
// kernel - so the crossing parameters may be off by one step.
// ---------------------------------------------------------------------------
// 0 = far outside the volume (every sample is out of bounds), 1 = block inside the block range, 2 = outside but
// adjacent to it (a rounding-boundary sample may still be in bounds: exact path)
__device__ __forceinline__ int block_class(const TsGrid& g, int bx, int by, int bz) {
  const int b0 = (-g.hN) >> TS_BSHIFT, b1 = (g.N - g.hN - 1) >> TS_BSHIFT;
  const int z0 = (-g.hNz) >> TS_BSHIFT, z1 = (g.Nz - g.hNz - 1) >> TS_BSHIFT;
  if (bx >= b0 && bx <= b1 && by >= b0 && by <= b1 && bz >= z0 && bz <= z1) return 1;
  if (bx >= b0 - 1 && bx <= b1 + 1 && by >= b0 - 1 && by <= b1 + 1 && bz >= z0 - 1 && bz <= z1 + 1) return 2;
  return 0;
}
__device__ __forceinline__ void gen_append(const TsMarchWs& w, int* err, uint32_t ray, int j, int cnt) {
  const int p = atomicAdd(&w.ctl->n_gen, 1);
  if ((uint32_t)p < w.gen_cap) w.gen[p] = TsSeg{ray, ((uint32_t)j << 12) | (uint32_t)cnt};
  else atomicOr(err, TS_ERR_RAYLIST_FULL);
}
__device__ __forceinline__ void dda_axis(float u, float t, int& b, float& jc, float& dj, int& sg) {
  b = __float2int_rn(u + t) >> TS_BSHIFT;  // block of the first sample (j = 1)
  if (u > 0.0f) { jc = ((float)(16 * b + 16) - 0.5f - t) / u; dj = 16.0f / u; sg = 1; }
  else if (u < 0.0f) { jc = ((float)(16 * b) - 0.5f - t) / u; dj = -16.0f / u; sg = -1; }
  else { jc = 3.0e38f; dj = 0.0f; sg = 0; }
}

#define WK_CHUNK 4096   // entries per chunk of the per-CTA segment lists
#define WK_MAXCH 1024   // chunks per CTA
struct WalkSmem {
  unsigned long long btab[RM_TAB];  // block coordinates -> pool index (direct mapped, rm_lookup)
  uint32_t chunk[WK_MAXCH];         // first entry of the CTA's k-th chunk in tmp_seg / tmp_key
  int n_chunk, cur, ovf, pad;
};

__global__ void __launch_bounds__(MS_THREADS, 3) k_seg_walk(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsBucket* buckets,
                                                             unsigned long long* bidx, uint32_t bucket_shift, const uint32_t* __restrict__ ray_list,
                                                             const int* __restrict__ n_rays_p, uint32_t ray_cap, TsMarchWs w, TsCounters* ctr) {
  extern __shared__ __align__(16) unsigned char ms_smem[];
  WalkSmem& S = *reinterpret_cast<WalkSmem*>(ms_smem);
  for (int e = threadIdx.x; e < RM_TAB; e += MS_THREADS) S.btab[e] = TS_EMPTY;
  if (threadIdx.x == 0) { S.n_chunk = 0; S.cur = 0; S.ovf = 0; }
  __syncthreads();
  const uint32_t nme = min((uint32_t)*n_rays_p, ray_cap);
  const uint32_t lane = threadIdx.x & 31u;
  const int max_seg_ray = (int)(in.max_steps * 0.125f) + 8;  // block crossings of the longest ray (+ slack)
  unsigned int my_oob = 0, my_rays = 0, my_fmax = 0;
  for (uint32_t base = blockIdx.x * MS_THREADS; base < nme; base += gridDim.x * MS_THREADS) {
    // room for this round's segments in the CTA's list
    if (threadIdx.x == 0 && !S.ovf) {
      const long long need = (long long)S.cur + (long long)MS_THREADS * max_seg_ray;
      while ((long long)S.n_chunk * WK_CHUNK < need) {
        uint32_t c0 = 0;
        bool okc = S.n_chunk < WK_MAXCH;
        if (okc) { c0 = atomicAdd(&w.ctl->tmp_cursor, (unsigned)WK_CHUNK); okc = c0 + WK_CHUNK <= w.seg_cap; }
        if (!okc) { S.ovf = 1; break; }  // workspace exhausted: the rest of this CTA's rays take the generic path
        S.chunk[S.n_chunk++] = c0;
      }
    }
    __syncthreads();
    const bool ovf = S.ovf != 0;
    const uint32_t r = base + threadIdx.x;
    float ux = 0.f, uy = 0.f, uz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
    int n = 0, s = 0;
    uint32_t f = 0;
    bool wide = false;
    if (r < nme) {
      const RaySetup rs = ray_setup(batch, in, g, S.btab, buckets, bidx, bucket_shift, ray_list[r], r, w, my_rays, my_fmax);
      ux = rs.ux; uy = rs.uy; uz = rs.uz; tx = rs.tx; ty = rs.ty; tz = rs.tz;
      n = rs.n; f = rs.f; s = rs.s; wide = rs.wide;
    }
    int bx, by, bz, sx, sy, sz;
    float jx, jy, jz, dx, dy, dz;
    dda_axis(ux, tx, bx, jx, dx, sx);
    dda_axis(uy, ty, by, jy, dy, sy);
    dda_axis(uz, tz, bz, jz, dz, sz);
    // whole ray inside the block range (both end points are; the cells in between lie between them)?
    bool inside = false;
    if (n > 0) {
      const float nf = (float)n;
      const int ex = __float2int_rn(__fmaf_rn(ux, nf, tx)) >> TS_BSHIFT, ey = __float2int_rn(__fmaf_rn(uy, nf, ty)) >> TS_BSHIFT,
                ez = __float2int_rn(__fmaf_rn(uz, nf, tz)) >> TS_BSHIFT;
      inside = block_class(g, bx, by, bz) == 1 && block_class(g, ex, ey, ez) == 1;
    }
    const uint32_t rj = (f << 24);  // frame of the ray rides in the top byte of the segment word
    int j = 1;
    const int max_iter = 3 * (n / 16 + 4);
    for (int it = 0;; ++it) {
      const bool act = j <= n && it < max_iter;
      if (!__any_sync(FULL, act)) break;
      int cnt = 0, jl = 0;
      if (act) {
        const float je = fminf(fminf(jx, jy), jz);
        jl = min((int)ceilf(je) - 1, n);  // last step strictly before the crossing
        if (jl >= j) cnt = min(jl - j + 1, 32);
      }
      int key = -1;  // block * SEG_CLS + length class
      if (cnt > 0) {
        const int cls = inside ? 1 : block_class(g, bx, by, bz);
        if (wide || cls == 2 || ovf) gen_append(w, g.err, r, j, cnt);
        else if (cls == 0) my_oob += (unsigned)cnt;
        else {
          const int blk = rm_lookup(g, S.btab, ts_pack_key(s, bx, by, bz), bx, by, bz);  // activates + marks the block dirty
          if (blk >= 0) key = (blk * SEG_CLS + ((cnt - 1) >> 1)) * SEG_REP + (int)(blockIdx.x & (SEG_REP - 1));  // (< 0: pool exhausted)
        }
      }
      const bool rec = key >= 0;
      const unsigned mrec = __ballot_sync(FULL, rec);
      if (mrec) {
        const int leader = __ffs(mrec) - 1;
        int p0 = 0;
        if ((int)lane == leader) p0 = atomicAdd(&S.cur, __popc(mrec));
        p0 = __shfl_sync(FULL, p0, leader);
        if (rec) {
          const int p = p0 + __popc(mrec & ((1u << lane) - 1u));
          const uint32_t gp = S.chunk[p >> 12] + (uint32_t)(p & (WK_CHUNK - 1));
          w.tmp_seg[gp] = TsSeg{r, rj | ((uint32_t)j << 8) | (uint32_t)cnt};
          w.tmp_key[gp] = (uint32_t)key;
          const unsigned grp = __match_any_sync(mrec, key);
          if ((int)lane == __ffs(grp) - 1) red_add_u32((unsigned int*)&w.seg_count[key], (unsigned)__popc(grp));
        }
      }
      j += cnt;
      if (act && j > jl) {  // the run up to the crossing is out: step into the next block
        if (jx <= jy && jx <= jz) { bx += sx; jx += dx; }
        else if (jy <= jz) { by += sy; jy += dy; }
        else { bz += sz; jz += dz; }
      }
    }
    // non-finite pose: hand the rest to the exact path in bounded pieces
    while (j <= n) { const int c = min(n - j + 1, 4095); gen_append(w, g.err, r, j, c); j += c; }
  }
  __syncthreads();
  const uint32_t cta = blockIdx.x;
  uint32_t* cch = w.cta_chunk + (size_t)cta * WK_MAXCH;
  for (int k = threadIdx.x; k < S.n_chunk; k += MS_THREADS) cch[k] = S.chunk[k];
  if (threadIdx.x == 0) w.cta_n[cta] = S.cur;
  for (int o = 16; o > 0; o >>= 1) {
    my_oob += __shfl_xor_sync(FULL, my_oob, o);
    my_rays += __shfl_xor_sync(FULL, my_rays, o);
    my_fmax = max(my_fmax, __shfl_xor_sync(FULL, my_fmax, o));
  }
  if (lane == 0) {
    if (my_oob) atomicAdd(&ctr->n_oob, (unsigned long long)my_oob);
    if (my_rays) atomicAdd(&ctr->n_rays, (unsigned long long)my_rays);
    if (my_fmax) atomicMax(&w.ctl->fmax_bits, my_fmax);
  }
}

// ---------------------------------------------------------------------------
// K2c: per block with segments: class counts -> offsets inside the block's run (seg_rel), list of touched blocks.
// seg_count is handed on zeroed: k_seg_place uses it as the fill cursor.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_seg_class(TsGrid g, TsMarchWs w) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  const uint32_t lane = threadIdx.x & 31u;
  const int stride = gridDim.x * 256;
  for (int b0 = blockIdx.x * 256; b0 < nb; b0 += stride) {
    const int b = b0 + (int)threadIdx.x;
    unsigned tot = 0, cost = 0;  // segments / samples (upper estimate: class q holds segments of 2q+1 or 2q+2 steps)
    if (b < nb && g.dirty_flag[b]) {
      uint4* c4 = reinterpret_cast<uint4*>(&w.seg_count[(size_t)b * SEG_KEYS]);
      uint4* r4 = reinterpret_cast<uint4*>(&w.seg_rel[(size_t)b * SEG_KEYS]);
#pragma unroll 4
      for (int q = 0; q < SEG_KEYS / 4; q++) {
        const uint4 c = c4[q];
        if ((c.x | c.y | c.z | c.w) == 0u) { continue; }
        r4[q] = make_uint4(tot, tot + c.x, tot + c.x + c.y, tot + c.x + c.y + c.z);
        tot += c.x + c.y + c.z + c.w;
        cost += (c.x + c.y + c.z + c.w) * (unsigned)(2 * (q / (SEG_REP / 4)) + 2);
        c4[q] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    const unsigned mt = __ballot_sync(FULL, tot > 0u);
    if (mt) {
      const int leader = __ffs(mt) - 1;
      int p0 = 0;
      if ((int)lane == leader) p0 = atomicAdd(&w.ctl->n_touched, __popc(mt));
      p0 = __shfl_sync(FULL, p0, leader);
      if (tot > 0u) {
        const int p = p0 + __popc(mt & ((1u << lane) - 1u));
        w.touched[p] = b;
        w.blk_total[p] = tot;
        w.blk_cost[p] = cost;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// K2d: scan over the touched blocks -> run offsets + work items (one CTA).  The items are listed LONGEST FIRST
// (estimated samples, 32 bins): the persistent march kernel takes them in list order, so its tail is made of the
// cheapest items (measured before: SMs idle 18 % of the kernel's time with the items in block order).
// ---------------------------------------------------------------------------
#define SC_BINS 32
__device__ __forceinline__ int item_bin(uint32_t cnt, uint32_t cost, uint32_t q, uint32_t nq) {
  // a block's run is ordered by segment length: chunk q of nq holds longer segments than chunk q-1 (ramp .5 .. 1.5)
  const uint32_t nseg = min((uint32_t)MR_CHUNK, cnt - q * MR_CHUNK);
  const float est = (float)nseg * ((float)cost / (float)cnt) * (0.5f + ((float)q + 0.5f) / (float)nq);
  return min(SC_BINS - 1, (int)(est * (1.0f / 4096.0f)));
}
__global__ void __launch_bounds__(1024) k_seg_scan(TsMarchWs w, TsCounters* ctr) {
  __shared__ uint32_t s_a[32];
  __shared__ uint32_t s_run;
  __shared__ unsigned int s_hist[SC_BINS], s_base[SC_BINS];
  __shared__ int s_ovf;
  const int nt = w.ctl->n_touched;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  if (threadIdx.x < SC_BINS) s_hist[threadIdx.x] = 0u;
  __syncthreads();
  // pass A: run offsets (exclusive scan of the blocks' segment counts), then the histogram of the items' estimated cost
  for (int base = 0; base < nt; base += 1024) {
    const int i = base + threadIdx.x;
    int blk = -1;
    uint32_t cnt = 0;
    if (i < nt) { blk = w.touched[i]; cnt = w.blk_total[i]; }
    uint32_t a = cnt;  // inclusive warp scan
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(FULL, a, o);
      if (lane >= o) a += ta;
    }
    if (lane == 31) s_a[wid] = a;
    __syncthreads();
    if (wid == 0) {
      uint32_t ta = s_a[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(FULL, ta, o);
        if (lane >= o) ta += ua;
      }
      s_a[lane] = ta;  // inclusive over warps
    }
    __syncthreads();
    if (i < nt) w.seg_off[blk] = s_run + (wid ? s_a[wid - 1] : 0u) + a - cnt;
    __syncthreads();
    if (threadIdx.x == 0) s_run += s_a[31];
    __syncthreads();
  }
  for (int i = wid; i < nt; i += 32) {  // one warp per block, one lane per chunk: the busiest blocks have dozens of chunks
    const uint32_t cnt = w.blk_total[i], cost = w.blk_cost[i];
    const uint32_t nq = (cnt + MR_CHUNK - 1) / MR_CHUNK;
    for (uint32_t q = lane; q < nq; q += 32) atomicAdd(&s_hist[item_bin(cnt, cost, q, nq)], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot_items = 0;
    for (int b = SC_BINS - 1; b >= 0; b--) { s_base[b] = tot_items; tot_items += s_hist[b]; s_hist[b] = 0u; }  // s_hist becomes the fill cursor
    const uint32_t tot = s_run;
    const bool ovf = tot_items > w.item_cap;  // (segments cannot exceed seg_cap: they come out of the same pool)
    s_ovf = ovf ? 1 : 0;
    w.ctl->total_segs = tot;
    w.ctl->overflow = ovf ? 1 : 0;
    w.ctl->n_full = 0;
    w.ctl->n_items = ovf ? 0 : (int)tot_items;
    // fixed-point scale: the largest k with fmax * 2^k < 2^30, clamped to [0, 48]
    const float fm = __uint_as_float(w.ctl->fmax_bits);
    int k = 30;
    if (fm > 0.0f) {
      int ex;
      frexpf(fm, &ex);  // fm = mant * 2^ex, mant in [0.5, 1)  ->  fm < 2^ex
      k = 30 - ex;
    }
    w.ctl->scale_k = max(0, min(48, k));
    atomicAdd(&ctr->n_segs, (unsigned long long)tot);
    atomicAdd(&ctr->n_items, (unsigned long long)(ovf ? 0u : tot_items));
  }
  __syncthreads();
  if (s_ovf) return;
  // pass B: the items, longest bin first
  for (int i = wid; i < nt; i += 32) {
    const int blk = w.touched[i];
    const uint32_t cnt = w.blk_total[i], cost = w.blk_cost[i], ex_seg = w.seg_off[blk];
    const uint32_t nq = (cnt + MR_CHUNK - 1) / MR_CHUNK;
    for (uint32_t q = lane; q < nq; q += 32) {
      const int b = item_bin(cnt, cost, q, nq);
      const uint32_t pos = s_base[b] + atomicAdd(&s_hist[b], 1u);
      w.items[pos] = TsItem{blk, ex_seg + q * MR_CHUNK, min((uint32_t)MR_CHUNK, cnt - q * MR_CHUNK), 0u};
    }
  }
}

// ---------------------------------------------------------------------------
// K2e: placement.  Same grid as k_seg_walk: CTA i moves the segments of its list into the blocks' runs:
// seg_off[block] + seg_rel[block, class] + cursor (warp-aggregated global atomic on seg_count, zero on entry).
// ---------------------------------------------------------------------------
#define PL_SPLIT 2
#define PL_THREADS 1024  // the kernel waits on one returning atomic per (warp, key): more warps per list hide it
__global__ void __launch_bounds__(PL_THREADS) k_seg_place(TsMarchWs w) {
  __shared__ uint32_t s_chunk[WK_MAXCH];
  const uint32_t cta = blockIdx.x / PL_SPLIT, part = blockIdx.x % PL_SPLIT;  // PL_SPLIT CTAs share one walk list
  const int total = w.cta_n[cta];
  if (total == 0) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t* cch = w.cta_chunk + (size_t)cta * WK_MAXCH;
  const int nch = (total + WK_CHUNK - 1) / WK_CHUNK;
  for (int k = threadIdx.x; k < nch; k += PL_THREADS) s_chunk[k] = cch[k];
  __syncthreads();
  const bool overflow = w.ctl->overflow != 0;  // work list did not fit: every segment goes to the generic path
  for (int p0 = (int)(part * PL_THREADS + (threadIdx.x & ~31)); p0 < total; p0 += PL_SPLIT * PL_THREADS) {
    const int p = p0 + (int)lane;
    const bool have = p < total;
    TsSeg sg = TsSeg{0u, 0u};
    int key = -1;
    if (have) {
      const uint32_t gp = s_chunk[p >> 12] + (uint32_t)(p & (WK_CHUNK - 1));
      sg = w.tmp_seg[gp];
      key = (int)w.tmp_key[gp];
    }
    if (overflow) {
      if (have) {
        const int pg = atomicAdd(&w.ctl->n_gen, 1);
        if ((uint32_t)pg < w.gen_cap) w.gen[pg] = TsSeg{sg.ray, (((sg.jc >> 8) & 0xFFFFu) << 12) | (sg.jc & 255u)};
      }
      continue;
    }
    const unsigned mh = __ballot_sync(FULL, have);
    if (have) {
      const unsigned grp = __match_any_sync(mh, key);
      const int leader = __ffs(grp) - 1;
      int c0 = 0;
      if ((int)lane == leader) c0 = atomicAdd(&w.seg_count[key], __popc(grp));
      c0 = __shfl_sync(grp, c0, leader) + __popc(grp & ((1u << lane) - 1u));
      w.seg[w.seg_off[key / SEG_KEYS] + w.seg_rel[key] + (uint32_t)c0] = sg;
    }
  }
}

// ---------------------------------------------------------------------------
// exact sample index, as the reference states it (:253-254)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void exact_index(float ux, float uy, float uz, const float* T, int j, float vs, float rvs, int& xi, int& yi, int& zi) {
  const float jf = (float)j;
  const float x = (ux * jf) * vs + T[0], y = (uy * jf) * vs + T[1], z = (uz * jf) * vs + T[2];
  xi = iroundf(div_vs(x, vs, rvs));
  yi = iroundf(div_vs(y, vs, rvs));
  zi = iroundf(div_vs(z, vs, rvs));
}

// one sample through the global path: bounds, block lookup (activating), one reduction
__device__ __forceinline__ void global_sample(const TsGrid& g, int s, int xi, int yi, int zi, float a, float wgt, unsigned int& upd,
                                              unsigned int& oob) {
  if ((unsigned)(xi + g.hN) < (unsigned)g.N && (unsigned)(yi + g.hN) < (unsigned)g.N && (unsigned)(zi + g.hNz) < (unsigned)g.Nz) {
    const int blk = ts_get_or_alloc_cached(g, ts_pack_key(s, xi >> TS_BSHIFT, yi >> TS_BSHIFT, zi >> TS_BSHIFT));
    if (blk >= 0) {
      ts_mark_dirty(g, blk);
      red_add_f32x2(&g.acc[(size_t)blk * TS_B3 + ts_voxel_off(xi, yi, zi)], a, wgt);
      upd++;
    }
  } else {
    oob++;
  }
}

// ---------------------------------------------------------------------------
// K2f: the march proper.  Persistent CTAs, one work item (block, <= MR_CHUNK segments) at a time.
//
// A block's run is ordered by segment length (k_seg_place), so the 32 segments a warp takes - ONE PER LANE, ray
// parameters in registers - have (nearly) equal length and are walked in lock step.  Lane l starts at step
// (7*l mod cnt) of its segment and wraps around: neighbouring rays are nearly parallel, so with every lane at a
// different depth the 32 shared-memory atomics of a step hit 32 different voxels (same-step marching puts 4-16 rays
// into one voxel half way to the surface: serialised atomics).
//
// Shared accumulators per voxel: two 64-bit sums (A = sum w*ds, B = sum w, units 2^-scale_k) kept as
//   lo  = sum x      mod 2^32
//   hi' = sum (x >> 16)            (arithmetic shift)
// Both are plain no-return ATOMS.ADD - no carry logic, no dependent read of the old value.  With fewer than 2^16
// samples per voxel and item (MR_CHUNK segments x <= 2 samples of a ray in one voxel) the true sum S is the unique
// value with S = lo (mod 2^32) in [hi' * 2^16, hi' * 2^16 + 2^32):  S = hi' * 2^16 + ((lo - (hi' << 16)) mod 2^32).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double mb_sum(unsigned int lo, int hi) {
  return (double)((long long)hi * 65536ll + (long long)(unsigned int)(lo - ((unsigned int)hi << 16)));
}

template <bool VERIFY, int MB_THREADS, int MB_MINB>
__global__ void __launch_bounds__(MB_THREADS, MB_MINB) k_march_blocks(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsMarchWs w,
                                                                 TsCounters* ctr) {
  extern __shared__ __align__(16) unsigned int mb_smem[];
  unsigned int* const a_lo = mb_smem;
  int* const a_hi = (int*)(mb_smem + MB_WORDS);
  unsigned int* const b_lo = mb_smem + 2 * MB_WORDS;
  int* const b_hi = (int*)(mb_smem + 3 * MB_WORDS);
  __shared__ int s_item;
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const float vs = in.vs, rvs = in.rvs, eps_hi = 0.5f - w.near_eps;
  const int scale_k = w.ctl->scale_k;
  const float fix = __uint_as_float((unsigned)(127 + scale_k) << 23);   // 2^k
  const double unfix = 1.0 / (double)fix;
  unsigned int my_upd = 0, my_oob = 0, my_slow = 0, my_fb = 0, my_bad = 0;
  for (int e = threadIdx.x; e < 4 * MB_WORDS; e += MB_THREADS) mb_smem[e] = 0u;
  const int n_items = w.ctl->n_items;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(&w.ctl->item_cursor, 1);
    __syncthreads();
    const int it = s_item;
    if (it >= n_items) break;
    const TsItem item = w.items[it];  // longest first (k_seg_scan)
    int sm, kx, ky, kz;
    ts_unpack_key(g.block_key[item.blk], sm, kx, ky, kz);
    const int ox = kx << TS_BSHIFT, oy = ky << TS_BSHIFT, oz = kz << TS_BSHIFT;
    const float fox = (float)ox, foy = (float)oy, foz = (float)oz;
    // blocks cut by the volume boundary (N not a multiple of 16) need the per-sample bounds test
    const bool partial = ox < -g.hN || ox + TS_B > g.N - g.hN || oy < -g.hN || oy + TS_B > g.N - g.hN || oz < -g.hNz || oz + TS_B > g.Nz - g.hNz;
    const int nseg = (int)item.nseg;
    const TsSeg* const segs = w.seg + item.seg0;
    // ---- 32 segments per warp, one per lane; the next batch's records are fetched while this one is marched ----
    TsSeg nx = TsSeg{0u, 0u};
    float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
    {
      const int i0 = (int)wid * 32 + (int)lane;
      if (i0 < nseg) {
        nx = segs[i0];
        const float4* src = reinterpret_cast<const float4*>(&w.rays[nx.ray]);
        na = __ldg(src); nb = __ldg(src + 1);
      }
    }
    constexpr int MB_WARPS = MB_THREADS / 32;
    for (int sb = (int)wid * 32; sb < nseg; sb += MB_WARPS * 32) {
      const TsSeg sg = nx;
      const float4 a = na, b = nb;
      {
        const int i1 = sb + MB_WARPS * 32 + (int)lane;
        nx = TsSeg{0u, 0u};
        if (i1 < nseg) {
          nx = segs[i1];
          const float4* src = reinterpret_cast<const float4*>(&w.rays[nx.ray]);
          na = __ldg(src); nb = __ldg(src + 1);
        }
      }
      const int cnt = (int)(sg.jc & 255u);  // 0 for lanes past the end of the item
      const int j0 = (int)((sg.jc >> 8) & 0xFFFFu);
      const uint32_t frame = sg.jc >> 24;
      const int maxc = __reduce_max_sync(FULL, cnt);
      const float jf = (float)j0;
      const float ux = a.x, uy = a.y, uz = a.z;
      // first sample of the segment in block-local voxel units: one rounding at magnitude <= ~20
      const float x0 = __fmaf_rn(a.x, jf, b.x - fox), y0 = __fmaf_rn(a.y, jf, b.y - foy), z0 = __fmaf_rn(a.z, jf, b.z - foz);
      const float L0 = __fmaf_rn(-jf, vs, a.w);
      const float wf = b.w * fix;  // w in fixed-point units (exact: power of two)
      const int wq = __float2int_rn(wf);
      const int wqh = wq >> 16;
      // weight far below the launch's largest (f32 reductions keep its relative precision), blocks cut by the volume
      // boundary, VERIFY: every sample of the segment takes the exact path below
      const bool all_slow = wq < MB_MIN_WQ || partial || VERIFY;
      const float cntf = (float)cnt;
      int k0 = 0;
      if (cnt > 1) k0 = (int)(lane * 7u) % cnt;  // staggered start
      float kf = (float)k0;
      my_upd += (unsigned)cnt;
      // ---- fast loop: branch-free.  Iteration i of a lane handles step (k0 + i) mod cnt of its segment.  A sample
      // whose rounding is not certain (|fraction - .5| < near_eps) or whose voxel is not in this block only sets bit i
      // of slow_mask; those are redone exactly after the loop.
      unsigned slow_mask = 0u;
      if (all_slow) slow_mask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
      else {
        for (int i = 0; i < maxc; ++i) {
          const float gx = __fmaf_rn(ux, kf, x0), gy = __fmaf_rn(uy, kf, y0), gz = __fmaf_rn(uz, kf, z0);
          // round to nearest + distance of the fraction from .5 (magic-number rounding: |g| < 2^22): the low mantissa
          // bits of t hold the block-local voxel coordinate
          const float tx = gx + 12582912.0f, ty = gy + 12582912.0f, tz = gz + 12582912.0f;
          const float rx = gx - (tx - 12582912.0f), ry = gy - (ty - 12582912.0f), rz = gz - (tz - 12582912.0f);
          const unsigned qx = __float_as_uint(tx), qy = __float_as_uint(ty), qz = __float_as_uint(tz);
          const bool inb = ((qx | qy | qz) - 0x4B400000u) <= 15u;  // all three coordinates in [0, 16)
          const bool sure = fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz)) <= eps_hi;
          const bool pred = i < cnt;
          const float avq = __fmaf_rn(-kf, vs, L0) * wf;  // w * ds, ds = L - j*vs (:258-264), in fixed-point units
          if (pred && inb && sure) {
            const unsigned e = qx * (unsigned)MB_SX + qy * (unsigned)MB_SY + qz - 0x4B400000u * (unsigned)(MB_SX + MB_SY + 1);
            const int xq = __float2int_rn(avq);
            atomicAdd(&a_lo[e], (unsigned int)xq);
            atomicAdd(&a_hi[e], xq >> 16);
            atomicAdd(&b_lo[e], (unsigned int)wq);
            atomicAdd(&b_hi[e], wqh);
          } else if (pred) {
            slow_mask |= 1u << i;
          }
          kf += 1.0f;
          if (kf >= cntf) kf = 0.0f;
        }
      }
      // ---- exact path (:253-254) for the flagged samples; a sample that really lies in another block, is out of
      // bounds or carries a tiny weight goes through the global path
      if (__any_sync(FULL, slow_mask != 0u)) {
        const bool tiny = wq < MB_MIN_WQ;
        while (slow_mask) {
          const int i = __ffs(slow_mask) - 1;
          slow_mask &= slow_mask - 1u;
          int k = k0 + i;
          if (k >= cnt) k -= cnt;
          const float kk = (float)k;
          const float avq = __fmaf_rn(-kk, vs, L0) * wf;
          int xi, yi, zi;
          exact_index(ux, uy, uz, batch.f[frame].T, j0 + k, vs, rvs, xi, yi, zi);
          const int lx = xi - ox, ly = yi - oy, lz = zi - oz;
          if (VERIFY || !all_slow) {
            const float gx = __fmaf_rn(ux, kk, x0), gy = __fmaf_rn(uy, kk, y0), gz = __fmaf_rn(uz, kk, z0);
            const float tx = gx + 12582912.0f, ty = gy + 12582912.0f, tz = gz + 12582912.0f;
            const float rx = gx - (tx - 12582912.0f), ry = gy - (ty - 12582912.0f), rz = gz - (tz - 12582912.0f);
            const bool near = fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz)) > eps_hi;
            if (near) my_slow++;
            if (VERIFY && !near && (lx != __float_as_int(tx) - 0x4B400000 || ly != __float_as_int(ty) - 0x4B400000 || lz != __float_as_int(tz) - 0x4B400000)) my_bad++;
          }
          bool local = !tiny && ((unsigned)(lx | ly | lz)) <= 15u;
          if (!local) my_fb++;
          if (local && partial &&
              !((unsigned)(xi + g.hN) < (unsigned)g.N && (unsigned)(yi + g.hN) < (unsigned)g.N && (unsigned)(zi + g.hNz) < (unsigned)g.Nz)) {
            local = false;  // global_sample counts it as out of bounds
          }
          if (local) {
            const int e = lx * MB_SX + ly * MB_SY + lz;
            const int xq = __float2int_rn(avq);
            atomicAdd(&a_lo[e], (unsigned int)xq);
            atomicAdd(&a_hi[e], xq >> 16);
            atomicAdd(&b_lo[e], (unsigned int)wq);
            atomicAdd(&b_hi[e], wqh);
          } else {
            my_upd--;
            global_sample(g, sm, xi, yi, zi, (float)((double)avq * unfix), (float)((double)wf * unfix), my_upd, my_oob);
          }
        }
      }
    }
    __syncthreads();
    // ---- flush: one coalesced reduction per touched voxel, accumulators back to zero ----
    float2* const acc = g.acc + (size_t)item.blk * TS_B3;
    for (int v = threadIdx.x; v < TS_B3; v += MB_THREADS) {
      const int e = (v >> 8) * MB_SX + ((v >> 4) & 15) * MB_SY + (v & 15);
      const unsigned int blo = b_lo[e];
      const int bhi = b_hi[e];
      if (blo == 0u && bhi == 0) continue;  // sum w > 0 for every touched voxel
      const unsigned int alo = a_lo[e];
      const int ahi = a_hi[e];
      a_lo[e] = 0u; a_hi[e] = 0; b_lo[e] = 0u; b_hi[e] = 0;
      red_add_f32x2(&acc[v], (float)(mb_sum(alo, ahi) * unfix), (float)(mb_sum(blo, bhi) * unfix));
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_upd += __shfl_xor_sync(FULL, my_upd, o);
    my_oob += __shfl_xor_sync(FULL, my_oob, o);
    my_slow += __shfl_xor_sync(FULL, my_slow, o);
    my_fb += __shfl_xor_sync(FULL, my_fb, o);
    my_bad += __shfl_xor_sync(FULL, my_bad, o);
  }
  // one global reduction per CTA and counter (same-address atomics serialise in L2)
  __shared__ unsigned int s_stat[5];
  if (threadIdx.x < 5) s_stat[threadIdx.x] = 0u;
  __syncthreads();
  if (lane == 0) {
    if (my_upd) atomicAdd(&s_stat[0], my_upd);
    if (my_oob) atomicAdd(&s_stat[1], my_oob);
    if (my_slow) atomicAdd(&s_stat[2], my_slow);
    if (my_fb) atomicAdd(&s_stat[3], my_fb);
    if (my_bad) atomicAdd(&s_stat[4], my_bad);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_stat[0]) atomicAdd(&ctr->n_updates, (unsigned long long)s_stat[0]);
    if (s_stat[1]) atomicAdd(&ctr->n_oob, (unsigned long long)s_stat[1]);
    if (s_stat[2]) atomicAdd(&ctr->n_slow, (unsigned long long)s_stat[2]);
    if (s_stat[3]) atomicAdd(&ctr->n_fallback, (unsigned long long)s_stat[3]);
    if (s_stat[4]) atomicAdd(&ctr->n_verify_bad, (unsigned long long)s_stat[4]);
  }
}

// ---------------------------------------------------------------------------
// K2g: generic path - exact index and one global reduction per sample
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_march_generic(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsMarchWs w, TsCounters* ctr) {
  const int n_gen = min(w.ctl->n_gen, (int)w.gen_cap);
  unsigned int my_upd = 0, my_oob = 0, my_gen = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_gen; i += gridDim.x * blockDim.x) {
    const TsSeg sg = w.gen[i];
    const float4* src = reinterpret_cast<const float4*>(&w.rays[sg.ray]);
    const float4 a = src[0], b = src[1];
    const TsFrame& fr = batch.f[(w.aux[sg.ray] >> 8) & 255u];
    const int j0 = (int)(sg.jc >> 12), cnt = (int)(sg.jc & 4095u);
    for (int q = 0; q < cnt; q++) {
      const int j = j0 + q;
      int xi, yi, zi;
      exact_index(a.x, a.y, a.z, fr.T, j, in.vs, in.rvs, xi, yi, zi);
      const float ds = __fmaf_rn(-(float)j, in.vs, a.w);
      global_sample(g, fr.submap, xi, yi, zi, b.w * ds, b.w, my_upd, my_oob);
      my_gen++;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_upd += __shfl_xor_sync(FULL, my_upd, o);
    my_oob += __shfl_xor_sync(FULL, my_oob, o);
    my_gen += __shfl_xor_sync(FULL, my_gen, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (my_upd) atomicAdd(&ctr->n_updates, (unsigned long long)my_upd);
    if (my_oob) atomicAdd(&ctr->n_oob, (unsigned long long)my_oob);
    if (my_gen) atomicAdd(&ctr->n_generic, (unsigned long long)my_gen);
  }
}

// end of a launch: fill cursors back to zero, control block cleared
__global__ void __launch_bounds__(256) k_march_reset(TsMarchWs w) {
  const int nt = w.ctl->n_touched;
  for (int q = blockIdx.x * 256 + threadIdx.x; q < nt * (SEG_KEYS / 4); q += gridDim.x * 256)
    reinterpret_cast<uint4*>(w.seg_count + (size_t)w.touched[q / (SEG_KEYS / 4)] * SEG_KEYS)[q % (SEG_KEYS / 4)] = make_uint4(0u, 0u, 0u, 0u);
  // the last CTA to finish clears the control block
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&w.ctl->ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < sizeof(TsMarchCtl) / 4) reinterpret_cast<int*>(w.ctl)[threadIdx.x] = 0;
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
int ts_march_alloc(tslam_tsdf* m) {
  TsMarchWs& w = m->mw;
  const size_t nr = m->ray_list_cap;
  const size_t nb = (size_t)m->g.max_blocks;
  TS_CUDA(cudaMalloc(&w.rays, nr * sizeof(TsRay)));
  TS_CUDA(cudaMalloc(&w.aux, nr * 4));
  size_t sc = nr * 16;  // ~8 segments per 90-step ray on the bench stream, 15 on 8 m rays
  if (sc < (8u << 20)) sc = 8u << 20;
  if (sc > (128u << 20)) sc = 128u << 20;
  const char* e = getenv("TSLAM_SEG_CAP");  // tests: force the overflow path
  if (e && atoll(e) > 0) sc = (size_t)atoll(e);
  w.seg_cap = (uint32_t)sc;
  TS_CUDA(cudaMalloc(&w.seg, sc * sizeof(TsSeg)));
  TS_CUDA(cudaMalloc(&w.tmp_seg, sc * sizeof(TsSeg)));
  TS_CUDA(cudaMalloc(&w.tmp_key, sc * 4));
  TS_CUDA(cudaMalloc(&w.seg_count, nb * SEG_KEYS * 4));
  TS_CUDA(cudaMemset(w.seg_count, 0, nb * SEG_KEYS * 4));
  TS_CUDA(cudaMalloc(&w.seg_rel, nb * SEG_KEYS * 4));
  TS_CUDA(cudaMalloc(&w.seg_off, nb * 4));
  TS_CUDA(cudaMalloc(&w.touched, nb * 4));
  TS_CUDA(cudaMalloc(&w.blk_total, nb * 4));
  TS_CUDA(cudaMalloc(&w.blk_cost, nb * 4));
  w.item_cap = (uint32_t)(sc / MR_CHUNK + nb + 16);
  TS_CUDA(cudaMalloc(&w.items, (size_t)w.item_cap * sizeof(TsItem)));
  w.gen_cap = (uint32_t)(nr > (1u << 20) ? nr : (1u << 20));
  TS_CUDA(cudaMalloc(&w.gen, (size_t)w.gen_cap * sizeof(TsSeg)));
  w.walk_x_max = m->sm_count * 3;  // one segment list per CTA (5 CTAs per SM measured slower: the per-(block, class) counters are the limit)
  const size_t ncta = (size_t)w.walk_x_max + TSLAM_MAX_BATCH + 8;
  TS_CUDA(cudaMalloc(&w.cta_n, ncta * 4));
  TS_CUDA(cudaMemset(w.cta_n, 0, ncta * 4));
  TS_CUDA(cudaMalloc(&w.cta_chunk, ncta * WK_MAXCH * 4));
  TS_CUDA(cudaFuncSetAttribute(k_seg_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WalkSmem)));
  TS_CUDA(cudaMalloc(&w.ctl, sizeof(TsMarchCtl)));
  TS_CUDA(cudaMemset(w.ctl, 0, sizeof(TsMarchCtl)));
  // |fast index - exact index| bound in voxel units (DESIGN.md "exact indices from a one-FMA fast path"):
  // 2^-24 * (4*max_steps + 2*|T/vs|) for the reference's own rounding chain + 2^-24 * (|T/vs| + max_steps + 100)
  // for ours, with |T/vs| <= N; doubled.
  const double ms = m->cfg.max_ray_length / m->cfg.voxel_scale;
  const double nmax = m->cfg.N > m->cfg.Nz ? m->cfg.N : m->cfg.Nz;
  double eps = 2.0 * 5.9604644775390625e-8 * (5.0 * ms + 3.0 * nmax + 100.0);
  if (eps < 1e-5) eps = 1e-5;
  if (eps > 0.25) eps = 1.0;  // absurd geometry: every sample takes the exact path
  const char* ee = getenv("TSLAM_NEAR_EPS");
  if (ee) eps = atof(ee);
  w.near_eps = (float)eps;
  TS_CUDA(cudaFuncSetAttribute(k_march_blocks<false, 320, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  TS_CUDA(cudaFuncSetAttribute(k_march_blocks<true, 320, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  return TSLAM_OK;
}

void ts_march_free(tslam_tsdf* m) {
  TsMarchWs& w = m->mw;
  cudaFree(w.rays); cudaFree(w.aux); cudaFree(w.seg); cudaFree(w.tmp_seg); cudaFree(w.tmp_key); cudaFree(w.seg_count); cudaFree(w.seg_rel);
  cudaFree(w.seg_off); cudaFree(w.touched); cudaFree(w.blk_total); cudaFree(w.blk_cost); cudaFree(w.items); cudaFree(w.gen); cudaFree(w.ctl); cudaFree(w.cta_n);
  cudaFree(w.cta_chunk);
}

// workspace of the binned march (allocated at the first launch that needs it)
int ts_march_setup(tslam_tsdf* m, cudaStream_t st, const TsBatch& batch, uint32_t bucket_shift) {
  (void)st; (void)batch; (void)bucket_shift;
  if (!m->mw.rays) { int rca = ts_march_alloc(m); if (rca) return rca; }
  return TSLAM_OK;
}

// sub_ev (profiling, may be null): 4 events recorded at the start, after walk + class + scan, placement and the march kernels
int ts_march_launch(tslam_tsdf* m, cudaStream_t st, const TsBatch& batch, uint32_t bucket_shift, cudaEvent_t* sub_ev) {
  const int sms = m->sm_count;
  const int gx = m->mw.walk_x_max;
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[0], st));
  k_seg_walk<<<gx, MS_THREADS, (int)sizeof(WalkSmem), st>>>(batch, m->in, m->g, m->buckets, m->bidx, bucket_shift, m->ray_list, m->n_rays, m->ray_list_cap,
                                                            m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  k_seg_class<<<sms * 2, 256, 0, st>>>(m->g, m->mw);
  TS_LAUNCH_CHECK(m);
  k_seg_scan<<<1, 1024, 0, st>>>(m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[1], st));
  k_seg_place<<<gx * PL_SPLIT, PL_THREADS, 0, st>>>(m->mw);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[2], st));
  const int grid = (sms - m->rm_reserve) * 3;
  if (m->march_verify) k_march_blocks<true, 320, 3><<<grid, 320, MB_SMEM, st>>>(batch, m->in, m->g, m->mw, m->counters);
  else k_march_blocks<false, 320, 3><<<grid, 320, MB_SMEM, st>>>(batch, m->in, m->g, m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  k_march_generic<<<sms * 2, 256, 0, st>>>(batch, m->in, m->g, m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  k_march_reset<<<32, 256, 0, st>>>(m->mw);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[3], st));
  return TSLAM_OK;
}
