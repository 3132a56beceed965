// libtslam.so - block-binned ray march (round 2).  sm_100a.
//
// Replaces the one-thread-per-ray march of round 1 (k_raymarch, one 8-byte global reduction per sample: bound by
// the L2 reduction rate, ~150 G/s, whatever the kernel did) for untextured maps.  process_new_pcl
// (dense_tsdf.py:236-270) is split into
//
//   k_ray_setup    per ray (= live bucket): exact bucket mean -> unit direction, length, weight, step count n
//                  (:240-251), occupy flag (:248), one 32-byte ray record; then the ray is cut into SEGMENTS -
//                  runs of consecutive steps that stay inside one 16^3 voxel block - and the segments are
//                  counted per block (warp-aggregated histogram; blocks are activated on the way);
//   k_seg_scan     exclusive scan of the per-block counts -> segment offsets and the work list: one item =
//                  <= MR_CHUNK segments of one block;
//   k_seg_fill     walks every ray again and drops its segments into their block's run;
//   k_march_blocks persistent CTAs take work items.  The block's pending sums live in SHARED MEMORY as exact
//                  64-bit fixed-point words (2^-24, lo/hi 32-bit halves, native ATOMS.ADD - measured 2.4-6 T
//                  lane-atomics/s against 0.15-0.19 T/s for global reductions, tools/ubench/atoms_bench.cu);
//                  the samples of a warp's 32 segments are flattened over the lanes (prefix sum), so ray-length
//                  divergence costs nothing; each sample's voxel index comes from ONE fused multiply-add per
//                  axis in block-local voxel units and is accepted only when its fractional part is at least
//                  near_eps away from .5 - otherwise (~1e-3 of the samples) the lane recomputes
//                  round((u*j*vs + T)/vs) exactly as the reference states it (:253-254), so voxel indices stay
//                  bit-exact; a sample whose exact voxel is not in the item's block takes a global reduction.
//                  At the end of the item every touched voxel is flushed with ONE coalesced REDG.ADD.F32x2 into
//                  the block's `acc` plane (what k_commit folds into TSDF/W, :264-267).
//   k_march_generic  the exact one-reduction-per-sample march for what the binned path does not take: segments
//                  next to the volume boundary, rays whose weight leaves the fixed-point range, and everything
//                  when the segment workspace overflows.
//
// Sample value: ds = L - j*vs (the reference forms |P - x| * sign((P - x).m), :258-260, which equals it up to
// f32 rounding of x: <= 3e-6 m on a 25 m map, tolerance 1e-4).
#include <cstdio>
#include <cstdlib>
#include "tslam_internal.cuh"

#define FULL 0xffffffffu
#define MS_THREADS 256
#define MB_THREADS 512
#define MB_WARPS (MB_THREADS / 32)
// shared accumulator layout: word index = x*MB_SX + y*MB_SY + z (skewed strides: consecutive samples of a ray
// spread over the banks whatever its direction - 3.6-way average conflict vs 5.7 for x<<8|y<<4|z)
#define MB_SX 277
#define MB_SY 17
#define MB_WORDS (15 * MB_SX + 15 * MB_SY + 16)  // 4426
#define MB_MIN_WQ 16384                          // a ray whose weight is < 2^14 fixed-point units (relative rounding error > 3e-5)
                                                // is applied with f32 global reductions instead
struct MbWarp {
  unsigned short map[1024];  // flattened sample -> (segment lane << 8 | step within the segment)
  float4 ra[32];             // ux, uy, uz, first-sample x (block-local voxel units)
  float4 rb[32];             // first-sample y, z, L - j0*vs, w
  int j0[32];
  uint32_t rid[32];
};
#define MB_SMEM (4 * MB_WORDS * 4 + MB_WARPS * (int)sizeof(MbWarp))

__device__ __forceinline__ void lohi_add(unsigned int* lo, int* hi, int x) {
  const unsigned int ux = (unsigned int)x;
  const unsigned int old = atomicAdd(lo, ux);
  const int c = (x >> 31) + ((old + ux) < old ? 1 : 0);  // sign extension + carry into the high word
  if (c) atomicAdd(hi, c);
}

// ---------------------------------------------------------------------------
// segment walk: the next run of steps [j, j+cnt) of a ray that stays in one block (approximately: fast index)
// ---------------------------------------------------------------------------
struct RayWalk {
  float ux, uy, uz, tx, ty, tz, iux, iuy, iuz;
};
__device__ __forceinline__ void walk_init(RayWalk& k, const TsRay& r) {
  k.ux = r.ux; k.uy = r.uy; k.uz = r.uz; k.tx = r.tx; k.ty = r.ty; k.tz = r.tz;
  k.iux = 1.0f / r.ux; k.iuy = 1.0f / r.uy; k.iuz = 1.0f / r.uz;
}
__device__ __forceinline__ float walk_axis(float u, float iu, float t, int b) {
  // last parameter j for which round(t + u*j) stays in cells [16b, 16b+15]
  if (u > 0.0f) return ((float)(16 * b + 16) - 0.5f - t) * iu;
  if (u < 0.0f) return ((float)(16 * b) - 0.5f - t) * iu;
  return 3.0e38f;
}
__device__ __forceinline__ int walk_next(const RayWalk& k, int j, int n, int& bx, int& by, int& bz) {
  const float jf = (float)j;
  bx = __float2int_rn(__fmaf_rn(k.ux, jf, k.tx)) >> TS_BSHIFT;
  by = __float2int_rn(__fmaf_rn(k.uy, jf, k.ty)) >> TS_BSHIFT;
  bz = __float2int_rn(__fmaf_rn(k.uz, jf, k.tz)) >> TS_BSHIFT;
  const float je = fminf(fminf(walk_axis(k.ux, k.iux, k.tx, bx), walk_axis(k.uy, k.iuy, k.ty, by)), walk_axis(k.uz, k.iuz, k.tz, bz));
  int jl = (int)ceilf(je) - 1;  // last step strictly before the crossing (a miss by one is caught per sample)
  jl = max(jl, j);
  jl = min(jl, min(n, j + 31));
  return jl - j + 1;
}
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

// FILL = false: count segments per block (histogram + touched list);  FILL = true: write them.
// Must be called by all 32 lanes; `live` masks lanes without a ray.
template <bool FILL>
__device__ __forceinline__ void walk_ray(const TsGrid& g, const TsMarchWs& w, unsigned long long* btab, bool live, const TsRay& ry,
                                         uint32_t rid, int n, int s, bool wide, bool overflow, unsigned int& oob) {
  RayWalk k;
  walk_init(k, ry);
  const uint32_t lane = threadIdx.x & 31u;
  int j = 1;
  while (true) {
    const bool act = live && j <= n;
    if (!__any_sync(FULL, act)) break;
    int bx = 0, by = 0, bz = 0, cnt = 0, blk = -1;
    bool binned = false;
    if (act) {
      cnt = walk_next(k, j, n, bx, by, bz);
      const int cls = block_class(g, bx, by, bz);
      if (wide || cls == 2) {
        if (!FILL) gen_append(w, g.err, rid, j, cnt);
      } else if (cls == 0) {
        if (!FILL) oob += (unsigned)cnt;
      } else {
        blk = rm_lookup(g, btab, ts_pack_key(s, bx, by, bz), bx, by, bz);
        binned = blk >= 0;  // < 0: pool exhausted (error flag raised, samples dropped)
      }
    }
    if (FILL && overflow) {
      if (binned) gen_append(w, g.err, rid, j, cnt);
    } else {
      const unsigned mb = __ballot_sync(FULL, binned);
      if (binned) {
        const unsigned grp = __match_any_sync(mb, blk);
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if ((int)lane == leader) {
          base = atomicAdd(&w.seg_count[blk], __popc(grp));
          if (!FILL && base == 0) w.touched[atomicAdd(&w.ctl->n_touched, 1)] = blk;
        }
        if (FILL) {
          base = __shfl_sync(grp, base, leader);
          const uint32_t pos = w.seg_off[blk] + (uint32_t)base + (uint32_t)__popc(grp & ((1u << lane) - 1u));
          w.seg[pos] = TsSeg{rid, ((uint32_t)j << 8) | (uint32_t)cnt};
        }
      }
    }
    j += cnt;
  }
}

// ---------------------------------------------------------------------------
// K2a: ray set-up (process_new_pcl :240-251) + segment count
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(MS_THREADS) k_ray_setup(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsBucket* buckets,
                                                           uint32_t bucket_shift, const uint32_t* __restrict__ ray_list,
                                                           const int* __restrict__ n_rays_p, uint32_t ray_cap, TsMarchWs w, TsCounters* ctr) {
  __shared__ unsigned long long btab[RM_TAB];
  for (int e = threadIdx.x; e < RM_TAB; e += MS_THREADS) btab[e] = TS_EMPTY;
  __syncthreads();
  const uint32_t n_rays = min((uint32_t)*n_rays_p, ray_cap);
  const float vs = in.vs;
  unsigned int my_rays = 0, my_oob = 0, my_fmax = 0;
  for (uint32_t base = blockIdx.x * MS_THREADS; base < n_rays; base += gridDim.x * MS_THREADS) {
    const uint32_t r = base + threadIdx.x;
    bool live = r < n_rays;
    int cnt = 0, n = 0, s = 0;
    long long sx = 0, sy = 0, sz = 0, sd = 0;
    uint32_t f = 0;
    if (live) {
      const uint32_t id = ray_list[r];
      f = id >> bucket_shift;
      TsBucket* bk = &buckets[id];
      cnt = bk->cnt;
      sx = bk->sx; sy = bk->sy; sz = bk->sz; sd = bk->sd;
      // PCLroot.deactivate_all() / new_pcl_count = 0 (:163, :270): hand the slot back zeroed
      const uint4 z4 = make_uint4(0, 0, 0, 0);
      uint4* q = reinterpret_cast<uint4*>(bk);
      q[0] = z4; q[1] = z4; q[2] = z4; q[3] = z4;
      live = cnt > 0;  // :240
    }
    TsRay ry;
    ry.ux = ry.uy = ry.uz = ry.L = ry.tx = ry.ty = ry.tz = ry.w = 0.0f;
    bool wide = false;
    float fmax = 0.0f;
    if (live) {
      my_rays++;
      const TsFrame& fr = batch.f[f];
      s = fr.submap;
      const double den = (double)cnt * FIXQ_D;
      const float mx = (float)((double)sx / den);  // pos_s2p = sum/c (:243), exact mean
      const float my = (float)((double)sy / den);
      const float mz = (float)((double)sz / den);
      const float zc = (float)((double)sd / den);  // z = new_pcl_z/c (:247)
      const float L = sqrtf((mx * mx + my * my) + mz * mz);  // :244
      if (L > 0.0f) {
        ry.ux = mx / L; ry.uy = my / L; ry.uz = mz / L;  // :245
        ry.L = L;
        const float Px = mx + fr.T[0], Py = my + fr.T[1], Pz = mz + fr.T[2];  // :246
        // occupy[sxyz_to_ijk(pos_p)] = 1 (:248)
        const int oi = iroundf(Px / vs), oj = iroundf(Py / vs), ok = iroundf(Pz / vs);
        if (ts_in_bounds(g, oi, oj, ok)) {
          const int bx = oi >> TS_BSHIFT, by = oj >> TS_BSHIFT, bz = ok >> TS_BSHIFT;
          const int blk = rm_lookup(g, btab, ts_pack_key(s, bx, by, bz), bx, by, bz);  // marks the block dirty
          if (blk >= 0) g.occ[(size_t)blk * TS_B3 + ts_voxel_off(oi, oj, ok)] = 1;
        }
        n = (int)fminf(L / vs + (float)in.internal_voxels, in.max_steps);  // :249-251
        ry.w = 1.0f / (zc * zc);  // w_x_p(d>=0, z) (:216-225, :262)
        ry.tx = (float)((double)fr.T[0] / (double)vs);
        ry.ty = (float)((double)fr.T[1] / (double)vs);
        ry.tz = (float)((double)fr.T[2] / (double)vs);
        // fixed-point scale of the launch: every |w| and |w*ds| must fit 31 bits (|ds| <= max(L, n*vs - L) + vs)
        const float dmax = fmaxf(L, (float)n * vs - L) + vs;
        fmax = fmaxf(ry.w, ry.w * dmax);
        wide = n > 65535;
      }
    }
    fmax = fminf(fmax, 3.0e38f);
    {
      unsigned fb = __float_as_uint(fmax);  // non-negative floats order like their bit patterns
      fb = __reduce_max_sync(FULL, fb);
      if ((threadIdx.x & 31) == 0 && fb > my_fmax) my_fmax = fb;
    }
    if (r < n_rays) {
      float4* dst = reinterpret_cast<float4*>(&w.rays[r]);
      dst[0] = make_float4(ry.ux, ry.uy, ry.uz, ry.L);
      dst[1] = make_float4(ry.tx, ry.ty, ry.tz, ry.w);
      w.aux[r] = ((uint32_t)min(n, 65535) << 16) | (f << 8) | (wide ? TS_AUX_WIDE : 0u);
    }
    if (wide && n > 65535) n = 65535;  // (never with sane configurations; keeps the aux word consistent)
    walk_ray<false>(g, w, btab, live && n > 0, ry, r, n, s, wide, false, my_oob);
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_rays += __shfl_xor_sync(FULL, my_rays, o);
    my_oob += __shfl_xor_sync(FULL, my_oob, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (my_rays) atomicAdd(&ctr->n_rays, (unsigned long long)my_rays);
    if (my_oob) atomicAdd(&ctr->n_oob, (unsigned long long)my_oob);
    if (my_fmax) atomicMax(&w.ctl->fmax_bits, my_fmax);
  }
}

// ---------------------------------------------------------------------------
// K2b: scan of the per-block segment counts -> offsets + work items (one CTA)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_seg_scan(TsMarchWs w, TsCounters* ctr) {
  __shared__ uint32_t s_a[32], s_b[32];
  __shared__ uint32_t s_run[2];
  const int nt = w.ctl->n_touched;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run[0] = 0; s_run[1] = 0; }
  __syncthreads();
  for (int base = 0; base < nt; base += 1024) {
    const int i = base + threadIdx.x;
    int blk = -1;
    uint32_t cnt = 0;
    if (i < nt) { blk = w.touched[i]; cnt = (uint32_t)w.seg_count[blk]; }
    const uint32_t items = (cnt + MR_CHUNK - 1) / MR_CHUNK;
    uint32_t a = cnt, b = items;  // inclusive warp scans
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(FULL, a, o), tb = __shfl_up_sync(FULL, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (lane == 31) { s_a[wid] = a; s_b[wid] = b; }
    __syncthreads();
    if (wid == 0) {
      uint32_t ta = s_a[lane], tb = s_b[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(FULL, ta, o), ub = __shfl_up_sync(FULL, tb, o);
        if (lane >= o) { ta += ua; tb += ub; }
      }
      s_a[lane] = ta; s_b[lane] = tb;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t run_seg = s_run[0], run_item = s_run[1];
    const uint32_t ex_seg = run_seg + (wid ? s_a[wid - 1] : 0u) + a - cnt;
    const uint32_t ex_item = run_item + (wid ? s_b[wid - 1] : 0u) + b - items;
    if (i < nt) {
      w.seg_off[blk] = ex_seg;
      w.seg_count[blk] = 0;  // becomes the fill cursor
      for (uint32_t q = 0; q < items; q++) {
        const uint32_t p = ex_item + q;
        if (p < w.item_cap) w.items[p] = TsItem{blk, ex_seg + q * MR_CHUNK, min((uint32_t)MR_CHUNK, cnt - q * MR_CHUNK), 0u};
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_run[0] = run_seg + s_a[31]; s_run[1] = run_item + s_b[31]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t tot = s_run[0], ni = s_run[1];
    const bool ovf = tot > w.seg_cap || ni > w.item_cap;
    w.ctl->total_segs = tot;
    w.ctl->overflow = ovf ? 1 : 0;
    w.ctl->n_items = ovf ? 0 : (int)ni;
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
    atomicAdd(&ctr->n_items, (unsigned long long)(ovf ? 0u : ni));
  }
}

// ---------------------------------------------------------------------------
// K2c: segment fill
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(MS_THREADS) k_seg_fill(const __grid_constant__ TsBatch batch, TsGrid g, const int* __restrict__ n_rays_p,
                                                          uint32_t ray_cap, TsMarchWs w) {
  __shared__ unsigned long long btab[RM_TAB];
  for (int e = threadIdx.x; e < RM_TAB; e += MS_THREADS) btab[e] = TS_EMPTY;
  __syncthreads();
  const uint32_t n_rays = min((uint32_t)*n_rays_p, ray_cap);
  const bool overflow = w.ctl->overflow != 0;
  unsigned int dummy = 0;
  for (uint32_t base = blockIdx.x * MS_THREADS; base < n_rays; base += gridDim.x * MS_THREADS) {
    const uint32_t r = base + threadIdx.x;
    TsRay ry;
    ry.ux = ry.uy = ry.uz = ry.L = ry.tx = ry.ty = ry.tz = ry.w = 0.0f;
    int n = 0, s = 0;
    bool wide = false;
    if (r < n_rays) {
      const float4* src = reinterpret_cast<const float4*>(&w.rays[r]);
      const float4 a = src[0], b = src[1];
      ry.ux = a.x; ry.uy = a.y; ry.uz = a.z; ry.L = a.w; ry.tx = b.x; ry.ty = b.y; ry.tz = b.z; ry.w = b.w;
      const uint32_t ax = w.aux[r];
      n = (int)(ax >> 16);
      s = batch.f[(ax >> 8) & 255u].submap;
      wide = (ax & TS_AUX_WIDE) != 0;
    }
    walk_ray<true>(g, w, btab, n > 0, ry, r, n, s, wide, overflow, dummy);
  }
}

// ---------------------------------------------------------------------------
// exact sample index, as the reference states it (:253-254)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void exact_index(const TsRay& ry, const float* T, int j, float vs, float rvs, int& xi, int& yi, int& zi) {
  const float jf = (float)j;
  const float x = (ry.ux * jf) * vs + T[0], y = (ry.uy * jf) * vs + T[1], z = (ry.uz * jf) * vs + T[2];
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
// K2d: the march proper.  Persistent CTAs, one work item (block, <= MR_CHUNK segments) at a time.
// ---------------------------------------------------------------------------
template <bool VERIFY>
__global__ void __launch_bounds__(MB_THREADS, 2) k_march_blocks(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsMarchWs w,
                                                                 TsCounters* ctr) {
  extern __shared__ __align__(16) unsigned int mb_smem[];
  unsigned int* const a_lo = mb_smem;
  int* const a_hi = (int*)(mb_smem + MB_WORDS);
  unsigned int* const b_lo = mb_smem + 2 * MB_WORDS;
  int* const b_hi = (int*)(mb_smem + 3 * MB_WORDS);
  MbWarp* const ws = reinterpret_cast<MbWarp*>(mb_smem + 4 * MB_WORDS) + (threadIdx.x >> 5);
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
    const TsItem item = w.items[it];
    int sm, kx, ky, kz;
    ts_unpack_key(g.block_key[item.blk], sm, kx, ky, kz);
    const int ox = kx << TS_BSHIFT, oy = ky << TS_BSHIFT, oz = kz << TS_BSHIFT;
    const float fox = (float)ox, foy = (float)oy, foz = (float)oz;
    // blocks cut by the volume boundary (N not a multiple of 16) need the per-sample bounds test
    const bool partial = ox < -g.hN || ox + TS_B > g.N - g.hN || oy < -g.hN || oy + TS_B > g.N - g.hN || oz < -g.hNz || oz + TS_B > g.Nz - g.hNz;
    for (uint32_t sb = wid * 32u; sb < item.nseg; sb += MB_WARPS * 32u) {
      // ---- stage 32 segments: ray records in block-local coordinates, flattened sample map ----
      int cnt = 0;
      uint32_t rid = 0;
      int j0 = 0;
      if (sb + lane < item.nseg) {
        const TsSeg sg = w.seg[item.seg0 + sb + lane];
        rid = sg.ray; j0 = (int)(sg.jc >> 8); cnt = (int)(sg.jc & 255u);
      }
      int incl = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if ((int)lane >= o) incl += t;
      }
      const int total = __shfl_sync(FULL, incl, 31);
      const int excl = incl - cnt;
      if (cnt > 0) {
        const float4* src = reinterpret_cast<const float4*>(&w.rays[rid]);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        const float jf = (float)j0;
        // first sample of the segment in block-local voxel units: one rounding at magnitude <= ~20
        const float x0 = __fmaf_rn(a.x, jf, b.x - fox), y0 = __fmaf_rn(a.y, jf, b.y - foy), z0 = __fmaf_rn(a.z, jf, b.z - foz);
        ws->ra[lane] = make_float4(a.x, a.y, a.z, x0);
        ws->rb[lane] = make_float4(y0, z0, __fmaf_rn(-jf, vs, a.w), b.w);
        ws->j0[lane] = j0;
        ws->rid[lane] = rid;
        for (int q = 0; q < cnt; q++) ws->map[excl + q] = (unsigned short)((lane << 8) | (unsigned)q);
      }
      __syncwarp();
      // ---- the samples, flattened over the lanes ----
      for (int sidx = (int)lane; sidx - (int)lane < total; sidx += 32) {
        const bool have = sidx < total;
        const unsigned m = have ? ws->map[sidx] : 0u;
        const unsigned sl = m >> 8;
        const float kf = (float)(m & 255u);
        const float4 A = ws->ra[sl], B = ws->rb[sl];
        const float gx = __fmaf_rn(A.x, kf, A.w), gy = __fmaf_rn(A.y, kf, B.x), gz = __fmaf_rn(A.z, kf, B.y);
        // round to nearest + distance of the fraction from .5 (magic-number rounding: |g| < 2^22)
        const float tx = gx + 12582912.0f, ty = gy + 12582912.0f, tz = gz + 12582912.0f;
        int lx = __float_as_int(tx) - 0x4B400000, ly = __float_as_int(ty) - 0x4B400000, lz = __float_as_int(tz) - 0x4B400000;
        const float rx = gx - (tx - 12582912.0f), ry = gy - (ty - 12582912.0f), rz = gz - (tz - 12582912.0f);
        const bool near = fabsf(rx) > eps_hi || fabsf(ry) > eps_hi || fabsf(rz) > eps_hi;
        bool ok = have;
        bool slow = have && (near || (((unsigned)(lx | ly | lz)) > 15u) || VERIFY);
        const float ds = __fmaf_rn(-kf, vs, B.z);  // L - j*vs
        const float av = B.w * ds;                  // w * ds (:264)
        if (__any_sync(FULL, slow)) {
          if (slow) {  // exact index (:253-254); a sample that really lies in another block goes through the global path
            const uint32_t rid2 = ws->rid[sl];
            const int j = ws->j0[sl] + (int)(m & 255u);
            TsRay ryx;
            ryx.ux = A.x; ryx.uy = A.y; ryx.uz = A.z;
            const TsFrame& fr = batch.f[(w.aux[rid2] >> 8) & 255u];
            int xi, yi, zi;
            exact_index(ryx, fr.T, j, vs, rvs, xi, yi, zi);
            const int ex = xi - ox, ey = yi - oy, ez = zi - oz;
            if (VERIFY && !near && (ex != lx || ey != ly || ez != lz)) my_bad++;
            if (near) my_slow++;
            lx = ex; ly = ey; lz = ez;
            if (((unsigned)(lx | ly | lz)) > 15u) {
              ok = false;
              my_fb++;
              global_sample(g, sm, xi, yi, zi, av, B.w, my_upd, my_oob);
            }
          }
        }
        if (ok && partial) {
          const int xi = lx + ox, yi = ly + oy, zi = lz + oz;
          if (!((unsigned)(xi + g.hN) < (unsigned)g.N && (unsigned)(yi + g.hN) < (unsigned)g.N && (unsigned)(zi + g.hNz) < (unsigned)g.Nz)) {
            ok = false;
            my_oob++;
          }
        }
        if (ok) {
          const int wq = __float2int_rn(B.w * fix);
          if (wq >= MB_MIN_WQ) {
            const int e = lx * MB_SX + ly * MB_SY + lz;
            lohi_add(&a_lo[e], &a_hi[e], __float2int_rn(av * fix));
            lohi_add(&b_lo[e], &b_hi[e], wq);
            my_upd++;
          } else {  // weight far below the launch's largest: f32 reduction keeps its relative precision
            my_fb++;
            global_sample(g, sm, lx + ox, ly + oy, lz + oz, av, B.w, my_upd, my_oob);
          }
        }
      }
      __syncwarp();
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
      const float Av = (float)(((double)ahi * 4294967296.0 + (double)alo) * unfix);
      const float Bv = (float)(((double)bhi * 4294967296.0 + (double)blo) * unfix);
      red_add_f32x2(&acc[v], Av, Bv);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my_upd += __shfl_xor_sync(FULL, my_upd, o);
    my_oob += __shfl_xor_sync(FULL, my_oob, o);
    my_slow += __shfl_xor_sync(FULL, my_slow, o);
    my_fb += __shfl_xor_sync(FULL, my_fb, o);
    my_bad += __shfl_xor_sync(FULL, my_bad, o);
  }
  if (lane == 0) {
    if (my_upd) atomicAdd(&ctr->n_updates, (unsigned long long)my_upd);
    if (my_oob) atomicAdd(&ctr->n_oob, (unsigned long long)my_oob);
    if (my_slow) atomicAdd(&ctr->n_slow, (unsigned long long)my_slow);
    if (my_fb) atomicAdd(&ctr->n_fallback, (unsigned long long)my_fb);
    if (my_bad) atomicAdd(&ctr->n_verify_bad, (unsigned long long)my_bad);
  }
}

// ---------------------------------------------------------------------------
// K2e: generic path - exact index and one global reduction per sample
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_march_generic(const __grid_constant__ TsBatch batch, TsIntrin in, TsGrid g, TsMarchWs w, TsCounters* ctr) {
  const int n_gen = min(w.ctl->n_gen, (int)w.gen_cap);
  unsigned int my_upd = 0, my_oob = 0, my_gen = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_gen; i += gridDim.x * blockDim.x) {
    const TsSeg sg = w.gen[i];
    const float4* src = reinterpret_cast<const float4*>(&w.rays[sg.ray]);
    const float4 a = src[0], b = src[1];
    TsRay ry;
    ry.ux = a.x; ry.uy = a.y; ry.uz = a.z; ry.L = a.w; ry.w = b.w;
    const TsFrame& fr = batch.f[(w.aux[sg.ray] >> 8) & 255u];
    const int j0 = (int)(sg.jc >> 12), cnt = (int)(sg.jc & 4095u);
    for (int q = 0; q < cnt; q++) {
      const int j = j0 + q;
      int xi, yi, zi;
      exact_index(ry, fr.T, j, in.vs, in.rvs, xi, yi, zi);
      const float ds = __fmaf_rn(-(float)j, in.vs, ry.L);
      global_sample(g, fr.submap, xi, yi, zi, ry.w * ds, ry.w, my_upd, my_oob);
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

// end of a launch: per-block counters back to zero, control block cleared
__global__ void __launch_bounds__(256) k_march_reset(TsMarchWs w) {
  const int nt = w.ctl->n_touched;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = i; q < nt; q += gridDim.x * blockDim.x) w.seg_count[w.touched[q]] = 0;
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
  size_t sc = nr * 4;
  if (sc < (8u << 20)) sc = 8u << 20;
  if (sc > (64u << 20)) sc = 64u << 20;
  const char* e = getenv("TSLAM_SEG_CAP");  // tests: force the overflow path
  if (e && atoll(e) > 0) sc = (size_t)atoll(e);
  w.seg_cap = (uint32_t)sc;
  TS_CUDA(cudaMalloc(&w.seg, sc * sizeof(TsSeg)));
  TS_CUDA(cudaMalloc(&w.seg_count, nb * 4));
  TS_CUDA(cudaMemset(w.seg_count, 0, nb * 4));
  TS_CUDA(cudaMalloc(&w.seg_off, nb * 4));
  TS_CUDA(cudaMalloc(&w.touched, nb * 4));
  w.item_cap = (uint32_t)(sc / MR_CHUNK + nb + 16);
  TS_CUDA(cudaMalloc(&w.items, (size_t)w.item_cap * sizeof(TsItem)));
  w.gen_cap = (uint32_t)(nr > (1u << 20) ? nr : (1u << 20));
  TS_CUDA(cudaMalloc(&w.gen, (size_t)w.gen_cap * sizeof(TsSeg)));
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
  TS_CUDA(cudaFuncSetAttribute(k_march_blocks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  TS_CUDA(cudaFuncSetAttribute(k_march_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  return TSLAM_OK;
}

void ts_march_free(tslam_tsdf* m) {
  TsMarchWs& w = m->mw;
  cudaFree(w.rays); cudaFree(w.aux); cudaFree(w.seg); cudaFree(w.seg_count); cudaFree(w.seg_off); cudaFree(w.touched);
  cudaFree(w.items); cudaFree(w.gen); cudaFree(w.ctl);
}

// sub_ev (profiling, may be null): 4 events recorded after set-up, scan, fill and the march kernels
int ts_march_launch(tslam_tsdf* m, cudaStream_t st, const TsBatch& batch, uint32_t bucket_shift, cudaEvent_t* sub_ev) {
  const int sms = m->sm_count;
  k_ray_setup<<<sms * 4, MS_THREADS, 0, st>>>(batch, m->in, m->g, m->buckets, bucket_shift, m->ray_list, m->n_rays, m->ray_list_cap, m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[0], st));
  k_seg_scan<<<1, 1024, 0, st>>>(m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[1], st));
  k_seg_fill<<<sms * 4, MS_THREADS, 0, st>>>(batch, m->g, m->n_rays, m->ray_list_cap, m->mw);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[2], st));
  const int grid = (sms - m->rm_reserve) * 2;
  if (m->march_verify) k_march_blocks<true><<<grid, MB_THREADS, MB_SMEM, st>>>(batch, m->in, m->g, m->mw, m->counters);
  else k_march_blocks<false><<<grid, MB_THREADS, MB_SMEM, st>>>(batch, m->in, m->g, m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  k_march_generic<<<sms * 2, 256, 0, st>>>(batch, m->in, m->g, m->mw, m->counters);
  TS_LAUNCH_CHECK(m);
  k_march_reset<<<32, 256, 0, st>>>(m->mw);
  TS_LAUNCH_CHECK(m);
  if (sub_ev) TS_CUDA(cudaEventRecord(sub_ev[3], st));
  return TSLAM_OK;
}
