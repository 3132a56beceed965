// Internal structures of libtslam.so (sm_100a).  Not part of the ABI.
//
// HBM layout of a TSDF map (replaces Taichi's pointer->dense SNode tree,
// dense_tsdf.py:108-118):
//   * one open-addressing hash table  slot -> (key40 | block24)   8 B/slot
//   * a pool of 16^3-voxel blocks stored as per-field PLANES (SoA across the pool,
//     4096 consecutive voxels of one block are contiguous in every plane):
//         acc  float2[4096]  pending (sum w*d, sum w) of not-yet-committed frames
//         tw   float2[4096]  committed (TSDF, W_TSDF)   (f16 x2 in the reference)
//         obs  u8[4096]      TSDF_observed
//         occ  i8[4096]      occupy
//     A voxel update of the ray-march kernel is ONE 8-byte vector reduction
//     (REDG.E.ADD.F32x2) into `acc`; whole blocks are streamed by the commit /
//     marching-cubes / export kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tslam.h"

#define TS_B 16
#define TS_B3 4096
#define TS_BSHIFT 4
#define TS_BMASK 15

// packed hash word:  [63:24] key (s:10 | bx:10 | by:10 | bz:10), [23:0] block index
#define TS_EMPTY 0xFFFFFFFFFFFFFFFFull
#define TS_IDX_MASK 0xFFFFFFull
#define TS_IDX_PENDING 0xFFFFFEull
#define TS_IDX_OVERFLOW 0xFFFFFDull
#define TS_MAX_BLOCKS 0xFFFFF0

#define TS_ERR_POOL_FULL 1
#define TS_ERR_TABLE_FULL 2
#define TS_ERR_RAYLIST_FULL 4
#define TS_ERR_BUCKET_RANGE 8   // a point further than 4095 voxels from the sensor origin: beyond the bucket index's key range
#define TS_PROF_RING 512
#define TS_PROF_EV 8   // events per profiled integrate launch

struct TsGrid {
  unsigned long long* table;  // hash words
  uint32_t table_mask;        // capacity-1 (power of two)
  int max_blocks;
  int* n_blocks;              // device allocation counter
  unsigned long long* block_key;  // [max_blocks] reverse map (key40)
  float2* acc;                // [max_blocks*4096]
  float2* tw;                 // [max_blocks*4096]
  uint8_t* obs;               // [max_blocks*4096]
  int8_t* occ;                // [max_blocks*4096]
  unsigned long long* cword;  // [max_blocks*4096] texture: winning colour word per voxel (null unless texture_enabled)
  float4* col;                // [max_blocks*4096] texture: committed colour rgb (+pad)
  uint8_t* ghost;             // [max_blocks] 1 = halo copy of a block owned by another rank (multi-GPU tiling)
  float* esdf;                // [max_blocks*4096] (allocated lazily by the ESDF path)
  int* esdf_dirty;            // [max_blocks] block committed since the last ESDF update (seeds the incremental wave)
  int* dirty_flag;            // [max_blocks]
  int* dirty_list;            // [max_blocks]
  int* n_dirty;
  int* err;                   // device error flags
  int hN, hNz, N, Nz;         // bounds: i in [-hN, N-hN)
};

__host__ __device__ __forceinline__ unsigned long long ts_pack_key(int s, int bx, int by, int bz) {
  return ((unsigned long long)(s & 1023) << 30) | ((unsigned long long)((bx + 512) & 1023) << 20) |
         ((unsigned long long)((by + 512) & 1023) << 10) | (unsigned long long)((bz + 512) & 1023);
}
__host__ __device__ __forceinline__ void ts_unpack_key(unsigned long long k, int& s, int& bx, int& by, int& bz) {
  s = (int)((k >> 30) & 1023);
  bx = (int)((k >> 20) & 1023) - 512;
  by = (int)((k >> 10) & 1023) - 512;
  bz = (int)(k & 1023) - 512;
}
__device__ __forceinline__ uint32_t ts_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}
template <class G>
__device__ __forceinline__ bool ts_in_bounds(const G& g, int i, int j, int k) {
  return i >= -g.hN && i < g.N - g.hN && j >= -g.hN && j < g.N - g.hN && k >= -g.hNz && k < g.Nz - g.hNz;
}
__device__ __forceinline__ int ts_voxel_off(int i, int j, int k) {
  return (((i & TS_BMASK) << TS_BSHIFT) | (j & TS_BMASK)) << TS_BSHIFT | (k & TS_BMASK);
}
__device__ __forceinline__ unsigned long long ts_ld_volatile(const unsigned long long* p) {
  return *(const volatile unsigned long long*)p;
}

// read-only lookup: block index or -1   (G = TsGrid or OcGrid)
template <class G>
__device__ __forceinline__ int ts_find(const G& g, unsigned long long key) {
  uint32_t slot = ts_hash(key) & g.table_mask;
  for (uint32_t probe = 0; probe <= g.table_mask; ++probe) {
    unsigned long long cur = ts_ld_volatile(&g.table[slot]);
    if (cur == TS_EMPTY) return -1;
    if ((cur >> 24) == key) {
      unsigned idx = (unsigned)(cur & TS_IDX_MASK);
      return (idx >= TS_IDX_OVERFLOW) ? -1 : (int)idx;
    }
    slot = (slot + 1) & g.table_mask;
  }
  return -1;
}

// lookup, activating a zero-filled block on miss (Taichi "write activates").
// Returns -1 when the pool is exhausted (error flag raised, sample dropped).
template <class G>
__device__ __forceinline__ int ts_get_or_alloc(const G& g, unsigned long long key) {
  uint32_t slot = ts_hash(key) & g.table_mask;
  for (uint32_t probe = 0; probe <= g.table_mask; ++probe) {
    unsigned long long cur = ts_ld_volatile(&g.table[slot]);
    if (cur == TS_EMPTY) {
      unsigned long long prev = atomicCAS(&g.table[slot], TS_EMPTY, (key << 24) | TS_IDX_PENDING);
      if (prev == TS_EMPTY) {
        int idx = atomicAdd(g.n_blocks, 1);
        unsigned long long word;
        if (idx >= g.max_blocks) {
          atomicSub(g.n_blocks, 1);
          atomicOr(g.err, TS_ERR_POOL_FULL);
          word = (key << 24) | TS_IDX_OVERFLOW;
          idx = -1;
        } else {
          g.block_key[idx] = key;
          word = (key << 24) | (unsigned long long)idx;
        }
        __threadfence();
        atomicExch(&g.table[slot], word);
        return idx;
      }
      cur = prev;
    }
    if ((cur >> 24) == key) {
      while ((cur & TS_IDX_MASK) == TS_IDX_PENDING) {
        __nanosleep(32);
        cur = ts_ld_volatile(&g.table[slot]);
      }
      unsigned idx = (unsigned)(cur & TS_IDX_MASK);
      return (idx >= TS_IDX_OVERFLOW) ? -1 : (int)idx;
    }
    slot = (slot + 1) & g.table_mask;
  }
  atomicOr(g.err, TS_ERR_TABLE_FULL);
  return -1;
}

// Touched blocks are flagged with a fire-and-forget store (no load on the ray-march critical
// path); k_collect_dirty compacts the flags into dirty_list before a commit.
__device__ __forceinline__ void ts_mark_dirty(const TsGrid& g, int blk) { g.dirty_flag[blk] = 1; }

// Fast lookup for the hot loops: first probe through L1 (plain cached load).  Entries are only
// ever inserted while kernels run (the table is cleared between launches, where L1 is
// invalidated), so a cached hit is always valid; anything else falls back to the coherent path.
template <class G>
__device__ __forceinline__ int ts_get_or_alloc_cached(const G& g, unsigned long long key) {
  const uint32_t slot = ts_hash(key) & g.table_mask;
  const unsigned long long cur = __ldca(&g.table[slot]);
  if ((cur >> 24) == key) {
    const unsigned idx = (unsigned)(cur & TS_IDX_MASK);
    if (idx < TS_IDX_OVERFLOW) return (int)idx;
  }
  return ts_get_or_alloc(g, key);
}

// ---------------------------------------------------------------------------
// arithmetic helpers shared by the integrate kernels (tslam_tsdf.cu, tslam_march.cu)
// ---------------------------------------------------------------------------
#define WMAX 1000.0f          // dense_tsdf.py:8
#define FIXQ 1048576.0f       // 2^20: fixed-point quantum of the per-frame bucket sums
#define FIXQ_D 1048576.0

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int iroundf(float x) { return (int)roundf(x); }  // ti.round(x, i32) mapping_common.py:263-266
__device__ __forceinline__ float sgnf(float v) { return (float)((0.0f < v) - (v < 0.0f)); }  // mapping_common.py:5-7

// x / vs, correctly rounded, without the slow-path check of the generic IEEE division: rvs = RN(1/vs) comes
// from the host; q1 = fma(fma(-q0, vs, x), rvs, q0) is the correctly rounded quotient (Markstein) for the
// operand ranges that occur here (|x| < 1e4 m, vs ~ 1e-2..1 m: no overflow / underflow / denormals).
__device__ __forceinline__ float div_vs(float x, float vs, float rvs) {
  const float q0 = __fmul_rn(x, rvs);
  const float e = __fmaf_rn(-q0, vs, x);
  return __fmaf_rn(e, rvs, q0);
}

__device__ __forceinline__ void red_add_f32x2(float2* addr, float a, float b) {
  // one 8-byte vector reduction without return value: REDG.E.ADD.F32x2 (sm_90+).  Spelled in PTX because
  // atomicAdd(float2*) was lowered to ATOMG (response sector per update) inside the divergent march loop.
  asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void red_add_u64(unsigned long long* addr, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(unsigned int* addr, unsigned int v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

#define RM_TAB 4096          // shared block-lookup table: 16^3 entries indexed by the low 4 bits of the block coords
// block lookup for the march loop: shared-memory table first (global loads queue behind the reduction traffic in
// the in-order L1TEX pipe: measured as the top stall), hash grid on a miss.
__device__ __forceinline__ int rm_lookup(const TsGrid& g, unsigned long long* tab, unsigned long long key, int bx, int by, int bz) {
  const int h = ((bx & 15) << 8) | ((by & 15) << 4) | (bz & 15);
  const unsigned long long w = tab[h];
  if ((w >> 24) == key) return (int)(w & TS_IDX_MASK);
  const int blk = ts_get_or_alloc_cached(g, key);
  if (blk >= 0) {
    ts_mark_dirty(g, blk);
    tab[h] = (key << 24) | (unsigned long long)blk;  // benign race: any writer stores a valid word
  }
  return blk;
}


// ---------------------------------------------------------------------------
// per-frame parameters of one integrate launch, passed BY VALUE as a
// __grid_constant__ kernel parameter (no H2D copy, no lifetime hazards)
// ---------------------------------------------------------------------------
struct TsFrame {
  float R[9];
  float T[3];
  int submap;
  unsigned int seq;  // frame sequence number of the map (texture: later frames overwrite earlier colours)
};
struct TsBatch {
  TsFrame f[TSLAM_MAX_BATCH];  // 64 * 56 B = 3584 B
};

struct TsIntrin {
  float fx, fy, cx, cy;
  float dmin_mm, dmax_mm;  // f32(min_ray*1000), f32(max_ray*1000)
  float vs;
  float rvs;               // RN(1/vs), for the exact FMA division in the march loop
  float max_steps;         // f32(max_ray_length / voxel_scale)
  float max_ray;           // f32(max_ray_length)
  int internal_voxels;
  int step;
  // texture (dense_tsdf.py:204-211, mapping_common.py:43-58)
  float fxc, fyc, cxc, cyc;
  int tex;         // texture_enabled
  int same_proj;   // color_same_proj
};

// per-frame bucket entry (dense_tsdf.py:64-70 new_pcl_count / new_pcl_sum_pos / new_pcl_z),
// exact fixed-point sums (2^-20 m).  64-byte stride = two 32-byte sectors.
struct __align__(64) TsBucket {
  unsigned long long key;  // packed (bx,by,bz)+1, 0 = empty
  long long sx, sy, sz, sd;
  int cnt;
  unsigned int cr, cg, cb;  // new_pcl_sum_color: exact integer channel sums
  int pad[2];
};

struct TsCounters {  // device-side statistics (tslam_tsdf_get_stats)
  unsigned long long n_px, n_valid, n_rays, n_updates, n_oob;
  // block-binned ray march (tslam_march.cu): diagnostics, not part of the parity counters
  unsigned long long n_segs, n_items, n_slow, n_fallback, n_generic, n_verify_bad;
};

// ---------------------------------------------------------------------------
// block-binned ray march (tslam_march.cu)
// ---------------------------------------------------------------------------
// One record per live ray (= per-frame bucket), written by k_ray_setup: unit direction, length, sensor origin in
// VOXEL units (T/vs) and the ray's sample weight 1/z^2 (dense_tsdf.py:243-247,262).
struct __align__(16) TsRay {
  float ux, uy, uz, L;
  float tx, ty, tz, w;
};
// aux word per ray: [31:16] n = number of march steps (dense_tsdf.py:249-251), [15:8] frame of the batch, [0] wide
// (more than 65535 steps: generic path)
#define TS_AUX_WIDE 1u
// A segment = run of consecutive march steps of one ray whose samples fall (approximately - every sample is
// re-checked exactly by the march kernel) into one 16^3 voxel block.  jc = frame << 24 | j0 << 8 | count (count <= 32).
struct TsSeg { uint32_t ray, jc; };
// generic-list records use jc = j0 << 12 | count (count <= 4095)
struct TsItem { int blk; uint32_t seg0, nseg, pad; };  // work item of k_march_blocks: <= MR_CHUNK segments of one block
#define MR_CHUNK 4096
struct TsMarchCtl {  // device-side control block of one launch (zeroed by k_march_reset at its end)
  int n_touched;     // blocks that received at least one segment
  int n_items;       // work items built by k_seg_scan
  int n_full;        // (unused since the items are listed longest first)
  int item_cursor;   // persistent-CTA work cursor of k_march_blocks
  int n_gen;         // generic-list records
  int overflow;      // work list did not fit: every segment goes through the generic path this launch
  unsigned int total_segs;
  unsigned int fmax_bits;  // float bits of max over rays of max(w, w*|ds|max): sizes the fixed-point scale of the launch
  int scale_k;       // shared-memory sums are kept in units of 2^-scale_k (k_seg_scan: largest k with max * 2^k < 2^30)
  int ticket;
  unsigned int tmp_cursor;  // next free entry of the segment-list pool
  int pad0, pad1;
  int pad[3];
};
struct TsMarchWs {
  TsRay* rays;         // [ray_list_cap]
  uint32_t* aux;       // [ray_list_cap]
  TsSeg* seg;          // [seg_cap] segments grouped by block, ordered by length inside a block
  uint32_t seg_cap;
  TsSeg* tmp_seg;      // [seg_cap] per-CTA segment lists of k_seg_walk (chunks of 4096 entries from one pool)
  uint32_t* tmp_key;   // [seg_cap] block * 16 + length class of the listed segment
  int* seg_count;      // [max_blocks * 16] segments per (block, class); then the fill cursor of k_seg_place
  uint32_t* seg_rel;   // [max_blocks * 16] start of the class inside the block's run
  uint32_t* seg_off;   // [max_blocks] start of the block's run
  int* touched;        // [max_blocks] blocks with segments
  uint32_t* blk_total; // [max_blocks] segments of touched[i]
  uint32_t* blk_cost;  // [max_blocks] samples of touched[i] (upper estimate from the length classes)
  TsItem* items;       // [item_cap]
  uint32_t item_cap;
  TsSeg* gen;          // [gen_cap] generic-path records (volume boundary, over-long rays, overflow)
  uint32_t gen_cap;
  TsMarchCtl* ctl;
  int* cta_n;          // [walk CTAs] segments in the CTA's list
  uint32_t* cta_chunk; // [walk CTAs * 1024] first entry of the CTA's k-th list chunk
  int walk_x_max;      // walk CTAs of a one-frame launch
  float near_eps;      // |frac - 0.5| below this -> exact index path
};

// host-side handle
struct tslam_tsdf {
  tslam_tsdf_config_t cfg;
  int device;
  TsGrid g;
  TsIntrin in;
  size_t table_cap;
  // integrate workspace
  // per-frame bucket grid, split in two: a small open-addressing INDEX (8 B per entry: record slot + 39-bit key) that
  // the bucket kernel probes, and the 64-byte RECORDS with the exact sums (every warp of the bucket kernel owns 32 of them), filled in the order the buckets
  // are opened - so the ray set-up reads and zeroes them in short runs instead of 64 random bytes per ray
  TsBucket* buckets;   // records [TSLAM_MAX_BATCH * bucket_cap]; frame f owns [f * bucket_cap, (f + 1) * bucket_cap)
  unsigned long long* bidx;  // index [TSLAM_MAX_BATCH * bucket_cap]
  uint32_t bucket_cap; // power of two
  uint32_t* ray_list;  // [TSLAM_MAX_BATCH * max_rays_per_frame]
  uint32_t ray_list_cap;
  int* n_rays;         // device counter
  uint16_t* depth_stage;  // device staging for host depth input [TSLAM_MAX_BATCH * max_image_pixels]
  uint8_t* tex_stage;     // device staging for host textures [2 * TSLAM_MAX_BATCH * max_image_pixels * 3] (texture_enabled only)
  uint8_t* rgb_stage;     // device staging for point-cloud colours
  unsigned int frame_seq; // frames integrated since the last renormalisation of the colour words (ts_seq_renorm)
  int q_th, q_tw, q_has_tex;
  // pinned host frames are not copied: q_hptr[q] = their device alias; the sampled rows are fetched over PCIe by
  // k_gather_rows in groups of TS_GATHER_GROUP frames (q_gathered = frames already handed to a gather launch)
  const uint16_t* q_hptr[TSLAM_MAX_BATCH];
  int q_gathered;
  int zero_copy;  // 1 = borrow page-locked frames (tslam_tsdf_set_frame_mode / TSLAM_ZERO_COPY=1); 0 = copy (default)
  // copy mode, pageable frames: the call copies the SAMPLED ROWS of the frame into this page-locked ring on the host (the
  // caller may reuse its buffer at once) and the GPU fetches them from there like a borrowed frame
  uint16_t* h_ring;       // [2][TSLAM_MAX_BATCH][ring_frame_cap] host, mapped
  uint16_t* h_ring_dev;   // its device alias
  size_t ring_frame_cap;  // elements per frame slot
  int stage_mode;         // pageable frames: 1 = through the ring (default), 0 = through cudaMemcpyAsync (TSLAM_FRAME_COPY=dma)
  int pinned_mode;        // copy mode, page-locked source: 2 = awaited DMA copy (default), 0 = through the ring like pageable frames, 1 = awaited row fetch
  int q_open, q_await;    // a frame's hand-over has begun (tslam_tsdf_queue_depth_begin) / its DMA copy is still to be awaited
  int q_sstride[TSLAM_MAX_BATCH];  // source row stride of the queued frame in 16-byte units
  int queue_launch[2];  // frames per queue launch, alternating (TSLAM_QUEUE_LAUNCH="a,b", default 32,32)
  int q_phase;          // which of the two thresholds the current launch uses (back to 0 at every flush)
  int trace;         // TSLAM_TRACE=1: print a per-launch timeline of the queue (debug)
  cudaEvent_t tr_ev[2][4];  // per staging buffer: copy stream first op / last op, main stream launch begin / end
  cudaEvent_t tr_base;
  int tr_have[2];
  int rm_reserve;    // SMs left out of the ray-march grid (2 once page-locked frames are being gathered, else 0)
  float* points_stage;    // device staging for host point clouds
  TsCounters* counters;
  float* pose_R;       // device pose table [max_submaps*9]
  float* pose_T;       // [max_submaps*3]
  float* colormap;     // device jet LUT [1024*3]
  int* scratch_i;      // small device scratch (counters for gather etc.)
  void* esdf_aux;      // ESDF: neighbour table, epochs, per-sweep change flags, class bytes (allocated with g.esdf)
  int esdf_full_needed, esdf_submap;  // the next update recomputes everything (first use, reset, bulk load) / submap of the kept state
  void* mc_scratch;    // marching cubes: per-block triangle counts / offsets (allocated on first use)
  long long launches;
  int n_integrate_calls;
  int profiling;
  cudaEvent_t* ev;          // profiling ring: TS_PROF_RING launches x 4 events (created lazily)
  long long prof_launches;  // integrate launches recorded since profiling was switched on
  int sm_count;
  int tile_cuts_set, tile_cuts_tiles[3], tile_cuts[3][17];  // data-driven tile boundaries of a tiled global map (tslam_dist.cu)
  int* tile_hist;      // device scratch of tslam_tsdf_dirty_hist (allocated on first use)
  TsMarchWs mw;        // block-binned ray march workspace (tslam_march.cu)
  int march_mode;      // 0 = legacy k_raymarch, 1 = block-binned (default for untextured maps), env TSLAM_MARCH
  int march_verify;    // TSLAM_MARCH_VERIFY=1: every fast-path index is re-computed exactly, mismatches counted
  bool clamp_on_commit;
  // frame queue of the per-frame API (tslam_tsdf_queue_depth): double-buffered device staging fed by a copy stream
  int q_n, q_buf, q_h, q_w;
  float q_R[TSLAM_MAX_BATCH * 9];
  float q_T[TSLAM_MAX_BATCH * 3];
  int q_s[TSLAM_MAX_BATCH];
  cudaStream_t copy_stream;
  cudaEvent_t ev_copied[2], ev_free[2];
  bool ev_free_valid[2];
};

// ---------------------------------------------------------------------------
// Octomap: hash grid of 8^3 blocks of u32 hit counters (taichi_octomap.py:63-84
// builds a K-ary pointer tree with one allocation per leaf instead)
// ---------------------------------------------------------------------------
#define OC_B 8
#define OC_B3 512
#define OC_BSHIFT 3
#define OC_BMASK 7
struct OcGrid {
  unsigned long long* table;
  uint32_t table_mask;
  int max_blocks;
  int* n_blocks;
  unsigned long long* block_key;
  unsigned int* cnt;  // [max_blocks*512]
  unsigned long long* cw;  // texture_enabled: [max_blocks*512] colour word (integrate seq << 24 | r << 16 | g << 8 | b), atomicMax
  int* err;
  int hN, hNz, N, Nz;
};
__device__ __forceinline__ int oc_voxel_off(int i, int j, int k) {
  return (((i & OC_BMASK) << OC_BSHIFT) | (j & OC_BMASK)) << OC_BSHIFT | (k & OC_BMASK);
}
struct tslam_octo {
  tslam_octo_config_t cfg;
  OcGrid g;
  TsIntrin in;
  size_t table_cap;
  uint16_t* depth_stage;
  float* points_stage;
  uint8_t* tex_stage;     // texture_enabled: device staging of a host colour image / point colours
  unsigned int frame_seq; // integrate calls so far
  float* pose_R;
  float* pose_T;
  int* scratch_i;
  long long launches;
  int sm_count;
};

// 64-bit max without a return value: RED.E.MAX.64 (atomicMax(unsigned long long*) with an unused result was lowered
// to ATOMG - a response sector per update, seen as 57 M "L2 ATOM sectors" per textured launch in ncu)
__device__ __forceinline__ void ts_red_max_u64(unsigned long long* addr, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// colour pixel of depth pixel (i, j): texture[j, i] (color_same_proj) or color_ind_from_depth_pt
// (mapping_common.py:43-58).  The reference tests color_i against h and color_j against w (swapped, :56); what passes
// that test but lies outside the image is an out-of-bounds read there - pixel (0,0) here.  Returns false when even the
// same-projection pixel is outside the colour image.
__device__ __forceinline__ bool ts_color_pixel(const TsIntrin& in, int i, int j, int th, int tw, int& ti, int& tj) {
  tj = j; ti = i;
  if (!in.same_proj) {
    ti = (int)((((float)i - in.cx) / in.fx) * in.fxc + in.cxc);
    tj = (int)((((float)j - in.cy) / in.fy) * in.fyc + in.cyc);
    if (ti < 0 || ti >= th || tj < 0 || tj >= tw || tj >= th || ti >= tw) { ti = 0; tj = 0; }
  }
  return tj < th && ti < tw;
}

// error plumbing (tslam_tsdf.cu)
void ts_set_error(const char* fmt, ...);
int ts_cuda_fail(cudaError_t e, const char* what);
#define TS_CUDA(x)                                    \
  do {                                                \
    cudaError_t _e = (x);                             \
    if (_e != cudaSuccess) return ts_cuda_fail(_e, #x); \
  } while (0)
#define TS_LAUNCH_CHECK(m)                                              \
  do {                                                                  \
    (m)->launches++;                                                    \
    cudaError_t _e = cudaGetLastError();                                \
    if (_e != cudaSuccess) return ts_cuda_fail(_e, "kernel launch");    \
  } while (0)

// shared across translation units
int ts_flush_pending(tslam_tsdf* m, cudaStream_t st);   // commit if anything is pending
int ts_check_deferred(tslam_tsdf* m);                    // read + translate device error flags (synchronises)
int ts_check_deferred_async(tslam_tsdf* m, cudaStream_t st);  // same, after the work queued on st
// tslam_march.cu: block-binned ray march of the rays listed in m->ray_list (replaces k_raymarch<false>)
int ts_march_alloc(tslam_tsdf* m);
void ts_march_free(tslam_tsdf* m);
int ts_march_setup(tslam_tsdf* m, cudaStream_t st, const TsBatch& batch, uint32_t bucket_shift);  // rays listed since the last call
int ts_march_launch(tslam_tsdf* m, cudaStream_t st, const TsBatch& batch, uint32_t bucket_shift, cudaEvent_t* sub_ev);
