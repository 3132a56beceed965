// libtslam.so - Octomap: per-voxel hit counter on a hash grid of 8^3 u32 blocks.
// Reference: taichi_slam/mapping/taichi_octomap.py.  The reference "Octomap" has no
// log-odds / ray casting: occupy[ijk] += 1 per point (:116-119) and a threshold
// test (:86-88); counts are exact integers, compared bit-for-bit with the oracle.
#include <cstring>
#include "tslam_internal.cuh"

__device__ __forceinline__ int oc_iroundf(float x) { return (int)roundf(x); }

// cw != 0: colour word of the hit (textured maps): color[ijk] = rgb (:120-124), largest word wins
__device__ __forceinline__ void oc_hit(const OcGrid& g, int s, int i, int j, int k, unsigned int add, unsigned long long cw = 0ull) {
  if (!ts_in_bounds(g, i, j, k)) return;
  const int blk = ts_get_or_alloc(g, ts_pack_key(s, i >> OC_BSHIFT, j >> OC_BSHIFT, k >> OC_BSHIFT));
  if (blk < 0) return;
  const size_t o = (size_t)blk * OC_B3 + oc_voxel_off(i, j, k);
  atomicAdd(&g.cnt[o], add);  // occupy[ijk] += 1 (:119)
  if (cw) ts_red_max_u64(&g.cw[o], cw);
}
// "Stupid OpenCV is BGR" (:121): colour channel 0 <- rgb[2], 2 <- rgb[0]
__device__ __forceinline__ unsigned long long oc_color_word(unsigned int seq, const uint8_t* p) {
  return ((unsigned long long)seq << 24) | ((unsigned long long)p[2] << 16) | ((unsigned long long)p[1] << 8) | (unsigned long long)p[0];
}
__device__ __forceinline__ void oc_word_to_rgb(unsigned long long w, float* out) {
  out[0] = (float)((w >> 16) & 255) / 255.0f; out[1] = (float)((w >> 8) & 255) / 255.0f; out[2] = (float)(w & 255) / 255.0f;
}

// recast_pcl_to_map_kernel (taichi_octomap.py:134-145)
__global__ void __launch_bounds__(256) k_octo_points(OcGrid g, const float* __restrict__ xyz, int n, TsFrame fr, float vs,
                                                      const uint8_t* __restrict__ rgb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float x = xyz[3 * (size_t)t], y = xyz[3 * (size_t)t + 1], z = xyz[3 * (size_t)t + 2];
  const float px = ((fr.R[0] * x + fr.R[1] * y) + fr.R[2] * z) + fr.T[0];  // :141
  const float py = ((fr.R[3] * x + fr.R[4] * y) + fr.R[5] * z) + fr.T[1];
  const float pz = ((fr.R[6] * x + fr.R[7] * y) + fr.R[8] * z) + fr.T[2];
  const unsigned long long cw = rgb ? oc_color_word(fr.seq, rgb + 3 * (size_t)t) : 0ull;
  oc_hit(g, fr.submap, oc_iroundf(px / vs), oc_iroundf(py / vs), oc_iroundf(pz / vs), 1u, cw);  // xyz_to_sijk mapping_common.py:251-255
}

// recast_depth_to_map_kernel (taichi_octomap.py:147-169)
__global__ void __launch_bounds__(256) k_octo_depth(OcGrid g, const uint16_t* __restrict__ depth, int h, int w, int hh, int ww,
                                                     TsFrame fr, TsIntrin in, const uint8_t* __restrict__ tex, int th, int tw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hh * ww) return;
  const int jj = t / ww, ii = t - jj * ww;
  const int j = jj * in.step, i = ii * in.step;
  const uint16_t d = depth[(size_t)j * w + i];
  const float df = (float)d;
  if (d == 0 || df > in.dmax_mm || df < in.dmin_mm) return;  // :155
  const float dep = df / 1000.0f;                              // :157
  const float x = ((float)i - in.cx) * dep / in.fx;
  const float y = ((float)j - in.cy) * dep / in.fy;
  const float px = ((fr.R[0] * x + fr.R[1] * y) + fr.R[2] * dep) + fr.T[0];  // :159
  const float py = ((fr.R[3] * x + fr.R[4] * y) + fr.R[5] * dep) + fr.T[1];
  const float pz = ((fr.R[6] * x + fr.R[7] * y) + fr.R[8] * dep) + fr.T[2];
  unsigned long long cw = 0ull;
  if (tex) {  // :160-167
    int ti, tj;
    const uint8_t zero3[3] = {0, 0, 0};
    cw = oc_color_word(fr.seq, ts_color_pixel(in, i, j, th, tw, ti, tj) ? tex + ((size_t)tj * tw + ti) * 3 : zero3);
  }
  oc_hit(g, fr.submap, oc_iroundf(px / in.vs), oc_iroundf(py / in.vs), oc_iroundf(pz / in.vs), 1u, cw);
}

__device__ __forceinline__ long long oc_warp_append(bool want, unsigned long long* counter) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (!m) return -1;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return want ? (long long)(base + __popc(m & ((1u << lane) - 1))) : -1;
}

__global__ void __launch_bounds__(256) k_octo_gather(OcGrid g, int submap, long long cap, int32_t* idx, unsigned int* count,
                                                      float* color, unsigned long long* counter) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  const long long total = (long long)nb * OC_B3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {  // uniform trip count: the warp ballot below needs all lanes
    const long long e = it * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    unsigned int c = 0;
    int s = 0, bx = 0, by = 0, bz = 0, v = 0;
    if (e < total) {
      const int b = (int)(e >> 9);
      v = (int)(e & 511);
      ts_unpack_key(g.block_key[b], s, bx, by, bz);
      c = g.cnt[e];
      want = (s == submap) && c > 0;
    }
    const long long row = oc_warp_append(want, counter);
    if (want && row < cap) {
      idx[3 * row] = bx * OC_B + (v >> 6);
      idx[3 * row + 1] = by * OC_B + ((v >> 3) & 7);
      idx[3 * row + 2] = bz * OC_B + (v & 7);
      count[row] = c;
      if (color) {
        if (g.cw) oc_word_to_rgb(g.cw[e], color + 3 * row);
        else color[3 * row] = color[3 * row + 1] = color[3 * row + 2] = 0.0f;
      }
    }
  }
}

// cvt_occupy_to_voxels(level) / cvt_occupy_voxels_to (taichi_octomap.py:90-114).
// occupy.parent(level) visits the active ancestor cells, reported at the group's base
// coordinate, and is_occupy (:86-88) is evaluated AT that coordinate: a group is
// exported iff its corner voxel exists and has count > min_occupy_thres.
__global__ void __launch_bounds__(256) k_octo_extract(OcGrid g, int submap, int group, float thres, const float* pR, const float* pT,
                                                       float vs, long long cap, float* xyz, float* rgb, int* counter) {
  const int nb = min(*g.n_blocks, g.max_blocks);
  const long long total = (long long)nb * OC_B3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long e = it * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    int i = 0, j = 0, k = 0, s = 0;
    if (e < total) {
      const int b = (int)(e >> 9);
      const int v = (int)(e & 511);
      int bx, by, bz;
      ts_unpack_key(g.block_key[b], s, bx, by, bz);
      i = bx * OC_B + (v >> 6); j = by * OC_B + ((v >> 3) & 7); k = bz * OC_B + (v & 7);
      const unsigned int c = g.cnt[e];
      // tree offset is -N/2 (taichi_octomap.py:72): groups are aligned in offset coordinates
      const bool aligned = ((i + g.hN) % group == 0) && ((j + g.hN) % group == 0) && ((k + g.hNz) % group == 0);
      want = (s == submap) && c > 0 && aligned && ((float)c > thres);
    }
    const unsigned m = __ballot_sync(0xffffffffu, want);
    int row = -1;
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(counter, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      row = base + __popc(m & ((1u << lane) - 1));
    }
    if (want && row < cap) {
      const float lx = (float)i * vs, ly = (float)j * vs, lz = (float)k * vs;  // sijk_to_xyz mapping_common.py:234-238
      const float* R = pR + 9 * s;
      const float* T = pT + 3 * s;
      xyz[3 * (size_t)row] = ((R[0] * lx + R[1] * ly) + R[2] * lz) + T[0];
      xyz[3 * (size_t)row + 1] = ((R[3] * lx + R[4] * ly) + R[5] * lz) + T[1];
      xyz[3 * (size_t)row + 2] = ((R[6] * lx + R[7] * ly) + R[8] * lz) + T[2];
      if (rgb && g.cw) oc_word_to_rgb(g.cw[e], rgb + 3 * (size_t)row);  // export_color[index] = color[sijk] (:101-102)
    }
  }
}

// fuse_submaps_kernel (taichi_octomap.py:171-189)
__global__ void __launch_bounds__(256) k_octo_fuse(OcGrid dst, OcGrid src, float thres, const float* pR, const float* pT, float vs) {
  const int nb = min(*src.n_blocks, src.max_blocks);
  const long long total = (long long)nb * OC_B3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const unsigned int c = src.cnt[e];
    if (!((float)c > thres)) continue;  // :181
    const int b = (int)(e >> 9), v = (int)(e & 511);
    int s, bx, by, bz;
    ts_unpack_key(src.block_key[b], s, bx, by, bz);
    const float lx = (float)(bx * OC_B + (v >> 6)) * vs, ly = (float)(by * OC_B + ((v >> 3) & 7)) * vs,
                lz = (float)(bz * OC_B + (v & 7)) * vs;
    const float* R = pR + 9 * s;
    const float* T = pT + 3 * s;
    const float x = ((R[0] * lx + R[1] * ly) + R[2] * lz) + T[0];  // :182
    const float y = ((R[3] * lx + R[4] * ly) + R[5] * lz) + T[1];
    const float z = ((R[6] * lx + R[7] * ly) + R[8] * lz) + T[2];
    const unsigned long long cw = (dst.cw && src.cw) ? src.cw[e] : 0ull;  // color[ijk_] = submap_color[s,i,j,k] (:189)
    oc_hit(dst, 0, oc_iroundf(x / vs), oc_iroundf(y / vs), oc_iroundf(z / vs), c, cw);  // :183-186
  }
}

// ---------------------------------------------------------------------------
static size_t oc_next_pow2(size_t v) {
  size_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

extern "C" int tslam_octo_create(const tslam_octo_config_t* cfg, tslam_octo_t** out) {
  if (!cfg || !out) return TSLAM_E_INVALID;
  *out = nullptr;
  if (cfg->voxel_scale <= 0 || cfg->N <= 0 || cfg->Nz <= 0 || cfg->N > 8192 || cfg->Nz > 8192 || cfg->recast_step <= 0) {
    ts_set_error("invalid octomap config");
    return TSLAM_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    ts_set_error("no CUDA device visible: libtslam has no CPU fallback");
    return TSLAM_E_NOGPU;
  }
  tslam_octo* m = new tslam_octo();
  memset(m, 0, sizeof(*m));
  m->cfg = *cfg;
  int dev = 0;
  TS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TS_CUDA(cudaGetDeviceProperties(&prop, dev));
  m->sm_count = prop.multiProcessorCount;
  if (m->cfg.max_submaps <= 0 || m->cfg.max_submaps > 1024) m->cfg.max_submaps = 1024;
  if (m->cfg.max_image_pixels <= 0) m->cfg.max_image_pixels = 640 * 480;
  if (m->cfg.max_points <= 0) m->cfg.max_points = 1 << 20;
  if (m->cfg.max_blocks <= 0) m->cfg.max_blocks = 1 << 18;  // 262144 blocks x 2 KB = 512 MB
  if (m->cfg.max_blocks > TS_MAX_BLOCKS) m->cfg.max_blocks = TS_MAX_BLOCKS;
  m->in.fx = (float)cfg->fx; m->in.fy = (float)cfg->fy; m->in.cx = (float)cfg->cx; m->in.cy = (float)cfg->cy;
  m->in.dmin_mm = (float)(cfg->min_ray_length * 1000.0);
  m->in.dmax_mm = (float)(cfg->max_ray_length * 1000.0);
  m->in.vs = (float)cfg->voxel_scale;
  m->in.step = cfg->recast_step;
  OcGrid& g = m->g;
  g.max_blocks = m->cfg.max_blocks;
  g.N = cfg->N; g.Nz = cfg->Nz; g.hN = cfg->N / 2; g.hNz = cfg->Nz / 2;
  m->table_cap = oc_next_pow2((size_t)g.max_blocks * 2 + 64);
  g.table_mask = (uint32_t)(m->table_cap - 1);
  TS_CUDA(cudaMalloc(&g.table, m->table_cap * 8));
  TS_CUDA(cudaMemset(g.table, 0xFF, m->table_cap * 8));
  TS_CUDA(cudaMalloc(&g.block_key, (size_t)g.max_blocks * 8));
  TS_CUDA(cudaMalloc(&g.cnt, (size_t)g.max_blocks * OC_B3 * 4));
  TS_CUDA(cudaMemset(g.cnt, 0, (size_t)g.max_blocks * OC_B3 * 4));
  g.cw = nullptr;
  if (cfg->texture_enabled) {
    TS_CUDA(cudaMalloc(&g.cw, (size_t)g.max_blocks * OC_B3 * 8));
    TS_CUDA(cudaMemset(g.cw, 0, (size_t)g.max_blocks * OC_B3 * 8));
    const size_t tb = (size_t)m->cfg.max_image_pixels > (size_t)m->cfg.max_points ? (size_t)m->cfg.max_image_pixels : (size_t)m->cfg.max_points;
    TS_CUDA(cudaMalloc(&m->tex_stage, tb * 3));
  }
  m->in.same_proj = 1;
  TS_CUDA(cudaMalloc(&m->scratch_i, 64 * 4));
  TS_CUDA(cudaMemset(m->scratch_i, 0, 64 * 4));
  g.n_blocks = m->scratch_i + 0;
  g.err = m->scratch_i + 2;
  TS_CUDA(cudaMalloc(&m->depth_stage, (size_t)m->cfg.max_image_pixels * 2));
  TS_CUDA(cudaMalloc(&m->points_stage, (size_t)m->cfg.max_points * 12));
  TS_CUDA(cudaMalloc(&m->pose_R, (size_t)m->cfg.max_submaps * 36));
  TS_CUDA(cudaMalloc(&m->pose_T, (size_t)m->cfg.max_submaps * 12));
  TS_CUDA(cudaMemset(m->pose_R, 0, (size_t)m->cfg.max_submaps * 36));
  TS_CUDA(cudaMemset(m->pose_T, 0, (size_t)m->cfg.max_submaps * 12));
  TS_CUDA(cudaDeviceSynchronize());
  *out = m;
  return TSLAM_OK;
}

extern "C" int tslam_octo_destroy(tslam_octo_t* m) {
  if (!m) return TSLAM_OK;
  cudaDeviceSynchronize();
  cudaFree(m->g.table); cudaFree(m->g.block_key); cudaFree(m->g.cnt); cudaFree(m->scratch_i);
  if (m->g.cw) cudaFree(m->g.cw);
  if (m->tex_stage) cudaFree(m->tex_stage);
  cudaFree(m->depth_stage); cudaFree(m->points_stage); cudaFree(m->pose_R); cudaFree(m->pose_T);
  delete m;
  return TSLAM_OK;
}

extern "C" int tslam_octo_reset(tslam_octo_t* m, void* stream) {
  if (!m) return TSLAM_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int nb = 0;
  TS_CUDA(cudaMemcpyAsync(&nb, m->g.n_blocks, 4, cudaMemcpyDeviceToHost, st));
  TS_CUDA(cudaStreamSynchronize(st));
  if (nb > m->g.max_blocks) nb = m->g.max_blocks;
  TS_CUDA(cudaMemsetAsync(m->g.table, 0xFF, m->table_cap * 8, st));
  if (nb) TS_CUDA(cudaMemsetAsync(m->g.cnt, 0, (size_t)nb * OC_B3 * 4, st));
  if (nb && m->g.cw) TS_CUDA(cudaMemsetAsync(m->g.cw, 0, (size_t)nb * OC_B3 * 8, st));
  TS_CUDA(cudaMemsetAsync(m->scratch_i, 0, 4 * 4, st));
  return TSLAM_OK;
}

extern "C" int tslam_octo_set_submap_pose(tslam_octo_t* m, int32_t s, const float* R9, const float* T3) {
  if (!m || !R9 || !T3 || s < 0 || s >= m->cfg.max_submaps) return TSLAM_E_INVALID;
  TS_CUDA(cudaMemcpy(m->pose_R + 9 * (size_t)s, R9, 36, cudaMemcpyHostToDevice));
  TS_CUDA(cudaMemcpy(m->pose_T + 3 * (size_t)s, T3, 12, cudaMemcpyHostToDevice));
  return TSLAM_OK;
}

extern "C" int tslam_octo_set_intrinsics(tslam_octo_t* m, double fx, double fy, double cx, double cy) {
  if (!m) return TSLAM_E_INVALID;
  m->cfg.fx = fx; m->cfg.fy = fy; m->cfg.cx = cx; m->cfg.cy = cy;
  m->in.fx = (float)fx; m->in.fy = (float)fy; m->in.cx = (float)cx; m->in.cy = (float)cy;
  return TSLAM_OK;
}

extern "C" int tslam_octo_set_color_intrinsics(tslam_octo_t* m, double fx, double fy, double cx, double cy, int color_same_proj) {
  if (!m) return TSLAM_E_INVALID;
  m->in.fxc = (float)fx; m->in.fyc = (float)fy; m->in.cxc = (float)cx; m->in.cyc = (float)cy;
  m->in.same_proj = color_same_proj ? 1 : 0;
  return TSLAM_OK;
}

static void oc_fill_frame(tslam_octo* m, TsFrame& fr, const float* R9, const float* T3, int submap) {
  memcpy(fr.R, R9, 36); memcpy(fr.T, T3, 12); fr.submap = submap;
  if (m->frame_seq < 0xFFFFFFFEu) m->frame_seq++;  // the colour word keeps 40 bits for it (seq << 24 | rgb): no practical limit
  fr.seq = m->frame_seq;
}

static int oc_deferred(tslam_octo* m) {
  int err = 0;
  TS_CUDA(cudaMemcpy(&err, m->g.err, 4, cudaMemcpyDeviceToHost));
  if (err) {
    int zero = 0;
    cudaMemcpy(m->g.err, &zero, 4, cudaMemcpyHostToDevice);
    ts_set_error("octomap block pool / table exhausted (max_blocks=%d, flags 0x%x): hits were dropped", m->g.max_blocks, err);
    return TSLAM_E_POOL_FULL;
  }
  return TSLAM_OK;
}

extern "C" int tslam_octo_integrate_points(tslam_octo_t* m, const float* xyz, int mem, int32_t n, const float* R9, const float* T3,
                                           int32_t submap, void* stream) {
  return tslam_octo_integrate_points_rgb(m, xyz, nullptr, mem, n, R9, T3, submap, stream);
}
extern "C" int tslam_octo_integrate_points_rgb(tslam_octo_t* m, const float* xyz, const uint8_t* rgb, int mem, int32_t n, const float* R9,
                                               const float* T3, int32_t submap, void* stream) {
  if (m && rgb && !m->g.cw) { ts_set_error("point colours given but the octomap was created with texture_enabled=0"); return TSLAM_E_INVALID; }
  if (!m || (!xyz && n > 0) || !R9 || !T3 || n < 0 || submap < 0 || submap >= m->cfg.max_submaps) return TSLAM_E_INVALID;
  if (n == 0) return TSLAM_OK;
  if (n > m->cfg.max_points && mem == TSLAM_MEM_HOST) { ts_set_error("n=%d exceeds max_points=%d", n, m->cfg.max_points); return TSLAM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const float* src = xyz;
  if (mem == TSLAM_MEM_HOST) {
    TS_CUDA(cudaMemcpyAsync(m->points_stage, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    src = m->points_stage;
  }
  const uint8_t* csrc = rgb;
  if (rgb && mem == TSLAM_MEM_HOST) {
    TS_CUDA(cudaMemcpyAsync(m->tex_stage, rgb, (size_t)n * 3, cudaMemcpyHostToDevice, st));
    csrc = m->tex_stage;
  }
  TsFrame fr;
  oc_fill_frame(m, fr, R9, T3, submap);
  k_octo_points<<<(n + 255) / 256, 256, 0, st>>>(m->g, src, n, fr, m->in.vs, csrc);
  m->launches++;
  TS_CUDA(cudaGetLastError());
  return TSLAM_OK;
}

extern "C" int tslam_octo_integrate_depth(tslam_octo_t* m, const uint16_t* depth, int mem, int32_t h, int32_t w, const float* R9,
                                          const float* T3, int32_t submap, void* stream) {
  return tslam_octo_integrate_depth_tex(m, depth, nullptr, mem, h, w, 0, 0, R9, T3, submap, stream);
}
extern "C" int tslam_octo_integrate_depth_tex(tslam_octo_t* m, const uint16_t* depth, const uint8_t* tex, int mem, int32_t h, int32_t w,
                                              int32_t th, int32_t tw, const float* R9, const float* T3, int32_t submap, void* stream) {
  if (m && tex) {
    if (!m->g.cw) { ts_set_error("colour image given but the octomap was created with texture_enabled=0"); return TSLAM_E_INVALID; }
    if (th <= 0 || tw <= 0 || (long long)th * tw > m->cfg.max_image_pixels) { ts_set_error("texture exceeds max_image_pixels"); return TSLAM_E_INVALID; }
  }
  if (!m || !depth || !R9 || !T3 || h <= 0 || w <= 0 || submap < 0 || submap >= m->cfg.max_submaps) return TSLAM_E_INVALID;
  if ((long long)h * w > m->cfg.max_image_pixels) { ts_set_error("frame exceeds max_image_pixels"); return TSLAM_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* src = depth;
  if (mem == TSLAM_MEM_HOST) {
    TS_CUDA(cudaMemcpyAsync(m->depth_stage, depth, (size_t)h * w * 2, cudaMemcpyHostToDevice, st));
    src = m->depth_stage;
  }
  const int step = m->cfg.recast_step;
  const int hh = (int)((double)h / step), ww = (int)((double)w / step);
  if (hh <= 0 || ww <= 0) return TSLAM_OK;
  const uint8_t* tsrc = tex;
  if (tex && mem == TSLAM_MEM_HOST) {
    TS_CUDA(cudaMemcpyAsync(m->tex_stage, tex, (size_t)th * tw * 3, cudaMemcpyHostToDevice, st));
    tsrc = m->tex_stage;
  }
  TsFrame fr;
  oc_fill_frame(m, fr, R9, T3, submap);
  k_octo_depth<<<(hh * ww + 255) / 256, 256, 0, st>>>(m->g, src, h, w, hh, ww, fr, m->in, tsrc, th, tw);
  m->launches++;
  TS_CUDA(cudaGetLastError());
  return TSLAM_OK;
}

extern "C" int tslam_octo_gather(tslam_octo_t* m, int32_t submap, int64_t cap, int32_t* idx, uint32_t* count, int64_t* n_out,
                                 void* stream) {
  return tslam_octo_gather2(m, submap, cap, idx, count, nullptr, n_out, stream);
}
extern "C" int tslam_octo_gather2(tslam_octo_t* m, int32_t submap, int64_t cap, int32_t* idx, uint32_t* count, float* color,
                                  int64_t* n_out, void* stream) {
  if (!m || !n_out) return TSLAM_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* ctr = (unsigned long long*)(m->scratch_i + 8);
  TS_CUDA(cudaMemsetAsync(ctr, 0, 8, st));
  k_octo_gather<<<m->sm_count * 8, 256, 0, st>>>(m->g, submap, cap, idx, count, color, ctr);
  m->launches++;
  TS_CUDA(cudaGetLastError());
  unsigned long long v = 0;
  TS_CUDA(cudaMemcpyAsync(&v, ctr, 8, cudaMemcpyDeviceToHost, st));
  TS_CUDA(cudaStreamSynchronize(st));
  *n_out = (int64_t)v;
  int rc = oc_deferred(m);
  if (rc) return rc;
  if ((int64_t)v > cap) { ts_set_error("octo gather: %lld voxels > capacity %lld", (long long)v, (long long)cap); return TSLAM_E_CAPACITY; }
  return TSLAM_OK;
}

extern "C" int tslam_octo_extract(tslam_octo_t* m, int32_t submap, int32_t level, int64_t cap, float* xyz, int32_t* count_dev,
                                  void* stream) {
  return tslam_octo_extract2(m, submap, level, cap, xyz, nullptr, count_dev, stream);
}
extern "C" int tslam_octo_extract2(tslam_octo_t* m, int32_t submap, int32_t level, int64_t cap, float* xyz, float* rgb,
                                   int32_t* count_dev, void* stream) {
  if (!m || !xyz || !count_dev || level < 1) return TSLAM_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  long long group = 1;
  for (int l = 1; l < level; l++) group *= m->cfg.K;
  if (group > (1 << 20)) group = 1 << 20;
  k_octo_extract<<<m->sm_count * 8, 256, 0, st>>>(m->g, submap, (int)group, (float)m->cfg.min_occupy_thres, m->pose_R, m->pose_T,
                                                  m->in.vs, cap, xyz, rgb, count_dev);
  m->launches++;
  TS_CUDA(cudaGetLastError());
  return TSLAM_OK;
}

extern "C" int tslam_octo_fuse(tslam_octo_t* dst, tslam_octo_t* src, void* stream) {
  if (!dst || !src || dst == src) return TSLAM_E_INVALID;
  int rc = tslam_octo_reset(dst, stream);  // taichi_octomap.py:195-196
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  k_octo_fuse<<<dst->sm_count * 8, 256, 0, st>>>(dst->g, src->g, (float)dst->cfg.min_occupy_thres, dst->pose_R, dst->pose_T,
                                                 dst->in.vs);
  dst->launches++;
  TS_CUDA(cudaGetLastError());
  return TSLAM_OK;
}

extern "C" int tslam_octo_sync(tslam_octo_t* m, void* stream) {
  if (!m) return TSLAM_E_INVALID;
  TS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return oc_deferred(m);
}
extern "C" int64_t tslam_octo_launch_count(tslam_octo_t* m) { return m ? m->launches : 0; }
