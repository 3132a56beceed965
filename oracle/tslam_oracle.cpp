// ============================================================================
// oracle/tslam_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the TaichiSLAM dense-mapping hot path,
// read off the reference sources; every function cites the reference
// file:line it follows (paths relative to /root/reference/taichi_slam/mapping).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library.  The product
// (taichislam_b200/) never links, imports or calls it.
//
// PARITY STATUS: *unpinned at the Taichi boundary* - Taichi itself cannot be
// imported or installed in the build container, the reference ships no golden
// vectors for integrate / octomap / marching cubes / ESDF and its kernels are
// racy (non-atomic RMW from concurrent rays, f16 atomics in arbitrary order), so
// the oracle is a *deterministic canonicalisation* of the source text.
// What IS pinned:
//   * by the reference's KERNEL SOURCE, EXECUTED: oracle/taichi_emu.py runs the
//     unmodified kernels of /root/reference in Python with Taichi's f16/f32 value
//     typing (one legal serial schedule); tools/make_golden_ref.py stores what they
//     compute (tests/golden/ref_exec.npz) and tests/test_oracle_vs_reference_exec.py
//     replays the inputs here: MODE_F16_FAITHFUL reproduces depth / point-cloud
//     integration with identical voxel sets and > 99.8 % bit-equal TSDF / W, the
//     exporters, load/save, marching cubes (same triangles), submap fusion (same
//     voxels, same NaN pattern) and the Octomap counts exactly.  ESDF (dead code in
//     the reference) stays unpinned.  Residual: the emulation is not Taichi's LLVM
//     backend (no FMA contraction, no races).
//   * by reference artefacts: the two exported maps in data/*.npy (schema, counts,
//     load->export round trip), the marching-cubes case tables (sha256), and the
//     host-side code of the classes run directly (tests/golden/host_reference.json).
//
// Numeric modes (TSDF integrate):
//   MODE_CANONICAL (0)  the mode the CUDA path is compared against bit-for-bit
//       on indices/flags and to 1e-4 on values: per-frame bucket sums are
//       accumulated EXACTLY (2^-20 m fixed point, order independent - the
//       reference uses racy f16 atomics there, dense_tsdf.py:230-232), all
//       other arithmetic is strict IEEE f32 in source order (no FMA
//       contraction), the per-voxel weighted average of one frame is applied
//       as one commit  T' = (T*W + sum w*d)/(W + sum w),  W' = min(W+sum w,1000)
//       which equals the reference's sequential RMW (dense_tsdf.py:264-267)
//       in exact arithmetic whenever the Wmax clamp does not bind mid-frame.
//   MODE_F32_LITERAL (1)  literal sequential restatement with f32 state:
//       f32 bucket sums in row-major pixel order, per-sample RMW in
//       lexicographic bucket order.
//   MODE_F16_FAITHFUL (2)  as (1) but every value the reference types f16
//       (dense_tsdf.py:64-66,92-93) is rounded to IEEE binary16 after every
//       op, following SURVEY Appendix A.1/A.2.
// Modes 1/2 exist to bound the distance between the canonical mode and the
// literal source text (tests/test_oracle_modes.py); they are never compared
// with the GPU bit-for-bit.
//
// Build: oracle/Makefile  (g++ -O2 -ffp-contract=off, no -ffast-math).
// ============================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <thread>
#include "../include/tslam_mc_cases.h"

namespace {

constexpr int OB = 8;            // oracle block edge (deliberately != the GPU's 16)
constexpr int OB3 = OB * OB * OB;
constexpr float WMAX = 1000.0f;  // dense_tsdf.py:8
constexpr double FIX = 1048576.0;  // 2^20 fixed point quantum for canonical bucket sums

enum { MODE_CANONICAL = 0, MODE_F32_LITERAL = 1, MODE_F16_FAITHFUL = 2 };

inline float h16(float x) { return (float)(_Float16)x; }  // round to binary16 (RN-even)

inline int iround(float x) { return (int)roundf(x); }  // ti.round(x, ti.i32): mapping_common.py:263-266
inline int sgn(float v) { return (0.0f < v) - (v < 0.0f); }  // mapping_common.py:5-7

inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int fmod_(int a, int b) { return a - fdiv(a, b) * b; }

struct Key {
  int s, x, y, z;
  bool operator==(const Key& o) const { return s == o.s && x == o.x && y == o.y && z == o.z; }
  bool operator<(const Key& o) const {
    if (s != o.s) return s < o.s;
    if (x != o.x) return x < o.x;
    if (y != o.y) return y < o.y;
    return z < o.z;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = (uint64_t)(uint32_t)k.s * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)k.x * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (uint64_t)(uint32_t)k.y * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (uint64_t)(uint32_t)k.z * 0x27D4EB2F165667C5ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

// One OB^3 voxel block: the four fields the reference places in the same dense
// cell (dense_tsdf.py:92-101) + the canonical per-frame accumulators + ESDF state
// (dense_esdf.py:88-92).
struct Block {
  float T[OB3], W[OB3];
  double A[OB3], Bw[OB3];  // pending sum(w*d), sum(w) (canonical mode): EXACT-ish double sums of the f32 terms, order independent
  uint8_t obs[OB3];
  int occ[OB3];           // i8 in the reference (dense_tsdf.py:95); wider here, saturated on export
  float esdf[OB3];
  uint8_t eobs[OB3];
  uint64_t cword[OB3];    // texture: winning (frame seq | closeness | rgb10) word per voxel (see color_word)
  float col[OB3][3];      // texture: committed colour
  bool pending;           // has non-zero accumulators (listed in Tsdf::dirty)
  Block() { memset(this, 0, sizeof(*this)); }
};

struct Pose {
  float R[9];
  float T[3];
  Pose() : R{1, 0, 0, 0, 1, 0, 0, 0, 1}, T{0, 0, 0} {}
};

// rows of the submap pose tables start at ZERO (ti fields, mapping_common.py:104-105)
static inline Pose pose_or_zero(const std::unordered_map<int, Pose>& t, int s) {
  auto it = t.find(s);
  if (it != t.end()) return it->second;
  Pose p;
  memset(p.R, 0, sizeof(p.R));
  return p;
}

struct TsdfCfg {  // mirrors the C struct in tslam_oracle.h (ctypes side)
  double voxel_scale;
  int N, Nz;
  double max_ray_length, min_ray_length;
  int internal_voxels;
  int recast_step;
  double fx, fy, cx, cy;
  int mode;
  int is_global_map;
  double disp_floor, disp_ceiling;
};

struct Stats {
  int64_t n_px, n_valid, n_rays, n_updates, n_oob;
};

struct Tsdf {
  TsdfCfg c;
  float vs;
  std::unordered_map<Key, Block*, KeyHash> blocks;
  std::vector<Block*> dirty;  // blocks with pending accumulators
  std::unordered_map<int, Pose> submap_pose;
  Stats st{};
  float colormap[1024][3];
  // texture path (dense_tsdf.py:75-77,96-103): colour intrinsics, projection mode, frame sequence number
  bool tex_enabled = false;
  bool color_same_proj = true;
  float fxc = 1, fyc = 1, cxc = 0, cyc = 0;
  uint32_t frame_seq = 0;

  ~Tsdf() {
    for (auto& kv : blocks) delete kv.second;
  }
  bool in_bounds(int i, int j, int k) const {
    // field offset [0,-N/2,-N/2,-Nz/2] (dense_tsdf.py:90) -> valid i in [-N/2, N/2).
    // The reference does not clamp (mapping_common.py:263-266: UB when outside);
    // canonical behaviour = skip the sample.
    int h = c.N / 2, hz = c.Nz / 2;
    return i >= -h && i < c.N - h && j >= -h && j < c.N - h && k >= -hz && k < c.Nz - hz;
  }
  Block* find(int s, int i, int j, int k, int* off) const {
    Key key{s, fdiv(i, OB), fdiv(j, OB), fdiv(k, OB)};
    auto it = blocks.find(key);
    if (it == blocks.end()) return nullptr;
    *off = (fmod_(i, OB) * OB + fmod_(j, OB)) * OB + fmod_(k, OB);
    return it->second;
  }
  Block* touch(int s, int i, int j, int k, int* off) {  // write activates a zero-filled block
    Key key{s, fdiv(i, OB), fdiv(j, OB), fdiv(k, OB)};
    auto it = blocks.find(key);
    Block* b;
    if (it == blocks.end()) {
      b = new Block();
      blocks.emplace(key, b);
    } else {
      b = it->second;
    }
    *off = (fmod_(i, OB) * OB + fmod_(j, OB)) * OB + fmod_(k, OB);
    return b;
  }
  // reads of inactive cells return 0 (Taichi pointer-SNode semantics)
  float readT(int s, int i, int j, int k) const {
    int o;
    Block* b = find(s, i, j, k, &o);
    return b ? b->T[o] : 0.0f;
  }
  int readObs(int s, int i, int j, int k) const {
    int o;
    Block* b = find(s, i, j, k, &o);
    return b ? b->obs[o] : 0;
  }
};

// ---------------------------------------------------------------------------
// jet colour LUT.  The reference fills colormap[i] = matplotlib.cm.jet(i/1024)
// (mapping_common.py:158-163).  matplotlib is absent here; cm.jet is the
// 256-entry LUT of the piecewise-linear "jet" segment data, so
// colormap[i] = LUT[i//4].  (adjacent/export path, not value-pinned)
// ---------------------------------------------------------------------------
static float seg_interp(const float (*d)[2], int n, float x) {
  for (int i = 1; i < n; i++)
    if (x <= d[i][0]) {
      float t = (x - d[i - 1][0]) / (d[i][0] - d[i - 1][0]);
      return d[i - 1][1] + t * (d[i][1] - d[i - 1][1]);
    }
  return d[n - 1][1];
}
static void fill_jet(float cm[1024][3]) {
  static const float r[][2] = {{0, 0}, {0.35f, 0}, {0.66f, 1}, {0.89f, 1}, {1, 0.5f}};
  static const float g[][2] = {{0, 0}, {0.125f, 0}, {0.375f, 1}, {0.64f, 1}, {0.91f, 0}, {1, 0}};
  static const float b[][2] = {{0, 0.5f}, {0.11f, 1}, {0.34f, 1}, {0.65f, 0}, {1, 0}};
  for (int i = 0; i < 1024; i++) {
    float x = (float)(i / 4) / 255.0f;
    cm[i][0] = seg_interp(r, 5, x);
    cm[i][1] = seg_interp(g, 6, x);
    cm[i][2] = seg_interp(b, 5, x);
  }
}

// ---------------------------------------------------------------------------
// Per-frame bucket grid  (dense_tsdf.py:64-70: new_pcl_count i32,
// new_pcl_sum_pos 3xf16, new_pcl_z f16; keyed by round(p/vs)).
// Points are appended in pixel order; finalize() stable-sorts them by cell and folds every cell's points in
// that (pixel) order, so iteration is lexicographic in (i,j,k) = the canonical bucket order, exactly what a
// std::map keyed by the cell would give - without one heap node per cell (the multi-threaded CPU baseline
// spent its time in malloc).
// ---------------------------------------------------------------------------
struct Bucket {
  int count = 0;
  int64_t fx = 0, fy = 0, fz = 0, fd = 0;  // canonical exact sums
  float sx = 0, sy = 0, sz = 0, sd = 0;    // literal float sums (f32 or f16-rounded)
  int64_t cr = 0, cg = 0, cb = 0;          // new_pcl_sum_color (dense_tsdf.py:76,234): exact integer channel sums
};
struct BucketGrid {
  struct Raw { std::array<int, 3> key; float p[3]; float z; int rgb[3]; };
  std::vector<Raw> raw;
  std::vector<std::pair<std::array<int, 3>, Bucket>> cells;
  void finalize(int mode) {
    std::vector<uint32_t> ord(raw.size());
    for (uint32_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return raw[a].key < raw[b].key; });
    cells.clear();
    for (uint32_t o : ord) {
      const Raw& r = raw[o];
      if (cells.empty() || cells.back().first != r.key) cells.emplace_back(r.key, Bucket());
      Bucket& q = cells.back().second;
      q.count += 1;
      q.cr += r.rgb[0]; q.cg += r.rgb[1]; q.cb += r.rgb[2];
      if (mode == MODE_CANONICAL) {
        q.fx += llrintf(r.p[0] * (float)FIX);
        q.fy += llrintf(r.p[1] * (float)FIX);
        q.fz += llrintf(r.p[2] * (float)FIX);
        q.fd += llrintf(r.z * (float)FIX);
      } else if (mode == MODE_F32_LITERAL) {
        q.sx += r.p[0];
        q.sy += r.p[1];
        q.sz += r.p[2];
        q.sd += r.z;
      } else {  // f16 atomic add: operand cast to f16, sum rounded to f16
        q.sx = h16(q.sx + h16(r.p[0]));
        q.sy = h16(q.sy + h16(r.p[1]));
        q.sz = h16(q.sz + h16(r.p[2]));
        q.sd = h16(q.sd + h16(r.z));
      }
    }
  }
  std::vector<std::pair<std::array<int, 3>, Bucket>>::iterator begin() { return cells.begin(); }
  std::vector<std::pair<std::array<int, 3>, Bucket>>::iterator end() { return cells.end(); }
};

// process_point  (dense_tsdf.py:227-234)
static void bucket_add(const Tsdf& m, BucketGrid& g, const float p[3], float z, const int* rgb = nullptr) {
  const float vs = m.vs;
  BucketGrid::Raw r;
  r.rgb[0] = rgb ? rgb[0] : 0; r.rgb[1] = rgb ? rgb[1] : 0; r.rgb[2] = rgb ? rgb[2] : 0;
  r.key = {iround(p[0] / vs), iround(p[1] / vs), iround(p[2] / vs)};  // xyz_to_ijk mapping_common.py:240-243
  r.p[0] = p[0]; r.p[1] = p[1]; r.p[2] = p[2];
  r.z = z;
  g.raw.push_back(r);
}

static inline void rot(const float R[9], const float v[3], float o[3]) {  // input_R[None] @ pt
  o[0] = (R[0] * v[0] + R[1] * v[1]) + R[2] * v[2];
  o[1] = (R[3] * v[0] + R[4] * v[1]) + R[5] * v[2];
  o[2] = (R[6] * v[0] + R[7] * v[1]) + R[8] * v[2];
}

// Texture.  The reference overwrites color[xi] with the ray's mean colour at every sample (dense_tsdf.py:268-269):
// racing rays, last writer wins.  Canonical rule: the winner is the sample with the largest word
//   [ frame sequence number : 22 | closeness to the ray's surface point, 4095 - min(4095, |ds|/vs*16) : 12 | rgb 3x10 ]
// i.e. later frames overwrite earlier ones (as in the reference) and, inside a frame, the sample nearest to its
// surface point wins (then the larger packed colour).  Colours are the bucket mean / 255 quantised to 10 bits
// (the reference stores f16: 11-bit significand).
static inline uint64_t color_word(uint32_t seq, float ds, float vs, const Bucket& q) {
  const float c = (float)q.count;
  auto q10 = [&](int64_t sum) { float v = ((float)sum / c) / 255.0f; int k = (int)(v * 1023.0f + 0.5f); return (uint64_t)(k < 0 ? 0 : (k > 1023 ? 1023 : k)); };
  int cl = (int)(fabsf(ds) / vs * 16.0f);
  if (cl > 4095) cl = 4095;
  return ((uint64_t)seq << 42) | ((uint64_t)(4095 - cl) << 30) | (q10(q.cr) << 20) | (q10(q.cg) << 10) | q10(q.cb);
}

// process_new_pcl  (dense_tsdf.py:236-270) - the ray-march hot loop.
static void raymarch(Tsdf& m, BucketGrid& g, const float Tin[3], int s) {
  const float vs = m.vs;
  const int mode = m.c.mode;
  const float max_steps = (float)(m.c.max_ray_length / m.c.voxel_scale);  // python-float constant, dense_tsdf.py:249
  g.finalize(mode);
  for (auto& kv : g) {
    Bucket& q = kv.second;
    if (q.count == 0) continue;  // :240
    m.st.n_rays++;
    float mx, my, mz, z, L, ux, uy, uz;
    if (mode == MODE_CANONICAL) {
      double den = (double)q.count * FIX;
      mx = (float)((double)q.fx / den);
      my = (float)((double)q.fy / den);
      mz = (float)((double)q.fz / den);
      z = (float)((double)q.fd / den);
      L = sqrtf((mx * mx + my * my) + mz * mz);
      ux = mx / L; uy = my / L; uz = mz / L;
    } else if (mode == MODE_F32_LITERAL) {
      float c = (float)q.count;
      mx = q.sx / c; my = q.sy / c; mz = q.sz / c;  // :243
      L = sqrtf((mx * mx + my * my) + mz * mz);     // :244
      ux = mx / L; uy = my / L; uz = mz / L;         // :245
      z = q.sd / c;                                  // :247
    } else {
      float c = h16((float)q.count);  // :242 ti.cast(count, f16)
      mx = h16(q.sx / c); my = h16(q.sy / c); mz = h16(q.sz / c);
      L = h16(sqrtf(h16(h16(h16(mx * mx) + h16(my * my)) + h16(mz * mz))));
      ux = h16(mx / L); uy = h16(my / L); uz = h16(mz / L);
      z = h16(q.sd / c);
    }
    if (!(L > 0.0f)) { q.count = 0; continue; }  // canonical: degenerate bucket (reference: NaN indices)
    const float Px = mx + Tin[0], Py = my + Tin[1], Pz = mz + Tin[2];  // :246
    {  // :248  occupy[sxyz_to_ijk(P)] = 1
      int oi = iround(Px / vs), oj = iround(Py / vs), ok = iround(Pz / vs);
      if (m.in_bounds(oi, oj, ok)) {
        int o; Block* b = m.touch(s, oi, oj, ok, &o);
        b->occ[o] = 1;
      }
    }
    int n = (int)fminf(L / vs + (float)m.c.internal_voxels, max_steps);  // :249-251 range(float) truncates
    float w;
    if (mode == MODE_F16_FAITHFUL) w = 1.0f / h16(z * z); else w = 1.0f / (z * z);  // w_x_p :216-225 with d>=0
    float jf = 0.0f;
    Block* last_b = nullptr;
    Key last_key{0, 0, 0, 0};
    for (int it = 0; it < n; it++) {
      jf += 1.0f;  // :252
      float x = (ux * jf) * vs + Tin[0], y = (uy * jf) * vs + Tin[1], zz = (uz * jf) * vs + Tin[2];  // :253
      int xi = iround(x / vs), yi = iround(y / vs), zi = iround(zz / vs);  // :254
      float vx = Px - x, vy = Py - y, vz = Pz - zz;  // :258
      float d = sqrtf((vx * vx + vy * vy) + vz * vz);  // :259
      float ds = d * (float)sgn((vx * mx + vy * my) + vz * mz);  // :260
      if (!m.in_bounds(xi, yi, zi)) { m.st.n_oob++; continue; }
      m.st.n_updates++;
      int o; Block* b;
      {  // consecutive samples of a ray mostly stay in one block: skip the hash lookup then
        const Key bk{s, fdiv(xi, OB), fdiv(yi, OB), fdiv(zi, OB)};
        if (last_b && bk == last_key) { b = last_b; o = (fmod_(xi, OB) * OB + fmod_(yi, OB)) * OB + fmod_(zi, OB); }
        else { b = m.touch(s, xi, yi, zi, &o); last_b = b; last_key = bk; }
      }
      if (mode == MODE_CANONICAL) {
        if (!b->pending) { b->pending = true; m.dirty.push_back(b); }
        b->A[o] += (double)(w * ds);
        b->Bw[o] += (double)w;
        if (m.tex_enabled) { const uint64_t cw = color_word(m.frame_seq, ds, vs, q); if (cw > b->cword[o]) b->cword[o] = cw; }  // :268-269
      } else if (mode == MODE_F32_LITERAL) {
        float T0 = b->T[o], W0 = b->W[o];
        b->T[o] = (T0 * W0 + w * ds) / (W0 + w);  // :264
        b->obs[o] = 1;                             // :265
        b->W[o] = fminf(W0 + w, WMAX);             // :267
      } else {
        float T0 = b->T[o], W0 = b->W[o];
        b->T[o] = h16((h16(T0 * W0) + w * ds) / (W0 + w));
        b->obs[o] = 1;
        b->W[o] = h16(fminf(W0 + w, WMAX));
      }
    }
    q.count = 0;  // :270
  }
}

// canonical commit of pending accumulators (see header).
static void commit(Tsdf& m, bool clamp) {
  for (Block* b : m.dirty) {
    b->pending = false;
    for (int o = 0; o < OB3; o++) {
      if (b->Bw[o] > 0.0) {
        float T0 = b->T[o], W0 = b->W[o];
        const float Af = (float)b->A[o], Bf = (float)b->Bw[o];  // exact sums, rounded once
        float Wn = W0 + Bf;
        b->T[o] = (T0 * W0 + Af) / Wn;
        b->W[o] = clamp ? fminf(Wn, WMAX) : Wn;
        b->obs[o] = 1;
        b->A[o] = 0.0f;
        b->Bw[o] = 0.0f;
        if (m.tex_enabled && b->cword[o]) {
          b->col[o][0] = (float)((b->cword[o] >> 20) & 1023) / 1023.0f;
          b->col[o][1] = (float)((b->cword[o] >> 10) & 1023) / 1023.0f;
          b->col[o][2] = (float)(b->cword[o] & 1023) / 1023.0f;
        }
      }
    }
  }
  m.dirty.clear();
}

// ---------------------------------------------------------------------------
// Marching cubes tables (marching_cube_mesher.py:196-241, :244-499)
// ---------------------------------------------------------------------------
static const uint64_t MC_CASES[256] = TSLAM_MC_CASE_WORDS;
static const int GRID[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};  // :196-206
static const int EDGE[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};  // :208-221
static inline int mc_tri(int ci, int q) { int v = (int)((MC_CASES[ci] >> (4 * q)) & 0xF); return v == 0xF ? -1 : v; }
static inline int mc_edge_mask(int ci) {
  int m = 0;
  for (int q = 0; q < 16; q++) { int e = mc_tri(ci, q); if (e >= 0) m |= 1 << e; }
  return m;
}

// ---------------------------------------------------------------------------
// Octomap = per-voxel hit counter (taichi_octomap.py:116-119)
// ---------------------------------------------------------------------------
struct Octo {
  double voxel_scale;  // the CONSTRUCTOR's voxel_scale (kernels use voxel_scale_, mapping_common.py:22-23)
  float vs;
  int N, Nz, K;
  double max_ray, min_ray;
  int step;
  double fx, fy, cx, cy;
  int min_occupy_thres;
  std::map<Key, uint32_t> cnt;  // (s,i,j,k) -> hits; std::map => sorted export
  // texture (taichi_octomap.py:77-79, :120-124): the reference overwrites color[ijk] per point, racing, last writer
  // wins.  Canonical: per voxel the largest word [integrate-call sequence number : 22 | r:8 | g:8 | b:8] (colour
  // after the BGR->RGB swap of :121-124) - the latest call wins, inside a call the largest packed colour.
  bool tex_enabled = false, color_same_proj = true;
  float fxc = 1, fyc = 1, cxc = 0, cyc = 0;
  uint32_t frame_seq = 0;
  std::map<Key, uint64_t> cw;
  void hit(const Key& k, uint32_t add, uint64_t word) {
    cnt[k] += add;
    if (tex_enabled && word) { uint64_t& w = cw[k]; if (word > w) w = word; }
  }
  std::unordered_map<int, Pose> submap_pose;
  bool in_bounds(int i, int j, int k) const {
    int h = N / 2, hz = Nz / 2;
    return i >= -h && i < N - h && j >= -h && j < N - h && k >= -hz && k < Nz - hz;
  }
};

}  // namespace

// ============================================================================
// C ABI (loaded by tests/ through ctypes; see oracle/oracle.py)
// ============================================================================
extern "C" {

// the jet LUT on its own (tests/test_oracle_cpu.py::test_jet_colormap_golden pins it to tests/golden/jet_1024.json)
void orc_colormap(float* out3072) {
  static float cm[1024][3];
  fill_jet(cm);
  memcpy(out3072, cm, sizeof(cm));
}

void* orc_tsdf_create(const TsdfCfg* cfg) {
  Tsdf* m = new Tsdf();
  m->c = *cfg;
  m->vs = (float)cfg->voxel_scale;
  fill_jet(m->colormap);
  return m;
}
void orc_tsdf_destroy(void* h) { delete (Tsdf*)h; }

// reset(): B.parent().deactivate_all()  (dense_tsdf.py:309-310)
void orc_tsdf_reset(void* h) {
  Tsdf* m = (Tsdf*)h;
  for (auto& kv : m->blocks) delete kv.second;
  m->blocks.clear();
  m->dirty.clear();
}

void orc_tsdf_set_submap_pose(void* h, int s, const float* R9, const float* T3) {  // mapping_common.py:121-131
  Tsdf* m = (Tsdf*)h;
  Pose p;
  memcpy(p.R, R9, sizeof(p.R));
  memcpy(p.T, T3, sizeof(p.T));
  m->submap_pose[s] = p;
}

void orc_tsdf_get_stats(void* h, int64_t* out5) {
  Tsdf* m = (Tsdf*)h;
  out5[0] = m->st.n_px; out5[1] = m->st.n_valid; out5[2] = m->st.n_rays; out5[3] = m->st.n_updates; out5[4] = m->st.n_oob;
}
void orc_tsdf_clear_stats(void* h) { ((Tsdf*)h)->st = Stats{}; }

// recast_depth_to_map_kernel  (dense_tsdf.py:188-214) with unproject_point_dep
// (mapping_common.py:31-41).  R9/T3 = input_R/input_T AFTER set_pose's
// convert_by_base + f32 cast (mapping_common.py:149-156) - host plumbing is the caller's.
void orc_tsdf_integrate_depth_tex(void* h, const uint16_t* depth, const uint8_t* tex, int TH, int TW, int H, int Wd, const float* R9,
                                  const float* T3, int submap, int do_commit);
void orc_tsdf_integrate_depth(void* h, const uint16_t* depth, int H, int Wd, const float* R9, const float* T3, int submap, int do_commit) {
  orc_tsdf_integrate_depth_tex(h, depth, nullptr, 0, 0, H, Wd, R9, T3, submap, do_commit);
}
void orc_tsdf_set_color(void* h, int enabled, int same_proj, double fx, double fy, double cx, double cy) {
  Tsdf* m = (Tsdf*)h;
  m->tex_enabled = enabled != 0; m->color_same_proj = same_proj != 0;
  m->fxc = (float)fx; m->fyc = (float)fy; m->cxc = (float)cx; m->cyc = (float)cy;
}
void orc_tsdf_integrate_depth_tex(void* h, const uint16_t* depth, const uint8_t* tex, int TH, int TW, int H, int Wd, const float* R9,
                                  const float* T3, int submap, int do_commit) {
  Tsdf* m = (Tsdf*)h;
  if (m->frame_seq < (1u << 22) - 1) m->frame_seq++;
  BucketGrid g;
  const int step = m->c.recast_step;
  const int hh = (int)((double)H / step), ww = (int)((double)Wd / step);  // range(0, h/step): float bound truncated
  const float fx = (float)m->c.fx, fy = (float)m->c.fy, cx = (float)m->c.cx, cy = (float)m->c.cy;
  const float dmax = (float)(m->c.max_ray_length * 1000.0), dmin = (float)(m->c.min_ray_length * 1000.0);
  for (int jj = 0; jj < hh; jj++) {
    int j = jj * step;
    for (int ii = 0; ii < ww; ii++) {
      int i = ii * step;
      m->st.n_px++;
      uint16_t d = depth[(size_t)j * Wd + i];
      if (d == 0) continue;                                  // :196
      if ((float)d > dmax || (float)d < dmin) continue;      // :198
      m->st.n_valid++;
      float dep = (float)d / 1000.0f;                        // :201
      float pt[3] = {((float)i - cx) * dep / fx, ((float)j - cy) * dep / fy, dep};  // mapping_common.py:37-40
      float p[3];
      rot(R9, pt, p);                                        // :203 (rotation only)
      int rgb[3] = {0, 0, 0};
      if (m->tex_enabled && tex) {
        int tj = j, ti = i;                                  // :206 color_same_proj: texture[j, i]
        if (!m->color_same_proj) {                           // :209 color_ind_from_depth_pt (mapping_common.py:43-58)
          ti = (int)((((float)i - cx) / fx) * m->fxc + m->cxc);
          tj = (int)((((float)j - cy) / fy) * m->fyc + m->cyc);
          // the reference tests color_i against h and color_j against w (swapped, :56); indices that pass that test
          // but fall outside the texture are an out-of-bounds read there - canonical: pixel (0,0) as well
          if (ti < 0 || ti >= TH || tj < 0 || tj >= TW || tj >= TH || ti >= TW) { ti = 0; tj = 0; }
        }
        if (tj < TH && ti < TW) { const uint8_t* px = tex + ((size_t)tj * TW + ti) * 3; rgb[0] = px[0]; rgb[1] = px[1]; rgb[2] = px[2]; }
      }
      bucket_add(*m, g, p, dep, rgb);                        // :207/:211/:213 process_point(pt_map, dep[, color])
    }
  }
  raymarch(*m, g, T3, submap);                               // :214
  if (m->c.mode == MODE_CANONICAL && do_commit) commit(*m, true);
}

// recast_pcl_to_map_kernel  (dense_tsdf.py:167-186)
void orc_tsdf_integrate_points_rgb(void* h, const float* xyz, const uint8_t* rgb, int n, const float* R9, const float* T3, int submap,
                                   int do_commit);
void orc_tsdf_integrate_points(void* h, const float* xyz, int n, const float* R9, const float* T3, int submap, int do_commit) {
  orc_tsdf_integrate_points_rgb(h, xyz, nullptr, n, R9, T3, submap, do_commit);
}
void orc_tsdf_integrate_points_rgb(void* h, const float* xyz, const uint8_t* rgbs, int n, const float* R9, const float* T3, int submap,
                                   int do_commit) {
  Tsdf* m = (Tsdf*)h;
  if (m->frame_seq < (1u << 22) - 1) m->frame_seq++;
  BucketGrid g;
  const float maxr = (float)m->c.max_ray_length;
  for (int idx = 0; idx < n; idx++) {
    m->st.n_px++;
    float pt[3] = {xyz[3 * idx], xyz[3 * idx + 1], xyz[3 * idx + 2]};
    float p[3];
    rot(R9, pt, p);                                          // :175
    float len = sqrtf((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]);  // :176
    if (len < maxr) {                                        // :177
      m->st.n_valid++;
      int rgb[3] = {0, 0, 0};
      if (m->tex_enabled && rgbs) { rgb[0] = rgbs[3 * idx]; rgb[1] = rgbs[3 * idx + 1]; rgb[2] = rgbs[3 * idx + 2]; }  // :179-183
      bucket_add(*m, g, p, len, rgb);                        // :183/:185 process_point(pt, pt.norm()[, rgb])
    }
  }
  raymarch(*m, g, T3, submap);
  if (m->c.mode == MODE_CANONICAL && do_commit) commit(*m, true);
}

void orc_tsdf_commit(void* h) { commit(*(Tsdf*)h, true); }

// count_active  (dense_tsdf.py:412-423)
int64_t orc_tsdf_count_active(void* h, int submap) {
  Tsdf* m = (Tsdf*)h;
  int64_t n = 0;
  for (auto& kv : m->blocks)
    if (kv.first.s == submap)
      for (int o = 0; o < OB3; o++) n += kv.second->obs[o] > 0;
  return n;
}

// to_numpy  (dense_tsdf.py:425-440); rows sorted lexicographically by (i,j,k).
int64_t orc_tsdf_gather(void* h, int submap, int64_t cap, int32_t* idx, float* tsdf, float* wts, int32_t* occ) {
  Tsdf* m = (Tsdf*)h;
  std::vector<Key> keys;
  for (auto& kv : m->blocks) if (kv.first.s == submap) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  struct Row { int i, j, k; float t, w; int occ; };
  std::vector<Row> rows;
  for (auto& key : keys) {
    Block* b = m->blocks[key];
    for (int o = 0; o < OB3; o++)
      if (b->obs[o] > 0) {
        int lx = o / (OB * OB), ly = (o / OB) % OB, lz = o % OB;
        rows.push_back({key.x * OB + lx, key.y * OB + ly, key.z * OB + lz, b->T[o], b->W[o], b->occ[o]});
      }
  }
  std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
    if (a.i != b.i) return a.i < b.i;
    if (a.j != b.j) return a.j < b.j;
    return a.k < b.k;
  });
  int64_t n = 0;
  for (auto& r : rows) {
    if (n < cap) {
      idx[3 * n] = r.i; idx[3 * n + 1] = r.j; idx[3 * n + 2] = r.k;
      tsdf[n] = r.t; wts[n] = r.w; occ[n] = r.occ;
    }
    n++;
  }
  return n;
}

// colours of the observed voxels, same row order as orc_tsdf_gather (to_numpy data_color, dense_tsdf.py:437-440)
int64_t orc_tsdf_gather_color(void* h, int submap, int64_t cap, float* col3) {
  Tsdf* m = (Tsdf*)h;
  struct Row { int i, j, k; float c[3]; };
  std::vector<Row> rows;
  for (auto& kv : m->blocks) {
    if (kv.first.s != submap) continue;
    Block* b = kv.second;
    for (int o = 0; o < OB3; o++)
      if (b->obs[o] > 0) rows.push_back({kv.first.x * OB + o / (OB * OB), kv.first.y * OB + (o / OB) % OB, kv.first.z * OB + o % OB,
                                         {b->col[o][0], b->col[o][1], b->col[o][2]}});
  }
  std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
    if (a.i != b.i) return a.i < b.i;
    if (a.j != b.j) return a.j < b.j;
    return a.k < b.k;
  });
  int64_t n = 0;
  for (auto& r : rows) {
    if (n < cap) { col3[3 * n] = r.c[0]; col3[3 * n + 1] = r.c[1]; col3[3 * n + 2] = r.c[2]; }
    n++;
  }
  return n;
}
void orc_tsdf_scatter_color(void* h, int submap, int64_t n, const int32_t* idx, const float* col3) {  // load_numpy :450-453
  Tsdf* m = (Tsdf*)h;
  for (int64_t r = 0; r < n; r++) {
    int i = idx[3 * r], j = idx[3 * r + 1], k = idx[3 * r + 2];
    if (!m->in_bounds(i, j, k)) continue;
    int o; Block* b = m->touch(submap, i, j, k, &o);
    b->col[o][0] = col3[3 * r]; b->col[o][1] = col3[3 * r + 1]; b->col[o][2] = col3[3 * r + 2];
  }
}

// load_numpy  (dense_tsdf.py:442-454)
void orc_tsdf_scatter(void* h, int submap, int64_t n, const int32_t* idx, const float* tsdf, const float* wts, const int32_t* occ) {
  Tsdf* m = (Tsdf*)h;
  for (int64_t r = 0; r < n; r++) {
    int i = idx[3 * r], j = idx[3 * r + 1], k = idx[3 * r + 2];
    if (!m->in_bounds(i, j, k)) continue;
    int o; Block* b = m->touch(submap, i, j, k, &o);
    b->T[o] = tsdf[r]; b->W[o] = wts[r]; b->occ[o] = occ[r]; b->obs[o] = 1;
  }
}

// fuse_submaps_kernel + fuse_with_interploation  (dense_tsdf.py:272-307), preceded by
// reset() (:312-313).  Poses: dst's submap table (set by set_base_pose_submap on the
// global map, submap_mapping.py:137).  Source voxels visited in lexicographic order.
void orc_tsdf_fuse(void* hdst, void* hsrc) {
  Tsdf* D = (Tsdf*)hdst;
  Tsdf* S = (Tsdf*)hsrc;
  orc_tsdf_reset(D);
  const float vs = D->vs;
  std::vector<Key> keys;
  for (auto& kv : S->blocks) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  for (auto& key : keys) {
    Block* sb = S->blocks[key];
    Pose P = pose_or_zero(D->submap_pose, key.s);
    for (int o = 0; o < OB3; o++) {
      if (!(sb->obs[o] > 0)) continue;  // :292
      int i = key.x * OB + o / (OB * OB), j = key.y * OB + (o / OB) % OB, k = key.z * OB + o % OB;
      float l[3] = {(float)i * vs, (float)j * vs, (float)k * vs};  // ijk_to_xyz mapping_common.py:221-223
      float r[3];
      rot(P.R, l, r);
      float g[3] = {(r[0] + P.T[0]) / vs, (r[1] + P.T[1]) / vs, (r[2] + P.T[2]) / vs};  // :293-294
      int lo[3] = {(int)floorf(g[0]), (int)floorf(g[1]), (int)floorf(g[2])};  // :296
      for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
          for (int dk = 0; dk < 2; dk++) {
            if (di + dj + dk == 0) continue;  // :300 (reference quirk: low corner skipped)
            int c[3] = {lo[0] + di, lo[1] + dj, lo[2] + dk};
            float wt = (1.0f - fabsf((float)c[0] - g[0])) * (1.0f - fabsf((float)c[1] - g[1])) * (1.0f - fabsf((float)c[2] - g[2]));  // :303
            if (!D->in_bounds(c[0], c[1], c[2])) continue;
            float w = sb->W[o] * wt;  // :307
            int oo; Block* db = D->touch(0, c[0], c[1], c[2], &oo);
            float w_new = w + db->W[oo];                                    // :274
            db->T[oo] = (db->W[oo] * db->T[oo] + w * sb->T[o]) / w_new;    // :275
            if (D->tex_enabled)                                             // :276-277
              for (int ch = 0; ch < 3; ch++) db->col[oo][ch] = (db->W[oo] * db->col[oo][ch] + w * sb->col[o][ch]) / w_new;
            db->W[oo] = w_new;                                              // :278 (no Wmax clamp)
            db->obs[oo] = 1;                                                // :279
            db->occ[oo] = db->occ[oo] + sb->occ[o];                         // :280
          }
    }
  }
}

// cvt_TSDF_surface_to_voxels_kernel  (dense_tsdf.py:339-365).  Output rows in
// lexicographic voxel order; returns the true demand (may exceed cap).
int64_t orc_tsdf_surface(void* h, int submap, int64_t cap, float* xyz, float* rgb) {
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  const float thres = (float)(m->c.voxel_scale * 1.8);  // :39
  const float fl = (float)m->c.disp_floor, ce = (float)m->c.disp_ceiling;
  std::vector<Key> keys;
  for (auto& kv : m->blocks) if (kv.first.s == submap) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  Pose P = pose_or_zero(m->submap_pose, submap);
  int64_t n = 0;
  for (auto& key : keys) {
    Block* b = m->blocks[key];
    for (int o = 0; o < OB3; o++) {
      if (b->obs[o] != 1) continue;                 // :349
      if (!(fabsf(b->T[o]) < thres)) continue;      // :350
      int i = key.x * OB + o / (OB * OB), j = key.y * OB + (o / OB) % OB, k = key.z * OB + o % OB;
      float l[3] = {(float)i * vs, (float)j * vs, (float)k * vs};
      float p[3];
      if (m->c.is_global_map) { p[0] = l[0]; p[1] = l[1]; p[2] = l[2]; }  // :352-353
      else { rot(P.R, l, p); p[0] += P.T[0]; p[1] += P.T[1]; p[2] += P.T[2]; }  // :355
      if (p[2] > ce || p[2] < fl) continue;         // :356
      if (n < cap) {
        xyz[3 * n] = p[0]; xyz[3 * n + 1] = p[1]; xyz[3 * n + 2] = p[2];
        int ci = (int)fmaxf(fminf(((p[2] - fl) / (ce - fl)) * 1023.0f, 1023.0f), 0.0f);  // mapping_common.py:216-219
        if (m->tex_enabled) { rgb[3 * n] = b->col[o][0]; rgb[3 * n + 1] = b->col[o][1]; rgb[3 * n + 2] = b->col[o][2]; }  // :360-362
        else { rgb[3 * n] = m->colormap[ci][0]; rgb[3 * n + 1] = m->colormap[ci][1]; rgb[3 * n + 2] = m->colormap[ci][2]; }
      }
      n++;
    }
  }
  return n;
}

// cvt_TSDF_to_voxels_slice_kernel  (dense_tsdf.py:367-385)
int64_t orc_tsdf_slice(void* h, int submap, float z, float dz, int64_t cap, float* xyz, float* val) {
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  // slice_z is an f16 field (dense_tsdf.py:72); _index = int(z/voxel_scale)
  int index = (int)(h16(z) / vs);
  std::vector<Key> keys;
  for (auto& kv : m->blocks) if (kv.first.s == submap) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  Pose P = pose_or_zero(m->submap_pose, submap);
  int64_t n = 0;
  for (auto& key : keys) {
    Block* b = m->blocks[key];
    for (int o = 0; o < OB3; o++) {
      if (!(b->obs[o] > 0)) continue;
      int i = key.x * OB + o / (OB * OB), j = key.y * OB + (o / OB) % OB, k = key.z * OB + o % OB;
      if (!((float)index - dz < (float)k && (float)k < (float)index + dz)) continue;  // :377
      float l[3] = {(float)i * vs, (float)j * vs, (float)k * vs};
      float p[3];
      if (m->c.is_global_map) { p[0] = l[0]; p[1] = l[1]; p[2] = l[2]; }
      else { rot(P.R, l, p); p[0] += P.T[0]; p[1] += P.T[1]; p[2] += P.T[2]; }
      if (n < cap) { xyz[3 * n] = p[0]; xyz[3 * n + 1] = p[1]; xyz[3 * n + 2] = p[2]; val[n] = b->T[o]; }
      n++;
    }
  }
  return n;
}

// generate_mesh_kernel / marching_on_a_cube / add_triangle / generate_normal
// (marching_cube_mesher.py:84-187).  Triangles emitted in lexicographic
// (block, cell, t) order; returns the true triangle demand.
int64_t orc_mc2(void* h, int step, float thres, int64_t cap_tri, float* verts, float* normals, float* colors);
int64_t orc_mc(void* h, int step, float thres, int64_t cap_tri, float* verts, float* normals) {
  return orc_mc2(h, step, thres, cap_tri, verts, normals, nullptr);
}
// colors (optional, textured maps): vertexInterp_color + add_triangle_color (marching_cube_mesher.py:62-82, :104-108)
int64_t orc_mc2(void* h, int step, float thres, int64_t cap_tri, float* verts, float* normals, float* colors) {
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  const float EPS = 1e-6f;  // :6
  std::vector<Key> keys;
  for (auto& kv : m->blocks) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  int64_t ntri = 0;
  for (auto& key : keys) {
    Block* b = m->blocks[key];
    const int s = key.s;
    for (int o = 0; o < OB3; o++) {
      if (!(b->obs[o] > 0 && b->T[o] < thres)) continue;  // :184
      int i = key.x * OB + o / (OB * OB), j = key.y * OB + (o / OB) % OB, k = key.z * OB + o % OB;
      float val[8];
      bool end = false;
      for (int c = 0; c < 8; c++) {  // :133-138
        int ci = i + GRID[c][0] * step, cj = j + GRID[c][1] * step, ck = k + GRID[c][2] * step;
        val[c] = m->readT(s, ci, cj, ck);
        if (m->readObs(s, ci, cj, ck) == 0) end = true;
      }
      if (end) continue;  // :140
      int cube = 0;
      for (int c = 0; c < 8; c++) if (val[c] < 0.0f) cube |= 1 << c;  // :141-144
      int mask = mc_edge_mask(cube);  // :146
      if (mask == 0) continue;
      float vl[12][3];
      float vc[12][3];
      for (int e = 0; e < 12; e++) {  // :151-172
        if (!(mask & (1 << e))) continue;
        int a = EDGE[e][0], bb = EDGE[e][1];
        float p1[3] = {(float)(i + GRID[a][0] * step), (float)(j + GRID[a][1] * step), (float)(k + GRID[a][2] * step)};
        float p2[3] = {(float)(i + GRID[bb][0] * step), (float)(j + GRID[bb][1] * step), (float)(k + GRID[bb][2] * step)};
        float v1 = val[a], v2 = val[bb];
        // vertexInterp :44-60, isolevel = 0
        if (fabsf(0.0f - v1) < EPS) { vl[e][0] = p1[0]; vl[e][1] = p1[1]; vl[e][2] = p1[2]; }
        else if (fabsf(0.0f - v2) < EPS) { vl[e][0] = p2[0]; vl[e][1] = p2[1]; vl[e][2] = p2[2]; }
        else {
          float mu = (0.0f - v1) / (v2 - v1);
          vl[e][0] = p1[0] + mu * (p2[0] - p1[0]);
          vl[e][1] = p1[1] + mu * (p2[1] - p1[1]);
          vl[e][2] = p1[2] + mu * (p2[2] - p1[2]);
        }
        if (colors) {  // vertexInterp_color :62-82 (mu stays 0 in the two snap branches; the "is zero" tests only look at channel 0)
          float mu = 0.0f;
          if (!(fabsf(0.0f - v1) < EPS) && !(fabsf(0.0f - v2) < EPS)) mu = (0.0f - v1) / (v2 - v1);
          float c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
          int oo; Block* cb;
          if ((cb = m->find(s, i + GRID[a][0] * step, j + GRID[a][1] * step, k + GRID[a][2] * step, &oo))) for (int ch = 0; ch < 3; ch++) c1[ch] = cb->col[oo][ch];
          if ((cb = m->find(s, i + GRID[bb][0] * step, j + GRID[bb][1] * step, k + GRID[bb][2] * step, &oo))) for (int ch = 0; ch < 3; ch++) c2[ch] = cb->col[oo][ch];
          for (int ch = 0; ch < 3; ch++) {
            float pc = c1[ch];
            if (c1[0] == 0.0f) pc = c2[ch];
            else if (!(c2[0] == 0.0f)) pc = c1[ch] + mu * (c2[ch] - c1[ch]);
            vc[e][ch] = pc;
          }
        }
      }
      for (int t = 0; t < 5; t++) {  // :173-174 / :110-125
        int e0 = mc_tri(cube, 3 * t);
        if (e0 == -1) continue;
        int es[3] = {e0, mc_tri(cube, 3 * t + 1), mc_tri(cube, 3 * t + 2)};
        if (ntri < cap_tri) {
          for (int q = 0; q < 3; q++) {
            const float* p = vl[es[q]];
            float* vo = verts + (ntri * 3 + q) * 3;
            vo[0] = p[0] * vs; vo[1] = p[1] * vs; vo[2] = p[2] * vs;  // ijk_to_xyz :40-42
            if (colors) { float* co = colors + (ntri * 3 + q) * 3; co[0] = vc[es[q]][0]; co[1] = vc[es[q]][1]; co[2] = vc[es[q]][2]; }
            float* no = normals + (ntri * 3 + q) * 3;
            if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) {
              // NaN TSDF corner (the shipped fixtures contain some) -> NaN vertex; round(NaN) is undefined in the
              // reference (:86).  Canonical: the normal of a non-finite vertex is NaN.
              no[0] = no[1] = no[2] = NAN;
              continue;
            }
            int pi = iround(p[0]), pj = iround(p[1]), pk = iround(p[2]);  // generate_normal :84-93
            float nx = m->readT(s, pi + 1, pj, pk) - m->readT(s, pi - 1, pj, pk);
            float ny = m->readT(s, pi, pj + 1, pk) - m->readT(s, pi, pj - 1, pk);
            float nz = m->readT(s, pi, pj, pk + 1) - m->readT(s, pi, pj, pk - 1);
            float nn = sqrtf((nx * nx + ny * ny) + nz * nz);
            no[0] = nx / nn; no[1] = ny / nn; no[2] = nz / nn;  // normalized(): NaN when the gradient is 0
          }
        }
        ntri++;
      }
    }
  }
  return ntri;
}

// ---------------------------------------------------------------------------
// ESDF.  The reference's ESDF (dense_esdf.py:228-333) is dead code at HEAD and its
// lower-queue never re-inserts improved voxels (:292,:298 commented out), so
// there is no runnable behaviour to pin: PARITY UNPINNED.  Canonical definition
// (SURVEY Appendix A.6), computed here by multi-source Dijkstra:
//   * voxel set = TSDF-observed voxels of `submap`;
//   * fixed band |TSDF| < gamma = voxel_scale  (is_fixed :228-230) : ESDF = TSDF (:315-320);
//   * other voxels: sign = sign(TSDF) (:324,:328); positive side
//       ESDF(v) = min(max_ray, min over 26-neighbours h in fixed U positive of ESDF(h)+|dir|*vs)
//     negative side mirrored with max / -max_ray  (process_lower_queue :275-299,
//     iterated to convergence as the commented-out re-insertion intended).
// ---------------------------------------------------------------------------
int64_t orc_esdf_update(void* h, int submap) {
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  const float gamma = (float)m->c.voxel_scale;
  const float far = (float)m->c.max_ray_length;
  const float dis[4] = {0.0f, vs, sqrtf(2.0f) * vs, sqrtf(3.0f) * vs};
  struct Node { float d; int i, j, k; };
  auto cmp = [](const Node& a, const Node& b) { return a.d > b.d; };
  int64_t nobs = 0;
  for (int pass = 0; pass < 2; pass++) {  // pass 0: positive side, pass 1: negative side (on |.|)
    std::priority_queue<Node, std::vector<Node>, decltype(cmp)> pq(cmp);
    for (auto& kv : m->blocks) {
      if (kv.first.s != submap) continue;
      Block* b = kv.second;
      for (int o = 0; o < OB3; o++) {
        if (!b->obs[o]) continue;
        if (pass == 0) nobs++;
        float t = b->T[o];
        int i = kv.first.x * OB + o / (OB * OB), j = kv.first.y * OB + (o / OB) % OB, k = kv.first.z * OB + o % OB;
        if (fabsf(t) < gamma) {
          if (pass == 0) { b->esdf[o] = t; b->eobs[o] = 1; }
          pq.push({pass == 0 ? t : -t, i, j, k});
        } else if (pass == 0) {
          b->esdf[o] = (float)sgn(t) * far; b->eobs[o] = 1;
        }
      }
    }
    while (!pq.empty()) {
      Node nd = pq.top(); pq.pop();
      int o = 0; Block* b = m->find(submap, nd.i, nd.j, nd.k, &o);
      float cur = pass == 0 ? b->esdf[o] : -b->esdf[o];
      if (nd.d > cur) continue;  // stale
      for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) for (int dk = -1; dk <= 1; dk++) {
        int nz = (di != 0) + (dj != 0) + (dk != 0);
        if (nz == 0) continue;
        int oo; Block* nb = m->find(submap, nd.i + di, nd.j + dj, nd.k + dk, &oo);
        if (!nb || !nb->obs[oo]) continue;
        float t = nb->T[oo];
        if (fabsf(t) < gamma) continue;             // fixed voxels are never relaxed
        if (pass == 0 ? !(t > 0.0f) : !(t < 0.0f)) continue;
        float cand = nd.d + dis[nz];
        float ncur = pass == 0 ? nb->esdf[oo] : -nb->esdf[oo];
        if (cand < ncur) {
          nb->esdf[oo] = pass == 0 ? cand : -cand;
          pq.push({cand, nd.i + di, nd.j + dj, nd.k + dk});
        }
      }
    }
  }
  return nobs;
}

int64_t orc_esdf_gather(void* h, int submap, int64_t cap, int32_t* idx, float* esdf) {
  Tsdf* m = (Tsdf*)h;
  struct Row { int i, j, k; float e; };
  std::vector<Row> rows;
  for (auto& kv : m->blocks) {
    if (kv.first.s != submap) continue;
    Block* b = kv.second;
    for (int o = 0; o < OB3; o++)
      if (b->eobs[o]) rows.push_back({kv.first.x * OB + o / (OB * OB), kv.first.y * OB + (o / OB) % OB, kv.first.z * OB + o % OB, b->esdf[o]});
  }
  std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
    if (a.i != b.i) return a.i < b.i;
    if (a.j != b.j) return a.j < b.j;
    return a.k < b.k;
  });
  int64_t n = 0;
  for (auto& r : rows) {
    if (n < cap) { idx[3 * n] = r.i; idx[3 * n + 1] = r.j; idx[3 * n + 2] = r.k; esdf[n] = r.e; }
    n++;
  }
  return n;
}

// ============================================================================
// Octomap
// ============================================================================
struct OctoCfg {
  double voxel_scale;
  int N, Nz, K;
  double max_ray_length, min_ray_length;
  int recast_step;
  double fx, fy, cx, cy;
  int min_occupy_thres;
};

void* orc_octo_create(const OctoCfg* c) {
  Octo* m = new Octo();
  m->voxel_scale = c->voxel_scale; m->vs = (float)c->voxel_scale;
  m->N = c->N; m->Nz = c->Nz; m->K = c->K;
  m->max_ray = c->max_ray_length; m->min_ray = c->min_ray_length; m->step = c->recast_step;
  m->fx = c->fx; m->fy = c->fy; m->cx = c->cx; m->cy = c->cy;
  m->min_occupy_thres = c->min_occupy_thres;
  return m;
}
void orc_octo_destroy(void* h) { delete (Octo*)h; }
void orc_octo_reset(void* h) { ((Octo*)h)->cnt.clear(); ((Octo*)h)->cw.clear(); }  // root.deactivate_all() taichi_octomap.py:210-211
void orc_octo_set_color(void* h, int enabled, int same_proj, double fx, double fy, double cx, double cy) {
  Octo* m = (Octo*)h;
  m->tex_enabled = enabled != 0; m->color_same_proj = same_proj != 0;
  m->fxc = (float)fx; m->fyc = (float)fy; m->cxc = (float)cx; m->cyc = (float)cy;
}
static inline uint64_t octo_word(uint32_t seq, const uint8_t* p) {  // "Stupid OpenCV is BGR" (:121-124)
  return ((uint64_t)seq << 24) | ((uint64_t)p[2] << 16) | ((uint64_t)p[1] << 8) | (uint64_t)p[0];
}
static inline void octo_rgb(uint64_t w, float* out) {
  out[0] = (float)((w >> 16) & 255) / 255.0f; out[1] = (float)((w >> 8) & 255) / 255.0f; out[2] = (float)(w & 255) / 255.0f;
}
void orc_octo_set_submap_pose(void* h, int s, const float* R9, const float* T3) {
  Octo* m = (Octo*)h;
  Pose p; memcpy(p.R, R9, sizeof(p.R)); memcpy(p.T, T3, sizeof(p.T));
  m->submap_pose[s] = p;
}

// recast_pcl_to_map_kernel + process_point  (taichi_octomap.py:134-145, :116-119)
void orc_octo_integrate_points_rgb(void* h, const float* xyz, const uint8_t* rgb, int n, const float* R9, const float* T3, int submap);
void orc_octo_integrate_points(void* h, const float* xyz, int n, const float* R9, const float* T3, int submap) {
  orc_octo_integrate_points_rgb(h, xyz, nullptr, n, R9, T3, submap);
}
void orc_octo_integrate_points_rgb(void* h, const float* xyz, const uint8_t* rgb, int n, const float* R9, const float* T3, int submap) {
  Octo* m = (Octo*)h;
  const float vs = m->vs;
  if (m->frame_seq < (1u << 22) - 1) m->frame_seq++;
  for (int idx = 0; idx < n; idx++) {
    float pt[3] = {xyz[3 * idx], xyz[3 * idx + 1], xyz[3 * idx + 2]};
    float p[3];
    rot(R9, pt, p);
    p[0] += T3[0]; p[1] += T3[1]; p[2] += T3[2];                         // :141
    int i = iround(p[0] / vs), j = iround(p[1] / vs), k = iround(p[2] / vs);  // xyz_to_sijk mapping_common.py:251-255
    if (!m->in_bounds(i, j, k)) continue;
    m->hit(Key{submap, i, j, k}, 1, rgb ? octo_word(m->frame_seq, rgb + 3 * (size_t)idx) : 0);  // :119-124
  }
}

// recast_depth_to_map_kernel  (taichi_octomap.py:147-169)
void orc_octo_integrate_depth_tex(void* h, const uint16_t* depth, const uint8_t* tex, int TH, int TW, int H, int Wd, const float* R9,
                                  const float* T3, int submap);
void orc_octo_integrate_depth(void* h, const uint16_t* depth, int H, int Wd, const float* R9, const float* T3, int submap) {
  orc_octo_integrate_depth_tex(h, depth, nullptr, 0, 0, H, Wd, R9, T3, submap);
}
void orc_octo_integrate_depth_tex(void* h, const uint16_t* depth, const uint8_t* tex, int TH, int TW, int H, int Wd, const float* R9,
                                  const float* T3, int submap) {
  Octo* m = (Octo*)h;
  const float vs = m->vs;
  if (m->frame_seq < (1u << 22) - 1) m->frame_seq++;
  const int step = m->step;
  const int hh = (int)((double)H / step), ww = (int)((double)Wd / step);
  const float fx = (float)m->fx, fy = (float)m->fy, cx = (float)m->cx, cy = (float)m->cy;
  const float dmax = (float)(m->max_ray * 1000.0), dmin = (float)(m->min_ray * 1000.0);
  for (int jj = 0; jj < hh; jj++) {
    int j = jj * step;
    for (int ii = 0; ii < ww; ii++) {
      int i = ii * step;
      uint16_t d = depth[(size_t)j * Wd + i];
      if (d == 0 || (float)d > dmax || (float)d < dmin) continue;   // :155
      float dep = (float)d / 1000.0f;                                // :157
      float pt[3] = {((float)i - cx) * dep / fx, ((float)j - cy) * dep / fy, dep};
      float p[3];
      rot(R9, pt, p);
      p[0] += T3[0]; p[1] += T3[1]; p[2] += T3[2];                   // :159
      int vi = iround(p[0] / vs), vj = iround(p[1] / vs), vk = iround(p[2] / vs);
      if (!m->in_bounds(vi, vj, vk)) continue;
      uint64_t word = 0;
      if (m->tex_enabled && tex) {  // :160-167
        int tj = j, ti = i;
        if (!m->color_same_proj) {  // color_ind_from_depth_pt (mapping_common.py:43-58), see orc_tsdf_integrate_depth_tex
          ti = (int)((((float)i - cx) / fx) * m->fxc + m->cxc);
          tj = (int)((((float)j - cy) / fy) * m->fyc + m->cyc);
          if (ti < 0 || ti >= TH || tj < 0 || tj >= TW || tj >= TH || ti >= TW) { ti = 0; tj = 0; }
        }
        const uint8_t zero3[3] = {0, 0, 0};
        word = octo_word(m->frame_seq, (tj < TH && ti < TW) ? tex + ((size_t)tj * TW + ti) * 3 : zero3);
      }
      m->hit(Key{submap, vi, vj, vk}, 1, word);
    }
  }
}

// every (i,j,k,count) of one submap, sorted
int64_t orc_octo_gather(void* h, int submap, int64_t cap, int32_t* idx, uint32_t* count) {
  Octo* m = (Octo*)h;
  int64_t n = 0;
  for (auto& kv : m->cnt) {
    if (kv.first.s != submap) continue;
    if (n < cap) { idx[3 * n] = kv.first.x; idx[3 * n + 1] = kv.first.y; idx[3 * n + 2] = kv.first.z; count[n] = kv.second; }
    n++;
  }
  return n;
}

// colours of the rows of orc_octo_gather, same order
void orc_octo_gather_color(void* h, int submap, int64_t cap, float* col) {
  Octo* m = (Octo*)h;
  int64_t n = 0;
  for (auto& kv : m->cnt) {
    if (kv.first.s != submap) continue;
    if (n < cap) {
      auto it = m->cw.find(kv.first);
      octo_rgb(it == m->cw.end() ? 0 : it->second, col + 3 * n);
    }
    n++;
  }
}

// cvt_occupy_to_voxels(level)  (taichi_octomap.py:90-102).  occupy.parent(level)
// iterates the ACTIVE cells of the ancestor `level` levels above the leaf field:
// level 1 = individual voxels, level L = K^(L-1)-aligned groups, reported at the
// group's base coordinate; is_occupy is evaluated AT that base coordinate
// (:97, :86-88: occupy > min_occupy_thres), so a group whose corner voxel was
// never hit is not exported.  xyz = sijk_to_xyz (mapping_common.py:234-238).
int64_t orc_octo_export2(void* h, int submap, int level, int64_t cap, float* xyz, float* rgb);
int64_t orc_octo_export(void* h, int submap, int level, int64_t cap, float* xyz) { return orc_octo_export2(h, submap, level, cap, xyz, nullptr); }
int64_t orc_octo_export2(void* h, int submap, int level, int64_t cap, float* xyz, float* rgb) {
  Octo* m = (Octo*)h;
  const float vs = m->vs;
  int g = 1;
  for (int l = 1; l < level; l++) g *= m->K;
  // offset of the tree: -N/2 (taichi_octomap.py:72): groups are aligned in offset coordinates
  const int h2 = m->N / 2, hz = m->Nz / 2;
  std::map<Key, int> cells;
  for (auto& kv : m->cnt) {
    if (kv.first.s != submap) continue;
    Key c{submap, fdiv(kv.first.x + h2, g) * g - h2, fdiv(kv.first.y + h2, g) * g - h2, fdiv(kv.first.z + hz, g) * g - hz};
    cells[c] = 1;
  }
  Pose P = pose_or_zero(m->submap_pose, submap);
  int64_t n = 0;
  for (auto& kv : cells) {
    auto it = m->cnt.find(kv.first);
    float occ = it == m->cnt.end() ? 0.0f : (float)it->second;
    if (!(occ > (float)m->min_occupy_thres)) continue;
    if (n < cap) {
      float l[3] = {(float)kv.first.x * vs, (float)kv.first.y * vs, (float)kv.first.z * vs};
      float p[3];
      rot(P.R, l, p);
      xyz[3 * n] = p[0] + P.T[0]; xyz[3 * n + 1] = p[1] + P.T[1]; xyz[3 * n + 2] = p[2] + P.T[2];
      if (rgb && m->tex_enabled) {  // export_color[index] = color[sijk] (:101-102)
        auto ic = m->cw.find(kv.first);
        octo_rgb(ic == m->cw.end() ? 0 : ic->second, rgb + 3 * n);
      }
    }
    n++;
  }
  return n;
}

// fuse_submaps_kernel  (taichi_octomap.py:171-189), after reset() (:195-196)
void orc_octo_fuse(void* hdst, void* hsrc) {
  Octo* D = (Octo*)hdst;
  Octo* S = (Octo*)hsrc;
  D->cnt.clear();
  D->cw.clear();
  const float vs = D->vs;
  for (auto& kv : S->cnt) {
    float occ = (float)kv.second;
    if (!(occ > (float)D->min_occupy_thres)) continue;  // :181
    Pose P = pose_or_zero(D->submap_pose, kv.first.s);
    float l[3] = {(float)kv.first.x * vs, (float)kv.first.y * vs, (float)kv.first.z * vs};
    float r[3];
    rot(P.R, l, r);
    int i = iround((r[0] + P.T[0]) / vs), j = iround((r[1] + P.T[1]) / vs), k = iround((r[2] + P.T[2]) / vs);  // :182-183
    if (!D->in_bounds(i, j, k)) continue;
    uint64_t word = 0;  // color[ijk_] = submap_color[s,i,j,k] (:189), racing: the most recently integrated source colour wins
    if (D->tex_enabled && S->tex_enabled) { auto ic = S->cw.find(kv.first); if (ic != S->cw.end()) word = ic->second; }
    D->hit(Key{0, i, j, k}, kv.second, word);  // :186
  }
}

// ---------------------------------------------------------------------------
// Map queries (BaseMap @ti.func helpers, mapping_common.py:165-204; DenseTSDF.is_occupy / is_unobserved,
// dense_tsdf.py:148-155: TSDF < 1.8*vs with NO observed test - inactive cells read 0, i.e. "occupied").
// ---------------------------------------------------------------------------
static void tsdf_probe(Tsdf* m, int s, float x, float y, float z, bool& occ, bool& unobs) {
  const float vs = m->vs;
  const int i = iround(x / vs), j = iround(y / vs), k = iround(z / vs);
  float t = 0.0f;
  int o = 0;
  if (m->in_bounds(i, j, k)) { t = m->readT(s, i, j, k); o = m->readObs(s, i, j, k); }
  occ = t < (float)(m->c.voxel_scale * 1.8);
  unobs = o == 0;
}
void orc_tsdf_query_points(void* h, int submap, int64_t n, const float* xyz, uint8_t* flags) {
  Tsdf* m = (Tsdf*)h;
  for (int64_t q = 0; q < n; q++) {
    bool occ, un;
    tsdf_probe(m, submap, xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2], occ, un);
    flags[q] = (uint8_t)((occ ? 1 : 0) | (un ? 2 : 0));
  }
}
void orc_tsdf_query_near(void* h, int submap, int64_t n, const float* xyz, int voxel, uint8_t* out) {  // mapping_common.py:193-204
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  for (int64_t q = 0; q < n; q++) {
    const int i0 = iround(xyz[3 * q] / vs), j0 = iround(xyz[3 * q + 1] / vs), k0 = iround(xyz[3 * q + 2] / vs);
    bool any = false;
    for (int i = -voxel; i < voxel; i++)
      for (int j = -voxel; j < voxel; j++)
        for (int k = -voxel; k < voxel; k++) {
          float t = 0.0f;
          if (m->in_bounds(i0 + i, j0 + j, k0 + k)) t = m->readT(submap, i0 + i, j0 + j, k0 + k);
          if (t < (float)(m->c.voxel_scale * 1.8)) any = true;
        }
    out[q] = any ? 1 : 0;
  }
}
void orc_tsdf_raycast(void* h, int submap, int64_t n, const float* pos, const float* dir, float max_dist, uint8_t* hit, float* xyz_out,
                      float* len_out) {  // mapping_common.py:165-178
  Tsdf* m = (Tsdf*)h;
  const float vs = m->vs;
  for (int64_t q = 0; q < n; q++) {
    const int steps = (int)(max_dist / vs);
    float x = 0.f, y = 0.f, z = 0.f, len = 0.f;
    bool succ = false;
    for (int j = 0; j < steps; j++) {
      len = (float)j * vs;
      x = dir[3 * q] * len + pos[3 * q]; y = dir[3 * q + 1] * len + pos[3 * q + 1]; z = dir[3 * q + 2] * len + pos[3 * q + 2];
      bool occ, un;
      tsdf_probe(m, submap, x, y, z, occ, un);
      if (occ) { succ = true; break; }
    }
    hit[q] = succ ? 1 : 0;
    xyz_out[3 * q] = x; xyz_out[3 * q + 1] = y; xyz_out[3 * q + 2] = z;
    len_out[q] = len;
  }
}
static bool octo_probe(Octo* m, int s, float x, float y, float z) {  // taichi_octomap.py:86-88
  const float vs = m->vs;
  const int i = iround(x / vs), j = iround(y / vs), k = iround(z / vs);
  uint32_t c = 0;
  if (m->in_bounds(i, j, k)) { auto it = m->cnt.find(Key{s, i, j, k}); if (it != m->cnt.end()) c = it->second; }
  return (float)c > (float)m->min_occupy_thres;
}
void orc_octo_query_points(void* h, int submap, int64_t n, const float* xyz, uint8_t* flags) {
  Octo* m = (Octo*)h;
  for (int64_t q = 0; q < n; q++) flags[q] = octo_probe(m, submap, xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2]) ? 1 : 0;
}
void orc_octo_raycast(void* h, int submap, int64_t n, const float* pos, const float* dir, float max_dist, uint8_t* hit, float* xyz_out,
                      float* len_out) {
  Octo* m = (Octo*)h;
  const float vs = m->vs;
  for (int64_t q = 0; q < n; q++) {
    const int steps = (int)(max_dist / vs);
    float x = 0.f, y = 0.f, z = 0.f, len = 0.f;
    bool succ = false;
    for (int j = 0; j < steps; j++) {
      len = (float)j * vs;
      x = dir[3 * q] * len + pos[3 * q]; y = dir[3 * q + 1] * len + pos[3 * q + 1]; z = dir[3 * q + 2] * len + pos[3 * q + 2];
      if (octo_probe(m, submap, x, y, z)) { succ = true; break; }
    }
    hit[q] = succ ? 1 : 0;
    xyz_out[3 * q] = x; xyz_out[3 * q + 1] = y; xyz_out[3 * q + 2] = z;
    len_out[q] = len;
  }
}

// ----------------------------------------------------------------------------
// CPU-baseline helper (bench.py `cpu_baseline` / `--impl reference`): integrate a
// stream of frames with `nthreads` host threads, thread t owning map handles[t]
// (independent submaps - the same submap-sharded decomposition the GPU bench
// uses across ranks; frames are handed out dynamically).  Returns frames done.
// ----------------------------------------------------------------------------
int orc_tsdf_integrate_stream_mt(void** handles, int nthreads, const uint16_t* depth, int nframes, int H, int Wd,
                                 const float* R9s, const float* T3s) {
  std::atomic<int> next(0);
  auto work = [&](int t) {
    for (;;) {
      int f = next.fetch_add(1);
      if (f >= nframes) break;
      orc_tsdf_integrate_depth(handles[t], depth + (size_t)f * H * Wd, H, Wd, R9s + 9 * f, T3s + 3 * f, 0, 1);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  return nframes;
}

}  // extern "C"
