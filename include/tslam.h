/* ============================================================================
 * tslam.h - C ABI of libtslam.so, the B200-native dense-mapping backend.
 *
 * The reference (xuhao1/TaichiSLAM) has no FFI layer: its callers use the Python
 * classes DenseTSDF / Octomap / MarchingCubeMesher directly, and those run Taichi
 * kernels.  This ABI is what a maintainer binds (ctypes) to
 * replace each Taichi kernel launch; every entry point names the reference
 * method/kernel it stands in for (paths relative to taichi_slam/mapping/).
 *
 * Conventions
 *  - plain C types only; opaque handles; no exceptions cross the boundary.
 *  - every call returns TSLAM_OK (0) or a negative TSLAM_E_* code;
 *    tslam_last_error() gives a thread-local message for the last failure.
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are
 *    ASYNCHRONOUS on that stream unless documented "synchronises".
 *  - `mem` arguments: TSLAM_MEM_DEVICE = the pointer is device memory on the
 *    handle's GPU (caller-owned, must stay alive until the stream reaches the
 *    call); TSLAM_MEM_HOST = host memory, the library copies it to the device
 *    itself (truly asynchronous only when the host memory is pinned).
 *  - handles are not thread-safe; distinct handles are independent.
 *  - voxel indices are signed and centred: valid i in [-N/2, N/2)
 *    (dense_tsdf.py:90).  Out-of-volume samples are skipped (the reference
 *    performs an unchecked access there, mapping_common.py:263-266).
 * ==========================================================================*/
#ifndef TSLAM_H
#define TSLAM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TSLAM_OK 0
#define TSLAM_E_INVALID (-1)   /* bad argument                                  */
#define TSLAM_E_CUDA (-2)      /* CUDA runtime error (see tslam_last_error)     */
#define TSLAM_E_POOL_FULL (-3) /* voxel-block pool exhausted; samples dropped   */
#define TSLAM_E_CAPACITY (-4)  /* caller buffer too small; output saturated     */
#define TSLAM_E_NOGPU (-5)     /* no CUDA device: there is NO CPU fallback      */

#define TSLAM_MEM_DEVICE 0
#define TSLAM_MEM_HOST 1

#define TSLAM_MAX_BATCH 64 /* frames per integrate launch (larger calls are split) */

const char* tslam_last_error(void);
int tslam_device_count(void);
/* ABI version of this header (tests check the library agrees). */
int tslam_abi_version(void);
#define TSLAM_ABI_VERSION 3

/* ----------------------------------------------------------------------------
 * TSDF map  (DenseTSDF, dense_tsdf.py)
 * --------------------------------------------------------------------------*/
typedef struct tslam_tsdf tslam_tsdf_t;

typedef struct tslam_tsdf_config {
  double voxel_scale;        /* dense_tsdf.py:13 voxel_scale                         */
  int32_t N, Nz;             /* grid edge in voxels, dense_tsdf.py:24-25             */
  double max_ray_length;     /* dense_tsdf.py:14                                     */
  double min_ray_length;     /* dense_tsdf.py:14                                     */
  int32_t internal_voxels;   /* dense_tsdf.py:15                                     */
  int32_t recast_step;       /* dense_tsdf.py:16                                     */
  double fx, fy, cx, cy;     /* K_cam_dep[0],[4],[2],[5]  mapping_common.py:33-36    */
  int32_t is_global_map;     /* dense_tsdf.py:15                                     */
  double disp_floor, disp_ceiling; /* dense_tsdf.py:16                               */
  int32_t max_submaps;       /* dense_tsdf.py:15 max_submap_num                      */
  int32_t max_blocks;        /* capacity of the 16^3 voxel-block pool (0 = derive)   */
  int32_t max_image_pixels;  /* largest h*w a depth frame may have (0 = 640*480)     */
  int32_t max_points;        /* largest point cloud for integrate_points (0 = 1<<20) */
  int32_t texture_enabled;   /* dense_tsdf.py:13 texture_enabled: colour planes + colour fusion        */
} tslam_tsdf_config_t;

/* DenseTSDF.__init__ / initialize_fields (dense_tsdf.py:13-118). */
int tslam_tsdf_create(const tslam_tsdf_config_t* cfg, tslam_tsdf_t** out);
int tslam_tsdf_destroy(tslam_tsdf_t* m);
/* DenseTSDF.reset(): B.parent().deactivate_all() (dense_tsdf.py:309-310). */
int tslam_tsdf_reset(tslam_tsdf_t* m, void* stream);
/* BaseMap.set_dep_camera_intrinsic (mapping_common.py:25-26). */
int tslam_tsdf_set_intrinsics(tslam_tsdf_t* m, double fx, double fy, double cx, double cy);
/* BaseMap.set_base_pose_submap_kernel (mapping_common.py:126-131): pose table row s. */
int tslam_tsdf_set_submap_pose(tslam_tsdf_t* m, int32_t s, const float* R9, const float* T3);

/* DenseTSDF.recast_depth_to_map -> recast_depth_to_map_kernel (dense_tsdf.py:162-165,
 * :188-214) for a batch of n_frames frames.  depth: uint16 [n_frames,h,w] millimetres;
 * R9s/T3s: HOST arrays [n_frames,9]/[n_frames,3] = input_R/input_T after set_pose's
 * convert_by_base + f32 cast (mapping_common.py:149-156); submap_ids: HOST int32
 * [n_frames] (active_submap_id per frame) or NULL for all-zero.
 * flags: TSLAM_F_COMMIT applies the pending weighted averages before returning
 * control to the stream (otherwise they stay pending until tslam_tsdf_commit or any
 * reader).  The per-frame bucket grid (PCLroot.deactivate_all, :163) is implicit. */
#define TSLAM_F_COMMIT 1
int tslam_tsdf_integrate_depth(tslam_tsdf_t* m, const uint16_t* depth, int mem, int32_t n_frames, int32_t h, int32_t w,
                               const float* R9s, const float* T3s, const int32_t* submap_ids, int flags, void* stream);
/* Per-frame form of the same call, for callers that hand over one frame at a time (the reference API):
 * copies ONE host frame to a device staging slot on an internal copy stream (overlapping the kernels of the
 * previous launch) and records its pose; every TSLAM_MAX_BATCH/2 queued frames (TSLAM_QUEUE_LAUNCH="a,b" in the
 * environment alternates other counts) - or tslam_tsdf_flush, which every reader calls implicitly - are integrated
 * and committed with one launch triple, so the kernels of one half run while the caller hands over the next.
 * The frame has been consumed when the call returns (the reference's semantics: its kernel launch copies the array
 * synchronously): a page-locked source is DMA-copied and the copy awaited; of a pageable source the sampled rows (every
 * recast_step-th row) are copied into a page-locked ring of the library with streaming stores, from where the GPU
 * fetches them.  (Environment, A/B only: TSLAM_FRAME_COPY=dma hands pageable frames to cudaMemcpyAsync;
 * TSLAM_PINNED_COPY=ring|fetch sends page-locked frames through the ring / awaits a row fetch inside the call.)
 * tslam_tsdf_set_frame_mode(m, 1) (or TSLAM_ZERO_COPY=1 in the environment) opts into BORROWING page-locked frames
 * instead: they are not copied at all, the GPU fetches their sampled rows (every recast_step-th row) straight from
 * host memory a few calls later - half the PCIe bytes for recast_step 2, no wait per frame - and they must then stay
 * valid and unchanged until tslam_tsdf_flush has returned and the stream has been synchronised, or until
 * 2*TSLAM_MAX_BATCH further frames have been queued (the call blocks rather than letting the host run further ahead). */
int tslam_tsdf_queue_depth(tslam_tsdf_t* m, const uint16_t* depth_host, int32_t h, int32_t w, const float* R9,
                           const float* T3, int32_t submap, void* stream);
int tslam_tsdf_flush(tslam_tsdf_t* m, void* stream);
int tslam_tsdf_set_frame_mode(tslam_tsdf_t* m, int borrow_pinned);
/* The same hand-over in two halves, so that the caller's own per-frame work (DenseTSDF.recast_depth_to_map: the f64 pose
 * arithmetic of convert_by_base, mapping_common.py:91-100) runs while the frame's DMA copy is in flight: _begin starts
 * the copy, _end awaits it (the frame has been consumed when _end returns), records the pose and launches a full
 * queue.  No other call on the handle between the two. */
int tslam_tsdf_queue_depth_begin(tslam_tsdf_t* m, const uint16_t* depth_host, int32_t h, int32_t w, void* stream);
int tslam_tsdf_queue_depth_end(tslam_tsdf_t* m, const float* R9, const float* T3, int32_t submap, void* stream);
/* Frames per queue launch, alternating a, b, a, ... (default TSLAM_MAX_BATCH/2 each).  Every launch ends with a commit:
 * the granule of the Wmax clamp (dense_tsdf.py:267).  (1, 1) = one commit per frame - what a frame-by-frame caller of
 * the reference sees below Wmax, and the closest a summed update gets to its per-sample clamp at Wmax
 * (tests/test_oracle_vs_reference_exec.py::test_commit_granularity_vs_per_sample_clamp_at_wmax) - at the price of
 * one launch sequence per frame. */
int tslam_tsdf_set_queue_launch(tslam_tsdf_t* m, int32_t frames_a, int32_t frames_b);
/* Textured maps (texture_enabled).  set_color_camera_intrinsic (mapping_common.py:28-29) + color_same_proj
 * (dense_tsdf.py:16).  The *_tex / *_rgb forms take the colour image uint8 [n_frames,th,tw,3] (channel order as
 * given - DenseTSDF does not swap BGR) or per-point colours uint8 [n,3]; tex == NULL integrates geometry only.
 * Colour semantics: the reference overwrites color[xi] at every sample, racing rays, last writer wins
 * (dense_tsdf.py:268-269); here the winner is deterministic: latest frame, then the sample closest to its ray's
 * surface point (DESIGN.md "Texture"). */
int tslam_tsdf_set_color_intrinsics(tslam_tsdf_t* m, double fx, double fy, double cx, double cy, int color_same_proj);
int tslam_tsdf_integrate_depth_tex(tslam_tsdf_t* m, const uint16_t* depth, const uint8_t* tex, int mem, int32_t n_frames, int32_t h,
                                   int32_t w, int32_t th, int32_t tw, const float* R9s, const float* T3s, const int32_t* submap_ids,
                                   int flags, void* stream);
int tslam_tsdf_queue_depth_tex(tslam_tsdf_t* m, const uint16_t* depth_host, const uint8_t* tex_host, int32_t h, int32_t w, int32_t th,
                               int32_t tw, const float* R9, const float* T3, int32_t submap, void* stream);
int tslam_tsdf_integrate_points_rgb(tslam_tsdf_t* m, const float* xyz, const uint8_t* rgb, int mem, int32_t n, const float* R9,
                                    const float* T3, int32_t submap, int flags, void* stream);
/* to_numpy / load_numpy with the colour column (dense_tsdf.py:437-440, :450-453): color f32[cap,3] DEVICE or NULL. */
int tslam_tsdf_gather2(tslam_tsdf_t* m, int32_t submap, int64_t cap, int32_t* idx, float* tsdf, float* w, int8_t* occ, float* color,
                       int64_t* n_out, void* stream);
int tslam_tsdf_scatter2(tslam_tsdf_t* m, int32_t submap, int64_t n, const int32_t* idx, const float* tsdf, const float* w,
                        const int8_t* occ, const float* color, void* stream);
/* DenseTSDF.recast_pcl_to_map -> recast_pcl_to_map_kernel (dense_tsdf.py:157-160, :167-186).
 * xyz: float32 [n,3]. */
int tslam_tsdf_integrate_points(tslam_tsdf_t* m, const float* xyz, int mem, int32_t n, const float* R9, const float* T3,
                                int32_t submap, int flags, void* stream);
/* Apply pending per-voxel weighted averages: T' = (T*W + sum w*d)/(W + sum w),
 * W' = min(W + sum w, Wmax=1000)  (dense_tsdf.py:264-267 applied once per voxel). */
int tslam_tsdf_commit(tslam_tsdf_t* m, void* stream);

/* DenseTSDF.count_active (dense_tsdf.py:412-423).  Synchronises. */
int tslam_tsdf_count_active(tslam_tsdf_t* m, int32_t submap, int64_t* n_out);
/* DenseTSDF.to_numpy (dense_tsdf.py:425-440): observed voxels of `submap` into
 * caller DEVICE arrays idx int32[cap,3], tsdf f32[cap], w f32[cap], occ int8[cap]
 * (row order unspecified, like the reference's atomic counter).  *n_out = demand.
 * Synchronises.  Returns TSLAM_E_CAPACITY (outputs saturated) when demand > cap. */
int tslam_tsdf_gather(tslam_tsdf_t* m, int32_t submap, int64_t cap, int32_t* idx, float* tsdf, float* w, int8_t* occ,
                      int64_t* n_out, void* stream);
/* DenseTSDF.load_numpy (dense_tsdf.py:442-454): DEVICE arrays, n rows. */
int tslam_tsdf_scatter(tslam_tsdf_t* m, int32_t submap, int64_t n, const int32_t* idx, const float* tsdf, const float* w,
                       const int8_t* occ, void* stream);
/* DenseTSDF.fuse_submaps -> reset() + fuse_submaps_kernel (dense_tsdf.py:272-318):
 * trilinear splat (7 corners, :300) of every observed voxel of every submap of `src`
 * into submap 0 of `dst`, using dst's pose table.  No Wmax clamp (:274-278). */
int tslam_tsdf_fuse(tslam_tsdf_t* dst, tslam_tsdf_t* src, void* stream);
/* cvt_TSDF_surface_to_voxels_kernel (dense_tsdf.py:339-365): append |TSDF|<1.8*vs
 * voxels of `submap` to xyz/rgb f32[cap,3] (DEVICE) starting at *count_dev (DEVICE
 * int32 counter, incremented; entries beyond cap are counted but not written). */
int tslam_tsdf_extract_surface(tslam_tsdf_t* m, int32_t submap, int64_t cap, float* xyz, float* rgb, int32_t* count_dev,
                               void* stream);
/* cvt_TSDF_to_voxels_slice_kernel (dense_tsdf.py:367-385). val = TSDF value. */
int tslam_tsdf_extract_slice(tslam_tsdf_t* m, int32_t submap, float z, float dz, int64_t cap, float* xyz, float* val,
                             float* rgb, int32_t* count_dev, void* stream);
/* Counters since the last clear: [0] sampled pixels, [1] valid pixels (range filter),
 * [2] rays (live buckets), [3] voxel updates applied, [4] out-of-volume samples,
 * [5] allocated voxel blocks, [6] device error flags, [7] integrate launches.
 * Synchronises. */
int tslam_tsdf_get_stats(tslam_tsdf_t* m, int64_t* out8, int clear);
/* Block until all work enqueued on `stream` for this map is done and report a
 * deferred device error (TSLAM_E_POOL_FULL) if one was raised. */
int tslam_tsdf_sync(tslam_tsdf_t* m, void* stream);
/* Kernels launched by this handle since creation (bench.py `gpu_launches`). */
int64_t tslam_tsdf_launch_count(tslam_tsdf_t* m);
/* CUDA-event timing of the integrate kernels, recorded on the launching stream when
 * profiling is on (ring of the last 512 integrate launches).  tslam_tsdf_kernel_ms writes
 * ms3[3*i+{0,1,2}] = bucket / ray-march / commit kernel time of the i-th most recent launch,
 * i < min(n, recorded); *n_out = rows written.  Synchronises. */
int tslam_tsdf_set_profiling(tslam_tsdf_t* m, int on);
int tslam_tsdf_kernel_ms(tslam_tsdf_t* m, int32_t n, float* ms3, int32_t* n_out);
/* Same ring, 7 columns per launch: bucket, ray march (all its kernels), commit, then the ray march split into
 * ray set-up + segment count / scan / segment fill / block march (tslam_march.cu; zeros for the legacy kernel). */
int tslam_tsdf_kernel_ms2(tslam_tsdf_t* m, int32_t n, float* ms7, int32_t* n_out);
/* Diagnostics of the block-binned ray march (process_new_pcl, dense_tsdf.py:236-270) since the last clear of
 * tslam_tsdf_get_stats: out6 = segments, work items, samples that took the exact-index path (fraction within
 * near_eps of .5), samples applied through a global reduction because their exact voxel lies outside the work
 * item's block, samples of the generic path, fast-path indices that disagreed with the exact index
 * (TSLAM_MARCH_VERIFY=1 only; must be 0).  Synchronises. */
int tslam_tsdf_get_march_stats(tslam_tsdf_t* m, int64_t* out6);

/* ----------------------------------------------------------------------------
 * Map queries for planners - batched forms of the @ti.func helpers of BaseMap (mapping_common.py:165-204) that
 * TopoGraphGen calls from its kernels (topo_graph.py:444-507).  All arrays DEVICE, n entries.
 * --------------------------------------------------------------------------*/
/* flags[i]: bit0 = is_pos_occupy (TSDF < 1.8*vs, dense_tsdf.py:152-155 - unobserved space reads 0, i.e. occupied),
 * bit1 = is_pos_unobserved (dense_tsdf.py:148-150). */
int tslam_tsdf_query_points(tslam_tsdf_t* m, int32_t submap, int64_t n, const float* xyz, uint8_t* flags, void* stream);
/* is_near_pos_occupy (mapping_common.py:193-204): any occupied cell in [-voxel, voxel)^3 around the point. */
int tslam_tsdf_query_near_occupy(tslam_tsdf_t* m, int32_t submap, int64_t n, const float* xyz, int32_t voxel, uint8_t* out, void* stream);
/* raycast (mapping_common.py:165-178): step voxel_scale along dir until an occupied cell; hit, last position, length. */
int tslam_tsdf_raycast(tslam_tsdf_t* m, int32_t submap, int64_t n, const float* pos, const float* dir, float max_dist, uint8_t* hit,
                       float* xyz_out, float* len_out, void* stream);

/* ----------------------------------------------------------------------------
 * Multi-GPU: spatially tiled global map (one process per GPU).  The global volume is cut into
 * tiles3[0] x tiles3[1] x tiles3[2] == world tiles of whole 16^3 blocks, tile t owned by rank t.  The library
 * packs / unpacks voxel blocks into caller DEVICE buffers; the caller exchanges them (NCCL all-to-all).
 * Keys are the library's packed block keys (opaque int64).  Planes per block: 4096 voxels.
 * --------------------------------------------------------------------------*/
/* fuse_submaps_kernel (dense_tsdf.py:282-307) of this rank's submaps, sums left PENDING (no commit). */
int tslam_tsdf_fuse_pending(tslam_tsdf_t* dst, tslam_tsdf_t* src, void* stream);
/* Fold pending fusion sums (no Wmax clamp, dense_tsdf.py:274-278). */
int tslam_tsdf_commit_fused(tslam_tsdf_t* m, void* stream);
/* Owner rank of block (bx,by,bz). */
int tslam_tiling_owner(tslam_tsdf_t* m, const int32_t* tiles3, int32_t world, int32_t bx, int32_t by, int32_t bz, int32_t* owner);
/* Tile boundaries.  By default the volume is cut into equal slices; tslam_tiling_set_cuts installs block-coordinate
 * cuts per axis (cuts_a[k] .. cuts_a[k+1]-1 = tile k, tiles3[a]+1 strictly increasing values; NULL tiles3 = back to
 * the default) - the host layer derives them from the occupied blocks (tslam_tsdf_dirty_hist: HOST int32[3][1024],
 * histogram of the block coordinates + 512 of the blocks touched since the last commit, per axis) so that a flat
 * flight volume does not leave half the ranks without surface.  Every rank must install the same cuts. */
int tslam_tiling_set_cuts(tslam_tsdf_t* m, const int32_t* tiles3, const int32_t* cuts_x, const int32_t* cuts_y, const int32_t* cuts_z);
int tslam_tsdf_dirty_hist(tslam_tsdf_t* m, int32_t* hist3x1024, void* stream);
/* counts_out[world] (HOST): touched blocks owned by every other rank.  Synchronises. */
int tslam_tsdf_foreign_count(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, int32_t* counts_out, void* stream);
/* Pack those blocks grouped by destination rank and clear them locally: keys int64[cap], acc f32[cap,4096,2]
 * (pending sums), obs u8[cap,4096], occ i8[cap,4096].  Synchronises. */
int tslam_tsdf_foreign_pack(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, const int32_t* counts,
                            int64_t cap, int64_t* keys, float* acc, uint8_t* obs, int8_t* occ, void* stream);
/* Add received blocks into the local map (sums stay pending until tslam_tsdf_commit_fused). */
int tslam_tsdf_unpack_add(tslam_tsdf_t* m, int64_t n, const int64_t* keys, const float* acc, const uint8_t* obs, const int8_t* occ,
                          void* stream);
/* Halo for meshing: blocks of the one-block boundary layer of this rank's tile, once per neighbouring tile. */
int tslam_tsdf_halo_count(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, int32_t* counts_out, void* stream);
int tslam_tsdf_halo_pack(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, const int32_t* counts, int64_t cap,
                         int64_t* keys, float* tw, uint8_t* obs, void* stream);
/* Insert received halo blocks as GHOSTS: read by marching cubes, never counted / exported / meshed as owners. */
int tslam_tsdf_ghost_unpack(tslam_tsdf_t* m, int64_t n, const int64_t* keys, const float* tw, const uint8_t* obs, void* stream);
/* Textured maps (colour is fused with the geometry, dense_tsdf.py:276-277): the same four calls with the colour
 * plane f32[cap,4096,4] riding along (pending sum of w*colour for the fusion exchange, committed colour for the halo).
 * col must be given for a textured map and NULL for an untextured one. */
int tslam_tsdf_foreign_pack2(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, const int32_t* counts,
                             int64_t cap, int64_t* keys, float* acc, uint8_t* obs, int8_t* occ, float* col, void* stream);
int tslam_tsdf_unpack_add2(tslam_tsdf_t* m, int64_t n, const int64_t* keys, const float* acc, const uint8_t* obs, const int8_t* occ,
                           const float* col, void* stream);
int tslam_tsdf_halo_pack2(tslam_tsdf_t* m, const int32_t* tiles3, int32_t rank, int32_t world, const int32_t* counts, int64_t cap,
                          int64_t* keys, float* tw, uint8_t* obs, float* col, void* stream);
int tslam_tsdf_ghost_unpack2(tslam_tsdf_t* m, int64_t n, const int64_t* keys, const float* tw, const uint8_t* obs, const float* col,
                             void* stream);

/* ----------------------------------------------------------------------------
 * Marching cubes  (MarchingCubeMesher, marching_cube_mesher.py)
 * --------------------------------------------------------------------------*/
/* generate_mesh_kernel (marching_cube_mesher.py:127-187): two-pass (count, scan,
 * emit) marching cubes over every active block of `m`.  verts/normals: DEVICE
 * f32 [3*cap_tri,3]; *n_tri_out = true triangle demand (host).  Synchronises. */
int tslam_mc_generate(tslam_tsdf_t* m, int32_t step, float tsdf_surface_thres, int64_t cap_tri, float* verts,
                      float* normals, int64_t* n_tri_out, void* stream);
/* ... with vertex colours (vertexInterp_color / add_triangle_color, marching_cube_mesher.py:62-82, :104-108):
 * colors f32[3*cap_tri,3] DEVICE or NULL. */
int tslam_mc_generate2(tslam_tsdf_t* m, int32_t step, float tsdf_surface_thres, int64_t cap_tri, float* verts, float* normals,
                       float* colors, int64_t* n_tri_out, void* stream);

/* ----------------------------------------------------------------------------
 * ESDF  (DenseSDF.propogate_esdf semantics, dense_esdf.py:228-333)
 * --------------------------------------------------------------------------*/
/* Converged 26-neighbour signed distance wavefront over the observed voxels of
 * `submap`.  *n_sweeps_out (host, optional) = global sweeps until no change. */
int tslam_esdf_update(tslam_tsdf_t* m, int32_t submap, int32_t* n_sweeps_out, void* stream);
/* Same with control and statistics.  mode 0: INCREMENTAL when the kept state allows it (same submap as the previous
 * update, no reset / load_numpy / ghost import since) - only the voxels that changed since the last update, the voxels
 * whose value may have been derived from them (raise wave) and whatever those lower again are recomputed; the result
 * is bit-identical to a full recompute.  mode 1: full recompute.  stats4 (HOST, may be NULL): lower sweeps,
 * raise sweeps (-1 = the update was full), changed voxels, suspect voxels.  Synchronises. */
int tslam_esdf_update2(tslam_tsdf_t* m, int32_t submap, int32_t mode, int32_t* stats4, void* stream);
/* ESDF of observed voxels: idx int32[cap,3], esdf f32[cap] (DEVICE). Synchronises. */
int tslam_esdf_gather(tslam_tsdf_t* m, int32_t submap, int64_t cap, int32_t* idx, float* esdf, int64_t* n_out,
                      void* stream);

/* ----------------------------------------------------------------------------
 * Octomap = per-voxel hit counter  (Octomap, taichi_octomap.py)
 * --------------------------------------------------------------------------*/
typedef struct tslam_octo tslam_octo_t;
typedef struct tslam_octo_config {
  double voxel_scale;       /* the CONSTRUCTOR's voxel_scale (mapping_common.py:22-23) */
  int32_t N, Nz, K;         /* taichi_octomap.py:19-27                                 */
  double max_ray_length, min_ray_length;
  int32_t recast_step;
  double fx, fy, cx, cy;
  int32_t min_occupy_thres; /* taichi_octomap.py:30,86-88                              */
  int32_t max_submaps;
  int32_t max_blocks;       /* capacity of the 8^3 counter-block pool (0 = derive)     */
  int32_t max_image_pixels;
  int32_t max_points;
  int32_t texture_enabled;  /* taichi_octomap.py:34,77-79: per-voxel colour                */
} tslam_octo_config_t;

int tslam_octo_create(const tslam_octo_config_t* cfg, tslam_octo_t** out);
int tslam_octo_destroy(tslam_octo_t* m);
int tslam_octo_reset(tslam_octo_t* m, void* stream); /* root.deactivate_all() taichi_octomap.py:210-211 */
int tslam_octo_set_submap_pose(tslam_octo_t* m, int32_t s, const float* R9, const float* T3);
/* BaseMap.set_dep_camera_intrinsic (mapping_common.py:25-26). */
int tslam_octo_set_intrinsics(tslam_octo_t* m, double fx, double fy, double cx, double cy);
/* recast_pcl_to_map_kernel (taichi_octomap.py:134-145): occupy[round((R p + T)/vs)] += 1. */
int tslam_octo_integrate_points(tslam_octo_t* m, const float* xyz, int mem, int32_t n, const float* R9, const float* T3,
                                int32_t submap, void* stream);
/* recast_depth_to_map_kernel (taichi_octomap.py:147-169). */
int tslam_octo_integrate_depth(tslam_octo_t* m, const uint16_t* depth, int mem, int32_t h, int32_t w, const float* R9,
                               const float* T3, int32_t submap, void* stream);
/* Textured octomaps (texture_enabled): process_point overwrites color[ijk] with the point's colour, channels swapped
 * BGR -> RGB, /255 (taichi_octomap.py:120-124); racing points, last writer wins.  Deterministic here: the latest
 * integrate call wins, inside a call the largest packed RGB.  rgb uint8 [n,3] / tex uint8 [th,tw,3] in `mem` space;
 * NULL colours integrate hits only.  Fusion copies the colour of the most recently integrated source voxel (:189). */
int tslam_octo_set_color_intrinsics(tslam_octo_t* m, double fx, double fy, double cx, double cy, int color_same_proj);
int tslam_octo_integrate_points_rgb(tslam_octo_t* m, const float* xyz, const uint8_t* rgb, int mem, int32_t n, const float* R9,
                                    const float* T3, int32_t submap, void* stream);
int tslam_octo_integrate_depth_tex(tslam_octo_t* m, const uint16_t* depth, const uint8_t* tex, int mem, int32_t h, int32_t w,
                                   int32_t th, int32_t tw, const float* R9, const float* T3, int32_t submap, void* stream);
/* gather / LoD export with the colour column: color f32[cap,3] / rgb f32[cap,3] DEVICE or NULL (:101-102, :113-114). */
int tslam_octo_gather2(tslam_octo_t* m, int32_t submap, int64_t cap, int32_t* idx, uint32_t* count, float* color, int64_t* n_out,
                       void* stream);
int tslam_octo_extract2(tslam_octo_t* m, int32_t submap, int32_t level, int64_t cap, float* xyz, float* rgb, int32_t* count_dev,
                        void* stream);
/* every (i,j,k,count>0) of `submap`: idx int32[cap,3], count uint32[cap] (DEVICE). Synchronises. */
int tslam_octo_gather(tslam_octo_t* m, int32_t submap, int64_t cap, int32_t* idx, uint32_t* count, int64_t* n_out,
                      void* stream);
/* cvt_occupy_to_voxels(level) / cvt_occupy_voxels_to (taichi_octomap.py:90-114): append to
 * xyz f32[cap,3] at *count_dev. */
int tslam_octo_extract(tslam_octo_t* m, int32_t submap, int32_t level, int64_t cap, float* xyz, int32_t* count_dev,
                       void* stream);
/* fuse_submaps_kernel (taichi_octomap.py:171-189) after reset(). */
int tslam_octo_fuse(tslam_octo_t* dst, tslam_octo_t* src, void* stream);
/* Octomap.is_occupy (taichi_octomap.py:86-88: occupy > min_occupy_thres) for points, and BaseMap.raycast. */
int tslam_octo_query_points(tslam_octo_t* m, int32_t submap, int64_t n, const float* xyz, uint8_t* flags, void* stream);
int tslam_octo_raycast(tslam_octo_t* m, int32_t submap, int64_t n, const float* pos, const float* dir, float max_dist, uint8_t* hit,
                       float* xyz_out, float* len_out, void* stream);
int tslam_octo_sync(tslam_octo_t* m, void* stream);
int64_t tslam_octo_launch_count(tslam_octo_t* m);

#ifdef __cplusplus
}
#endif
#endif /* TSLAM_H */
