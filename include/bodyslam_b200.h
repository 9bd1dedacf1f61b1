/*
 * bodyslam_b200.h -- C ABI of the B200-native BodySLAM depth->3D hot path.
 *
 * The reference (GuidoManni/BodySLAM) is pure Python and has no FFI of its own: the
 * drop-in boundary is the Python signatures of MDEM (`colorize`, `DepthEstimator`) and 3DM
 * (`TSDF`, `RGBD`, `update_map_after_pg`), which `bodyslam_b200/*.py` mirror.  This header is
 * what those mirrors bind (ctypes); each entry point cites the reference
 * interface it replaces (paths relative to /root/reference, R/ = BodySLAM_Refactored/,
 * N/ = BodySLAM_not_refactored/).
 *
 * Conventions
 *   - every function returns 0 on success, a negative BSLAM_E_* code on failure;
 *     bslam_last_error() returns the message of the calling thread's last failure;
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors),
 *     pointers named h_* are HOST pointers; nothing is allocated behind the caller's back
 *     except what bslam_tsdf_create is asked to allocate;
 *   - all work is enqueued on `stream` (a cudaStream_t); no entry point synchronises
 *     unless its comment says it returns a host value;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with BSLAM_E_CUDA.
 */
#ifndef BODYSLAM_B200_H
#define BODYSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSLAM_API __attribute__((visibility("default")))

typedef void *bslam_stream_t;           /* cudaStream_t */
typedef struct bslam_volume bslam_volume; /* opaque dense TSDF box (brick-ordered in HBM) */

enum {
    BSLAM_OK = 0,
    BSLAM_E_ARG = -1,     /* bad argument / shape / unsupported format (Open3D: "Unsupported image format.") */
    BSLAM_E_CUDA = -2,    /* CUDA runtime error, incl. "no device" */
    BSLAM_E_CAPACITY = -3 /* caller-provided output buffer too small */
};

/* brick edge in voxels, frames per integrate launch */
#define BSLAM_BRICK 8
#define BSLAM_MAX_BATCH 256

/* z-march modes of bslam_tsdf_integrate (see DESIGN.md "float32 recurrence") */
#define BSLAM_ZMARCH_BRICK 8   /* camera-space point re-evaluated at every brick base (oracle z_restart=8); in unit-activation
                                  mode at every UNIT base instead = Open3D's literal per-unit recurrence (oracle z_restart=0) */
#define BSLAM_ZMARCH_LITERAL 0 /* Open3D's literal per-column float32 recurrence from z=0 (validation kernel) */

BSLAM_API const char *bslam_last_error(void);
BSLAM_API int bslam_version(void);
/* number of CUDA devices visible (0 when none / no driver); never fails */
BSLAM_API int bslam_device_count(void);

/* ------------------------------------------------------------------ MDEM (K1)
 * Replaces the ZoeDepth `infer_pil(output_type="pil")` tail reached from
 * R/src/depth_estimation/interface.py:61 and N/MDEM/mdem_interface.py:68:
 *     u16 = (metres_f32 * scale_mul).astype(uint16)       (scale_mul = 256)
 * Values are truncated toward zero; out-of-range values saturate to [0, 65535]. */
BSLAM_API int bslam_scale_u16(const float *d_depth, int64_t n, float scale_mul, uint16_t *d_out,
                              bslam_stream_t stream);

/* bytes of device workspace bslam_colorize needs for a batch of B images */
BSLAM_API size_t bslam_colorize_workspace_bytes(int B);

/*
 * Fused metric scaling + colorization of a batch of B images of H*W pixels.
 * Replaces `colorize(value, vmin, vmax, cmap, invalid_val, ..., background_color)`
 * R/examples/depth_estimation/depth_map_scaling.py:12-45 (== batch_processing.py:12-45) for
 * 16-bit integer-valued input, per image:
 *   valid = value != invalid_val;  vmin/vmax = numpy percentile(valid, p_lo / p_hi) (linear);
 *   x = (value - vmin) / (vmax - vmin) in f64 (0 when vmin == vmax);
 *   index = matplotlib Colormap.__call__: trunc(x*256), <0 -> 0, ==256 -> 255, >255 -> 255;
 *   rgba = lut[index]; invalid pixels -> bg_rgba.
 * Input is EITHER d_depth_m (f32 metres; scaled to u16 with scale_mul first and, if
 * d_u16_out != NULL, written there) OR d_u16_in.  d_lut is 256*4 bytes (gamma already folded
 * in by the host).  has_invalid == 0 disables the invalid test (invalid_val outside u16).
 * h_vmin_vmax: optional HOST [B][2] f64 overrides (NaN = compute by percentile).
 * d_vmin_vmax_out: optional DEVICE [B][2] f64, the values used.
 * d_table_override: optional DEVICE [B][65536] u8 index table (value_transform support):
 *   when given, index = d_table_override[b][value] and percentiles are skipped.
 */
BSLAM_API int bslam_colorize(const float *d_depth_m, const uint16_t *d_u16_in, int B, int H, int W,
                             float scale_mul, uint16_t *d_u16_out, uint8_t *d_rgba,
                             const uint8_t *d_lut, double p_lo, double p_hi, int has_invalid,
                             uint16_t invalid_val, uint32_t bg_rgba, const double *h_vmin_vmax,
                             double *d_vmin_vmax_out, const uint8_t *d_table_override,
                             void *d_workspace, bslam_stream_t stream);

/*
 * `colorize` for a FLOAT32 image (metres as ZoeDepth predicts them, numpy or torch -- the reference accepts
 * both, depth_map_scaling.py:14-15): same steps as bslam_colorize, in NumPy's float32 arithmetic
 * (percentile virtual index, lerp, normalisation all float32, as NumPy 2.x evaluates them for a float32 array);
 * the order statistics come from an exact two-level radix select on the floats' order-preserving keys.
 * invalid_val is compared as float (`value == invalid_val`); NaN pixels get the colormap's "bad" colour (0,0,0,0).
 * d_workspace: bslam_colorize_f32_workspace_bytes(B).  h_vmin_vmax / d_vmin_vmax_out as in bslam_colorize.
 */
BSLAM_API size_t bslam_colorize_f32_workspace_bytes(int B);
BSLAM_API int bslam_colorize_f32(const float *d_value, int B, int H, int W, uint8_t *d_rgba, const uint8_t *d_lut,
                                 double p_lo, double p_hi, int has_invalid, float invalid_val, uint32_t bg_rgba,
                                 const double *h_vmin_vmax, double *d_vmin_vmax_out, void *d_workspace,
                                 bslam_stream_t stream);

/* min/max-normalised 8-bit depth, `np.uint8(255*(d-min)/(max-min))`, of
 * N/3DM/slam_utils.py:250-264 (then coloured through d_lut if d_rgb != NULL: 256*3 bytes,
 * e.g. cv2.COLORMAP_JET in BGR order). d_workspace: bslam_colorize_workspace_bytes(B). */
BSLAM_API int bslam_minmax_u8(const uint16_t *d_u16, int B, int H, int W, uint8_t *d_gray,
                              uint8_t *d_rgb, const uint8_t *d_lut3, void *d_workspace,
                              bslam_stream_t stream);

/* median of a u16 image set through the same histogram machinery; used for
 * compute_median_scale_factor (N/EVALUATION/MDEM_eval.py:114-127). d_out: DEVICE [B] f64. */
BSLAM_API int bslam_median_u16(const uint16_t *d_u16, int B, int64_t n_per_image, int has_invalid,
                               uint16_t invalid_val, double *d_out, void *d_workspace,
                               bslam_stream_t stream);

/* ------------------------------------------------------------------ 3DM depth scaling (a4)
 * Replaces Open3D RGBDImage.create_from_color_and_depth(depth_scale, depth_trunc) as called at
 * N/3DM/slam_utils.py:212-220:  d = (float)u16 / (float)depth_scale;  d >= depth_trunc -> 0.
 * depth_trunc <= 0 disables truncation (the cv2 variant, slam_utils.py:231-233). */
BSLAM_API int bslam_depth_from_u16(const uint16_t *d_in, int64_t n, float depth_scale,
                                   float depth_trunc, float *d_out, bslam_stream_t stream);

/* Host helper: n row-major 4x4 float64 matrices inverted by the cofactor (adjugate / determinant) formula, the one
 * Eigen's Matrix4d::inverse() evaluates for Open3D's `extrinsic.inverse()` -- the camera pose of back-projection and
 * unit activation.  No device work. */
BSLAM_API int bslam_invert4x4(const double *h_in, double *h_out, int n);

/* ------------------------------------------------------------------ back-projection (K2)
 * Replaces `pixel_to_3d` N/3DM/scaling_system.py:72-77 applied densely, i.e. Open3D
 * PointCloud.create_from_depth_image / create_from_rgbd_image(depth, intrinsic, extrinsic)
 * as called at N/3DM/mapping_module.py:37,41,42,173:
 *   z = d; x = (u-cx)*z/fx; y = (v-cy)*z/fy; p = cam_to_world * [x,y,z,1]  for d > 0,
 * rows visited with step `stride`, output compacted in row-major order per image.
 * h_K = {fx,fy,cx,cy} (f32, HOST); h_cam_to_world = [B][12] f32 row-major 3x4 (HOST;
 * inverse(extrinsic), computed by the caller in f64).  valid_only == 0 writes one row per
 * visited pixel (NaN for d <= 0) and no compaction happens.
 * d_xyz [capacity][3] f32; d_rgb optional [capacity][3] f32 = u8/255 (needs d_rgb_u8 [B][H][W][3]);
 * d_counts DEVICE [B+1] i64: rows per image and, last, the total.  Images are concatenated.
 * Rows beyond `capacity` are counted but not written (BSLAM_E_CAPACITY is NOT raised; compare). */
BSLAM_API size_t bslam_backproject_workspace_bytes(int B, int H, int W, int stride);
BSLAM_API int bslam_backproject(const float *d_depth, const uint8_t *d_rgb_u8, int B, int H, int W,
                                int stride, const float *h_K, const float *h_cam_to_world,
                                int valid_only, float *d_xyz, float *d_rgb, int64_t capacity,
                                int64_t *d_counts, void *d_workspace, bslam_stream_t stream);

/* ------------------------------------------------------------------ TSDF volume (K3)
 * Replaces `TSDF.__init__` N/3DM/tsdf.py:6-12 for the dense N^3 equivalent of the volume it
 * builds (Open3D UniformTSDFVolume semantics; box of nx*ny*nz voxels whose local z = 0 is
 * global plane gz0 of a larger grid -- z-slab sharding).  World centre of global voxel
 * (X,Y,Z) = origin + (X+.5, Y+.5, Z+.5)*voxel_length.  Storage is brick-ordered
 * (8x8x8 bricks, {tsdf,weight} float2 per voxel, optional f32 rgb planes).
 * d_storage: optional caller-owned device buffer of bslam_tsdf_storage_bytes(); NULL lets the
 * library cudaMalloc on `device`. */
BSLAM_API size_t bslam_tsdf_storage_bytes(int nx, int ny, int nz, int with_color);
BSLAM_API int bslam_tsdf_create(bslam_volume **out, int nx, int ny, int nz, int gz0,
                                double voxel_length, double sdf_trunc, const double *h_origin,
                                int with_color, int device, void *d_storage, bslam_stream_t stream);
BSLAM_API int bslam_tsdf_destroy(bslam_volume *vol);
BSLAM_API int bslam_tsdf_reset(bslam_volume *vol, bslam_stream_t stream);
/* deepcopy of N/3DM/tsdf.py:24 (build_copy_3D_map): dst must have identical geometry */
BSLAM_API int bslam_tsdf_copy(const bslam_volume *src, bslam_volume *dst, bslam_stream_t stream);

/*
 * Replaces `TSDF.build_3D_map(rgbd, intrinsic, extrinsic)` N/3DM/tsdf.py:14-22 (Open3D
 * integrate) for F frames in order (F = 1: the per-frame SLAM loop N/3DM/slam.py:117,179;
 * F > 1: the replay of `update_map_after_pg` N/3DM/slam_utils.py:124-135).
 * d_depth [F][H][W] f32 metres (0 = invalid); d_rgb optional [F][H][W][3] u8 (colour volumes; read
 * through the aligned 32-bit words that hold a pixel's bytes, so up to 3 bytes either side of the
 * buffer are touched -- inside the allocation granule of any CUDA allocator);
 * h_K = {fx,fy,cx,cy} f64 HOST; h_extrinsics [F][16] f64 HOST row-major world->camera.
 * zmarch: BSLAM_ZMARCH_BRICK (fast path) or BSLAM_ZMARCH_LITERAL (validation kernel).
 * d_update_counts: optional DEVICE [F] u64, += number of voxels updated per frame.
 * dry_run != 0 only counts (volume untouched).
 */
BSLAM_API int bslam_tsdf_integrate(bslam_volume *vol, const float *d_depth, const uint8_t *d_rgb,
                                   int F, int H, int W, const double *h_K,
                                   const double *h_extrinsics, int zmarch,
                                   unsigned long long *d_update_counts, int dry_run,
                                   bslam_stream_t stream);

/*
 * The replay shape of `update_map_after_pg` (N/3DM/slam_utils.py:124-135) with the frames as they
 * come off the PNG decoder: uint16 depth in 3DM units.  Fuses `RGBD._read_rgbd_for_tsdf`'s depth
 * conversion (N/3DM/slam_utils.py:212-220: f32(u16) / depth_scale, >= depth_trunc -> 0) into the
 * first pass of the integration; d_depth_scratch [F][H][W] f32 receives the converted frames
 * (what bslam_depth_from_u16 would have produced) and must stay valid until the call's work is done.
 */
BSLAM_API int bslam_tsdf_integrate_u16(bslam_volume *vol, const uint16_t *d_depth_u16, float depth_scale,
                                       float depth_trunc, float *d_depth_scratch, const uint8_t *d_rgb,
                                       int F, int H, int W, const double *h_K, const double *h_extrinsics,
                                       unsigned long long *d_update_counts, bslam_stream_t stream);

/*
 * The same launch in two halves, for callers that stream chunk after chunk (update_map_after_pg-style replays):
 * bslam_tsdf_prepare_u16 enqueues everything that does not touch the volume -- depth conversion + tile statistics,
 * unit marks, culling, claim order -- for F <= BSLAM_MAX_BATCH frames on `prep_stream`;
 * bslam_tsdf_integrate_prepared enqueues the integration of the OLDEST prepared launch on `stream` (it waits for
 * the preparation through an event; no host synchronisation).  Up to two launches may be prepared ahead, so that the
 * preparation of chunk k+1 runs in the SM slots the tail of chunk k's integration leaves idle.  Frame order =
 * preparation order.  The buffers handed to prepare must stay valid until the matching integration is done.
 */
BSLAM_API int bslam_tsdf_prepare_u16(bslam_volume *vol, const uint16_t *d_depth_u16, float depth_scale, float depth_trunc,
                                     float *d_depth_scratch, int F, int H, int W, const double *h_K, const double *h_extrinsics,
                                     bslam_stream_t prep_stream);
BSLAM_API int bslam_tsdf_integrate_prepared(bslam_volume *vol, const uint8_t *d_rgb, unsigned long long *d_update_counts,
                                            bslam_stream_t stream);

/* Round-robin z-sharding: this box holds every `stride_bricks`-th 8-voxel brick layer of the
 * grid, starting at global plane gz0 (= 8 * rank): local plane z is global plane
 * gz0 + (z / 8) * 8 * stride_bricks + z % 8.  Balances integration across ranks whatever the
 * camera looks at; extraction needs contiguous slabs again (re-shard brick layers first, see
 * bodyslam_b200/sharding.py).  stride_bricks = 1 (default) is a contiguous slab. */
BSLAM_API int bslam_tsdf_set_z_interleave(bslam_volume *vol, int stride_bricks);
/* byte offsets {voxels, colour, brick flags, total} inside the storage buffer; a brick layer
 * (all bricks of one bz) is contiguous in each region, which is what re-sharding moves */
BSLAM_API int bslam_tsdf_layout(const bslam_volume *vol, size_t *h_offsets /* [4] */);

/* frames per integrate launch (1..BSLAM_MAX_BATCH, 0 = library default).  Larger batches keep a
 * voxel in registers across more frames; smaller ones keep the batch's depth images L2-resident. */
BSLAM_API int bslam_tsdf_set_batch(bslam_volume *vol, int frames_per_launch);

/*
 * ScalableTSDFVolume semantics -- the volume `TSDF.__init__` literally constructs (N/3DM/tsdf.py:7-12:
 * volume_unit_resolution = 32, depth_sampling_stride = 8).  With unit_resolution > 0 a frame only
 * integrates the unit_resolution^3-voxel units activated by its stride-sampled, back-projected
 * depth points (+- sdf_trunc), and voxel centres are evaluated per unit like Open3D's per-unit
 * UniformTSDFVolume; everything else stays weight 0 exactly like the sparse reference.  The grid
 * must consist of whole units on the world unit grid (origin = k * unit_resolution * voxel_length);
 * z_total: plane count of the whole grid when this volume is a z-shard (0 = nz).  0 switches back
 * to the dense UniformTSDFVolume semantics (default).
 * depth_sampling_stride = -1: "dense with the reference's arithmetic" -- EVERY unit of the box is integrated by every
 * frame (no activation mask), with the per-unit voxel centres and the per-unit float32 z recurrence; on the voxels of
 * the units ScalableTSDFVolume would have activated the result equals the reference's bit for bit.
 */
BSLAM_API int bslam_tsdf_set_unit_activation(bslam_volume *vol, int unit_resolution, int depth_sampling_stride, int z_total);

/* The reference's ScalableTSDFVolume is unbounded; this box is not.  Every integrated frame's
 * stride-sampled depth points (the same sampling ScalableTSDFVolume::Integrate uses) are
 * back-projected and counted: h_stat3 = {points seen, points whose +-sdf_trunc neighbourhood lies
 * entirely outside the box, points partly outside}.  Always on in unit-activation mode; in dense
 * mode bslam_tsdf_set_clip_check(vol, stride > 0, z_total) turns it on (one extra small kernel per
 * launch; z_total = planes of the whole grid when this box is a z-shard, 0 = gz0 + nz).
 * bslam_tsdf_clip_stats synchronises the stream. */
BSLAM_API int bslam_tsdf_set_clip_check(bslam_volume *vol, int sampling_stride, int z_total);
BSLAM_API int bslam_tsdf_clip_stats(bslam_volume *vol, unsigned long long *h_stat3, int reset, bslam_stream_t stream);

/* Chain-length histogram of the last integrate launch: h_hist32[k] = active bricks with 8k+1 .. 8k+8
 * active frames (measurement aid; synchronises). */
BSLAM_API int bslam_tsdf_chain_histogram(bslam_volume *vol, unsigned int *h_hist32, bslam_stream_t stream);

/* z layers per integrate warp: a brick's 8 layers are shared by 2 * 8 / n warps.  8 is the most
 * instruction-efficient; 4 / 2 shorten the serial frame chain of a brick, which bounds the launch
 * time on small shards (8 GPUs).  0 (default) = chosen from the shard's brick count. */
BSLAM_API int bslam_tsdf_set_z_split(bslam_volume *vol, int z_layers_per_warp);

/* Culling statistics accumulated by dry runs (bslam_tsdf_integrate with dry_run = 1) since the last
 * reset: h_stat4 = {(warp, frame) pairs tested, pairs with a voxel projecting into the image, pairs
 * with an updated voxel, voxels tested}.  Measurement aid for DESIGN.md / bench.py; synchronises. */
BSLAM_API int bslam_tsdf_dry_stats(bslam_volume *vol, unsigned long long *h_stat4, int reset, bslam_stream_t stream);

/* Device self-test: n random operand triples through the kernels' shared-reciprocal division and
 * magic-number floor, compared with IEEE `/` and (int) casts; returns the number of mismatches
 * (must be 0).  Synchronises the stream. */
BSLAM_API int bslam_selftest(unsigned long long n, unsigned int seed, unsigned long long *h_mismatches,
                             bslam_stream_t stream);

/* Measurement hook (bench.py roofline): when enabled, the dominant kernel of every integrate
 * launch (brick_integrate_kernel) is bracketed by CUDA events on the launch stream.  Up to 1024
 * launches are buffered between reads; bslam_tsdf_profile_read synchronises on them and returns
 * the accumulated kernel milliseconds and launch count since bslam_tsdf_profile(vol, 1). */
BSLAM_API int bslam_tsdf_profile(bslam_volume *vol, int enable);
BSLAM_API int bslam_tsdf_profile_read(bslam_volume *vol, double *h_ms_total, long long *h_launches);
/* the same events, per stage of a launch: h_ms_stage3 = {depth statistics (+ fused a4), unit marks + culls + claim
 * order, brick_integrate_kernel} accumulated milliseconds (the timeline of one integrate launch) */
BSLAM_API int bslam_tsdf_profile_read_stages(bslam_volume *vol, double *h_ms_stage3, long long *h_launches);

/* brick order <-> Open3D order idx = (x*ny + y)*nz + z (parity / interchange).
 * d_color: [nx*ny*nz*3] f32 or NULL. */
BSLAM_API int bslam_tsdf_export(const bslam_volume *vol, float *d_tsdf, float *d_weight,
                                float *d_color, bslam_stream_t stream);
BSLAM_API int bslam_tsdf_import(bslam_volume *vol, const float *d_tsdf, const float *d_weight,
                                const float *d_color, bslam_stream_t stream);
/* {tsdf,weight} of local plane z as [nx][ny] float2 (halo exchange between z-slabs) */
BSLAM_API int bslam_tsdf_export_plane(const bslam_volume *vol, int z, float *d_plane_f2,
                                      bslam_stream_t stream);

/* ------------------------------------------------------------------ surface extraction (K4)
 * Replaces `TSDF.extract_mesh()` N/3DM/tsdf.py:42-43 (Open3D extract_triangle_mesh):
 * cubes with any zero-weight corner skipped, Bourke tables, vertices shared per edge
 * (x,y,z,axis), triangles (e0,e2,e1).  Two-phase: count (returns HOST counts; synchronises
 * the stream), then emit into caller buffers.
 * d_halo_lo / d_halo_hi: optional [nx][ny] float2 planes z = -1 / z = nz of the neighbouring
 * slabs (NULL = outside the grid = weight 0).  With d_halo_hi the cubes based at local
 * z = nz-1 are emitted here; vertices on edges owned by plane z = nz are NOT (the upper slab,
 * which receives this slab's top plane as its d_halo_lo, emits them).
 * Vertex numbering is deterministic (brick order, 32-voxel chunk, axis, voxel) but differs
 * from Open3D's serial first-seen order; d_keys (optional, [V][4] i32: x,y,z local,axis) lets
 * callers canonicalise.  Vertex ids in d_tri are local to this box unless the edge is owned
 * by plane z = nz, in which case the id is -(1 + (x*ny + y)*4 + axis) (resolved on gather).
 * d_vertices [cap_v][3] f32 world; d_colors optional [cap_v][3] f32 in [0,1] (colour volumes);
 * d_tri [cap_t][3] i32.  Rows beyond capacity are not written. */
BSLAM_API int bslam_mc_count(bslam_volume *vol, const float *d_halo_lo, const float *d_halo_hi,
                             int64_t *h_counts /* [2]: V, T */, bslam_stream_t stream);
BSLAM_API int bslam_mc_emit(bslam_volume *vol, const float *d_halo_lo, const float *d_halo_hi,
                            float *d_vertices, int32_t *d_keys, float *d_colors, int64_t cap_v,
                            int32_t *d_tri, int64_t cap_t, bslam_stream_t stream);

/* Gather side of the z-slab meshes ("per-slab meshes concatenated on gather", north_star): the slabs' d_keys [V][4]
 * and d_tris [T][3] are concatenated bottom slab first (h_nv / h_nt rows each).  In place: key z becomes global
 * (+ h_z_offsets[s]); triangle corners become global ids (+ the slab's vertex base), and the negative ids
 * -(1 + (x*ny + y)*4 + axis) bslam_mc_emit wrote for vertices owned by the next slab's plane 0 are resolved.
 * d_workspace: bslam_mesh_merge_workspace_bytes().  h_unresolved (optional; synchronises): number of
 * references that found no vertex (must be 0). */
BSLAM_API size_t bslam_mesh_merge_workspace_bytes(int n_slabs, int nx, int ny);
BSLAM_API int bslam_mesh_merge(int n_slabs, const int64_t *h_nv, const int64_t *h_nt, const int32_t *h_z_offsets, int nx, int ny,
                               int32_t *d_keys, int32_t *d_tris, void *d_workspace, int64_t *h_unresolved, bslam_stream_t stream);

/* ------------------------------------------------------------------ tensor-pipeline integrator (row f4)
 * Replaces `MAP.integrate(curr_rgbd, i, curr_global_pose)` N/3DM/tsdf.py:71-83, i.e. Open3D
 * `t.pipelines.slam.Model.integrate(frame, depth_scale, depth_max, trunc_voxel_multiplier)` on a VoxelBlockGrid of
 * 16^3 blocks (`MAP.__init__` N/3DM/tsdf.py:57-69), for F frames in order, on the bounded dense box `vol`
 * (origin on the world block grid: origin = k * 16 * voxel_size; created without unit activation):
 * blocks are activated per frame by the depth-touch rule (every 4th pixel, 4 points along the ray over
 * [d - trunc, d + trunc], trunc = trunc_voxel_multiplier * voxel_size), activated voxels get the PROJECTIVE update
 * sdf = depth - z with the depth_max cut (see csrc/bslam_vbg.cu).  d_depth_u16 [F][H][W] raw depth (divided by
 * depth_scale in the kernel, like Open3D), d_rgb optional [F][H][W][3] u8 (colour volumes), h_poses [F][16] f64
 * camera->world (T_frame_to_model), h_depth_max [F] f64 (`curr_rgbd.depth_max`).  d_workspace:
 * bslam_vbg_workspace_bytes() bytes, ZEROED by the caller once; its first 16 bytes accumulate {ray points
 * seen, ray points outside the box} (bslam_vbg_stats; the reference's hash map is unbounded).
 * `synthesize_model_frame` (ray casting, whose result MAP discards) is not provided. */
BSLAM_API size_t bslam_vbg_workspace_bytes(int nx, int ny, int nz);
BSLAM_API int bslam_vbg_integrate(bslam_volume *vol, const uint16_t *d_depth_u16, const uint8_t *d_rgb, int F, int H, int W,
                                  const double *h_K, const double *h_poses, const double *h_depth_max, float depth_scale,
                                  float trunc_voxel_multiplier, void *d_workspace, unsigned long long *d_update_counts,
                                  bslam_stream_t stream);
BSLAM_API int bslam_vbg_stats(const void *d_workspace, unsigned long long *h_stat2, bslam_stream_t stream);
/* Surface-extraction flavour of bslam_mc_* / bslam_points_*: a voxel is valid when weight >= weight_threshold
 * (0 = Open3D's legacy `weight != 0`; the tensor pipeline's extract_triangle_mesh / extract_point_cloud default is 3)
 * and vertices sit at (index + vertex_offset) * voxel_length (legacy 0.5 = voxel centres; tensor pipeline 0). */
BSLAM_API int bslam_tsdf_set_extract_flavour(bslam_volume *vol, float weight_threshold, double vertex_offset);

/* Replaces `TSDF.extract_pcd()` N/3DM/tsdf.py:39-40 (Open3D extract_point_cloud): interior
 * voxels with w != 0 and -0.98 <= f < 0.98, sign change towards +x/+y/+z neighbour, linear
 * zero crossing; normals from the 0.99-voxel central difference of the trilinear TSDF.
 * Single-box volumes only (gz0 = 0 and no halos). */
/* Incremental mode for the reference's per-frame cadence (extract_pcd after every build_3D_map, N/3DM/slam.py:126,195):
 * the integration kernels flag the bricks they change; a count / emit pair then re-extracts only the bricks whose 3x3x3
 * brick neighbourhood changed since the previous pair and copies every other brick's points from a per-brick cache
 * (<= 128 points per brick; larger bricks are always recomputed).  Same output, same order as the full extraction.
 * with_normals fixes whether normals are produced (bslam_points_emit must then be given / not given d_normals); colour
 * volumes need d_colors; every bslam_points_emit needs its own bslam_points_count first and room for all points. */
BSLAM_API int bslam_points_set_incremental(bslam_volume *vol, int enable, int with_normals);
/* {surface-candidate bricks, bricks recomputed} of the last bslam_points_count (host values, no synchronisation) */
BSLAM_API int bslam_points_last_stats(const bslam_volume *vol, long long *h_stat2);
BSLAM_API int bslam_points_count(bslam_volume *vol, int64_t *h_count, bslam_stream_t stream);
BSLAM_API int bslam_points_emit(bslam_volume *vol, float *d_points, float *d_normals,
                                float *d_colors, int32_t *d_keys, int64_t cap,
                                bslam_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BODYSLAM_B200_H */
