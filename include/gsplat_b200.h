/*
 * gsplat_b200.h -- C ABI of the B200-native differentiable 3DGS rasterizer.
 *
 * This is the drop-in boundary for the hot path of scomup/EasyGaussianSplatting: the seven
 * operators its pybind11 module `gsplatcu` exports (reference gsplatcu/ext.cpp:68-76, host
 * launchers gsplatcu/gausplat.cu).  The reference has no C ABI of its own -- its boundary
 * is `std::vector<torch::Tensor> f(torch::Tensor...)` -- so each entry point below is what
 * a maintainer would call from those launchers instead of the `<<<>>>` launches
 * (INTEGRATION.md shows the binding).  The Python package `gsplatcu/` in this repository
 * binds them with ctypes and reproduces the reference's Python operator surface.
 *
 * Conventions (all entry points):
 *   - plain C, no torch / C++ types; every pointer is a DEVICE pointer unless the name ends
 *     in `_host`; tensors are dense row-major float32 (int32 where stated);
 *   - caller owns every buffer, including workspaces (sizes from the *_workspace_bytes
 *     queries); nothing is allocated or freed inside;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*); no device-wide sync.
 *     The only host synchronisation is gsb_splat_bin's wait for the patch count P;
 *   - return 0 on success, non-zero (a cudaError_t value, or -1 for argument errors)
 *     otherwise; gsb_last_error() returns a thread-local description;
 *   - N == 0 / P == 0 are valid and produce empty / zero outputs (the reference reads out
 *     of bounds for N == 0, gausplat.cu:67, and drops the image for P == 1,
 *     kernel.cu:140-143; both are handled here).
 *   - "culled" Gaussians (depths[i] < 0.2) get all-zero outputs, exactly what the
 *     reference's zero-filled tensors hold for them.
 */
#ifndef GSPLAT_B200_H_
#define GSPLAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 1
#define GSB_TILE 16            /* reference common.cuh:13 BLOCK */
#define GSB_RECORD_BYTES 48    /* packed per-Gaussian record, see DESIGN.md "data layout" */

typedef void *gsb_stream_t; /* cudaStream_t */

int gsb_abi_version(void);
const char *gsb_last_error(void);

/* F.1.1/F.1.2/B.1.2.  Replaces `project` (ext.cpp:54-61, gausplat.cu:253-296, kernel.cu:553-617).
 * pws[N,3], Rcw[3,3], tcw[3] -> us[N,2], pcs[N,3], depths[N] (-1 where z < 0.2),
 * du_dpcs[N,2,3] (nullable = calc_J false). */
int gsb_project(int N, const float *pws, const float *Rcw, const float *tcw, float fx, float fy,
                float cx, float cy, float *us, float *pcs, float *depths, float *du_dpcs,
                gsb_stream_t stream);

/* F.2/B.2.  Replaces `computeCov3D` (ext.cpp:39-42, gausplat.cu:162-199, kernel.cu:326-423).
 * rots[N,4] (w,x,y,z; used un-normalised), scales[N,3], depths[N] -> cov3ds[N,6] (upper
 * triangle), dcov3d_drots[N,6,4], dcov3d_dscales[N,6,3] (both nullable together). */
int gsb_compute_cov3d(int N, const float *rots, const float *scales, const float *depths,
                      float *cov3ds, float *dcov3d_drots, float *dcov3d_dscales,
                      gsb_stream_t stream);

/* F.3/B.3.  Replaces `computeCov2D` (ext.cpp:44-52, gausplat.cu:201-251, kernel.cu:425-551).
 * width/height only feed the +-1.3*tan_fov clamp.  -> cov2ds[N,3], dcov2d_dcov3ds[N,3,6],
 * dcov2d_dpcs[N,3,3] (nullable together). */
int gsb_compute_cov2d(int N, const float *cov3ds, const float *pcs, const float *Rcw,
                      const float *depths, float fx, float fy, float width, float height,
                      float *cov2ds, float *dcov2d_dcov3ds, float *dcov2d_dpcs,
                      gsb_stream_t stream);

/* F.4.  Replaces `sh2Color` (ext.cpp:63-66, gausplat.cu:298-338, kernel.cu:619-807).
 * shs[N,3*k] laid out [coef][rgb], k = sh_dim3 in {1,4,9,16}; not gated by depth.
 * -> colors[N,3], dcolor_dshs[N,1,k], dcolor_dpws[N,3,3] (nullable together). */
int gsb_sh2color(int N, int sh_dim3, const float *shs, const float *pws, const float *twc,
                 float *colors, float *dcolor_dshs, float *dcolor_dpws, gsb_stream_t stream);

/* F.5.3/B.5.3.  Replaces `inverseCov2D` (ext.cpp:34-36, gausplat.cu:340-373, kernel.cu:274-324).
 * depths is READ-WRITE: a NaN 1/det sets depths[i] = -1.  -> cinv2ds[N,3], areas[N,2] int32,
 * dcinv2d_dcov2ds[N,3,3] (nullable). */
int gsb_inverse_cov2d(int N, const float *cov2ds, float *depths, float *cinv2ds, int32_t *areas,
                      float *dcinv2d_dcov2ds, gsb_stream_t stream);

/* ---- fused per-Gaussian path (extension).  No single reference entry point:
 * gsb_preprocess_forward == project + computeCov3D + computeCov2D + sh2Color + inverseCov2D
 * with calc_J = false (gsmodel.py:21-36 without the Jacobian outputs) in one kernel, and
 * gsb_preprocess_backward == the Jacobian chain of gsmodel.py:72-85 (torch.bmm over the saved
 * Jacobians) as analytic vector-Jacobian products recomputed from the parameters.
 * width/height feed the fov clamp exactly as in gsb_compute_cov2d.  Outputs of the forward
 * feed gsb_splat_bin / gsb_splat_render unchanged; the backward consumes gsb_splat_backward's
 * dloss_dus / dloss_dcinv2ds / dloss_dcolors (dloss_dalphas passes through untouched) and
 * writes dloss_dpws[N,3], dloss_dshs[N,3k], dloss_dscales[N,3], dloss_drots[N,4].
 * Two optional short cuts around passes that only re-format data:
 *   forward:  alphas[N] + records (both or neither) -- also writes the N packed 48-byte
 *             per-Gaussian records the rasterizers gather from; hand them to gsb_splat_render /
 *             gsb_splat_backward as `packed_records` and the separate pack pass is skipped;
 *   backward: moments[N,9] + cinv2ds[N,3] + dloss_dus_out[N,2] + dloss_dalphas_out[N] (all or
 *             none) -- the upstream gradients are taken as the raw moment rows
 *             gsb_splat_backward leaves in `moments_out`; the three dloss_d* inputs are then
 *             ignored (may be NULL) and dL/du, dL/dalpha are written to the two outputs. */
int gsb_preprocess_forward(int N, int sh_dim3, const float *pws, const float *rots, const float *scales,
                           const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                           float fy, float cx, float cy, float width, float height, float *us,
                           float *cinv2ds, float *colors, float *depths, int32_t *areas, const float *alphas,
                           void *records, gsb_stream_t stream);
int gsb_preprocess_backward(int N, int sh_dim3, const float *pws, const float *rots, const float *scales,
                            const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                            float fy, float cx, float cy, float width, float height, const float *dloss_dus,
                            const float *dloss_dcinv2ds, const float *dloss_dcolors, float *dloss_dpws,
                            float *dloss_dshs, float *dloss_dscales, float *dloss_drots, const float *moments,
                            const float *cinv2ds, float *dloss_dus_out, float *dloss_dalphas_out,
                            gsb_stream_t stream);

/* dL/dsh from per-view dL/dcolor (extension; multi-view data parallelism).
 * colour = 0.5 + sum_l Y_l(dir) sh_l is linear in sh and never clamped (kernel.cu:735-774), so
 * dloss_dshs[n, l, c] = sum_v Y_l((pws[n] - twcs[v]) / |.|) * dloss_dcolors[v, n, c].
 * gsb_preprocess_backward may then be called with dloss_dshs = NULL (it skips the 192-byte row),
 * the V views of a step exchange 12 B per Gaussian and view instead of one 192-byte row, and every
 * rank expands here.  twcs[V,3], dloss_dcolors[V,N,3] -> dloss_dshs[N,3*sh_dim3] (overwritten;
 * views summed in index order; V = 1 reproduces gsb_preprocess_backward's row to rounding). */
int gsb_sh_grad_expand(int N, int sh_dim3, int V, const float *pws, const float *twcs, const float *dloss_dcolors,
                       float *dloss_dshs, gsb_stream_t stream);

/* ---- splat, phase 1: tile rectangles + patch count.
 * Replaces getRects + thrust::inclusive_scan + the D2H read of the total
 * (gausplat.cu:54-67, kernel.cu:82-122).  depths and areas are READ-WRITE (Gaussians that
 * touch no tile get depths = -1, areas = 0).  Leaves rects/offsets in `bin_ws` for phase 2.
 * Writes the patch count to *P_host and the largest depth key (uint32)(depth*1000) of a
 * binned Gaussian to *depth_key_max_host (nullable) after synchronising `stream` (the one
 * host sync; the reference has the same read-back, gausplat.cu:67). */
size_t gsb_splat_bin_workspace_bytes(int N);
int gsb_splat_bin(int H, int W, int N, const float *us, float *depths, int32_t *areas,
                  void *bin_ws, size_t bin_ws_bytes, int64_t *P_host, uint32_t *depth_key_max_host,
                  gsb_stream_t stream);

/* ---- splat, phase 2: keys, duplicate-key radix sort, tile ranges, record packing, draw.
 * Replaces createKeys + thrust::sort_by_key + getRanges + draw (gausplat.cu:69-105,
 * kernel.cu:46-80,125-271).  `alphas` is [N] (or [N,1]).  Outputs: image[3,H,W] planar,
 * contrib[H,W] int32, final_tau[H,W], patch_range_per_tile[T,2] int32 (T = tiles),
 * gsid_per_patch[P] int32 (Gaussian id of every patch in (tile, depth-mm, id) order).
 * Every output element is written (no pre-zeroing needed).
 * depth_key_max: the value phase 1 returned (bounds the sort width; keys are packed into 32
 * bits when tile and depth bits fit), or 0xFFFFFFFF for the reference's full 64-bit layout --
 * the resulting order is the same.  packed_records: NULL, or the per-Gaussian records
 * gsb_preprocess_forward wrote.  With NULL they are built here, sit at
 * ws + gsb_splat_records_offset(...) after the call and may be handed to gsb_splat_backward as
 * `packed_records` while `ws` and the inputs are unchanged. */
size_t gsb_splat_workspace_bytes(int N, int H, int W, int64_t P);
size_t gsb_splat_records_offset(int N, int H, int W, int64_t P);
int gsb_splat_render(int H, int W, int N, int64_t P, uint32_t depth_key_max, const float *us,
                     const float *cinv2ds, const float *alphas, const float *depths, const float *colors,
                     const void *packed_records, const void *bin_ws, void *ws, size_t ws_bytes, float *image,
                     int32_t *contrib, float *final_tau, int32_t *patch_range_per_tile,
                     int32_t *gsid_per_patch, gsb_stream_t stream);

/* ---- splat, both phases without the mid-way host round trip (extension).
 * gsb_splat_bin + gsb_splat_render stall the GPU while the host reads P back to size the sort
 * workspace (the reference does the same, gausplat.cu:64-67).  Here the caller provides CAPACITIES
 * instead -- an upper bound P_cap on the patch count (e.g. 1.25 x the previous frame's P) and on the
 * depth keys (depth_key_cap >= every (uint32)(depth*1000); it fixes the key width) -- all kernels
 * read the actual P from device memory, and the host looks at [P, max depth key, flags] only
 * after the sort and the rasterizer have been enqueued behind the copy: no GPU bubble.
 * Workspaces: bin_ws as for gsb_splat_bin; ws >= gsb_splat_workspace_bytes(N, H, W, P_cap);
 * gsid_per_patch[P_cap] (the first P entries are valid).  status_host[3] (pinned memory for a truly
 * asynchronous copy) receives P, the largest depth key and the flags.
 * Returns 0, or GSB_CAPACITY_EXCEEDED when a bound was too small: the outputs are then undefined and
 * the caller repeats the frame with gsb_splat_bin / gsb_splat_render (or larger capacities);
 * depths / areas have been culled in place exactly as gsb_splat_bin does, which is idempotent. */
#define GSB_CAPACITY_EXCEEDED 2
int gsb_splat_forward(int H, int W, int N, const float *us, const float *cinv2ds, const float *alphas,
                      float *depths, const float *colors, int32_t *areas, const void *packed_records,
                      int64_t P_cap, uint32_t depth_key_cap, void *bin_ws, size_t bin_ws_bytes, void *ws,
                      size_t ws_bytes, float *image, int32_t *contrib, float *final_tau,
                      int32_t *patch_range_per_tile, int32_t *gsid_per_patch, uint32_t *status_host,
                      gsb_stream_t stream);

/* Same, but only ENQUEUES (no host wait, no event): for capture into a CUDA graph.  status_host is
 * written by a stream-ordered copy; the caller reads it after the work has completed (flags bit 1:
 * P_cap exceeded, bit 2: depth_key_cap exceeded, or P > P_cap -> the frame's outputs are invalid). */
int gsb_splat_forward_enqueue(int H, int W, int N, const float *us, const float *cinv2ds, const float *alphas,
                              float *depths, const float *colors, int32_t *areas, const void *packed_records,
                              int64_t P_cap, uint32_t depth_key_cap, void *bin_ws, size_t bin_ws_bytes, void *ws,
                              size_t ws_bytes, float *image, int32_t *contrib, float *final_tau,
                              int32_t *patch_range_per_tile, int32_t *gsid_per_patch, uint32_t *status_host,
                              gsb_stream_t stream);

/* ---- splatB.  Replaces `splatB` (ext.cpp:20-32, gausplat.cu:114-159, kernel.cu:809-950).
 * Consumes the forward's contrib / final_tau / patch_range_per_tile / gsid_per_patch.
 * -> dloss_dus[N,1,2], dloss_dcinv2ds[N,1,3], dloss_dalphas[N,1,1], dloss_dcolors[N,1,3]
 * (every element written).  packed_records: NULL (the per-Gaussian records are rebuilt from the
 * four attribute arrays into `ws`) or the forward's record array.
 * moments_out: NULL, or [N,9] -- the raw per-Gaussian moment rows are left there and the
 * conversion to the four gradient tensors is skipped (they may then be NULL): the caller feeds
 * the rows to gsb_preprocess_backward, which does the conversion in registers. */
size_t gsb_splat_backward_workspace_bytes(int N, int H, int W, int64_t P);
int gsb_splat_backward(int H, int W, int N, int64_t P, const float *us, const float *cinv2ds,
                       const float *alphas, const float *colors, const int32_t *contrib,
                       const float *final_tau, const int32_t *patch_range_per_tile,
                       const int32_t *gsid_per_patch, const float *dloss_dgammas,
                       const void *packed_records, void *ws,
                       size_t ws_bytes, float *dloss_dus, float *dloss_dcinv2ds,
                       float *dloss_dalphas, float *dloss_dcolors, float *moments_out, gsb_stream_t stream);

/* ---- batched tiny matmul (extension).  C[b] = A[b] (m x k) . B[b] (k x n), or B shared by all
 * batches when b_shared != 0; dense row-major f32.  This is the product the reference's
 * GSFunction.backward applies ~12 times over the per-Gaussian Jacobians (gsmodel.py:72-85);
 * ops.py routes torch.matmul on the Jacobian tensors it returns to this entry point so the
 * unmodified reference chain runs at HBM speed instead of through batched-GEMV library calls. */
int gsb_small_bmm(long long batch, int m, int k, int n, const float *A, const float *B, int b_shared, float *C,
                  gsb_stream_t stream);

/* ---- training loss (extension).  Replaces `gau_loss`
 * (gsplat/pytorch_ssim.py:64-67, called at train.py:52) and its autograd backward:
 *   loss = (1 - lambda) mean|image - gt| + lambda (1 - mean SSIM), 11x11 Gaussian window
 *   (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2, image/gt [3,H,W] planar f32.
 * Writes the scalar loss to *loss_out (device) and, if dloss_dimage != NULL, dloss/dimage
 * [3,H,W] -- the tensor loss.backward() would hand to splatB. */
size_t gsb_gau_loss_workspace_bytes(int H, int W);
int gsb_gau_loss(int H, int W, const float *image, const float *gt_image, float loss_lambda, float *loss_out,
                 float *dloss_dimage, void *ws, size_t ws_bytes, gsb_stream_t stream);

/* ---- density control (extension).  Replaces GSModel.update_density_info,
 * update_gaussian_density and reset_alpha with prune_params / update_params
 * (gsplat/gsmodel.py:132-166, 219-331): one classify pass, one scan, one pass that moves every
 * surviving row of the 6 parameter tensors and their Adam moments exactly once and appends the
 * clones and splits, instead of 18 boolean-index + torch.cat rebuilds.
 *
 * gsb_gaussians: the reference's six training tensors (gsmodel.py:114-127), device pointers
 * to dense row-major f32.  For the Adam moments pass two more of these (exp_avg, exp_avg_sq);
 * a NULL struct pointer means "the optimizer has no state yet" (update_params' else branch). */
typedef struct gsb_gaussians {
  float *pws;        /* [N,3]  */
  float *low_shs;    /* [N,3]  */
  float *high_shs;   /* [N,45] */
  float *alphas_raw; /* [N,1]  logit of opacity */
  float *scales_raw; /* [N,3]  log of scale     */
  float *rots_raw;   /* [N,4]  (w,x,y,z), un-normalised */
} gsb_gaussians;

/* update_density_info (:219-234): grad_accum[i] += |dloss_dus[i]|, cunt[i] += 1 where mask[i]
 * (uint8 0/1).  first != 0 reproduces the first call after a density update: grad_accum is
 * overwritten with the norm of EVERY Gaussian and cunt with the mask (:228-229). */
int gsb_density_accumulate(int64_t N, const float *dloss_dus, const uint8_t *mask, float *grad_accum,
                           int32_t *cunt, int first, gsb_stream_t stream);

/* Classification + output slots (:238-262).  Class per Gaussian into cls[N] (0 keep, 1 keep +
 * clone, 2 keep + split, 3 prune): prune if alphas_raw < alpha_raw_min or max(scales_raw) >
 * scale_raw_max; else g = grad_accum / cunt (0/0 -> 0); g >= grad_min selects it, and
 * max(exp(scales_raw)) <= scale_clone_max decides clone vs split.  slots[N,3] int32 = the
 * exclusive counts of survivors / clones / splits before each Gaussian.
 * counts_host[3] <- (K survivors, C clones, S splits).  HOST SYNCHRONISATION: waits for the
 * counts, like the reference's int(torch.sum(...)) reads -- the caller sizes the outputs
 * (K + C + S rows) and draws the S x 3 unit normals from them. */
size_t gsb_density_workspace_bytes(int64_t N);
int gsb_density_plan(int64_t N, const float *alphas_raw, const float *scales_raw, const float *grad_accum,
                     const int32_t *cunt, float alpha_raw_min, float scale_raw_max, float grad_min,
                     float scale_clone_max, void *ws, size_t ws_bytes, uint8_t *cls, int32_t *slots,
                     int64_t *counts_host, gsb_stream_t stream);

/* The rebuild (:236-318 + :132-166).  dst rows: survivors in order, then clones, then splits.
 * A clone row is (pws, shs, logit(sigmoid(alphas_raw)), log(exp(scales_raw)),
 * normalize(rots_raw)); a split row moves pws by R(q) (z * exp(scales_raw)) with z[S,3] the
 * caller's unit normals (the reference draws torch.normal(0, std=scales) = z * std, :276-277)
 * and shrinks the NEW Gaussian's scale by 0.6 (the original keeps its size, :281-282).  New
 * rows get zero Adam moments.  src_m/src_v/dst_m/dst_v: NULL, or all four set. */
int gsb_density_apply(int64_t N, const uint8_t *cls, const int32_t *slots, int64_t K, int64_t C, int64_t S,
                      const gsb_gaussians *src, const gsb_gaussians *src_m, const gsb_gaussians *src_v,
                      const float *z, const gsb_gaussians *dst, const gsb_gaussians *dst_m,
                      const gsb_gaussians *dst_v, gsb_stream_t stream);

/* reset_alpha (:320-331): alphas_raw = min(alphas_raw, reset_raw); the alpha group's Adam
 * moments (nullable) are zeroed. */
int gsb_reset_alpha(int64_t N, float *alphas_raw, float *exp_avg, float *exp_avg_sq, float reset_raw,
                    gsb_stream_t stream);

/* ---- Gaussian record conversion (extension; SURVEY 8f row N3, gsplat/gau_io.py).  A "gs row"
 * is the reference's .npy record (gau_io.py:7-12): pw[3] rot[4] scale[3] alpha sh[sh_dim],
 * 11 + sh_dim floats, activated values.
 * gsb_ply_rows_to_gs: load_ply's arithmetic (:60-107) on the PLY vertex block rows[N,stride]:
 *   colmap[11 + sh_dim] (device, int32) = source column of each output column (the host folds
 *   the f_rest channel-major -> coefficient-major transpose of :91 into it); rot / |rot|,
 *   exp(scale), sigmoid(opacity).
 * gsb_gs_to_params: get_training_params (gsmodel.py:95-113): gs rows -> the six raw tensors,
 *   high_shs padded with 0.001 up to 45 columns.   sh_dim in {3, 12, 27, 48}.
 * gsb_params_to_gs: save_training_params (gau_io.py:138-153): -> gs rows with sh_dim = 48. */
int gsb_ply_rows_to_gs(int64_t N, int stride, int sh_dim, const float *rows, const int32_t *colmap,
                       float *gs_rows, gsb_stream_t stream);
int gsb_gs_to_params(int64_t N, int sh_dim, const float *gs_rows, const gsb_gaussians *dst, gsb_stream_t stream);
int gsb_params_to_gs(int64_t N, const gsb_gaussians *src, float *gs_rows, gsb_stream_t stream);

/* ---- multi-GPU gradient exchange (extension; SURVEY 8e -- the reference is single-GPU).
 * Multi-view data parallelism: one process per GPU, Gaussians replicated, one camera per rank;
 * the parameter gradients (dshs, drots, dpws, dscales, dalphas = 3k + 11 floats per Gaussian)
 * must be summed over ranks.  Instead of an all-reduce after the backward, the fused
 * per-Gaussian backward PUSHES each gradient tile into peer memory while it computes
 * (gsb_preprocess_backward_push), and one kernel sums and broadcasts
 * (gsb_grad_reduce_broadcast).  Protocol, per step, on every rank, same `epoch` (1, 2, 3, ...):
 *     gsb_preprocess_backward_push(...);  gsb_grad_reduce_broadcast(...);
 * after which the summed gradients sit in the rank's own region at
 * gsb_exchange_result_offset(..., segment), segment 0..4 = dshs [R,3k], drots [R,4], dpws [R,3],
 * dscales [R,3], dalphas [R,1] with R = world * rows-per-rank >= N (rows >= N are padding).
 *
 * A "region" is device memory of gsb_exchange_region_bytes(N, sh_dim3, world) bytes obtained
 * from gsb_comm_alloc (cudaMalloc + zero fill + CUDA IPC handle, GSB_COMM_HANDLE_BYTES bytes);
 * the handles are exchanged between the processes by the caller (any transport) and opened
 * with gsb_comm_open.  regions_host[world]: HOST array of device pointers, entry r = rank r's
 * region as seen from this process (own pointer for r == rank).  world in {1, 2, 4, 8}.
 * gsb_preprocess_backward_push takes the upstream gradients either as the four dloss_d* tensors
 * or, like gsb_preprocess_backward, as moments[N,9] + cinv2ds[N,3] (+ dloss_dus_out[N,2]).
 * The kernels spin on flags in peer memory with a 5 s timeout; gsb_exchange_status returns 1
 * if a wait ever timed out (a peer missing) -- the results are then invalid. */
#define GSB_COMM_HANDLE_BYTES 64
size_t gsb_exchange_region_bytes(int N, int sh_dim3, int world);
size_t gsb_exchange_result_offset(int N, int sh_dim3, int world, int segment);
int gsb_comm_alloc(size_t bytes, void **ptr, void *handle_out);
int gsb_comm_open(const void *handle, void **peer_ptr);
int gsb_comm_close(void *peer_ptr);
int gsb_comm_free(void *ptr);
int gsb_exchange_status(const void *region, int *status_host);
int gsb_preprocess_backward_push(int N, int sh_dim3, const float *pws, const float *rots, const float *scales,
                                 const float *shs, const float *Rcw, const float *tcw, const float *twc,
                                 float fx, float fy, float cx, float cy, float width, float height,
                                 const float *dloss_dus, const float *dloss_dcinv2ds, const float *dloss_dcolors,
                                 const float *dloss_dalphas, const float *moments, const float *cinv2ds,
                                 float *dloss_dus_out, int world, int rank, void *const *regions_host,
                                 uint32_t epoch, gsb_stream_t stream);
int gsb_grad_reduce_broadcast(int N, int sh_dim3, int world, int rank, void *const *regions_host, uint32_t epoch,
                              gsb_stream_t stream);

/* ---- launch accounting and optional per-kernel timing (no reference counterpart; used by
 * bench.py for `gpu_launches` and the roofline's live CUDA-event kernel durations).
 * Kernel ids 0..gsb_profile_kernels()-1, names from gsb_profile_kernel_name().
 * gsb_profile_launches(id) counts launches since load (id < 0: all kernels).
 * While enabled, every kernel launch is bracketed by two events on its stream;
 * gsb_profile_read() waits for them, returns the summed duration and clears the list. */
void gsb_profile_enable(int on);
int gsb_profile_kernels(void);
const char *gsb_profile_kernel_name(int id);
long long gsb_profile_launches(int id);
int gsb_profile_read(int id, double *ms_total, long long *timed_launches);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H_ */
