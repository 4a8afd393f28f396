/*
 * nsig.h — C ABI of libnsig_b200.so: the B200 (sm_100a) native replacement for the
 * ray-batch render/train hot path of luo-ziyuan/NeRF_Signature.
 *
 * Every entry point
 *   - takes raw DEVICE pointers, sizes and scalars, and the CUDA stream to run on;
 *   - never allocates, frees or synchronises (stream-ordered, re-entrant);
 *   - returns 0 on success, a cudaError_t (>0) from the launch, or NSIG_EINVAL (-1)
 *     for arguments the kernels cannot honour.
 * All tensors are contiguous row-major; float = IEEE fp32, int = int32.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef NSIG_H_
#define NSIG_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSIG_EINVAL (-1)
#define NSIG_MAX_LEVELS 16      /* base encoder levels (hash_encoding.py:49)            */
#define NSIG_MAX_MSG_TABLES 256 /* 2*message_dim tables (hash_encoding_wtmk_bit.py:64)  */

typedef void* nsig_stream_t; /* cudaStream_t */

/* library identity: "nsig_b200 <version> sm_100a" */
const char* nsig_version(void);

/* ------------------------------------------------------------------------- */
/* raymarching utilities — raymarching/src/raymarching.h:7-11                 */
/* ------------------------------------------------------------------------- */

/* raymarching.h:7  near_far_from_aabb (kernel raymarching.cu:92-145) */
int nsig_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                            uint32_t N, float min_near, float* nears, float* fars,
                            nsig_stream_t stream);

/* raymarching.h:8  sph_from_ray (raymarching.cu:163-198) */
int nsig_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                      float* coords, nsig_stream_t stream);

/* raymarching.h:9  morton3D (raymarching.cu:214-226) */
int nsig_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, nsig_stream_t stream);

/* raymarching.h:10 morton3D_invert (raymarching.cu:237-254) */
int nsig_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords,
                         nsig_stream_t stream);

/* raymarching.h:11 packbits (raymarching.cu:268-289); N = number of output BYTES */
int nsig_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                  nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* training march / composite — raymarching/src/raymarching.h:13-15           */
/* ------------------------------------------------------------------------- */

/* Scratch (bytes) nsig_march_rays_train needs for N rays (per-ray counts + scan state). */
size_t nsig_march_rays_train_scratch_bytes(uint32_t N);

/* raymarching.h:13 march_rays_train (raymarching.cu:312-480).
 * Same arguments as the reference plus `scratch`.  Differences that are allowed by the
 * reference's own nondeterminism (atomicAdd order): ray n always owns row n of
 * `rays` and sample offsets are the exclusive prefix sum of the per-ray counts in ray
 * order, starting at counter[0]'s value on entry.  counter[0] += total samples,
 * counter[1] += N, exactly as the reference's atomics leave them.
 * xyzs/dirs/deltas rows past counter[0] are NOT touched (the reference wrapper zero-fills
 * them before the call, raymarching.py:205-207; use nsig_zero_sample_padding), except that
 * when rays are dropped for lack of room the rows after the last kept ray are cleared. */
int nsig_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                          float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                          uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas,
                          int32_t* rays, int32_t* counter, const float* noises,
                          void* scratch, nsig_stream_t stream);

/* nsig_march_rays_train on a grid of at most `max_blocks` CTAs of 8 warps (0 = no limit): each warp walks rays
 * n, n + 8 * max_blocks, ...  Same outputs bit for bit.  For a march that runs on a branch next to other kernels and
 * must leave them room on every SM (the training harness issues the march of batch t+1 beside the decoder of step t);
 * the reference has no counterpart (its march_rays_train always owns the device, raymarching.cu:312). */
int nsig_march_rays_train_limited(const float* rays_o, const float* rays_d, const uint8_t* grid,
                                  float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                  uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                  const float* fars, float* xyzs, float* dirs, float* deltas,
                                  int32_t* rays, int32_t* counter, const float* noises,
                                  void* scratch, uint32_t max_blocks, nsig_stream_t stream);

/* Zero rows [counter[0], end) of the three sample buffers — the rows the reference gets from
 * torch.zeros (raymarching.py:205-207): end = min(align_up(counter[0]), M) with
 * align_up(m) = m + align - m % align (adds a full `align` when already aligned,
 * raymarching.py:224-229), or end = M when align == 0 (mean_count mode returns all M rows). */
int nsig_zero_sample_padding(float* xyzs, float* dirs, float* deltas, const int32_t* counter,
                             uint32_t align, uint32_t M, nsig_stream_t stream);

/* raymarching.h:14 composite_rays_train_forward (raymarching.cu:501-577) */
int nsig_composite_rays_train_forward(const float* sigmas, const float* rgbs,
                                      const float* deltas, const int32_t* rays, uint32_t M,
                                      uint32_t N, float T_thresh, float* weights_sum,
                                      float* depth, float* image, nsig_stream_t stream);

/* raymarching.h:15 composite_rays_train_backward (raymarching.cu:602-682).
 * Every sample row owned by a (non-dropped) ray is written, zeros after early termination;
 * rows owned by no ray (padding, dropped rays) are left untouched — the reference wrapper
 * zero-fills the buffers first (raymarching.py:283-284). */
int nsig_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                       const float* sigmas, const float* rgbs,
                                       const float* deltas, const int32_t* rays,
                                       const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh,
                                       float* grad_sigmas, float* grad_rgbs,
                                       nsig_stream_t stream);

/* composite_rays_train_forward plus the epilogue NeRFRenderer.run_cuda applies to its outputs in the training
 * branch (renderer_wtmk.py:298-303), in the same kernel:
 *     image_out = image + (1 - weights_sum)[:, None] * bg_color      (scalar bg_color)
 *     depth_out = clamp(depth - nears, min=0) / (fars - nears)
 * weights_sum / depth / image still receive the raw composite outputs (the backward needs them).
 * The backward takes the gradient of image_out (and optionally of weights_sum, NULL = none): the background term
 * contributes -bg_color * sum_c grad_image_out[c] to d/d weights_sum.  depth carries no gradient, as in the
 * reference (raymarching.py:275). */
int nsig_composite_rays_train_blend_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                            const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                            float bg_color, const float* nears, const float* fars,
                                            float* weights_sum, float* depth, float* image,
                                            float* image_out, float* depth_out, nsig_stream_t stream);
int nsig_composite_rays_train_blend_backward(const float* grad_weights_sum, const float* grad_image_out,
                                             const float* sigmas, const float* rgbs, const float* deltas,
                                             const int32_t* rays, const float* weights_sum,
                                             const float* image, uint32_t M, uint32_t N, float T_thresh,
                                             float bg_color, float* grad_sigmas, float* grad_rgbs,
                                             nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* inference march / composite — raymarching/src/raymarching.h:17-18          */
/* ------------------------------------------------------------------------- */

/* raymarching.h:17 march_rays (raymarching.cu:701-805). xyzs/dirs/deltas must be
 * zero-filled by the caller (raymarching.py:333-335). */
int nsig_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive,
                    const float* rays_t, const float* rays_o, const float* rays_d,
                    float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                    const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                    float* dirs, float* deltas, const float* noises, nsig_stream_t stream);

/* raymarching.h:18 composite_rays (raymarching.cu:819-905); in-place on rays_alive,
 * rays_t, weights_sum, depth, image. */
int nsig_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs,
                        const float* deltas, float* weights_sum, float* depth, float* image,
                        nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* hash encoders — hash_encoding.py:48-111, hash_encoding_wtmk_bit.py:51-116   */
/* ------------------------------------------------------------------------- */

/* HashEmbedder.forward (hash_encoding.py:96-111): x[B,3] (already normalised to the
 * bounding box [0,1]) -> out[B, 2*n_levels] fp32.  tables[l] is embeddings[l].weight
 * ([2^log2_T, 2] fp32, device pointer; the array itself is a HOST array).
 * resolutions[l] is floor(base*b**l) computed by the caller with the reference's torch
 * fp32 expression (hash_encoding.py:100).  slots (optional, may be NULL):
 * int32 [B, n_levels, 8] hashed voxel indices (hash_encoding.py:43-44) for parity tests. */
int nsig_hash_encode_forward(const float* x, uint32_t B, const float* const* tables,
                             const float* resolutions, uint32_t n_levels, uint32_t log2_T,
                             float* out, int32_t* slots, nsig_stream_t stream);

/* d(loss)/d(embeddings[l].weight) of the above: scatter-add of grad_out[B, 2*n_levels]
 * into grad_tables[l] ([2^log2_T, 2] fp32, accumulated, caller zero-fills). */
int nsig_hash_encode_backward(const float* x, const float* grad_out, uint32_t B,
                              float* const* grad_tables, const float* resolutions,
                              uint32_t n_levels, uint32_t log2_T, nsig_stream_t stream);

/* Watermark-bit encoder, hash_encoding_wtmk_bit.py:99-116, in its algebraically equal
 * pre-summed form: all message_dim "levels" share one resolution, so
 *   out = trilerp(S)[x],  S = sum_i embeddings[2*i + bit_i].weight.
 * nsig_msg_table_sum builds S ([2^log2_T,2] fp32) from the DEVICE message vector
 * (float 0/1, read on device: no .item() sync, cf. hash_encoding_wtmk_bit.py:110). */
/* elem_begin / elem_count (floats, multiples of 8; count 0 = the whole table): the slice of S to build - with the
 * optimizer state sharded over the ranks of a data-parallel job every rank sums only the table slice it owns and the
 * slices of S are all-gathered (parallel.py). */
int nsig_msg_table_sum(const float* const* tables, uint32_t message_dim, const float* message,
                       uint32_t log2_T, float* S, uint32_t elem_begin, uint32_t elem_count,
                       nsig_stream_t stream);

/* The same encoder in the reference's literal per-bit form (message_dim gathers per
 * corner, summed in bit order) — kept for parity tests against the pre-summed form. */
int nsig_msg_encode_forward_perbit(const float* x, uint32_t B, const float* const* tables,
                                   uint32_t message_dim, const float* message,
                                   float resolution, uint32_t log2_T, float* out,
                                   nsig_stream_t stream);

/* Parity probe: the hashed slots (hash_encoding.py:43-44) exactly as the FUSED kernels
 * (nsig_field_forward / nsig_render_rays / nsig_grid_sweep) derive them — same device
 * function, csrc/hash_common.cuh locate_fused — in the layout of nsig_hash_encode_forward's
 * `slots` ([B, n_levels, 8] int32).  weights (optional): [B, n_levels, 3] interpolation
 * weights of the fused path.  Used by tests/test_hash_gpu.py to pin the fused path's
 * integer outputs to the reference-order encoder on cell-boundary inputs. */
int nsig_fused_hash_slots(const float* x, uint32_t B, const float* resolutions, uint32_t n_levels,
                          uint32_t log2_T, int32_t* slots, float* weights, nsig_stream_t stream);

/* half2 shadow copies of n_levels base tables for the fused kernels (BASELINE north_star kernel 2: "vectorised
 * half2 loads"): tables_h2[l][i] = fp16(tables[l][i] * 2^k_l) with 2^k_l the power of two that puts the level's
 * largest magnitude in [2^14, 2^15); inv_scale[l] = 2^-k_l (device float[n_levels]).  absmax_scratch: device
 * uint32[n_levels].  Both pointer arrays are HOST arrays of device pointers.  The base tables are frozen in
 * watermark training (network_wtmk_tcnn.py:90-95), so this runs once; clean training refreshes it per step. */
int nsig_tables_to_half2(const float* const* tables, uint32_t n_levels, uint32_t log2_T,
                         void* const* tables_h2, float* inv_scale, uint32_t* absmax_scratch,
                         nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* field network — nerf/network_wtmk_tcnn.py:97-124, nerf/network_hash.py:77-110 */
/* ------------------------------------------------------------------------- */

/* Weights of the two bias-free 64-wide MLPs that the reference instantiates through
 * tiny-cuda-nn (network_wtmk_tcnn.py:52-62, 78-88), fp16, row-major [out, in]:
 *   sigma: W0[64,32] W1[16,64]            (3072 halfs)
 *   color: W0[64,32] W1[64,64] W2[16,64]  (7168 halfs; input col 31 and output rows 3..15
 *                                           are padding)                                  */
#define NSIG_SIGMA_PARAMS 3072
#define NSIG_COLOR_PARAMS 7168

/* Fused field forward: NeRFNetwork.forward(x, d, message) (network_wtmk_tcnn.py:97-124).
 *   xyzs[M,3] in [-bound,bound], dirs[M,3] unit vectors
 *   base tables/resolutions as in nsig_hash_encode_forward (16 levels, F=2)
 *   S: pre-summed message table or NULL (message=None / clean network_hash.py)
 *   sigma_w / color_w: fp16 weights (layout above)
 *   density_scale: sigmas = density_scale * trunc_exp(h0)  (renderer_wtmk.py:294; pass 1 for the
 *          bare network output)
 *   M_dev (optional): device int32 holding the live sample count (the march counter); the
 *          kernel processes min(M, *M_dev) rows, so no host sync is needed to size the launch
 * outputs: sigmas[M] fp32, rgbs[M,3] fp32 (sigmoid applied)
 *          feat_out (optional) [M,32] fp16: encoder output incl. message feature, saved for
 *          nsig_field_backward / nsig_field_backward_tc (which recompute the MLPs from it).
 *          masks_out (optional) [M,4] x 8 bytes: the ReLU sign masks of the three hidden layers in the
 *          fragment order of the kernels (entry (row, q) = {m1s | m1c << 8, m2c} of quad thread q; within a
 *          layer's 16 bits, bit nt <-> hidden unit nt*8 + 2*q, bit 16 + nt <-> unit nt*8 + 2*q + 1), saved for
 *          nsig_field_backward_masks.
 *   tables_h2 / h2_inv_scale (optional, both or neither): half2 shadow copies of the 16 base tables and their
 *          device float[16] de-scaling factors, as nsig_tables_to_half2 writes them.  When given, the kernel
 *          gathers those (one 32-bit load per corner) instead of the fp32 tables; hash slots are unchanged, the
 *          features carry one extra fp16 rounding (<= 2^-11 of the largest corner value), within the path's 1e-3. */
int nsig_field_forward(const float* xyzs, const float* dirs, uint32_t M, float bound,
                       const float* const* tables, const float* resolutions, uint32_t log2_T,
                       const float* S, float msg_resolution, const void* sigma_w,
                       const void* color_w, float density_scale, const int32_t* M_dev,
                       float* sigmas, float* rgbs, void* feat_out, void* masks_out,
                       const void* const* tables_h2, const float* h2_inv_scale, nsig_stream_t stream);

/* Density-only variant: NeRFNetwork.density (network_wtmk_tcnn.py:126-143);
 * geo_feat (optional) [M,15] fp16. */
int nsig_field_density(const float* xyzs, uint32_t M, float bound, const float* const* tables,
                       const float* resolutions, uint32_t log2_T, const float* S,
                       float msg_resolution, const void* sigma_w, float density_scale,
                       float* sigmas, void* geo_feat, const void* const* tables_h2,
                       const float* h2_inv_scale, nsig_stream_t stream);

/* Colour branch only: NeRFNetwork.color (network_wtmk_tcnn.py:146-176): SH4(dirs) ++ geo_feat
 * ([M,15] fp16 from nsig_field_density) -> colour MLP -> sigmoid -> rgbs[M,3] fp32. */
int nsig_color_forward(const float* dirs, const void* geo_feat, uint32_t M, const void* color_w,
                       float* rgbs, nsig_stream_t stream);

/* Fused frame renderer = the whole inference branch of NeRFRenderer.run_cuda (nerf/renderer_wtmk.py:323-372:
 * near_far_from_aabb + the host loop over march_rays -> forward(x,d,message) -> composite_rays with alive-ray
 * compaction) for N rays in one persistent kernel; a warp owns a ray until it terminates (T < T_thresh) or
 * leaves the box.  Field arguments as in nsig_field_forward.  Outputs are the loop's accumulators before the
 * background blend (renderer_wtmk.py:369-370 stay in the caller): weights_sum[N], depth[N] = sum w*t,
 * image[N,3]; nears/fars[N] optional.  work_counter: device uint32, MUST be zero on entry (dynamic ray
 * scheduling).  sample_count (optional): device uint32, incremented by the number of samples evaluated.
 * noises (optional) [N]: start-offset noise for perturb=True (renderer_wtmk.py:353). */
int nsig_render_rays(const float* rays_o, const float* rays_d, uint32_t N, const float* aabb, float min_near,
                     float bound, const uint8_t* grid, uint32_t C, uint32_t H, float dt_gamma,
                     uint32_t max_steps, float T_thresh, const float* noises, const float* const* tables,
                     const float* resolutions, uint32_t log2_T, const float* S, float msg_resolution,
                     const void* sigma_w, const void* color_w, float density_scale, uint32_t* work_counter,
                     float* weights_sum, float* depth, float* image, float* nears, float* fars,
                     uint32_t* sample_count, const void* const* tables_h2, const float* h2_inv_scale,
                     nsig_stream_t stream);

/* Fused field backward (watermark mode: MLP dgrad only):
 * given dL/dsigma[M], dL/drgb[M,3] and the saved feat[M,32], recompute the MLP
 * activations, back-propagate to the encoder output and scatter-add the gradient of
 * channels 30,31 into G ([2^log2_T,2] fp32 = gradient of the pre-summed table S,
 * caller zero-fills).  grad_feat (optional) [M,32] fp32 receives the full encoder-output
 * gradient (needed for base-table training in clean mode).
 * grad_sigma_w / grad_color_w (optional, both or neither): fp32 [NSIG_SIGMA_PARAMS] / [NSIG_COLOR_PARAMS]
 * gradients of the MLP weights in the layout of sigma_w / color_w (clean-model training,
 * network_hash.py:154-161), ACCUMULATED into the buffers (caller zero-fills): tensor-core contraction over the
 * rows of each 16-sample tile, fp32 accumulation per CTA in shared memory, one global atomic per weight and CTA. */
int nsig_field_backward(const float* xyzs, const float* dirs, uint32_t M, float bound,
                        const void* feat, const float* grad_sigmas, const float* grad_rgbs,
                        const void* sigma_w, const void* color_w, float density_scale,
                        const int32_t* M_dev, float msg_resolution, uint32_t log2_T, float* G,
                        float* grad_feat, float* grad_sigma_w, float* grad_color_w,
                        nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* occupancy grid — nerf/renderer_wtmk.py:380-538; ray generation — nerf/utils_wtmk_disen.py:59-143 */
/* ------------------------------------------------------------------------- */

/* The density sweep of NeRFRenderer.update_extra_state (renderer_wtmk.py:456-514) in one launch for all
 * cascades: cell -> centre + jitter -> NeRFNetwork.density -> sigma * density_scale.
 *   cells == NULL (full update, n must be H^3): every cell of every cascade once, in Morton order, and
 *          the EMA density_grid = max(density_grid * decay, sigma) where both are >= 0 (renderer_wtmk.py:521-523)
 *          is applied in place; *sum (optional, device double) += sum of max(density_grid, 0).
 *   cells != NULL ([C, n] Morton indices, e.g. from nsig_grid_sample_cells): partial update; sigma is
 *          recorded with an atomic max in tmp_grid ([C, H^3], caller fills with -1; where the reference's
 *          index_put keeps an arbitrary one of duplicate draws, this keeps the largest) and
 *          nsig_grid_finalize applies the EMA.
 *   noise (optional) [C, n, 3] in [0,1): the torch.rand_like values of renderer_wtmk.py:479,505; NULL = an
 *          in-kernel Philox4x32-10 stream keyed by `seed` (the reference's stream is torch's global
 *          generator, which no other implementation can reproduce).
 * Field arguments as in nsig_field_density.  bound is a double because the reference evaluates
 * bound - bound/H in python floats before rounding to fp32. */
int nsig_grid_sweep(float* density_grid, float* tmp_grid, const int32_t* cells, uint32_t n,
                    const float* noise, uint64_t seed, uint32_t C, uint32_t H, double bound, float decay,
                    const float* const* tables, const float* resolutions, uint32_t log2_T,
                    const float* S, float msg_resolution, const void* sigma_w, float density_scale,
                    double* sum, const void* const* tables_h2, const float* h2_inv_scale,
                    nsig_stream_t stream);

/* EMA + mean of the partial update (renderer_wtmk.py:521-524) over all n_cells = C*H^3 cells:
 * density_grid = max(density_grid * decay, tmp_grid) where both >= 0; *sum += sum of max(density_grid, 0). */
int nsig_grid_finalize(float* density_grid, const float* tmp_grid, uint32_t n_cells, float decay,
                       double* sum, nsig_stream_t stream);

/* packbits (raymarching.cu:268-289) with the threshold of renderer_wtmk.py:524-530 computed on the device:
 * mean_density = *sum / n_cells, thresh = min(mean_density, density_thresh).  stats (optional, device
 * float[2]) receives (mean_density, thresh).  n_bytes = n_cells / 8. */
int nsig_grid_pack(const float* density_grid, uint32_t n_bytes, const double* sum, uint32_t n_cells,
                   float density_thresh, uint8_t* bitfield, float* stats, nsig_stream_t stream);

/* Cell selection of the partial update (renderer_wtmk.py:489-501) without torch.nonzero's host sync:
 * cells[c, 0:n_uniform) = morton3D(randint(0, H, 3)); cells[c, n_uniform:) = uniformly drawn members of
 * {i : density_grid[c, i] > 0} (ordered compaction, then a random pick).  cells: int32 [C, n_uniform+n_occupied]. */
size_t nsig_grid_sample_cells_scratch_bytes(uint32_t C, uint32_t H);
int nsig_grid_sample_cells(const float* density_grid, uint32_t C, uint32_t H, uint32_t n_uniform,
                           uint32_t n_occupied, uint64_t seed, int32_t* cells, void* scratch,
                           nsig_stream_t stream);

/* NeRFRenderer.mark_untrained_grid (renderer_wtmk.py:380-442): density_grid[c, i] = -1 for every cell whose
 * centre no camera sees (z > 0, |x| < cx/fx * z + 2*half, |y| < cy/fy * z + 2*half).  poses: [B,4,4] cam2world. */
int nsig_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy,
                             uint32_t C, uint32_t H, double bound, float* density_grid,
                             nsig_stream_t stream);

/* get_rays (utils_wtmk_disen.py:59-143): rays_o/rays_d [B,N,3] of pixels inds[b*inds_batch_stride + n]
 * (int64 pixel ids h*W + w; inds_batch_stride = 0 shares one index list over the batch, as the reference's
 * expand does; inds == NULL means all H*W pixels, N = H*W).  poses: [B,4,4] cam2world. */
int nsig_get_rays(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H,
                  uint32_t W, const int64_t* inds, int64_t inds_batch_stride, uint32_t N,
                  float* rays_o, float* rays_d, nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* HiDDeN message decoder — nerf/hidden_models.py:16-35, 104-137 */
/* ------------------------------------------------------------------------- */

/* HiddenDecoder_multi_views(num_blocks, num_bits, input_ch=3, channels=64, redundancy) applied to
 * normalize_img(image): forward and backward as a chain of tensor-core kernels (csrc/decoder.cu) with the
 * rounding points of torch.autocast(float16).  image: [B,H,W,3] fp32 (the rendered blocks, already clamped to
 * [0,1]; the ImageNet normalisation of hidden_models.normalize_img is applied inside).
 * params: HOST array of 4*(num_blocks+1)+2 device pointers (fp32): per ConvBNRelu block conv.weight [cout,cin,3,3],
 * conv.bias, bn.weight, bn.bias (the last block has num_bits*redundancy <= 8 outputs), then linear.weight,
 * linear.bias.  workspace: nsig_decoder_workspace_bytes() bytes, kept between forward and backward.
 * forward writes logits [B, num_bits] fp32 (HiddenDecoder_multi_views.forward's return value).
 * backward takes dlogits [B, num_bits] fp32, ACCUMULATES every parameter gradient into grads (same order and
 * shapes as params, fp32) and writes dimage [B,H,W,3] fp32 (optional).  Weight gradients are reduced in a fixed order
 * (per-CTA partials in the workspace + a ticket per tap): results are run-to-run deterministic.
 * prepared_weights (optional, may be NULL): nsig_decoder_weights_bytes() bytes filled by nsig_decoder_prepare_weights
 * from the SAME params - the fp16 copies of the conv weights the kernels consume.  With NULL the forward converts the
 * weights itself (one more kernel on the critical path of every call); a caller that knows when the weights change
 * (the optimizer) converts them once per update instead.  backward must get what forward got. */
size_t nsig_decoder_workspace_bytes(uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks);
size_t nsig_decoder_weights_bytes(uint32_t num_blocks);
int nsig_decoder_prepare_weights(const float* const* params, uint32_t num_blocks, uint32_t num_bits,
                                 uint32_t redundancy, void* weights, nsig_stream_t stream);

/* Deferred tail of nsig_decoder_backward.  After nsig_decoder_defer_weight_grads(1), a backward returns as soon as the INPUT
 * gradient (dimage) and the BatchNorm gradients are enqueued on `stream`; the conv weight / bias gradients are still being
 * computed on the library's side streams and `grads` / `workspace` must stay untouched and alive until
 * nsig_decoder_finish_backward(stream) has been called: it makes `stream` wait for them and enqueues the deterministic
 * reduction into `grads`.  A caller whose next kernels only need dimage (the renderer's backward,
 * utils_wtmk_disen.py:1175) overlaps them with the decoder's weight gradients this way.  No-op when nothing is pending; a new
 * backward finishes a forgotten tail first.  Process-wide switch, one pending backward per device. */
int nsig_decoder_defer_weight_grads(int on);
int nsig_decoder_finish_backward(nsig_stream_t stream);

/* Parity probe of the decoder kernels' activation arithmetic: gelu[i] = fp16(GELU(y[i])), gelu_grad[i] = fp16(GELU'(y[i]))
 * for n fp16 values, through the very device functions the conv kernels apply while staging their tiles (nn.GELU(), exact /
 * erf form: hidden_models.py:26).  Lets a test sweep all 2^16 fp16 inputs against torch and against float64. */
int nsig_decoder_gelu_probe(const void* y, uint32_t n, void* gelu, void* gelu_grad, nsig_stream_t stream);
int nsig_decoder_forward(const float* image, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks,
                         uint32_t num_bits, uint32_t redundancy, const float* const* params, void* workspace,
                         float* logits, const void* prepared_weights, nsig_stream_t stream);
int nsig_decoder_backward(const float* dlogits, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks,
                          uint32_t num_bits, uint32_t redundancy, const float* const* params,
                          float* const* grads, void* workspace, float* dimage, const void* prepared_weights,
                          nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* loss head of the watermark training step — nerf/utils_wtmk_disen.py:592-593, 636-644 */
/* ------------------------------------------------------------------------- */

/* image: the rendered [block rays | content rays] pixels as n_total floats, the first n_block of which belong to the
 * watermark blocks.  forward: pred = clamp(image[:n_block], 0, 1) (utils_wtmk_disen.py:593), content = image[n_block:].
 * backward: grad_image = [grad_pred where 0 <= image <= 1 else 0 | grad_content]; a NULL gradient reads as zeros. */
int nsig_split_clamp_forward(const float* image, uint32_t n_block, uint32_t n_total, float* pred, float* content,
                             nsig_stream_t stream);
int nsig_split_clamp_backward(const float* image, const float* grad_pred, const float* grad_content,
                              uint32_t n_block, uint32_t n_total, float* grad_image, nsig_stream_t stream);

/* lossi = mean((image - gt)^2) over n floats (utils_wtmk_disen.py:636), lossw = mean BCE-with-logits of
 * logits[md] * temp against message[md] (utils_wtmk_disen.py:441,641), loss = lambda_w*lossw + lambda_i*lossi (:644).
 * forward writes out[3] = (loss, lossi, lossw) and the gradients of `loss`: g_image[n], g_logits[md];
 * backward scales them by the device scalar *grad_out (the incoming gradient of `loss`, e.g. the GradScaler's
 * scale) into d_image / d_logits.  One CTA, no atomics: deterministic. */
int nsig_wtmk_loss_forward(const float* image, const float* gt, uint32_t n, const float* logits,
                           const float* message, uint32_t md, float lambda_w, float lambda_i, float temp,
                           float* out, float* g_image, float* g_logits, nsig_stream_t stream);
int nsig_wtmk_loss_backward(const float* g_image, const float* g_logits, uint32_t n, uint32_t md,
                            const float* grad_out, float* d_image, float* d_logits, nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* gradient exchange of the ray-sharded training path (no reference precedent) */
/* ------------------------------------------------------------------------- */

/* In-place all-reduce (MEAN) of a float bucket that lives in peer-mapped ("symmetric") memory on every rank of
 * one NVLink/NVSwitch node, as one kernel: entry barrier, two-shot reduce (rank r reduces and broadcasts slice r),
 * exit barrier.  bufs / flags: HOST arrays of `world` device pointers - every rank's bucket and flag buffer as
 * mapped into THIS process (bufs[rank] is the local bucket).  flags[r]: uint32[nsig_allreduce_grid() * world],
 * zero-initialised once; the barriers reset them.  multicast: NVSwitch multicast address of the bucket (the
 * reduction then happens in the switch: multimem.ld_reduce / multimem.st) or NULL (plain P2P loads and stores).
 * n: bucket length in floats, a multiple of 4.  Every rank must call it with the same n and the same launch
 * order; stream-ordered, CUDA-graph capturable. */
uint32_t nsig_allreduce_grid(void);
int nsig_allreduce_mean_inplace(void* const* bufs, void* const* flags, void* multicast, uint32_t n,
                                uint32_t rank, uint32_t world, nsig_stream_t stream);

/* Watermark-mode backward from what the forward already produced (the default hot path): `masks` as written by
 * nsig_field_forward(masks_out), the forward outputs sigmas / rgbs (sigmoid'(z) = rgb (1 - rgb), exp(logit) = sigma /
 * density_scale) and the incoming gradients.  No MLP recomputation: only the five data-gradient GEMMs run, then dL/dS is
 * scatter-added into G ([2^log2_T, 2] fp32, accumulated).  Same result as nsig_field_backward(G only) up to 1 ulp of the
 * exp() factor. */
int nsig_field_backward_masks(const float* xyzs, uint32_t M, float bound, const void* masks, const float* sigmas,
                              const float* rgbs, const float* grad_sigmas, const float* grad_rgbs,
                              const void* sigma_w, const void* color_w, float density_scale,
                              const int32_t* M_dev, float msg_resolution, uint32_t log2_T, float* G,
                              nsig_stream_t stream);

/* The watermark-mode case of nsig_field_backward (G only: no full feature gradient, no weight gradients - SURVEY F13:
 * everything but the message tables is frozen) on the 5th-generation tensor cores: every MLP layer of a 128-sample
 * tile is one tcgen05.mma chain with the accumulator in TMEM (csrc/field_tc.cu).  Same arguments, same result up to
 * fp32 summation order; `feat` must be 16-byte aligned. */
int nsig_field_backward_tc(const float* xyzs, const float* dirs, uint32_t M, float bound, const void* feat,
                           const float* grad_sigmas, const float* grad_rgbs, const void* sigma_w,
                           const void* color_w, float density_scale, const int32_t* M_dev,
                           float msg_resolution, uint32_t log2_T, float* G, nsig_stream_t stream);

/* nsig_field_backward_masks on tcgen05 + TMEM: five UMMA layers per 128-row tile, one thread per sample row, six
 * CTAs per SM (csrc/field_tc.cu).  Same arguments and result; `masks` must be 16-byte aligned. */
int nsig_field_backward_tc_masks(const float* xyzs, uint32_t M, float bound, const void* masks,
                                 const float* sigmas, const float* rgbs, const float* grad_sigmas,
                                 const float* grad_rgbs, const void* sigma_w, const void* color_w,
                                 float density_scale, const int32_t* M_dev, float msg_resolution,
                                 uint32_t log2_T, float* G, nsig_stream_t stream);

/* ------------------------------------------------------------------------- */
/* optimizer step of the message tables — nerf/utils_wtmk_disen.py:1175-1181   */
/* ------------------------------------------------------------------------- */

/* torch.optim.Adam (betas, eps, no weight decay, no amsgrad) over the parameters
 * network_wtmk_tcnn.py:179-188 hands it, restricted to the message tables: the message_dim tables
 * embeddings[2i + bit_i] selected by `message` (device float 0/1) are updated with the one gradient
 * G = dL/dS ([2^log2_T,2] fp32) they all share; unselected tables are untouched, exactly
 * like parameters whose .grad is None.
 *   ptr_table: DEVICE array uint64[3][n_tables] = addresses of (param, exp_avg, exp_avg_sq) per table
 *   steps    : device float[n_tables], per-table step counts (incremented for the selected tables)
 *   coef     : device float[n_tables][2] scratch (bias-correction scalars)
 *   grad_scale / found_inf: optional device scalars with torch.amp.GradScaler's meaning (G is divided
 *   by *grad_scale; the whole step is skipped when *found_inf != 0).
 *   lr_dev   : optional device float; when non-NULL the learning rate is read from it at execution
 *   time instead of `lr`, so a per-step scheduler (the reference's LambdaLR, scheduler_update_every_step)
 *   keeps acting on a step that is replayed from a CUDA graph.
 *   elem_begin / elem_count (floats, multiples of 4; count 0 = everything): the slice of every selected table (and of
 *   its moments) to update - ZeRO-style sharding of the optimizer over the ranks of a data-parallel job.
 *   steps_prepared: 1 when nsig_msg_adam_lookahead_sum was already called for THIS update (it advanced steps / coef);
 *   0 otherwise. */
int nsig_msg_adam_step(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                       const float* message, const float* G, float* steps, float* coef,
                       const float* grad_scale, const float* found_inf, float lr, float beta1,
                       float beta2, float eps, uint32_t log2_T, const float* lr_dev,
                       uint32_t elem_begin, uint32_t elem_count, uint32_t steps_prepared, nsig_stream_t stream);

/* Look-ahead table sum for a software-pipelined optimizer: S [2^log2_T,2] = sum_i table[2i + next_i] with the values the
 * tables WILL hold after the pending nsig_msg_adam_step(message = message_applied, same G / scalars / range) - computed
 * from (param, exp_avg, exp_avg_sq, G) with the update's own arithmetic, nothing but S (and steps / coef) written.
 * Bit-identical to running the update and then nsig_msg_table_sum(message_next).  The next step's field forward then
 * depends on ~0.26 GB of reads instead of the update's ~0.8 GB of traffic, and the update itself (call nsig_msg_adam_step
 * with steps_prepared = 1) may run any time before G is written again.  Replaces, for that schedule, the call pair
 * hash_encoding_wtmk_bit.py:99-116 (per-bit table selection) after optimizer.step() (utils_wtmk_disen.py:1178). */
int nsig_msg_adam_lookahead_sum(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                                const float* message_applied, const float* message_next, const float* G,
                                float* steps, float* coef, const float* grad_scale, const float* found_inf,
                                float lr, float beta1, float beta2, float eps, uint32_t log2_T,
                                const float* lr_dev, uint32_t elem_begin, uint32_t elem_count, float* S,
                                nsig_stream_t stream);

/* nsig_msg_adam_step(message = message_applied) AND nsig_msg_table_sum(message_next) in one pass over the tables: S is
 * accumulated while the update streams - from the freshly updated value where message_next selects the table being updated,
 * from one extra read of the sibling table otherwise - in nsig_msg_table_sum's order, so S is bit-identical to the call
 * pair.  A skipped update (*found_inf != 0) leaves the tables alone and sums them as they are.  Same arguments as
 * nsig_msg_adam_lookahead_sum; with a slice (elem_begin / elem_count) only that slice of S is written.  Replaces
 * optimizer.step() (utils_wtmk_disen.py:1178) + the next step's per-bit table selection
 * (hash_encoding_wtmk_bit.py:99-116) when the caller knows the next message before it applies the update. */
int nsig_msg_adam_step_sum(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                           const float* message_applied, const float* message_next, const float* G,
                           float* steps, float* coef, const float* grad_scale, const float* found_inf,
                           float lr, float beta1, float beta2, float eps, uint32_t log2_T,
                           const float* lr_dev, uint32_t elem_begin, uint32_t elem_count, float* S,
                           nsig_stream_t stream);

/* torch.amp.GradScaler's per-step work (utils_wtmk_disen.py:1175-1181) over the flat gradient bucket in one launch:
 * non-finite check of flat[0..n) -> *found_inf (0/1); *step_scale = the scale this step's gradients carry (what the
 * optimizer kernels divide by); then _amp_update_scale_ on *scale / *growth_tracker (backoff on overflow, growth after
 * growth_interval clean steps); *adam_step (optional) += 1 when the step is not skipped.  scratch: uint32[2], zero-
 * initialised once by the caller, left zeroed.  flat must be 16-byte aligned.
 * enabled (optional): device word; when it is 0 no optimizer step is pending (the harness' deferred-step mode after a
 * flush): *found_inf = 1 so that the optimizer kernels skip, scaler state untouched. */
int nsig_grad_check_update_scale(const float* flat, uint32_t n, float* scale, int32_t* growth_tracker,
                                 float growth_factor, float backoff_factor, int32_t growth_interval,
                                 float* found_inf, float* step_scale, float* adam_step, uint32_t* scratch,
                                 const uint32_t* enabled, nsig_stream_t stream);

/* torch.optim.Adam (no weight decay / amsgrad) over one flat fp32 parameter vector (the HiDDeN decoder's tensors viewed
 * as one buffer).  step: device float holding this step's count (>= 1); grad_scale / found_inf / lr_dev as in
 * nsig_msg_adam_step. */
int nsig_flat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n,
                        const float* step, const float* grad_scale, const float* found_inf, float lr,
                        const float* lr_dev, float beta1, float beta2, float eps, nsig_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NSIG_H_ */
