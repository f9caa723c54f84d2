/*
 * spv_b200.h -- C ABI of the B200-native (sm_100a) differentiable Gaussian rasterizer.
 *
 * This is the drop-in boundary ("B1"): one extern "C" entry point per
 * native function the reference binds through pybind11 in
 *   /root/reference/src/submodules/dptr/dptr/gs/src/ext.cpp:14-32
 * (18 m.def's taking torch::Tensor), restated with plain device pointers, sizes and a CUDA stream.
 * No torch types cross this boundary.  The Python package splatter_a_video_b200.gs (a mirror of
 * the reference's `dptr.gs` module) allocates every tensor with torch and passes data_ptr()s and
 * torch.cuda.current_stream().cuda_stream.
 *
 * Conventions
 *  - All pointers are DEVICE pointers (fp32 / int32 / uint8 as typed), dense row-major.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - Functions never allocate and never synchronise the host; they return 0 on success or a
 *    cudaError_t value (launch/config errors are checked after every launch, unlike the reference,
 *    which checks none -- include/utils.h:9-10).  spv_last_error() gives a readable message.
 *  - Outputs are fully written by the callee (the reference relies on torch::zeros; here the
 *    callee clears what it must, so callers may pass torch.empty buffers).
 *  - Tiles are 16x16 pixels, tile id = ty * ceil(W/16) + tx (include/config.h:7-10).
 */
#ifndef SPV_B200_H
#define SPV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPV_ABI_VERSION 1
#define SPV_TILE 16

/* the library is built with -fvisibility=hidden; only these entry points are exported */
#if defined(__GNUC__)
#define SPV_API __attribute__((visibility("default")))
#else
#define SPV_API
#endif

SPV_API int spv_abi_version(void);
SPV_API const char *spv_last_error(void);
/* number of CUDA kernels this library has launched since load (bench.py reports it as `gpu_launches`) */
SPV_API long long spv_launch_count(void);
/* Measurement hook (no reference counterpart): when enabled, every blend kernel launch is bracketed by a pair of CUDA events
 * on its own stream (slot 0 = forward kernel, 1 = backward kernel; also inside a captured graph).  `read` waits for the
 * slot's closing event and returns the device time of the most recent launch.  Disabled (default): no cost. */
/* Runtime switch of a kernel variant (0 = the default).  "bwd_variant" = 1: the chunk-barrier backward blend kernel of round 1
 * instead of the ring-staged one (also read once from the environment variable SPV_BWD_VARIANT). */
SPV_API int spv_set_option(const char *name, int value);
SPV_API int spv_kernel_timer_enable(int on);
SPV_API int spv_kernel_timer_read(int slot, float *ms);

/* ---- K1/K2: project_point_forward/backward (ext.cpp:15-16; src/project_point.cu:13-145) ---------- */
/* intr = [fx,fy,cx,cy]; extr = 12 floats, row-major 3x4 [R|t] (a 4x4 matrix's first 12 floats work). */
SPV_API int spv_project_point_forward(int P, const float *xyz, const float *intr, const float *extr,
                              int W, int H, float nearest, float extent,
                              float *uv /*[P,2]*/, float *depth /*[P,1]*/, void *stream);
/* dL_dintr[4] / dL_dextr[12] may be NULL (the reference passes nullptr unless requires_grad). */
SPV_API int spv_project_point_backward(int P, const float *xyz, const float *intr, const float *extr,
                               const float *depth, const float *dL_duv, const float *dL_ddepth,
                               float *dL_dxyz /*[P,3]*/, float *dL_dintr, float *dL_dextr, void *stream);

/* Orthographic projection of the video trainer (pointrix/renderer/dptr_ortho_enhanced.py:145-202;
 * torch ops in the reference, one kernel here).  backward: d(uv,depth)/d(xyz) through the 3x4 extr. */
SPV_API int spv_project_point_ortho_forward(int P, const float *xyz, const float *extr, int W, int H,
                                    float nearest, float extent, float *uv, float *depth, void *stream);
SPV_API int spv_project_point_ortho_backward(int P, const float *extr, int W, int H, const float *depth,
                                     const float *dL_duv, const float *dL_ddepth, float *dL_dxyz, void *stream);

/* ---- K3/K4: compute_cov3d_forward/backward (ext.cpp:17-18; src/compute_cov3d.cu:14-147) ----------- */
SPV_API int spv_compute_cov3d_forward(int P, const float *scales, const float *uquats, const uint8_t *visible,
                              float *cov3d /*[P,6]*/, void *stream);
SPV_API int spv_compute_cov3d_backward(int P, const float *scales, const float *uquats, const uint8_t *visible,
                               const float *dL_dcov3d, float *dL_dscales, float *dL_duquats, void *stream);

/* ---- K5/K6: ewa_project_forward/backward (ext.cpp:19-20; src/ewa_project.cu:16-252) --------------- */
SPV_API int spv_ewa_project_forward(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                            const float *uv, int W, int H, const uint8_t *visible,
                            float *conic /*[P,3]*/, int *radius /*[P]*/, int *tiles /*[P]*/, void *stream);
SPV_API int spv_ewa_project_backward(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                             const int *radius, const float *dL_dconic,
                             float *dL_dxyz, float *dL_dcov3d, float *dL_dintr /*nullable*/,
                             float *dL_dextr /*nullable*/, void *stream);
/* Orthographic EWA of the video trainer (dptr_ortho_enhanced.py:18-111), J = diag(W/2,H/2) rows. */
SPV_API int spv_ewa_project_ortho_forward(int P, const float *cov3d, const float *extr, const float *uv, int W, int H,
                                  const uint8_t *visible, float *conic, int *radius, int *tiles, void *stream);
SPV_API int spv_ewa_project_ortho_backward(int P, const float *cov3d, const float *extr, int W, int H, const int *radius,
                                   const float *dL_dconic, float *dL_dcov3d, void *stream);

/* ---- K7-K10: compute_sh(_free)_forward/backward (ext.cpp:23-24,29-30; src/compute_sh*.cu) --------- */
/* shs is read with a per-point stride of (deg+1)^2 * 3 floats exactly like the reference
 * (compute_sh.cu:45); S_alloc = shs.size(1) only sizes the zero-filled dL_dshs[P,S_alloc,3].
 * visible == NULL means "every point visible" (what the renderers pass, dptr_ortho_enhanced.py:272).
 * dL_ddirs == NULL (backward): the direction gradient is not wanted -- the coefficients are then not read at all.   */
SPV_API int spv_compute_sh_forward(int P, const float *shs, int deg, const float *dirs, const uint8_t *visible,
                           int free_variant, float *colors /*[P,3]*/, uint8_t *clamped /*[P,3], NULL if free*/,
                           void *stream);
SPV_API int spv_compute_sh_backward(int P, const float *shs, int deg, const float *dirs, const uint8_t *visible,
                            const uint8_t *clamped /*NULL if free*/, const float *dL_dcolors, int S_alloc,
                            float *dL_dshs, float *dL_ddirs, void *stream);
/* SH colour along the constant direction (0,0,1) from the coefficients of the bases 0, 2, 6, 12 alone (shs_z = [P,4,3]): what
 * DPTROrthoEnhancedRender evaluates (dptr_ortho_enhanced.py:270-271), bit-identical to spv_compute_sh_forward / _backward with
 * deg = 3 and dirs = (0,0,1) on the full tensor, at 48 instead of 192 bytes per Gaussian. */
SPV_API int spv_compute_sh_z_forward(int P, const float *shs_z, float *colors, uint8_t *clamped /*[P,3] or NULL*/, void *stream);
SPV_API int spv_compute_sh_z_backward(int P, const uint8_t *clamped /*or NULL*/, const float *dL_dcolors, float *dL_dshs_z, void *stream);

/* ---- K11-K14: sort_gaussian (ext.cpp:21-22; gs/sort_gaussian.py:41-54; src/sort_gaussian.cu) ------ */
/* Step 1: inclusive scan of tiles-touched (torch.cumsum in the reference).  offsets[P-1] = I.        */
SPV_API size_t spv_sort_scan_workspace_bytes(int P);
SPV_API int spv_sort_scan(int P, const int *tiles, int *offsets /*[P]*/, void *workspace, size_t ws_bytes, void *stream);
/* Step 2: emit (tile|depth) keys, radix-sort the significant bits, gather ids, tile ranges.          */
SPV_API size_t spv_sort_workspace_bytes(int P, int64_t I);
SPV_API int spv_sort_gaussian(int P, int64_t I, const float *uv, const float *depth, const int *radius,
                      const int *offsets, int W, int H, int *idx_sorted /*[I]*/,
                      int *tile_range /*[ceil(W/16)*ceil(H/16),2]*/, void *workspace, size_t ws_bytes, void *stream);

/* ---- K15-K20: alpha_blending{,_enhanced,_with_bias}_forward/backward (ext.cpp:25-28,31-32) -------- */
/* feature is [P,C] row-major (the reference transposes to [C,P] internally, alpha_blending.cu:282).
 * gs_idx ([H,W,K], -1 padded; K=0/NULL for the plain variant), opacity_bias (NULL unless with_bias). */
SPV_API int spv_alpha_blend_forward(int P, int C, int W, int H, int K, int enable_truncation,
                            const float *uv, const float *conic, const float *opacity, const float *feature,
                            const float *opacity_bias, const int *idx_sorted, const int *tile_range, float bg,
                            float *rendered /*[C,H,W]*/, float *final_T /*[H,W]*/, int *ncontrib /*[H,W]*/,
                            int *gs_idx, void *stream);
SPV_API size_t spv_alpha_blend_backward_workspace_bytes(int P, int C);
SPV_API int spv_alpha_blend_backward(int P, int C, int W, int H,
                             const float *uv, const float *conic, const float *opacity, const float *feature,
                             const float *opacity_bias, const int *idx_sorted, const int *tile_range, float bg,
                             const float *final_T, const int *ncontrib, const float *dL_drendered /*[C,H,W]*/,
                             float *dL_duv /*[P,2]*/, float *dL_dabs_uv /*[P,2]*/, float *dL_dconic /*[P,3]*/,
                             float *dL_dopacity /*[P,1]*/, float *dL_dfeature /*[P,C]*/,
                             float *dL_dopacity_bias /*[P,1] or NULL*/,
                             void *workspace, size_t ws_bytes, void *stream);

/* ---- Fused single-traversal blending of the trainer's three passes (dptr_ortho_enhanced.py:342-376) ----
 * feature = [rgb(3) | depth(1) | attributes(C-4)] row-major [P,C]; per-group backgrounds.  Results equal the three
 * separate reference calls: uv/conic gradients from all channels, opacity from rgb+depth only (the attribute pass
 * receives opacity.detach()), dL_duv_rgb / dL_dabs_uv_rgb from the RGB pass only (they feed ndc.grad / abs_ndc.grad;
 * the other passes receive ndc.detach()).  Extension of the ABI: no single reference entry point corresponds. */
SPV_API int spv_alpha_blend_groups_forward(int P, int C, int W, int H, int K, const float *uv, const float *conic,
                                   const float *opacity, const float *feature, const int *idx_sorted,
                                   const int *tile_range, float bg_rgb, float bg_depth, float bg_attr,
                                   float *rendered /*[C,H,W]*/, float *final_T, int *ncontrib, int *gs_idx /*[H,W,K]*/,
                                   void *stream);
SPV_API size_t spv_alpha_blend_groups_backward_workspace_bytes(int P);
SPV_API int spv_alpha_blend_groups_backward(int P, int C /*4..23*/, int W, int H, const float *uv, const float *conic,
                                    const float *opacity, const float *feature, const int *idx_sorted,
                                    const int *tile_range, float bg_rgb, float bg_depth, float bg_attr,
                                    const float *final_T, const int *ncontrib, const float *dL_drendered,
                                    float *dL_duv, float *dL_duv_rgb, float *dL_dabs_uv_rgb, float *dL_dconic,
                                    float *dL_dopacity, float *dL_dfeature, void *workspace, size_t ws_bytes,
                                    void *stream);

/* ---- Fused per-frame entry points (no host synchronisation; CUDA-graph capturable) -------------------------------
 * spv_bin_capacity: scan + (optionally exactly culled) key emission + radix sort + ranges for at most I_cap
 * intersections; status[0] = intersections kept, status[1] = 1 on overflow (device ints).
 * spv_frame_ortho_forward/backward: DPTROrthoEnhancedRender.render_iter (dptr_ortho_enhanced.py:205-383) and its full
 * backward as one call each.  images = [rgb(3) | depth(1) | attrs(A)] x H x W; A <= 19.  The workspace filled by the
 * forward call must be passed unchanged to the backward call. */
SPV_API size_t spv_bin_capacity_workspace_bytes(int P, int64_t I_cap);
SPV_API int spv_bin_capacity(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                     const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status,
                     void *workspace, size_t ws_bytes, void *stream);
/* spv_bin_tiles: same contract and bit-identical results, without the global radix sort: per-tile histogram in the count
 * pass, one-CTA scan over the tiles, scatter of (depth bits << 32 | id) keys into per-tile segments, per-tile bitonic sort in
 * shared memory (segments above 1024 / 25600 keys: 1024-thread CTAs with 200 KB dynamic shared memory / in place in global memory). */
SPV_API size_t spv_bin_tiles_workspace_bytes(int P, int64_t I_cap, int W, int H);
SPV_API int spv_bin_tiles(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                  const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status,
                  void *workspace, size_t ws_bytes, void *stream);
SPV_API size_t spv_frame_workspace_bytes(int P, int64_t I_cap, int W, int H, int A);
SPV_API int spv_frame_ortho_forward(int P, int W, int H, int n_groups /*<= 8*/, const float *const *attr_ptrs /*host array of device ptrs*/,
                            const int *attr_channels /*host*/, int K, int64_t I_cap, int cull, const float *position,
                            const float *scaling, const float *rotation, const float *opacity,
                            const float *shs /*[P,16,3], or [P,4,3] = the bases 0, 2, 6, 12 alone (sh_bases = 4): the renderer's
                                               constant view direction (0,0,1) reaches no other basis*/,
                            int sh_bases /*16 or 4*/, const float *extr, float nearest, float extent, float bg_rgb, float *images /*[4+A,H,W]*/,
                            int *gs_idx /*[H,W,K]*/, int *radii /*[P]*/, int *status /*[2]*/, void *workspace,
                            size_t ws_bytes, void *stream);
/* dL_dimage_planes: host array of 4+A device pointers, one [H,W] gradient plane per image channel (NULL = no gradient);
 * dL_dattr_ptrs: host array of n_groups device pointers receiving each attribute group's gradient (NULL = not needed). */
SPV_API int spv_frame_ortho_backward(int P, int W, int H, int n_groups, const int *attr_channels,
                             int n_grad_channels /* leading image channels (rgb, depth, then attribute groups) whose feature
                                                    gradient is wanted; gradient-free groups must come last */,
                             int64_t I_cap,
                             const float *scaling, const float *rotation, const float *opacity, const float *shs,
                             int sh_bases /*as in the forward call*/,
                             const float *extr, float bg_rgb, const float *const *dL_dimage_planes, float *dL_dposition,
                             float *dL_dscaling, float *dL_drotation, float *dL_dopacity, float *dL_dshs /*[P,sh_bases,3]*/,
                             float *const *dL_dattr_ptrs, float *dL_dndc /*[P,2] or NULL*/, float *dL_dabs_ndc /*[P,2] or NULL*/,
                             float *dL_drgb_out /*[P,3] or NULL.  Non-NULL defers the SH backward (frame-parallel training):
                                                  the gradient of the SH-evaluated colours is written here, dL_dshs is not
                                                  touched, and spv_compute_sh_backward is run by the caller on the
                                                  all-reduced colour gradient*/,
                             uint8_t *clamped_out /*[P,3] or NULL: the forward's clamp mask for that deferred call*/,
                             int first_backward /*1: first backward over this workspace (the forward call left the packed
                                                  gradient rows cleared, off the critical path); 0: clear them again*/,
                             void *workspace, size_t ws_bytes, void *stream);
/* Blend stage of the grouped backward with per-channel gradient planes; leaves 36-float packed rows in `packed`. */
SPV_API int spv_alpha_blend_groups_backward_packed(int P, int C, int W, int H, const float *uv, const float *conic,
                                           const float *opacity, const float *feature, const int *idx_sorted,
                                           const int *tile_range, float bg_rgb, float bg_depth, float bg_attr,
                                           const float *final_T, const int *ncontrib, const float *const *planes_host,
                                           int n_grad_channels, float *packed, void *stream);

/* ---- Per-frame deformation (next row f-1): cubic-spline position of the active model
 * (src/dynamic_gaussian_with_base_point_cloud.py:236-250).  coeff = pos_cubic_node viewed as [P,4,NI,3] (layout 0, the
 * reference's parameter layout) or stored interval-major as [P,NI,4,3] (layout 1: the 4 coefficients of one interval are 48
 * contiguous bytes -- 2 sectors instead of 4 scattered ones per evaluation; convert with spline layout helpers of gs.frame); the
 * interval index and in-interval distance are DEVICE scalars (graph-replayable).  Backward: gradient to the coefficients, same
 * layout. */
SPV_API int spv_deform_spline_forward(int P, int NI, int layout, const float *base, const float *coeff, const int *idx_dev,
                              const float *dist_dev, float *pos /*[P,3]*/, void *stream);
SPV_API int spv_deform_spline_backward(int P, int NI, int layout, const int *idx_dev, const float *dist_dev, const float *dL_dpos,
                               float *dL_dcoeff /*[P,4,NI,3]*/, int accumulate, void *stream);

/* Both frame times of a training step (ids1 rendered, ids2 as the `track_gs` attribute, trainer_fragGS.py:486-508) in one
 * pass.  backward2 writes into a gradient sink kept clean incrementally: `dirty` = int[17] on the device ([0] = count,
 * then interval indices holding non-zero gradient from earlier calls or gradient exchanges; zero-initialise it together
 * with the sink); listed intervals are zeroed unless re-written, afterwards the list is {idx1, idx2}. */
SPV_API int spv_deform_spline_forward2(int P, int NI, int layout, const float *base, const float *coeff, const int *idx1_dev,
                               const float *dist1_dev, const int *idx2_dev, const float *dist2_dev, float *pos1,
                               float *pos2, void *stream);
SPV_API int spv_deform_spline_backward2(int P, int NI, int layout, const int *idx1_dev, const float *dist1_dev, const int *idx2_dev,
                                const float *dist2_dev, const float *dL_dpos1, const float *dL_dpos2 /*or NULL*/,
                                int *dirty, float *dL_dcoeff /*[P,4,NI,3] sink*/, void *stream);

/* Deferred spline backward for frame-parallel training: the coefficient gradient is linear in dL_dpos, so ranks exchange the
 * 2 x 3 position-gradient floats per Gaussian (all-gather) instead of 2 x 12 coefficient-gradient floats and every rank
 * rebuilds all ranks' interval gradients locally, summed per interval in (rank, slot) order (bit-identical everywhere).
 * defer: payload = [dL_dpos1 (3P) | dL_dpos2 (3P, zeros if NULL) | idx1 bits, dist1, idx2 bits, dist2] (6P + 4 floats).
 * gathered: `world` payloads, `stride` floats apart.  `dirty` as in spv_deform_spline_backward2. */
SPV_API int spv_deform_defer(int P, const float *dL_dpos1, const float *dL_dpos2 /*or NULL*/, const int *idx1_dev,
                     const float *dist1_dev, const int *idx2_dev, const float *dist2_dev, float *payload, void *stream);
SPV_API int spv_deform_spline_backward_gathered(int P, int NI, int layout, int world, const float *gathered, long long stride,
                                        float scale /*applied to every dL_dpos, e.g. 1/world*/, int *dirty,
                                        float *dL_dcoeff /*[P,4,NI,3] sink*/, void *stream);

/* Rotation at frame time t (get_rotation, :184-198): normalize(rotation + detached poly/Fourier offsets); basis_dev holds
 * [t^0..t^3 | cos(t*pi*(1..4)) | sin(t*pi*(1..4))] on the device.  Backward: through the normalisation to `rotation`. */
SPV_API int spv_deform_rotation_forward(int P, const float *rotation, const float *rot_poly_feat /*[P,4,4]*/,
                                const float *rot_fourier_feat /*[P,8,4]*/, const float *basis_dev /*[12]*/,
                                float *out /*[P,4]*/, float *inv_norm /*[P]*/, void *stream);
SPV_API int spv_deform_rotation_backward(int P, const float *out, const float *inv_norm, const float *dL_dout,
                                 float *dL_drotation, void *stream);
/* Polynomial(4) + Fourier(8) POSITION of the alternative model (src/dynamic_gaussian_points.py:170-186, get_position):
 * pos = position + pos_poly_feat[P,4,3] . t^k + pos_fourier_feat[P,8,3] . {cos,sin}(t pi (1..4)); basis_dev as above.  The backward
 * WRITES its gradients; NULL outputs are skipped (dL_dposition under detach_pos). */
SPV_API int spv_deform_polyfourier_forward(int P, const float *position, const float *pos_poly_feat, const float *pos_fourier_feat,
                                   const float *basis_dev, float *pos, void *stream);
SPV_API int spv_deform_polyfourier_backward(int P, const float *basis_dev, const float *dL_dpos, float *dL_dposition,
                                    float *dL_dpos_poly_feat, float *dL_dpos_fourier_feat, void *stream);

/* ---- Fused Adam over the flat parameter buffer (next row f-3, optimizer half; torch.optim.Adam arithmetic) ---------- */
SPV_API int spv_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                  const long long *seg_end_host, const float *seg_lr_host, float beta1, float beta2, float eps, int step,
                  void *stream);
/* Same update with the optimizer clock and the learning rates in device memory (CUDA-graph capturable: every replay advances
 * the bias corrections).  state_dev = float[8] {step, 1-b1^step, sqrt(1-b2^step), -, b1^step and b2^step as two doubles},
 * zero-initialised once; lr_dev = float[nseg]. */
SPV_API int spv_adam_step_device(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                         const long long *seg_end_host, const float *lr_dev, float beta1, float beta2, float eps,
                         float *state_dev, void *stream);
/* Interval-lazy variant for ONE segment of spline coefficients (segment `lazy_seg` = [P, 4*NI*3], layout as in
 * spv_deform_spline_*): a step streams only the intervals that hold gradient (dirty_dev = the int[17] list of
 * spv_deform_spline_backward2 / _gathered) and every other interval's zero-gradient Adam updates are replayed -- same arithmetic,
 * same per-step constants (ring_dev = float[2*4096], zero-initialised) -- when the interval is next needed: spv_adam_lazy_prepare
 * (the two intervals a forward pass is about to read) or spv_adam_lazy_flush (all; before densification, checkpoints, rendering
 * other frames).  last_dev = int[NI], zero-initialised: the step each interval is current through.  Parameters are identical to
 * spv_adam_step_device's whenever they are observed through prepare / flush. */
SPV_API int spv_adam_step_lazy(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                       const long long *seg_end_host, const float *lr_dev, float beta1, float beta2, float eps, float *state_dev,
                       int lazy_seg, int P, int NI, int layout, const int *dirty_dev, int *last_dev, float *ring_dev, void *stream);
SPV_API int spv_adam_lazy_prepare(long long n, int nseg, const long long *seg_end_host, int lazy_seg, int P, int NI, int layout,
                          float *param, float *exp_avg, float *exp_avg_sq, const int *idx1_dev, const int *idx2_dev, int *last_dev,
                          const float *ring_dev, const float *state_dev, const float *lr_dev, float beta1, float beta2, float eps,
                          void *stream);
SPV_API int spv_adam_lazy_flush(long long n, int nseg, const long long *seg_end_host, int lazy_seg, int P, int NI, int layout, float *param,
                        float *exp_avg, float *exp_avg_sq, int *last_dev, const float *ring_dev, const float *state_dev,
                        const float *lr_dev, float beta1, float beta2, float eps, void *stream);

/* ---- Densification on the flat SoA (next row f-3, structure half) -------------------------------------------------
 * Role of AtlasGaussianSplattingOptimizer.update_structure / densification / prune / reset_opacity
 * (src/pointrix/optimizer/atlas_gs_optimizer.py:93-379) and PointCloud.extand_points / remove_points with their
 * optimizer-state surgery (src/pointrix/point_cloud/points.py:281-365).
 * stats: for visible points (visible == NULL: radii > 0) max_radii = max(., radii), grad_accum += |ndc_grad|, denom += 1.
 * flags: bit 0 clone, bit 1 split, bit 2 prune from grads = grad_accum / denom (NaN -> 0), max(scaling), opacity, max_radii;
 *        scaling / opacity are the STORED parameters (scaling_is_log / opacity_is_logit select exp / sigmoid activation);
 *        size_threshold <= 0 disables the two size tests like the reference's `if size_threshold:`.
 * flat_regather: every [P_old, width] block of `old_flat` moves to its [P_new, width] place in `new_flat`:
 *        new_row[j] = old_row[src[j]], zeros where src[j] < 0 (fresh Adam moments of new points).  One launch for all blocks.
 * split_children: for new rows with child_index[j] >= 0 (index into `samples`, the torch.normal draw of new_pos_scale):
 *        position = R(rotation[src]) @ sample + position[src]; scaling = inverse_activation(scaling[src] / div).
 * reset_opacity: opacity = inverse_sigmoid(min(sigmoid(opacity), cap)) and cleared moments (replace_optimizer). */
SPV_API int spv_densify_stats(int P, const float *ndc_grad /*[P,2]*/, const int *radii, const uint8_t *visible, float *grad_accum,
                      float *denom, float *max_radii, void *stream);
SPV_API int spv_densify_flags(int P, const float *grad_accum, const float *denom, const float *scaling /*[P,3]*/,
                      const float *opacity /*[P]*/, const float *max_radii, int scaling_is_log, int opacity_is_logit,
                      float grad_threshold, float dense_extent, float min_opacity, float size_threshold, float big_extent,
                      uint8_t *flags, void *stream);
SPV_API int spv_flat_regather(int n_blocks /*<= 32*/, const int *widths_host, const long long *old_offsets_host,
                      const long long *new_offsets_host, int P_new, const int *src /*[P_new] device*/, const float *old_flat,
                      float *new_flat, void *stream);
SPV_API int spv_split_children(int P_new, const int *src, const int *child_index, const float *samples /*[n_children,3]*/,
                       const float *old_position, const float *old_scaling, const float *old_rotation, int scaling_is_log,
                       float div, float *new_position, float *new_scaling, void *stream);
SPV_API int spv_reset_opacity(int P, float cap, int opacity_is_logit, float *opacity, float *exp_avg /*or NULL*/,
                      float *exp_avg_sq /*or NULL*/, void *stream);

/* ---- Gradient exchange packing for frame-parallel training (the reference has no gradient collective,
 * src/train.py:19-31,210-213) ----
 * The flat gradient buffer is described by a table of per-parameter segments; each parameter's per-Gaussian row is viewed as
 * [A, B, C] and slices are taken along B: mode 0 = dense (all of B, B <= 16), 1 = subset (sel[0..nsel) fixed, e.g. the 4 SH
 * bases that receive gradient under the constant view direction), 2 = sparse (at most one such segment; nsel <= 2 slice
 * indices read from device scalars, e.g. the spline intervals of this rank's two frame times).
 * pack: flat_grad -> comm_allreduce [n_allreduce floats] and comm_allgather [n_allgather floats = payload + 16 index words],
 * scaled.  The caller all-reduces the first and all-gathers the second (NCCL).  unpack: writes the reduced values back,
 * adds every rank's sparse slices at their intervals in rank order, and lists the touched intervals in `dirty`
 * (int[17], see spv_deform_spline_backward2; may be NULL). */
#define SPV_EXCHANGE_MAX_SEGMENTS 16
#define SPV_EXCHANGE_MAX_SELECT 16
typedef struct spv_exchange_segment {
    long long flat_offset;
    int A, B, C;
    int mode;
    int nsel;
    int sel[SPV_EXCHANGE_MAX_SELECT];
} spv_exchange_segment;
SPV_API int spv_exchange_sizes(int P, int nseg, const spv_exchange_segment *segs /*host*/, long long *n_allreduce,
                       long long *n_allgather);
SPV_API int spv_exchange_pack(int P, int nseg, const spv_exchange_segment *segs, const int *const *sparse_idx_dev /*host array of
                      nsel device pointers*/, const float *flat_grad, float scale, float *comm_allreduce,
                      float *comm_allgather, void *stream);
/* Local reduction behind a single all-gather: out[0..n) = scale * sum over ranks (rank order) of gathered[r*stride + 0..n);
 * n and stride multiples of 4 floats, buffers 16-byte aligned. */
SPV_API int spv_exchange_reduce(long long n, int world, const float *gathered, long long stride, float scale, float *out,
                        void *stream);
/* The same through NVLink peer loads from symmetric buffers (no collective library on the data path): peer_rows = host array
 * of `world` device pointers in rank order (own row included), each a row of n_row floats = [n_red floats that are summed |
 * floats that are gathered]; the gathered tails are copied into rows[r*row_stride + n_red ..).  The caller synchronises the
 * ranks before (all rows written) and must not let a row be overwritten before every rank ran this (double buffering). */
SPV_API int spv_exchange_reduce_peers(long long n_red, long long n_row, int world, const float *const *peer_rows, float scale,
                              float *reduced, float *rows, long long row_stride, void *stream);
/* The gather part alone: rows[r*row_stride + e] = peer_rows[r][e] for e in [n_red, n_row).  The three reduction entry points
 * skip their gather part when called with rows == NULL, so the gather (and the deferred spline backward behind it) can run on
 * a second stream next to the reduction. */
SPV_API int spv_exchange_gather_peers(long long n_red, long long n_row, int world, const float *const *peer_rows, float *rows,
                              long long row_stride, void *stream);
/* Two-phase form for larger groups (inbound volume 2(N-1)/N x the summed block instead of N-1 x): phase 1 sums this rank's
 * 1/world slice of [0, n_red) over the peers' rows into red_pub (its symmetric area) and `reduced`, and copies the gathered
 * tails into `rows`; the caller places a cross-rank barrier; phase 2 fetches the other slices from peer_red[owner]. */
SPV_API int spv_exchange_reduce_scatter_peers(long long n_red, long long n_row, int rank, int world, const float *const *peer_rows,
                                      float scale, float *red_pub, float *reduced, float *rows, long long row_stride,
                                      void *stream);
/* NVLS form: the NVSwitch sums.  mc_row / mc_red are the MULTICAST addresses of the symmetric row / red areas: this rank
 * multimem.ld_reduce's its 1/world slice (fp32 add in the switch), scales it and multimem.st's it into every rank's red area;
 * after the caller's second barrier each rank finds the whole summed block in its own red area. */
SPV_API int spv_exchange_nvls(long long n_red, long long n_row, int rank, int world, const float *mc_row, float *mc_red,
                      const float *const *peer_rows, float scale, float *rows, long long row_stride, void *stream);
/* NVLS mode, publish by push: the summed block row[0, n_red) is copied into this rank's symmetric staging row (sym_row, local
 * address; the peers' multimem.ld_reduce reads it), the gathered tail row[n_red, n_row) is multicast-stored into slot `rank` of
 * every rank's gathered area (mc_gathered_slot = multicast address of that slot).  After the barrier that follows every rank
 * holds all tails locally -- no peer loads. */
SPV_API int spv_exchange_publish(long long n_red, long long n_row, const float *row, float *sym_row, float *mc_gathered_slot,
                         void *stream);
SPV_API int spv_exchange_fetch_reduced(long long n_red, int rank, int world, const float *const *peer_red, float *reduced,
                               void *stream);
SPV_API int spv_exchange_unpack(int P, int nseg, const spv_exchange_segment *segs, int world, const float *comm_allreduce,
                        const float *gathered /*[world, n_allgather]*/, float *flat_grad, int *dirty, void *stream);

/* ---- Fused image losses (next row f-2): scalar loss + dL/d(rendered image), no host sync --------------------------- */
/* RGB: weight * ((1-lambda) * mean|p-g| + lambda * (1 - SSIM)) as the trainer computes it (trainer_fragGS.py:573-578 with
 * pointrix/model/loss.py:22-38 l1_loss and :62-112 ssim).  pred_chw = rendered rgb [3,H,W]; gt_hwc = ground truth [H,W,3]
 * (the trainer's layout).  The trainer passes [1,H,W,3] tensors to `ssim`, which therefore takes H as the channel count
 * (loss.py:83): the 11x11 window runs over (x, colour) inside every row -- reproduced exactly.  loss = [total, l1, ssim];
 * dL_dpred_chw [3,H,W] may be NULL (evaluation). */
SPV_API size_t spv_loss_rgb_workspace_bytes(int W, int H);
SPV_API int spv_loss_rgb(int W, int H, const float *pred_chw, const float *gt_hwc, float weight, float lambda_dssim,
                         float *loss /*[3]*/, float *dL_dpred_chw, void *workspace, size_t ws_bytes, void *stream);
/* Depth: depth_loss_dpt without weights (src/loss.py:184-206; trainer_fragGS.py:599-601): both maps shifted by their median
 * (torch.median: lower middle) and scaled by their mean absolute deviation, then MSE.  The gradient flows through the median
 * element and the deviation like autograd's; among elements tied at the median the lowest index receives it. */
SPV_API size_t spv_loss_depth_workspace_bytes(int n);
SPV_API int spv_loss_depth_dpt(int n, const float *pred, const float *gt, float weight, float *loss /*[1]*/,
                               float *dL_dpred /*[n] or NULL*/, void *workspace, size_t ws_bytes, void *stream);
/* Track: the "optical flow" loss of trainer_fragGS.py:531-571: rendered track image [>=2,H,W] (normalised x,y in channels
 * 0,1) sampled at the integer query pixels, denormalised (src/util.py:82), per-point mean |.| against target_xy, trimmed at
 * torch.quantile(q) over the visible points and weighted (src/criterion.py:46-51), divided by max(H,W).  Point i of
 * query_xy pairs with point i of target_xy/visible/weights (the reference relies on raster-ordered unique query pixels).
 * dL_dtrack_chw [2,H,W] is zero-filled by the call (may be NULL). */
SPV_API size_t spv_loss_track_workspace_bytes(int n_points);
SPV_API int spv_loss_track(int n_points, int W, int H, const float *track_chw, const int *query_xy /*[n,2] (x,y)*/,
                           const float *target_xy /*[n,2] pixels*/, const unsigned char *visible /*[n]*/,
                           const float *weights /*[n]*/, float quantile, float weight, float *loss /*[1]*/,
                           float *dL_dtrack_chw, void *workspace, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPV_B200_H */
