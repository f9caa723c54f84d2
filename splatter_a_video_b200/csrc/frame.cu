// Fused per-frame entry points for the video trainer's orthographic renderer: the whole hot path of
// DPTROrthoEnhancedRender.render_iter (/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py:205-383) forward,
// and its whole backward, each as ONE C call that only enqueues kernels on the caller's stream:
//   * no host synchronisation (the reference has two .item() calls per frame, sort_gaussian.cu:93,132): the number
//     of tile intersections stays on the device, buffers are sized by a caller-chosen capacity, overflow is flagged in
//     `status` -- so a whole training step can be captured in a CUDA graph;
//   * exact tile culling (sort.cu) shortens every list without changing any pixel;
//   * the three blend passes are one traversal (blend.cu, grouped kernels).
// Intermediates live in a caller-provided workspace that the backward call reuses.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {

constexpr int kThreads = 256;
inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct FrameWs {
    float *dirs, *rgb, *uv, *depth, *cov3d, *conic, *feature, *final_T;
    uint8_t *vis, *clamped;
    int *radius, *tiles, *idx_sorted, *tile_range, *tile_order, *ncontrib;
    void *bin_ws; size_t bin_bytes;
    // backward temporaries
    float *packed, *g_uv, *g_uv_rgb, *g_abs, *g_conic, *g_op, *g_feat, *g_rgb, *g_depth, *g_cov3d, *g_dirs;
    void *blend_ws; size_t blend_bytes;
    size_t total;
};

FrameWs carve(void *base, int P, int64_t I_cap, int W, int H, int A) {
    FrameWs f;
    char *p = (char *)base;
    const size_t Pn = (size_t)(P > 0 ? P : 1), C = 4 + (size_t)A, T = (size_t)spv::tiles_x(W) * spv::tiles_y(H), HW = (size_t)W * H;
    auto take = [&](size_t bytes) { char *r = p; p += al(bytes); return (void *)r; };
    f.dirs = (float *)take(Pn * 12); f.rgb = (float *)take(Pn * 12); f.uv = (float *)take(Pn * 8); f.depth = (float *)take(Pn * 4);
    f.cov3d = (float *)take(Pn * 24); f.conic = (float *)take(Pn * 12); f.feature = (float *)take(Pn * spv::kRecordFloats * 4);   // records
    f.final_T = (float *)take(HW * 4);
    f.vis = (uint8_t *)take(Pn); f.clamped = (uint8_t *)take(Pn * 3);
    f.radius = (int *)take(Pn * 4); f.tiles = (int *)take(Pn * 4);
    f.idx_sorted = (int *)take((size_t)(I_cap > 0 ? I_cap : 1) * 4); f.tile_range = (int *)take(T * 8);
    f.tile_order = (int *)take(T * 4);
    f.ncontrib = (int *)take(HW * 4);
    f.bin_bytes = spv_bin_capacity_workspace_bytes(P, I_cap);
    { const size_t tb = spv_bin_tiles_workspace_bytes(P, I_cap, W, H); if (tb > f.bin_bytes) f.bin_bytes = tb; }
    f.bin_ws = take(f.bin_bytes);
    f.blend_bytes = spv_alpha_blend_groups_backward_workspace_bytes(P); f.blend_ws = take(f.blend_bytes);
    f.packed = nullptr;
    f.g_uv = (float *)take(Pn * 8); f.g_uv_rgb = (float *)take(Pn * 8); f.g_abs = (float *)take(Pn * 8);
    f.g_conic = (float *)take(Pn * 12); f.g_op = (float *)take(Pn * 4); f.g_feat = (float *)take(Pn * C * 4);
    f.g_rgb = (float *)take(Pn * 12); f.g_depth = (float *)take(Pn * 4); f.g_cov3d = (float *)take(Pn * 24);
    f.g_dirs = (float *)take(Pn * 12);
    f.total = (size_t)(p - (char *)base);
    return f;
}

__global__ void __launch_bounds__(kThreads)
frame_prep_kernel(int P, float *__restrict__ dirs) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    dirs[3 * i] = 0.f; dirs[3 * i + 1] = 0.f; dirs[3 * i + 2] = 1.f;   // constant view direction (0,0,1), :270-271
}

// Attribute groups: the renderer receives the attribute stack as separately named per-Gaussian tensors
// (render_attributes_list, trainer_fragGS.py:510-512); they are consumed / differentiated in place, no torch.cat.
constexpr int kMaxGroups = 8;
struct AttrGroups { const float *in[kMaxGroups]; float *grad[kMaxGroups]; int ch[kMaxGroups]; int start[kMaxGroups]; int n; };

// Packed gradient rows of the grouped blend backward (spv::kPackedRowGroups layout) -> everything the rest of the chain and
// the caller need, in one pass: uv / conic (workspace), opacity, colour and depth gradients (workspace), the attribute
// groups' gradients straight into their own tensors, and the ndc dummies' gradients with their [W/2,H/2] scale.
__global__ void __launch_bounds__(kThreads)
unpack_frame_kernel(int P, int A, const float *__restrict__ packed, float2 *__restrict__ g_uv, float *__restrict__ g_conic,
                    float *__restrict__ g_op, float *__restrict__ g_rgb, float *__restrict__ g_depth, const AttrGroups gr,
                    float2 *__restrict__ g_ndc, float2 *__restrict__ g_abs_ndc, float hw, float hh) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float r[spv::kPackedRowGroups];
    const float4 *row = reinterpret_cast<const float4 *>(packed + (size_t)i * spv::kPackedRowGroups);
#pragma unroll
    for (int k = 0; k < spv::kPackedRowGroups / 4; ++k) {
        const float4 q = row[k];
        r[4 * k] = q.x; r[4 * k + 1] = q.y; r[4 * k + 2] = q.z; r[4 * k + 3] = q.w;
    }
    if (g_uv) {   // NULL: the fused geometry backward reads uv / conic / depth gradients straight from the packed rows
        g_uv[i] = make_float2(r[0], r[1]);
        g_conic[3 * i] = r[4]; g_conic[3 * i + 1] = r[5]; g_conic[3 * i + 2] = r[6];
        g_depth[i] = r[11];
    }
    g_op[i] = r[7];
    g_rgb[3 * i] = r[8]; g_rgb[3 * i + 1] = r[9]; g_rgb[3 * i + 2] = r[10];
#pragma unroll
    for (int q = 0; q < kMaxGroups; ++q) {
        if (q < gr.n && gr.grad[q]) {
            float *o = gr.grad[q] + (size_t)i * gr.ch[q];
#pragma unroll
            for (int c = 0; c < 19; ++c)
                if (c >= gr.start[q] && c < gr.start[q] + gr.ch[q]) o[c - gr.start[q]] = r[12 + c];
        }
    }
    if (g_ndc) g_ndc[i] = make_float2(r[31] * hw, r[32] * hh);
    if (g_abs_ndc) g_abs_ndc[i] = make_float2(r[2] * hw, r[3] * hh);
}

inline int make_groups(AttrGroups &g, int n, const float *const *in, float *const *grad, const int *ch) {
    g.n = n;
    int A = 0;
    for (int q = 0; q < kMaxGroups; ++q) {
        g.in[q] = (q < n && in) ? in[q] : nullptr;
        g.grad[q] = (q < n && grad) ? grad[q] : nullptr;
        g.ch[q] = q < n ? ch[q] : 0;
        g.start[q] = A;
        A += g.ch[q];
    }
    return A;
}

#define SPV_TRY_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

using spv::SideLane;
using spv::side_lane;   // the per-(thread, device) second stream for independent branches of a frame (runtime.cu)

}  // namespace

extern "C" {

size_t spv_frame_workspace_bytes(int P, int64_t I_cap, int W, int H, int A) {
    return carve(nullptr, P, I_cap, W, H, A).total;
}

int spv_frame_ortho_forward(int P, int W, int H, int n_groups, const float *const *attr_ptrs, const int *attr_channels,
                            int K, int64_t I_cap, int cull, const float *position, const float *scaling,
                            const float *rotation, const float *opacity, const float *shs, int sh_bases, const float *extr,
                            float nearest, float extent, float bg_rgb, float *images, int *gs_idx, int *radii,
                            int *status, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0 || W <= 0 || H <= 0) return 0;
    if (n_groups < 0 || n_groups > kMaxGroups) { spv::set_error(cudaErrorInvalidValue, "spv_frame_ortho_forward: at most 8 attribute groups"); return (int)cudaErrorInvalidValue; }
    AttrGroups gr;
    const int A = make_groups(gr, n_groups, attr_ptrs, nullptr, attr_channels);
    if (4 + A > 23 || K <= 0) { spv::set_error(cudaErrorInvalidValue, "spv_frame_ortho_forward: need at most 19 attribute channels and K > 0"); return (int)cudaErrorInvalidValue; }
    if (sh_bases != 16 && sh_bases != 4) { spv::set_error(cudaErrorInvalidValue, "spv_frame_ortho_forward: shs is [P,16,3] (degree 3) or [P,4,3] (bases 0, 2, 6, 12)"); return (int)cudaErrorInvalidValue; }
    FrameWs f = carve(workspace, P, I_cap, W, H, A);
    if (ws_bytes < f.total) { spv::set_error(cudaErrorInvalidValue, "spv_frame_ortho_forward: workspace too small"); return (int)cudaErrorInvalidValue; }
    const unsigned g = spv::cdiv(P, kThreads);
    const int C = 4 + A;
    // ---- side branch: SH colours (evaluated for every point: the renderer passes no visibility mask, :272), then -- once
    //      the main branch has the conics -- the per-Gaussian blend records and the -1 fill of the id image
    SideLane *lane = side_lane();
    if (!lane) { spv::set_error(cudaGetLastError(), "spv_frame_ortho_forward: side stream"); return (int)cudaErrorUnknown; }
    // The serial chain geometry -> count -> scan -> emit -> tile sort is the critical path up to the forward blend; it runs on the
    // lane's high-priority stream so that its (small, latency-bound) kernels are not queued behind the CTA backlog of the side
    // branch's streaming kernels (measured: tile_scan's single CTA waited 20 us for pack_records' 7000).
    const bool flat = spv::get_option("flat_chain") != 0;
    cudaStream_t hs = flat ? s : lane->hi;
    SPV_CUDA_TRY(cudaEventRecord(lane->fork, s), "spv_frame_ortho_forward/fork");
    SPV_CUDA_TRY(cudaStreamWaitEvent(lane->stream, lane->fork, 0), "spv_frame_ortho_forward/fork");
    if (!flat) SPV_CUDA_TRY(cudaStreamWaitEvent(hs, lane->fork, 0), "spv_frame_ortho_forward/fork");
    if (sh_bases == 4) {     // only the coefficients the constant view direction reaches: 48 B per Gaussian, no direction array
        SPV_TRY_RC(spv_compute_sh_z_forward(P, shs, f.rgb, f.clamped, (void *)lane->stream));
    } else {
        frame_prep_kernel<<<g, kThreads, 0, lane->stream>>>(P, f.dirs);      // only the SH kernels read it: off the main branch
        SPV_TRY_RC(spv::check_launch("spv_frame_ortho_forward/prep"));
        SPV_TRY_RC(spv_compute_sh_forward(P, shs, 3, f.dirs, nullptr, 0, f.rgb, f.clamped, (void *)lane->stream));
    }
    // ---- main branch: projection, visibility, covariance, conic / radius / tile rectangle in ONE pass (geometry.cu: the staged
    //      kernels' bodies back to back, bit-identical results), then culled binning + tile sort
    static const bool radix = [] { const char *e = getenv("SPV_BIN_RADIX"); return e && e[0] == '1'; }();
    if (!radix) SPV_TRY_RC(spv::bin_tiles_clear(P, I_cap, W, H, f.tile_range, status, f.bin_ws, (void *)hs));   // one launch, before the chain
    SPV_TRY_RC(spv::frame_geometry_forward(P, position, scaling, rotation, extr, W, H, nearest, extent, f.uv, f.depth, f.vis, f.cov3d,
                                           f.conic, f.radius, f.tiles, radii, (void *)hs));
    SPV_CUDA_TRY(cudaEventRecord(lane->mid, hs), "spv_frame_ortho_forward/mid");
    SPV_CUDA_TRY(cudaStreamWaitEvent(lane->stream, lane->mid, 0), "spv_frame_ortho_forward/mid");
    SPV_TRY_RC(spv::pack_records(P, A, f.uv, f.conic, opacity, f.radius, f.rgb, f.depth, n_groups, attr_ptrs, attr_channels,
                                 f.feature, (void *)lane->stream));
    SPV_CUDA_TRY(cudaMemsetAsync(gs_idx, 0xFF, sizeof(int) * (size_t)H * W * K, lane->stream), "spv_frame_ortho_forward");
    // the backward's packed gradient rows are cleared here, off the critical path (the workspace belongs to this frame)
    SPV_CUDA_TRY(cudaMemsetAsync(f.blend_ws, 0, sizeof(float) * (size_t)spv::kPackedRowGroups * P, lane->stream), "spv_frame_ortho_forward");
    SPV_CUDA_TRY(cudaEventRecord(lane->join, lane->stream), "spv_frame_ortho_forward/join");
    // binning: per-tile segments + shared-memory tile sort (SPV_BIN_RADIX=1 selects the global radix sort for A/B runs)
    if (radix) SPV_TRY_RC(spv_bin_capacity(P, I_cap, f.uv, f.depth, f.radius, f.conic, opacity, cull, W, H, f.idx_sorted, f.tile_range,
                                           status, f.bin_ws, f.bin_bytes, (void *)hs));
    else SPV_TRY_RC(spv::bin_tiles_ordered(P, I_cap, f.uv, f.depth, f.radius, f.conic, opacity, cull, W, H, f.idx_sorted, f.tile_range,
                                           status, f.tile_order, /*cleared=*/true, f.bin_ws, f.bin_bytes, (void *)hs));
    if (!flat) {
        SPV_CUDA_TRY(cudaEventRecord(lane->hi_join, hs), "spv_frame_ortho_forward/join");
        SPV_CUDA_TRY(cudaStreamWaitEvent(s, lane->hi_join, 0), "spv_frame_ortho_forward/join");
    }
    SPV_CUDA_TRY(cudaStreamWaitEvent(s, lane->join, 0), "spv_frame_ortho_forward/join");
    return spv::blend_records_forward(C, W, H, K, f.feature, f.idx_sorted, f.tile_range, radix ? nullptr : f.tile_order, bg_rgb, 1.0f,
                                      0.0f, images, f.final_T, f.ncontrib, gs_idx, stream);
}

int spv_frame_ortho_backward(int P, int W, int H, int n_groups, const int *attr_channels, int n_grad_channels, int64_t I_cap,
                             const float *scaling, const float *rotation, const float *opacity, const float *shs, int sh_bases,
                             const float *extr, float bg_rgb, const float *const *dL_dimage_planes, float *dL_dposition,
                             float *dL_dscaling, float *dL_drotation, float *dL_dopacity, float *dL_dshs,
                             float *const *dL_dattr_ptrs, float *dL_dndc, float *dL_dabs_ndc, float *dL_drgb_out,
                             uint8_t *clamped_out, int first_backward, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0 || W <= 0 || H <= 0) return 0;
    AttrGroups gr;
    const int A = make_groups(gr, n_groups, nullptr, dL_dattr_ptrs, attr_channels);
    FrameWs f = carve(workspace, P, I_cap, W, H, A);
    if (ws_bytes < f.total) { spv::set_error(cudaErrorInvalidValue, "spv_frame_ortho_backward: workspace too small"); return (int)cudaErrorInvalidValue; }
    const int C = 4 + A;
    const unsigned g = spv::cdiv(P, kThreads);
    float *packed = (float *)f.blend_ws;
    static const bool radix = [] { const char *e = getenv("SPV_BIN_RADIX"); return e && e[0] == '1'; }();
    SPV_TRY_RC(spv::blend_records_backward(P, C, W, H, f.feature, f.idx_sorted, f.tile_range, radix ? nullptr : f.tile_order, bg_rgb, 1.0f, 0.0f, f.final_T,
                                           f.ncontrib, dL_dimage_planes, n_grad_channels, /*want_abs=*/dL_dabs_ndc != nullptr, packed,
                                           /*packed_is_zero=*/first_backward != 0, stream));
    // Two independent readers of the packed rows: (side stream) the unpack of the opacity / colour / attribute / ndc gradients and,
    // behind it, colours -> SH coefficients (view direction is a constant: its gradient is discarded); (caller's stream)
    // uv, depth -> position ; conic -> cov3d -> scaling, rotation.
    // Deferred SH backward (frame-parallel training): the colour gradient and the clamp mask leave through the caller's
    // buffers, the SH coefficients' gradient is produced after the gradient exchange from the REDUCED colour gradient.
    const bool defer_sh = dL_drgb_out != nullptr;
    float *g_rgb = defer_sh ? dL_drgb_out : f.g_rgb;
    SideLane *lane = side_lane();
    if (!lane) { spv::set_error(cudaGetLastError(), "spv_frame_ortho_backward: side stream"); return (int)cudaErrorUnknown; }
    SPV_CUDA_TRY(cudaEventRecord(lane->fork, s), "spv_frame_ortho_backward/fork");
    SPV_CUDA_TRY(cudaStreamWaitEvent(lane->stream, lane->fork, 0), "spv_frame_ortho_backward/fork");
    unpack_frame_kernel<<<g, kThreads, 0, lane->stream>>>(P, A, packed, nullptr, nullptr, dL_dopacity, g_rgb, nullptr, gr,
                                                          (float2 *)dL_dndc, (float2 *)dL_dabs_ndc, 0.5f * (float)W, 0.5f * (float)H);
    SPV_TRY_RC(spv::check_launch("spv_frame_ortho_backward/unpack"));
    if (defer_sh) {
        if (clamped_out) SPV_CUDA_TRY(cudaMemcpyAsync(clamped_out, f.clamped, (size_t)P * 3, cudaMemcpyDeviceToDevice, lane->stream), "spv_frame_ortho_backward");
    } else {
        if (sh_bases == 4) SPV_TRY_RC(spv_compute_sh_z_backward(P, f.clamped, f.g_rgb, dL_dshs, (void *)lane->stream));
        else SPV_TRY_RC(spv_compute_sh_backward(P, shs, 3, f.dirs, nullptr, f.clamped, f.g_rgb, 16, dL_dshs, /*dL_ddirs=*/nullptr, (void *)lane->stream));   // constant view direction: its gradient is discarded
    }
    SPV_CUDA_TRY(cudaEventRecord(lane->join, lane->stream), "spv_frame_ortho_backward/join");
    SPV_TRY_RC(spv::frame_geometry_backward(P, packed, scaling, rotation, extr, W, H, f.depth, f.vis, f.cov3d, f.radius, dL_dposition,
                                            dL_dscaling, dL_drotation, stream));
    SPV_CUDA_TRY(cudaStreamWaitEvent(s, lane->join, 0), "spv_frame_ortho_backward/join");
    return 0;
}

}  // extern "C"
