// Densification on the flat per-Gaussian SoA (structure half): the role of
// AtlasGaussianSplattingOptimizer.update_structure / densification / prune / reset_opacity
// (/root/reference/src/pointrix/optimizer/atlas_gs_optimizer.py:93-379) and of PointCloud.extand_points / remove_points with
// their optimizer-state surgery (/root/reference/src/pointrix/point_cloud/points.py:281-365).
//
// The reference runs ~10 masked torch ops per step for the statistics and, every 100 steps, re-creates every parameter and
// both Adam moments tensor by tensor (index, cat, nn.Parameter, state dict surgery, empty_cache).  Here:
//   densify_stats_kernel   one pass: max radius, accumulated |grad ndc|, visit count            (:110-121)
//   densify_flags_kernel   one pass: clone / split / prune decisions per point                  (:199-252, :333-353)
//   regather_rows_kernel   ONE launch moves every parameter block (or Adam moment block) of the flat buffer to its place in
//                          the new population: new_row[j] = old_row[src[j]] (src < 0: zeros -- fresh optimizer state)
//   split_children_kernel  position / scaling of the split children                            (:254-283)
//   reset_opacity_kernel   opacity cap + cleared moments                                        (:185-197)
// The index lists (which old point every new point comes from) are small integer work done by the host wrapper
// (splatter_a_video_b200/densify.py) so the population order is exactly the reference's: kept, clones, split children.
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {
constexpr int kThreads = 256;
constexpr int kMaxBlocks = 32;

__global__ void __launch_bounds__(kThreads)
densify_stats_kernel(int P, const float2 *__restrict__ ndc_grad, const int *__restrict__ radii, const uint8_t *__restrict__ visible,
                     float *__restrict__ grad_accum, float *__restrict__ denom, float *__restrict__ max_radii) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (visible ? visible[i] != 0 : r > 0) {
        max_radii[i] = fmaxf(max_radii[i], (float)r);
        const float2 g = ndc_grad[i];
        grad_accum[i] += sqrtf(g.x * g.x + g.y * g.y);
        denom[i] += 1.0f;
    }
}

// bit 0 clone, bit 1 split, bit 2 prune.  `scaling` / `opacity` are the stored (pre-activation) parameters.
__global__ void __launch_bounds__(kThreads)
densify_flags_kernel(int P, const float *__restrict__ grad_accum, const float *__restrict__ denom, const float *__restrict__ scaling,
                     const float *__restrict__ opacity, const float *__restrict__ max_radii, int scaling_is_log, int opacity_is_logit,
                     float grad_threshold, float dense_extent, float min_opacity, float size_threshold, float big_extent,
                     uint8_t *__restrict__ flags) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float g = grad_accum[i] / denom[i];
    if (isnan(g)) g = 0.0f;                                              // grads[grads.isnan()] = 0.0
    float s0 = scaling[3 * i], s1 = scaling[3 * i + 1], s2 = scaling[3 * i + 2];
    if (scaling_is_log) { s0 = expf(s0); s1 = expf(s1); s2 = expf(s2); }
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    float o = opacity[i];
    if (opacity_is_logit) o = 1.0f / (1.0f + expf(-o));
    const bool hot = fabsf(g) >= grad_threshold;                          // torch.norm over the size-1 last dim
    uint8_t f = 0;
    if (hot && smax <= dense_extent) f |= 1;
    if (g >= grad_threshold && smax > dense_extent) f |= 2;               // split tests the padded scalar, not its norm
    if (o < min_opacity || (size_threshold > 0.f && (max_radii[i] > size_threshold || smax > big_extent))) f |= 4;
    flags[i] = f;
}

struct Blocks { long long old_off[kMaxBlocks], new_off[kMaxBlocks]; int width[kMaxBlocks]; int n; };

// one warp per (block, new row): lanes over the row's floats -> coalesced on both sides, no integer division
__global__ void __launch_bounds__(kThreads)
regather_rows_kernel(int P_new, Blocks b, const int *__restrict__ src, const float *__restrict__ old_flat, float *__restrict__ new_flat) {
    const int q = blockIdx.y;
    const int j = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= P_new) return;
    const int w = b.width[q], s = src[j];
    float *dst = new_flat + b.new_off[q] + (long long)j * w;
    if (s < 0) {
        for (int c = lane; c < w; c += 32) dst[c] = 0.f;
    } else {
        const float *row = old_flat + b.old_off[q] + (long long)s * w;
        for (int c = lane; c < w; c += 32) dst[c] = row[c];
    }
}

// new_pos = R(rotation[parent]) @ sample + position[parent]; new_scaling = inverse_activation(scaling[parent] / div)
__global__ void __launch_bounds__(kThreads)
split_children_kernel(int P_new, const int *__restrict__ src, const int *__restrict__ child_index, const float *__restrict__ samples,
                      const float *__restrict__ old_position, const float *__restrict__ old_scaling, const float4 *__restrict__ old_rotation,
                      int scaling_is_log, float div, float *__restrict__ new_position, float *__restrict__ new_scaling) {
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= P_new) return;
    const int ci = child_index[j];
    if (ci < 0) return;
    const int p = src[j];
    const float4 q0 = old_rotation[p];
    const float nrm = sqrtf(q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w);     // build_rotation normalises
    const float r = q0.x / nrm, x = q0.y / nrm, y = q0.z / nrm, z = q0.w / nrm;
    const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                           {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                           {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
    const float sx = samples[3 * ci], sy = samples[3 * ci + 1], sz = samples[3 * ci + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        new_position[3 * j + k] = (R[k][0] * sx + R[k][1] * sy + R[k][2] * sz) + old_position[3 * p + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float s = old_scaling[3 * p + k];
        if (scaling_is_log) s = logf(expf(s) / div); else s = s / div;
        new_scaling[3 * j + k] = s;
    }
}

__global__ void __launch_bounds__(kThreads)
reset_opacity_kernel(int P, float cap, int opacity_is_logit, float *__restrict__ opacity, float *__restrict__ exp_avg,
                     float *__restrict__ exp_avg_sq) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float o = opacity[i];
    if (opacity_is_logit) {
        o = fminf(1.0f / (1.0f + expf(-o)), cap);
        o = logf(o / (1.0f - o));                                         // inverse_sigmoid
    } else o = fminf(o, cap);
    opacity[i] = o;
    if (exp_avg) exp_avg[i] = 0.f;
    if (exp_avg_sq) exp_avg_sq[i] = 0.f;
}
}  // namespace

extern "C" {

int spv_densify_stats(int P, const float *ndc_grad, const int *radii, const uint8_t *visible, float *grad_accum, float *denom,
                      float *max_radii, void *stream) {
    if (P <= 0) return 0;
    densify_stats_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, (const float2 *)ndc_grad, radii, visible, grad_accum,
                                                                                      denom, max_radii);
    return spv::check_launch("spv_densify_stats");
}

int spv_densify_flags(int P, const float *grad_accum, const float *denom, const float *scaling, const float *opacity,
                      const float *max_radii, int scaling_is_log, int opacity_is_logit, float grad_threshold, float dense_extent,
                      float min_opacity, float size_threshold, float big_extent, uint8_t *flags, void *stream) {
    if (P <= 0) return 0;
    densify_flags_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P, grad_accum, denom, scaling, opacity, max_radii, scaling_is_log, opacity_is_logit, grad_threshold, dense_extent, min_opacity,
        size_threshold, big_extent, flags);
    return spv::check_launch("spv_densify_flags");
}

int spv_flat_regather(int n_blocks, const int *widths_host, const long long *old_offsets_host, const long long *new_offsets_host,
                      int P_new, const int *src, const float *old_flat, float *new_flat, void *stream) {
    if (P_new <= 0 || n_blocks <= 0) return 0;
    if (n_blocks > kMaxBlocks) { spv::set_error(cudaErrorInvalidValue, "spv_flat_regather: at most 32 blocks"); return (int)cudaErrorInvalidValue; }
    Blocks b;
    b.n = n_blocks;
    for (int q = 0; q < kMaxBlocks; ++q) {
        b.width[q] = q < n_blocks ? widths_host[q] : 0;
        b.old_off[q] = q < n_blocks ? old_offsets_host[q] : 0;
        b.new_off[q] = q < n_blocks ? new_offsets_host[q] : 0;
    }
    const dim3 grid(spv::cdiv(P_new, kThreads / 32), n_blocks);
    regather_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(P_new, b, src, old_flat, new_flat);
    return spv::check_launch("spv_flat_regather");
}

int spv_split_children(int P_new, const int *src, const int *child_index, const float *samples, const float *old_position,
                       const float *old_scaling, const float *old_rotation, int scaling_is_log, float div, float *new_position,
                       float *new_scaling, void *stream) {
    if (P_new <= 0) return 0;
    split_children_kernel<<<spv::cdiv(P_new, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P_new, src, child_index, samples, old_position, old_scaling, (const float4 *)old_rotation, scaling_is_log, div, new_position,
        new_scaling);
    return spv::check_launch("spv_split_children");
}

int spv_reset_opacity(int P, float cap, int opacity_is_logit, float *opacity, float *exp_avg, float *exp_avg_sq, void *stream) {
    if (P <= 0) return 0;
    reset_opacity_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, cap, opacity_is_logit, opacity, exp_avg, exp_avg_sq);
    return spv::check_launch("spv_reset_opacity");
}

}  // extern "C"
