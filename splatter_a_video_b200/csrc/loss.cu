// Fused image losses of the training step: the step immediately after the rasterizer.  Each entry
// produces the scalar loss AND dL/d(rendered image) in a handful of launches, with no host synchronisation, so the whole
// render -> loss -> backward chain stays on one stream (and inside one CUDA graph).
//
//   spv_loss_rgb        0.8 L1 + 0.2 (1 - SSIM)     trainer_fragGS.py:573-578, pointrix/model/loss.py:22-112
//   spv_loss_depth_dpt  median/MAD-normalised MSE   trainer_fragGS.py:599-601, src/loss.py:184-206
//   spv_loss_track      quantile-trimmed, confidence-weighted L1 on the rendered track image at the query pixels
//                                                   trainer_fragGS.py:531-571, src/criterion.py:46-51, src/util.py:75-82
//
// SSIM quirk kept on purpose: the trainer hands `ssim` tensors of shape [1,H,W,3], and the reference takes
// `channel = img.size(-3)` = H, so its depthwise 11x11 Gaussian window slides over the (x, colour) plane of every image ROW:
// 11 taps along x and 11 taps along the 3-wide colour axis (zero padded), rows never mix.
#include "common.cuh"
#include "../../include/spv_b200.h"

#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int kRow = 128;               // pixels of one image row per CTA
constexpr int kHalo = 5;                // window_size // 2
constexpr int kSpan = kRow + 2 * kHalo;
constexpr int kRed = 256;               // threads of the streaming reduction kernels
constexpr int kRedBlocks = 296;         // 2 CTAs per SM
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// gaussian(11, 1.5) of pointrix/model/loss.py:58-60, evaluated in float32 like the reference (exp in double, sum in float)
__constant__ float kG[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f, 2.130055279e-01f,
                             2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f, 3.600077331e-02f, 7.598758209e-03f,
                             1.028380124e-03f};

__device__ __forceinline__ float sgnf(float d) { return (float)((d > 0.f) - (d < 0.f)); }

template <int NT>
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    // fixed-order tree: warp shuffles, then warp 0 over the per-warp sums; every thread of warp 0 gets the total
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < NT / 32 ? scratch[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        t = __shfl_sync(0xffffffffu, t, 0);
    }
    __syncthreads();
    return t;
}

// Sums `k` interleaved columns of a [n][k] array of per-CTA partials in a fixed order; result valid in every thread.
template <int NT, int K>
__device__ __forceinline__ void total_of_partials(const double *__restrict__ part, int n, double (&out)[K], double *scratch,
                                                  double *bcast) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double v = 0.0;
        for (int i = threadIdx.x; i < n; i += NT) v += __ldcg(part + (size_t)i * K + j);   // L2: also valid for partials written by
                                                                                           // other CTAs of the SAME launch (ticket pattern)
        v = block_sum<NT>(v, scratch);
        if (threadIdx.x == 0) bcast[j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) out[j] = bcast[j];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ RGB: L1 + SSIM
// ONE kernel, one CTA per image row (rows never mix, see the quirk above): the row of prediction and ground truth is staged
// in shared memory, every pixel's SSIM value and its derivatives w.r.t. the three window sums that depend on the prediction
// (E[p], E[p^2], E[p g]; pre-multiplied by dLoss/dSSIM) stay in shared memory, and after one barrier the transposed depthwise
// convolution (the window is symmetric: the same zero-padded convolution) + the L1 sign term give
//     dL/dp = G*fA + 2 p (G*fB) + g (G*fC) + l1scale sign(p - g).
// 2 x 4.9 MB read and 4.9 MB written per 854x480 frame instead of the 9 derivative maps (14.7 MB) going through HBM twice.
// The last CTA to finish (ticket) sums the per-row partials in a fixed order and writes the three scalars.
constexpr int kRgbThreads = 256;

__global__ void __launch_bounds__(kRgbThreads)
rgb_row_kernel(int W, int H, const float *__restrict__ pred, const float *__restrict__ gt, float dscale, float l1scale,
               float weight, float lambda, float *__restrict__ dL_dpred, double *__restrict__ partials, unsigned *__restrict__ ticket,
               float *__restrict__ loss) {
    extern __shared__ __align__(16) float s_row[];
    const int span = W + 2 * kHalo;
    float *sp = s_row, *sg = s_row + 3 * span, *sf = s_row + 6 * span;   // [3][span], [3][span], [9][span]
    __shared__ double scratch[kRgbThreads / 32], bcast[2];
    __shared__ bool s_last;
    const int y = blockIdx.x;
    const size_t HW = (size_t)H * W, row = (size_t)y * W;
    for (int i = threadIdx.x; i < 3 * span; i += kRgbThreads) {
        const int c = i / span, k = i - c * span, x = k - kHalo;
        sp[i] = (x >= 0 && x < W) ? pred[c * HW + row + x] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * W; i += kRgbThreads) {     // ground truth is [H,W,3]: one contiguous row
        const int x = i / 3, c = i - 3 * x;
        sg[c * span + x + kHalo] = gt[row * 3 + i];
    }
    for (int i = threadIdx.x; i < 3 * 2 * kHalo; i += kRgbThreads) {
        const int c = i / (2 * kHalo), k = i - c * 2 * kHalo;
        sg[c * span + (k < kHalo ? k : W + k)] = 0.f;
    }
    for (int i = threadIdx.x; i < 9 * 2 * kHalo; i += kRgbThreads) {
        const int m = i / (2 * kHalo), k = i - m * 2 * kHalo;
        sf[m * span + (k < kHalo ? k : W + k)] = 0.f;
    }
    __syncthreads();
    float ssim_sum = 0.f, l1_sum = 0.f;
    for (int x = threadIdx.x; x < W; x += kRgbThreads) {
        float h[3][5];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = kG[k], p = sp[c * span + x + k], g = sg[c * span + x + k];
                a0 += w * p; a1 += w * g; a2 += w * (p * p); a3 += w * (g * g); a4 += w * (p * g);
            }
            h[c][0] = a0; h[c][1] = a1; h[c][2] = a2; h[c][3] = a3; h[c][4] = a4;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const float w = kG[5 + cc - c];
#pragma unroll
                for (int q = 0; q < 5; ++q) m[q] += w * h[cc][q];
            }
            const float mu1 = m[0], mu2 = m[1];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = m[2] - mu1_sq, s2 = m[3] - mu2_sq, s12 = m[4] - mu12;
            const float a1 = 2.f * mu12 + kC1, a2 = 2.f * s12 + kC2, b1 = mu1_sq + mu2_sq + kC1, b2 = s1 + s2 + kC2;
            const float inv = 1.f / (b1 * b2);
            ssim_sum += (a1 * a2) * inv;
            // partials with (mu1, s1, s12) independent, then chained onto the raw window sums A=E[p], B=E[p^2], C=E[p g]
            const float d_s12 = 2.f * a1 * inv;
            const float d_s1 = -(a1 * a2) * inv / b2;
            const float d_mu1 = 2.f * mu2 * a2 * inv - 2.f * mu1 * (a1 * a2) * inv / b1;
            sf[(0 * 3 + c) * span + x + kHalo] = dscale * (d_mu1 - 2.f * mu1 * d_s1 - mu2 * d_s12);
            sf[(1 * 3 + c) * span + x + kHalo] = dscale * d_s1;
            sf[(2 * 3 + c) * span + x + kHalo] = dscale * d_s12;
            l1_sum += fabsf(sp[c * span + x + kHalo] - sg[c * span + x + kHalo]);
        }
    }
    __syncthreads();
    if (dL_dpred) {
        for (int x = threadIdx.x; x < W; x += kRgbThreads) {
            float h[3][3];   // [kind][colour]: horizontal pass
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) a += kG[k] * sf[m * span + x + k];
                h[m / 3][m % 3] = a;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float GA = 0.f, GB = 0.f, GC = 0.f;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float w = kG[5 + cc - c];
                    GA += w * h[0][cc]; GB += w * h[1][cc]; GC += w * h[2][cc];
                }
                const float p = sp[c * span + x + kHalo], g = sg[c * span + x + kHalo];
                dL_dpred[c * HW + row + x] = GA + 2.f * p * GB + g * GC + l1scale * sgnf(p - g);
            }
        }
    }
    const double t0 = block_sum<kRgbThreads>((double)ssim_sum, scratch);
    const double t1 = block_sum<kRgbThreads>((double)l1_sum, scratch);
    if (threadIdx.x == 0) {
        partials[2 * y] = t0; partials[2 * y + 1] = t1;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)(H - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double t[2];
    total_of_partials<kRgbThreads, 2>(partials, H, t, scratch, bcast);
    if (threadIdx.x == 0) {
        const double n_elems = 3.0 * (double)H * (double)W;
        const float ssim = (float)(t[0] / n_elems), l1 = (float)(t[1] / n_elems);
        loss[0] = weight * ((1.f - lambda) * l1 + lambda * (1.f - ssim));
        loss[1] = l1;
        loss[2] = ssim;
        *ticket = 0u;   // ready for the next call (also inside a replayed CUDA graph)
    }
}

// ------------------------------------------------------------------------------------------------ depth: depth_loss_dpt
// torch.median of the two maps = the element of rank (n-1)/2: a three-pass radix SELECT (11 + 11 + 10 bits of the order-preserving
// key) over both maps at once instead of two full 32-bit sorts.  Per pass: shared-memory histogram of the digit of every element
// that still matches the prefix found so far (warp-aggregated: background pixels share one value), merged into a global
// histogram; the last CTA of each map (ticket) walks the 2048 bins in order, fixes the digit and the remaining rank, and
// clears the histogram for the next pass.
constexpr int kSelBins = 2048;
struct SelState { unsigned prefix, rank, ticket, pad; };

__device__ __forceinline__ unsigned order_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

template <int SHIFT, int BITS>
__global__ void __launch_bounds__(kRed)
select_pass_kernel(int n, const float *__restrict__ a0, const float *__restrict__ a1, unsigned *__restrict__ hist /*[2][kSelBins]*/,
                   SelState *__restrict__ state /*[2]*/, float *__restrict__ med /*[2]*/) {
    __shared__ unsigned s_hist[kSelBins];
    __shared__ unsigned s_scan[kRed];
    __shared__ bool s_last;
    const int which = blockIdx.y;
    const float *__restrict__ src = which ? a1 : a0;
    unsigned *gh = hist + which * kSelBins;
    SelState *st = state + which;
    constexpr unsigned kDigits = 1u << BITS;
    const unsigned prefix = SHIFT + BITS < 32 ? __ldcg(&st->prefix) : 0u;
    constexpr unsigned hi_mask = SHIFT + BITS < 32 ? (0xffffffffu << (SHIFT + BITS)) : 0u;
    for (int i = threadIdx.x; i < kSelBins; i += kRed) s_hist[i] = 0u;
    __syncthreads();
    for (int i0 = blockIdx.x * kRed; i0 < n; i0 += gridDim.x * kRed) {   // whole warps iterate together (match_any below)
        const int i = i0 + threadIdx.x;
        unsigned digit = 0xffffffffu;
        if (i < n) {
            const unsigned k = order_key(src[i]);
            if ((k & hi_mask) == prefix) digit = (k >> SHIFT) & (kDigits - 1u);
        }
        // one atomic per warp when the whole warp shares a digit (background pixels, narrow value ranges), else one per lane
        int same;
        __match_all_sync(0xffffffffu, digit, &same);
        if (same) { if (digit != 0xffffffffu && (threadIdx.x & 31) == 0) atomicAdd(&s_hist[digit], 32u); }
        else if (digit != 0xffffffffu) atomicAdd(&s_hist[digit], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)kDigits; i += kRed) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&gh[i], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last CTA of this map: find the digit whose cumulative count crosses the rank
    constexpr int per = kDigits / kRed;
    static_assert(kDigits % kRed == 0, "bins per thread");
    unsigned loc[per], sum = 0;
#pragma unroll
    for (int q = 0; q < per; ++q) { loc[q] = __ldcg(&gh[threadIdx.x * per + q]); sum += loc[q]; }
    // exclusive scan of the per-thread sums: warp shuffles, then the kRed / 32 warp totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned run = incl - sum;
#pragma unroll
    for (int w = 0; w < kRed / 32; ++w) if (w < warp) run += s_scan[w];
    const unsigned rank = SHIFT + BITS < 32 ? __ldcg(&st->rank) : (unsigned)((n - 1) / 2);
#pragma unroll
    for (int q = 0; q < per; ++q) {
        if (rank >= run && rank < run + loc[q]) {     // exactly one (thread, q) over the CTA
            const unsigned np = prefix | ((unsigned)(threadIdx.x * per + q) << SHIFT);
            st->prefix = np;
            st->rank = rank - run;
            if (SHIFT == 0) med[which] = key_value(np);
        }
        run += loc[q];
    }
#pragma unroll
    for (int q = 0; q < per; ++q) gh[threadIdx.x * per + q] = 0u;
    if (threadIdx.x == 0) st->ticket = 0u;
}

// A: sum |p - t_p|, sum sign(p - t_p), sum |g - t_g|; index of the (first) element that equals the median of p
__global__ void __launch_bounds__(kRed)
depth_stats_kernel(int n, const float *__restrict__ pred, const float *__restrict__ gt, const float *__restrict__ med,
                   double *__restrict__ partA, int *__restrict__ med_idx) {
    __shared__ double scratch[kRed / 32];
    const float tp = med[0], tg = med[1];   // torch.median: the lower of the two middles (rank (n-1)/2)
    float a = 0.f, sg = 0.f, b = 0.f;
    int first = 0x7fffffff;
    for (int i = blockIdx.x * kRed + threadIdx.x; i < n; i += gridDim.x * kRed) {
        const float p = pred[i], d = p - tp;
        a += fabsf(d); sg += sgnf(d); b += fabsf(gt[i] - tg);
        if (p == tp) first = min(first, i);
    }
    first = __reduce_min_sync(0xffffffffu, first);   // background pixels tie at the median: one atomic per warp, not per pixel
    if ((threadIdx.x & 31) == 0 && first != 0x7fffffff) atomicMin(med_idx, first);
    const double t0 = block_sum<kRed>((double)a, scratch), t1 = block_sum<kRed>((double)sg, scratch),
                 t2 = block_sum<kRed>((double)b, scratch);
    if (threadIdx.x == 0) { partA[3 * blockIdx.x] = t0; partA[3 * blockIdx.x + 1] = t1; partA[3 * blockIdx.x + 2] = t2; }
}

// B: residuals r = (p - t_p)/s_p - (g - t_g)/s_g: sum r^2, sum r, sum r * (p - t_p)/s_p
__global__ void __launch_bounds__(kRed)
depth_resid_kernel(int n, const float *__restrict__ pred, const float *__restrict__ gt, const float *__restrict__ med,
                   const double *__restrict__ partA, double *__restrict__ partB) {
    __shared__ double scratch[kRed / 32], bcast[3];
    double A[3];
    total_of_partials<kRed, 3>(partA, gridDim.x, A, scratch, bcast);
    const float tp = med[0], tg = med[1];
    const float sp = (float)(A[0] / n), sg = (float)(A[2] / n);
    float r2 = 0.f, r1 = 0.f, rd = 0.f;
    for (int i = blockIdx.x * kRed + threadIdx.x; i < n; i += gridDim.x * kRed) {
        const float dn = (pred[i] - tp) / sp, gn = (gt[i] - tg) / sg, r = dn - gn;
        r2 += r * r; r1 += r; rd += r * dn;
    }
    const double t0 = block_sum<kRed>((double)r2, scratch), t1 = block_sum<kRed>((double)r1, scratch),
                 t2 = block_sum<kRed>((double)rd, scratch);
    if (threadIdx.x == 0) { partB[3 * blockIdx.x] = t0; partB[3 * blockIdx.x + 1] = t1; partB[3 * blockIdx.x + 2] = t2; }
}

// C: loss and gradient.  With a_j = 2 r_j / N, S_a = sum a_j, S_ad = sum a_j dn_j, m = median index:
//   dL/dp_i = a_i/s - [i==m] S_a/s - (S_ad/s) * (1/N) (sign(p_i - t) - [i==m] sum_k sign(p_k - t))
__global__ void __launch_bounds__(kRed)
depth_grad_kernel(int n, const float *__restrict__ pred, const float *__restrict__ gt, const float *__restrict__ med,
                  const double *__restrict__ partA, const double *__restrict__ partB,
                  const int *__restrict__ med_idx, float weight, float *__restrict__ loss, float *__restrict__ dL_dpred) {
    __shared__ double scratch[kRed / 32], bcast[3];
    double A[3], B[3];
    total_of_partials<kRed, 3>(partA, gridDim.x, A, scratch, bcast);
    total_of_partials<kRed, 3>(partB, gridDim.x, B, scratch, bcast);
    const float tp = med[0], tg = med[1];
    const float sp = (float)(A[0] / n), sg = (float)(A[2] / n);
    const float invn = 1.f / (float)n;
    const float Sa = (float)(2.0 * B[1] / n), Sad = (float)(2.0 * B[2] / n), Ssgn = (float)A[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] = weight * (float)(B[0] / n);
    if (dL_dpred == nullptr) return;
    const int m = *med_idx;
    for (int i = blockIdx.x * kRed + threadIdx.x; i < n; i += gridDim.x * kRed) {
        const float p = pred[i], dn = (p - tp) / sp, gn = (gt[i] - tg) / sg;
        const float a = 2.f * (dn - gn) * invn;
        float g = a / sp - (Sad / sp) * (sgnf(p - tp) * invn);
        if (i == m) g += -Sa / sp + (Sad / sp) * (Ssgn * invn);
        dL_dpred[i] = weight * g;
    }
}

// ---- the whole depth loss as ONE kernel ------------------------------------------------------------------------------------
// The six kernels above are a chain of global reductions (3 select passes -> MAD -> residual sums -> gradient), each a few
// microseconds of work behind a launch boundary: 66 us on the step's critical path at 854x480.  Here one grid of <= one CTA per
// SM walks the same phases and meets at a device-side barrier between them (a monotone arrival counter in global memory; all
// CTAs are co-resident -- the grid never exceeds the SM count and a CTA needs 17 KB of shared memory and 512 threads).  Both
// maps (3.3 MB) stay in L2 between the phases.  Same arithmetic per element as the kernels above; every CTA derives the select
// result and the totals from the shared histograms / per-CTA partials itself, in the same fixed order.
constexpr int kDepthThreads = 512;
struct DepthCtl { unsigned arrived; unsigned bail; unsigned med_pos; unsigned pad; };   // med_pos = n - (first index with p == median), 0 = none yet

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the grid: arrive, then wait until `target` arrivals have been counted since the launch's memset
__device__ __forceinline__ void grid_barrier(DepthCtl *ctl, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                               // this CTA's histogram atomics / partials before its arrival
        atomicAdd(&ctl->arrived, 1u);
        unsigned spins = 0;
        while (ld_acquire_gpu(&ctl->arrived) < target) {
            if (++spins > (1u << 24)) { ctl->bail = 1u; break; }   // ~0.1 s of polling: give up (the loss comes out as NaN) instead of hanging
        }
    }
    __syncthreads();
}

// one select pass over both maps: digit histograms of the elements matching the prefixes found so far, merged into the pass's
// global histograms; after the barrier EVERY CTA walks the 2 x kDigits bins and fixes (prefix, rank) of both maps.
// cp / cg: this thread's elements (index start + q * stride), loaded once by the caller -- the passes never touch memory again.
constexpr int kDepthPerThread = 8;
template <int SHIFT, int BITS>
__device__ __forceinline__ void depth_select_pass(int n, int start, int stride, const float (&cp)[kDepthPerThread],
                                                  const float (&cg)[kDepthPerThread],
                                                  unsigned *__restrict__ ghist /*[2][kSelBins] of this pass*/, DepthCtl *ctl,
                                                  unsigned barrier_target, unsigned (&prefix)[2], unsigned (&rank)[2],
                                                  unsigned (*s_hist)[kSelBins], unsigned *s_scan, unsigned (*s_sel)[2]) {
    constexpr unsigned kDigits = 1u << BITS;
    constexpr unsigned hi_mask = SHIFT + BITS < 32 ? (0xffffffffu << (SHIFT + BITS)) : 0u;
    for (int i = threadIdx.x; i < 2 * kSelBins; i += kDepthThreads) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kDepthPerThread; ++q) {          // every thread runs all rounds (match_all below is warp-wide)
        const bool live = start + q * stride < n;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            unsigned digit = 0xffffffffu;
            if (live) {
                const unsigned k = order_key(m ? cg[q] : cp[q]);
                if ((k & hi_mask) == prefix[m]) digit = (k >> SHIFT) & (kDigits - 1u);
            }
            int same;
            __match_all_sync(0xffffffffu, digit, &same);     // background pixels share one value: one atomic per warp
            if (same) { if (digit != 0xffffffffu && (threadIdx.x & 31) == 0) atomicAdd(&s_hist[m][digit], 32u); }
            else if (digit != 0xffffffffu) atomicAdd(&s_hist[m][digit], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (int)kSelBins; i += kDepthThreads) {
        const unsigned c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
    grid_barrier(ctl, barrier_target);
    constexpr int per = (kDigits + kDepthThreads - 1) / kDepthThreads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        unsigned loc[per], sum = 0;
#pragma unroll
        for (int q = 0; q < per; ++q) {
            const int bin = threadIdx.x * per + q;
            loc[q] = bin < (int)kDigits ? __ldcg(&ghist[m * kSelBins + bin]) : 0u;
            sum += loc[q];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        unsigned run = incl - sum;
#pragma unroll
        for (int w = 0; w < kDepthThreads / 32; ++w) if (w < warp) run += s_scan[w];
#pragma unroll
        for (int q = 0; q < per; ++q) {
            if (rank[m] >= run && rank[m] < run + loc[q]) {     // exactly one (thread, q) over the CTA
                s_sel[m][0] = prefix[m] | ((unsigned)(threadIdx.x * per + q) << SHIFT);
                s_sel[m][1] = rank[m] - run;
            }
            run += loc[q];
        }
        __syncthreads();
        prefix[m] = s_sel[m][0]; rank[m] = s_sel[m][1];
        __syncthreads();
    }
}

// n <= gridDim.x * kDepthThreads * kDepthPerThread (the launcher falls back to the kernel chain above for larger images)
__global__ void __launch_bounds__(kDepthThreads)
depth_fused_kernel(int n, const float *__restrict__ pred, const float *__restrict__ gt, float weight, DepthCtl *__restrict__ ctl,
                   unsigned *__restrict__ hist /*[3][2][kSelBins]*/, double *__restrict__ partA, double *__restrict__ partB,
                   float *__restrict__ loss, float *__restrict__ dL_dpred) {
    __shared__ unsigned s_hist[2][kSelBins];
    __shared__ unsigned s_scan[kDepthThreads / 32];
    __shared__ unsigned s_sel[2][2];
    __shared__ double scratch[kDepthThreads / 32], bcast[3];
    const unsigned G = gridDim.x;
    const int start = blockIdx.x * kDepthThreads + threadIdx.x, stride = G * kDepthThreads;
    float cp[kDepthPerThread], cg[kDepthPerThread];
#pragma unroll
    for (int q = 0; q < kDepthPerThread; ++q) {      // 2 x 8 independent loads in flight; the maps are not read again
        const int i = start + q * stride;
        cp[q] = i < n ? pred[i] : 0.f;
        cg[q] = i < n ? gt[i] : 0.f;
    }
    // torch.median = the element of rank (n-1)/2 (the lower of the two middles)
    unsigned prefix[2] = {0u, 0u}, rank[2] = {(unsigned)((n - 1) / 2), (unsigned)((n - 1) / 2)};
    depth_select_pass<21, 11>(n, start, stride, cp, cg, hist, ctl, G, prefix, rank, s_hist, s_scan, s_sel);
    depth_select_pass<10, 11>(n, start, stride, cp, cg, hist + 2 * kSelBins, ctl, 2 * G, prefix, rank, s_hist, s_scan, s_sel);
    depth_select_pass<0, 10>(n, start, stride, cp, cg, hist + 4 * kSelBins, ctl, 3 * G, prefix, rank, s_hist, s_scan, s_sel);
    const float tp = key_value(prefix[0]), tg = key_value(prefix[1]);
    // A: sum |p - t_p|, sum sign(p - t_p), sum |g - t_g|; the (first) element that equals the median of p
    {
        float a = 0.f, sg = 0.f, b = 0.f;
        int first = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < kDepthPerThread; ++q) {
            const int i = start + q * stride;
            if (i < n) {
                const float p = cp[q], d = p - tp;
                a += fabsf(d); sg += sgnf(d); b += fabsf(cg[q] - tg);
                if (p == tp) first = min(first, i);
            }
        }
        first = __reduce_min_sync(0xffffffffu, first);
        if ((threadIdx.x & 31) == 0 && first != 0x7fffffff) atomicMax(&ctl->med_pos, (unsigned)(n - first));
        const double t0 = block_sum<kDepthThreads>((double)a, scratch), t1 = block_sum<kDepthThreads>((double)sg, scratch),
                     t2 = block_sum<kDepthThreads>((double)b, scratch);
        if (threadIdx.x == 0) { partA[3 * blockIdx.x] = t0; partA[3 * blockIdx.x + 1] = t1; partA[3 * blockIdx.x + 2] = t2; }
    }
    grid_barrier(ctl, 4 * G);
    double A[3], B[3];
    total_of_partials<kDepthThreads, 3>(partA, (int)G, A, scratch, bcast);
    const float sp = (float)(A[0] / n), sgt = (float)(A[2] / n);
    // B: residuals r = (p - t_p)/s_p - (g - t_g)/s_g: sum r^2, sum r, sum r * (p - t_p)/s_p
    {
        float r2 = 0.f, r1 = 0.f, rd = 0.f;
#pragma unroll
        for (int q = 0; q < kDepthPerThread; ++q) {
            if (start + q * stride < n) {
                const float dn = (cp[q] - tp) / sp, gn = (cg[q] - tg) / sgt, r = dn - gn;
                r2 += r * r; r1 += r; rd += r * dn;
            }
        }
        const double t0 = block_sum<kDepthThreads>((double)r2, scratch), t1 = block_sum<kDepthThreads>((double)r1, scratch),
                     t2 = block_sum<kDepthThreads>((double)rd, scratch);
        if (threadIdx.x == 0) { partB[3 * blockIdx.x] = t0; partB[3 * blockIdx.x + 1] = t1; partB[3 * blockIdx.x + 2] = t2; }
    }
    grid_barrier(ctl, 5 * G);
    total_of_partials<kDepthThreads, 3>(partB, (int)G, B, scratch, bcast);
    // C: loss and gradient (formula at depth_grad_kernel)
    const bool bailed = __ldcg(&ctl->bail) != 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] = bailed ? __int_as_float(0x7fc00000) : weight * (float)(B[0] / n);
    if (dL_dpred == nullptr) return;
    const float invn = 1.f / (float)n;
    const float Sa = (float)(2.0 * B[1] / n), Sad = (float)(2.0 * B[2] / n), Ssgn = (float)A[1];
    const unsigned mp = __ldcg(&ctl->med_pos);
    const int m = mp ? n - (int)mp : 0x7fffffff;
#pragma unroll
    for (int q = 0; q < kDepthPerThread; ++q) {
        const int i = start + q * stride;
        if (i < n) {
            const float p = cp[q], dn = (p - tp) / sp, gn = (cg[q] - tg) / sgt;
            const float a = 2.f * (dn - gn) * invn;
            float g = a / sp - (Sad / sp) * (sgnf(p - tp) * invn);
            if (i == m) g += -Sa / sp + (Sad / sp) * (Ssgn * invn);
            dL_dpred[i] = weight * g;
        }
    }
}

// ------------------------------------------------------------------------------------------------ track: trimmed masked L1
// per point: mean_c |denormalised rendered track - target| (invisible points: +inf so they sort last); count of visible
__global__ void __launch_bounds__(kRed)
track_point_kernel(int n, int W, int H, const float *__restrict__ track, const int *__restrict__ query_xy,
                   const float *__restrict__ gt_xy, const uint8_t *__restrict__ visible, float *__restrict__ vals,
                   int *__restrict__ n_visible) {
    const int i = blockIdx.x * kRed + threadIdx.x;
    if (i >= n) return;
    float v = __int_as_float(0x7f800000);
    const int qx = query_xy[2 * i], qy = query_xy[2 * i + 1];
    if (visible[i] && qx >= 0 && qx < W && qy >= 0 && qy < H) {
        const size_t o = (size_t)qy * W + qx, HW = (size_t)H * W;
        const float px = (track[o] + 1.f) * (float)W / 2.f, py = (track[HW + o] + 1.f) * (float)H / 2.f;   // util.py:82
        v = (fabsf(px - gt_xy[2 * i]) + fabsf(py - gt_xy[2 * i + 1])) / 2.f;
        atomicAdd(n_visible, 1);
    }
    vals[i] = v;
}

// one CTA: quantile threshold (torch.quantile, linear interpolation, float32 rank), the two masked sums in a fixed order,
// the loss, then the scatter of the gradient onto the two coordinate planes (atomics: query pixels may repeat)
__global__ void __launch_bounds__(1024)
track_reduce_kernel(int n, int W, int H, const float *__restrict__ track, const int *__restrict__ query_xy,
                    const float *__restrict__ gt_xy, const float *__restrict__ weights, const float *__restrict__ vals,
                    const float *__restrict__ sorted, const int *__restrict__ n_visible, float quantile, float weight,
                    float *__restrict__ loss, float *__restrict__ dL_dtrack) {
    __shared__ double scratch[32], bcast[2];
    const int M = *n_visible;
    if (M <= 0) {   // trainer_fragGS.py:569-570: no visible track -> zero loss, no gradient
        if (threadIdx.x == 0) loss[0] = 0.f;
        return;
    }
    const float pos = quantile * (float)(M - 1);
    const int lo = (int)floorf(pos), hi = (int)ceilf(pos);
    const float fr = pos - (float)lo, a = sorted[lo], b = sorted[hi];
    const float thr = fr < 0.5f ? a + fr * (b - a) : b - (b - a) * (1.f - fr);   // at::lerp
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float v = vals[i];
        if (v <= thr) { num += (double)(v * weights[i]); den += (double)weights[i]; }
    }
    num = block_sum<1024>(num, scratch);
    if (threadIdx.x == 0) bcast[0] = num;
    den = block_sum<1024>(den, scratch);
    if (threadIdx.x == 0) bcast[1] = den;
    __syncthreads();
    const float numf = (float)bcast[0], denf = (float)bcast[1] + 1e-8f;     // ndim = 1 (criterion.py:49-51)
    const float hw = (float)max(H, W);
    if (threadIdx.x == 0) loss[0] = weight * (numf / denf) / hw;
    if (dL_dtrack == nullptr) return;
    const size_t HW = (size_t)H * W;
    const float k = weight / (denf * hw);
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float v = vals[i];
        if (!(v <= thr)) continue;
        const int qx = query_xy[2 * i], qy = query_xy[2 * i + 1];
        const size_t o = (size_t)qy * W + qx;
        const float px = (track[o] + 1.f) * (float)W / 2.f, py = (track[HW + o] + 1.f) * (float)H / 2.f;
        const float w = k * weights[i] * 0.5f;
        atomicAdd(&dL_dtrack[o], w * sgnf(px - gt_xy[2 * i]) * ((float)W / 2.f));
        atomicAdd(&dL_dtrack[HW + o], w * sgnf(py - gt_xy[2 * i + 1]) * ((float)H / 2.f));
    }
}

// Block-wide radix SELECT over `n` floats in shared memory: the value of rank `rank` (0-based, ascending) -- three passes over the
// order-preserving keys (11 + 11 + 10 bits), shared-memory histogram + block scan per pass.  All threads of a 1024-thread CTA
// call it; every thread gets the result.  s_hist: 2048 counters, s_aux: 36 words.
__device__ float block_select_1024(const float *vals, int n, unsigned rank, unsigned *s_hist, unsigned *s_aux) {
    unsigned prefix = 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0), bits = pass == 2 ? 10 : 11;
        const unsigned hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + bits)), dmask = (1u << bits) - 1u;
        s_hist[threadIdx.x] = 0u; s_hist[threadIdx.x + 1024] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 1024) {
            const unsigned k = order_key(vals[i]);
            if ((k & hi_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & dmask], 1u);
        }
        __syncthreads();
        // exclusive scan of the 2048 bins: two per thread, warp shuffles, then the 32 warp totals
        const unsigned c0 = s_hist[2 * threadIdx.x], c1 = s_hist[2 * threadIdx.x + 1];
        unsigned incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_aux[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_aux[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            s_aux[lane] = w;
        }
        __syncthreads();
        const unsigned before = (incl - (c0 + c1)) + (warp ? s_aux[warp - 1] : 0u);
        if (rank >= before && rank < before + c0) { s_aux[32] = prefix | ((unsigned)(2 * threadIdx.x) << shift); s_aux[33] = rank - before; }
        else if (rank >= before + c0 && rank < before + c0 + c1) { s_aux[32] = prefix | ((unsigned)(2 * threadIdx.x + 1) << shift); s_aux[33] = rank - before - c0; }
        __syncthreads();
        prefix = s_aux[32]; rank = s_aux[33];
        __syncthreads();
    }
    return key_value(prefix);
}

// Up to kTrackFused query points (the TAPIR grid of one frame): everything above in ONE CTA -- per-point values in shared memory,
// the two order statistics of the quantile by radix select (no sort), the two masked sums, the loss and the gradient scatter.
constexpr int kTrackFused = 16384;

__global__ void __launch_bounds__(1024)
track_fused_kernel(int n, int W, int H, const float *__restrict__ track, const int *__restrict__ query_xy,
                   const float *__restrict__ gt_xy, const uint8_t *__restrict__ visible, const float *__restrict__ weights,
                   float quantile, float weight, float *__restrict__ loss, float *__restrict__ dL_dtrack) {
    extern __shared__ __align__(16) float s_vals[];
    __shared__ double scratch[32], bcast[2];
    __shared__ unsigned s_hist[2048], s_aux[36];
    __shared__ int s_M;
    const float inf = __int_as_float(0x7f800000);
    const size_t HW = (size_t)H * W;
    if (threadIdx.x == 0) s_M = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        float v = inf;
        const int qx = query_xy[2 * i], qy = query_xy[2 * i + 1];
        if (visible[i] && qx >= 0 && qx < W && qy >= 0 && qy < H) {
            const size_t o = (size_t)qy * W + qx;
            const float px = (track[o] + 1.f) * (float)W / 2.f, py = (track[HW + o] + 1.f) * (float)H / 2.f;   // util.py:82
            v = (fabsf(px - gt_xy[2 * i]) + fabsf(py - gt_xy[2 * i + 1])) / 2.f;
            ++mine;
        }
        s_vals[i] = v;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_M, mine);
    __syncthreads();
    const int M = s_M;
    if (M <= 0) {   // trainer_fragGS.py:569-570: no visible track -> zero loss, no gradient
        if (threadIdx.x == 0) loss[0] = 0.f;
        return;
    }
    const float pos = quantile * (float)(M - 1);
    const int lo = (int)floorf(pos), hi = (int)ceilf(pos);
    const float fr = pos - (float)lo;
    const float a = block_select_1024(s_vals, n, (unsigned)lo, s_hist, s_aux);
    const float b = hi == lo ? a : block_select_1024(s_vals, n, (unsigned)hi, s_hist, s_aux);
    const float thr = fr < 0.5f ? a + fr * (b - a) : b - (b - a) * (1.f - fr);   // at::lerp
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float v = s_vals[i];
        if (v <= thr) { num += (double)(v * weights[i]); den += (double)weights[i]; }
    }
    num = block_sum<1024>(num, scratch);
    if (threadIdx.x == 0) bcast[0] = num;
    den = block_sum<1024>(den, scratch);
    if (threadIdx.x == 0) bcast[1] = den;
    __syncthreads();
    const float numf = (float)bcast[0], denf = (float)bcast[1] + 1e-8f;     // ndim = 1 (criterion.py:49-51)
    const float hw = (float)max(H, W);
    if (threadIdx.x == 0) loss[0] = weight * (numf / denf) / hw;
    if (dL_dtrack == nullptr) return;
    const float kk = weight / (denf * hw);
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float v = s_vals[i];
        if (!(v <= thr)) continue;
        const int qx = query_xy[2 * i], qy = query_xy[2 * i + 1];
        const size_t o = (size_t)qy * W + qx;
        const float px = (track[o] + 1.f) * (float)W / 2.f, py = (track[HW + o] + 1.f) * (float)H / 2.f;
        const float w = kk * weights[i] * 0.5f;
        atomicAdd(&dL_dtrack[o], w * sgnf(px - gt_xy[2 * i]) * ((float)W / 2.f));
        atomicAdd(&dL_dtrack[HW + o], w * sgnf(py - gt_xy[2 * i + 1]) * ((float)H / 2.f));
    }
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
inline unsigned row_blocks(int W) { return spv::cdiv(W, kRow); }

struct DepthWs { unsigned *hist; SelState *state; float *med; double *partA, *partB; int *med_idx; size_t head_bytes, total;
                 DepthCtl *ctl; unsigned *fhist; size_t fused_head_bytes; };
DepthWs carve_depth(void *base, int n) {
    (void)n;
    DepthWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = base ? (char *)base + off : nullptr; off += align256(bytes); return p; };
    // [hist | state | med | med_idx] are cleared / preset by ONE memset + one tiny copy per call
    w.hist = (unsigned *)take(sizeof(unsigned) * 2 * kSelBins);
    w.state = (SelState *)take(sizeof(SelState) * 2);
    w.med = (float *)take(sizeof(float) * 2);
    w.head_bytes = off;
    w.med_idx = (int *)take(sizeof(int));
    w.partA = (double *)take(sizeof(double) * 3 * kRedBlocks);
    w.partB = (double *)take(sizeof(double) * 3 * kRedBlocks);
    // single-kernel path: [barrier / flags | 3 passes x 2 maps of histograms] cleared by ONE memset per call
    const size_t f0 = off;
    w.ctl = (DepthCtl *)take(sizeof(DepthCtl));
    w.fhist = (unsigned *)take(sizeof(unsigned) * 6 * kSelBins);
    w.fused_head_bytes = off - f0;
    w.total = off;
    return w;
}

struct TrackWs { float *vals, *sorted; int *n_visible; void *cub; size_t cub_bytes, total; };
TrackWs carve_track(void *base, int n) {
    TrackWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = base ? (char *)base + off : nullptr; off += align256(bytes); return p; };
    w.vals = (float *)take(sizeof(float) * (size_t)n);
    w.sorted = (float *)take(sizeof(float) * (size_t)n);
    w.n_visible = (int *)take(sizeof(int));
    size_t cb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cb, (const float *)nullptr, (float *)nullptr, n);
    w.cub_bytes = cb > 0 ? cb : ((size_t)n * 16 + (1 << 20));
    w.cub = take(w.cub_bytes);
    w.total = off;
    return w;
}

}  // namespace

extern "C" {

size_t spv_loss_rgb_workspace_bytes(int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    return align256(sizeof(double) * 2 * (size_t)H) + 256;
}

int spv_loss_rgb(int W, int H, const float *pred_chw, const float *gt_hwc, float weight, float lambda_dssim, float *loss,
                 float *dL_dpred_chw, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) { spv::set_error(cudaErrorInvalidValue, "spv_loss_rgb: empty image"); return (int)cudaErrorInvalidValue; }
    if (ws_bytes < spv_loss_rgb_workspace_bytes(W, H)) { spv::set_error(cudaErrorInvalidValue, "spv_loss_rgb: workspace too small"); return (int)cudaErrorInvalidValue; }
    const size_t dyn = sizeof(float) * 15 * (size_t)(W + 2 * kHalo);
    if (dyn > 200 * 1024) { spv::set_error(cudaErrorInvalidValue, "spv_loss_rgb: image rows wider than 3400 pixels are not supported"); return (int)cudaErrorInvalidValue; }
    double *partials = (double *)workspace;
    unsigned *ticket = (unsigned *)((char *)workspace + align256(sizeof(double) * 2 * (size_t)H));
    SPV_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), s), "spv_loss_rgb");
    static std::atomic<unsigned long long> configured{0};
    spv::opt_in_dynamic_smem(rgb_row_kernel, 200 * 1024, configured);
    const double n_elems = 3.0 * (double)H * (double)W;
    // d(loss)/d(ssim_map element) = -weight * lambda / N ; d(loss)/d|p-g| = weight * (1 - lambda) / N
    rgb_row_kernel<<<H, kRgbThreads, dyn, s>>>(W, H, pred_chw, gt_hwc, (float)(-(double)weight * lambda_dssim / n_elems),
                                               (float)((double)weight * (1.0 - lambda_dssim) / n_elems), weight, lambda_dssim,
                                               dL_dpred_chw, partials, ticket, loss);
    return spv::check_launch("spv_loss_rgb", 1);
}

size_t spv_loss_depth_workspace_bytes(int n) { return n > 0 ? carve_depth(nullptr, n).total : 0; }

int spv_loss_depth_dpt(int n, const float *pred, const float *gt, float weight, float *loss, float *dL_dpred,
                       void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= 0) { spv::set_error(cudaErrorInvalidValue, "spv_loss_depth_dpt: empty image"); return (int)cudaErrorInvalidValue; }
    DepthWs w = carve_depth(workspace, n);
    if (ws_bytes < w.total) { spv::set_error(cudaErrorInvalidValue, "spv_loss_depth_dpt: workspace too small"); return (int)cudaErrorInvalidValue; }
    // one grid of one CTA per SM (co-resident: the phases meet at device-side barriers) holding the image in registers
    const unsigned sms = (unsigned)spv::sm_count();
    const unsigned G = spv::cdiv(n, kDepthThreads) < sms ? spv::cdiv(n, kDepthThreads) : sms;
    if (spv::get_option("depth_staged") == 0 && (long long)n <= (long long)G * kDepthThreads * kDepthPerThread && G <= (unsigned)kRedBlocks) {
        SPV_CUDA_TRY(cudaMemsetAsync(w.ctl, 0, w.fused_head_bytes, s), "spv_loss_depth_dpt");
        // An SM keeps its shared-memory carve-out while CTAs are resident: with the default (small) preference this kernel's CTAs,
        // one on every SM, kept the rgb loss (52 KB of dynamic shared memory per CTA) off the GPU until they were done
        // (measured: rgb_row_kernel started 48 us late).  Ask for the same maximal carve-out the other loss kernels run with.
        static std::atomic<unsigned long long> carve{0};
        spv::prefer_max_carveout(depth_fused_kernel, carve);
        depth_fused_kernel<<<G, kDepthThreads, 0, s>>>(n, pred, gt, weight, w.ctl, w.fhist, w.partA, w.partB, loss, dL_dpred);
        return spv::check_launch("spv_loss_depth_dpt", 1);
    }
    SPV_CUDA_TRY(cudaMemsetAsync(w.hist, 0, w.head_bytes, s), "spv_loss_depth_dpt");      // histograms, select state, tickets
    SPV_CUDA_TRY(cudaMemsetAsync(w.med_idx, 0x7f, sizeof(int), s), "spv_loss_depth_dpt");
    const unsigned sel_blocks = spv::cdiv(n, kRed * 4);       // measured: 296 x 2 CTAs (12 / 12 / 8 us per pass at 854x480) beat 74 x 2 (24 / 18 / 10)
    const dim3 sel_grid(sel_blocks < (unsigned)kRedBlocks ? sel_blocks : (unsigned)kRedBlocks, 2);
    select_pass_kernel<21, 11><<<sel_grid, kRed, 0, s>>>(n, pred, gt, w.hist, w.state, w.med);
    select_pass_kernel<10, 11><<<sel_grid, kRed, 0, s>>>(n, pred, gt, w.hist, w.state, w.med);
    select_pass_kernel<0, 10><<<sel_grid, kRed, 0, s>>>(n, pred, gt, w.hist, w.state, w.med);
    depth_stats_kernel<<<kRedBlocks, kRed, 0, s>>>(n, pred, gt, w.med, w.partA, w.med_idx);
    depth_resid_kernel<<<kRedBlocks, kRed, 0, s>>>(n, pred, gt, w.med, w.partA, w.partB);
    depth_grad_kernel<<<kRedBlocks, kRed, 0, s>>>(n, pred, gt, w.med, w.partA, w.partB, w.med_idx, weight, loss, dL_dpred);
    return spv::check_launch("spv_loss_depth_dpt", 6);
}

size_t spv_loss_track_workspace_bytes(int n_points) { return n_points > 0 ? carve_track(nullptr, n_points).total : 0; }

int spv_loss_track(int n_points, int W, int H, const float *track_chw, const int *query_xy, const float *target_xy,
                   const unsigned char *visible, const float *weights, float quantile, float weight, float *loss,
                   float *dL_dtrack_chw, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) { spv::set_error(cudaErrorInvalidValue, "spv_loss_track: empty image"); return (int)cudaErrorInvalidValue; }
    if (dL_dtrack_chw) SPV_CUDA_TRY(cudaMemsetAsync(dL_dtrack_chw, 0, sizeof(float) * 2 * (size_t)H * W, s), "spv_loss_track");
    if (n_points <= 0) {
        SPV_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), s), "spv_loss_track");
        return 0;
    }
    if (n_points <= kTrackFused) {
        const size_t dyn = sizeof(float) * (size_t)((n_points + 3) / 4 * 4);
        static std::atomic<unsigned long long> configured{0};
        spv::opt_in_dynamic_smem(track_fused_kernel, sizeof(float) * kTrackFused, configured);
        track_fused_kernel<<<1, 1024, dyn, s>>>(n_points, W, H, track_chw, query_xy, target_xy, visible, weights, quantile, weight, loss,
                                                dL_dtrack_chw);
        return spv::check_launch("spv_loss_track", 1);
    }
    TrackWs w = carve_track(workspace, n_points);
    if (ws_bytes < w.total) { spv::set_error(cudaErrorInvalidValue, "spv_loss_track: workspace too small"); return (int)cudaErrorInvalidValue; }
    SPV_CUDA_TRY(cudaMemsetAsync(w.n_visible, 0, sizeof(int), s), "spv_loss_track");
    track_point_kernel<<<spv::cdiv(n_points, kRed), kRed, 0, s>>>(n_points, W, H, track_chw, query_xy, target_xy, visible, w.vals, w.n_visible);
    SPV_CUDA_TRY(cub::DeviceRadixSort::SortKeys(w.cub, w.cub_bytes, w.vals, w.sorted, n_points, 0, 32, s), "spv_loss_track/sort");
    track_reduce_kernel<<<1, 1024, 0, s>>>(n_points, W, H, track_chw, query_xy, target_xy, weights, w.vals, w.sorted, w.n_visible,
                                           quantile, weight, loss, dL_dtrack_chw);
    return spv::check_launch("spv_loss_track", 2 + 4);
}

}  // extern "C"
