// Fused Adam over the flat per-Gaussian parameter buffer (optimizer half): one streaming pass
// over (param, grad, exp_avg, exp_avg_sq) with per-segment learning rates, instead of torch.optim.Adam's per-tensor
// kernel sequence (the reference builds one param group per attribute, src/pointrix/optimizer/__init__.py:27-62).
// Arithmetic follows torch.optim.Adam (no amsgrad, no weight decay): m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
// p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps).  Pure HBM stream: 16 B read + 12 B written per element.
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {
constexpr int kThreads = 256;
constexpr int kMaxSeg = 16;
struct Segs { long long end[kMaxSeg]; float lr[kMaxSeg]; int n; };

// Device-resident optimizer clock for the graph-capturable entry: state = [step, 1 - b1^step, sqrt(1 - b2^step), unused].
__global__ void adam_tick_kernel(float *__restrict__ state, float b1, float b2) {
    const double t = (double)state[0] + 1.0;
    state[0] = (float)t;
    state[1] = (float)(1.0 - pow((double)b1, t));
    state[2] = (float)sqrt(1.0 - pow((double)b2, t));
}

__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, float lr_over_bc1, float b1, float b2, float eps,
                                            float bc2_sqrt) {
    m = m + (g - m) * (1.f - b1);                 // lerp_
    v = v * b2 + (1.f - b2) * g * g;              // mul_().addcmul_()
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - lr_over_bc1 * (m / denom);            // addcdiv_(value=-step_size)
}

// One float4 of (param, grad, exp_avg, exp_avg_sq) per thread: 16 B read + 12 B written per element, nothing else.  The
// per-segment learning rate is looked up ONCE per float4 (segment boundaries almost never fall inside one); streaming loads /
// stores keep the 1 GB pass out of the way of the L2-resident Gaussian records.
__global__ void __launch_bounds__(kThreads)
adam_kernel(long long n4, long long n, float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
            float4 *__restrict__ v, Segs segs, float b1, float b2, float eps, float bc1, float bc2_sqrt,
            const float *__restrict__ state, const float *__restrict__ lr_dev) {
    const long long k = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (k >= n4) return;
    if (state) { bc1 = state[1]; bc2_sqrt = state[2]; }   // device clock (graph replay): advanced by adam_tick_kernel
    float4 P = __ldcs(p + k), M = __ldcs(m + k), V = __ldcs(v + k);
    const float4 G = __ldcs(g + k);
    const long long e0 = 4 * k;
    int s0 = segs.n - 1;
#pragma unroll
    for (int s = kMaxSeg - 2; s >= 0; --s)
        if (s < segs.n && e0 < segs.end[s]) s0 = s;
    const bool one = e0 + 3 < segs.end[s0] && e0 + 3 < n;   // the whole float4 lies in segment s0
    if (one) {
        const float lr = (lr_dev ? lr_dev[s0] : segs.lr[s0]) / bc1;
        adam_update(P.x, G.x, M.x, V.x, lr, b1, b2, eps, bc2_sqrt);
        adam_update(P.y, G.y, M.y, V.y, lr, b1, b2, eps, bc2_sqrt);
        adam_update(P.z, G.z, M.z, V.z, lr, b1, b2, eps, bc2_sqrt);
        adam_update(P.w, G.w, M.w, V.w, lr, b1, b2, eps, bc2_sqrt);
    } else {
        float *pp = &P.x, *mm = &M.x, *vv = &V.x;
        const float *gg = &G.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long e = e0 + c;
            if (e >= n) break;
            int sc = s0;
            while (sc + 1 < segs.n && e >= segs.end[sc]) ++sc;
            const float lr = (lr_dev ? lr_dev[sc] : segs.lr[sc]) / bc1;
            adam_update(pp[c], gg[c], mm[c], vv[c], lr, b1, b2, eps, bc2_sqrt);
        }
    }
    __stcs(p + k, P); __stcs(m + k, M); __stcs(v + k, V);
}
}  // namespace

extern "C" {
/* n elements (buffers padded to a multiple of 4 floats and 16-byte aligned); segment s covers [end[s-1], end[s]) with
 * learning rate lr[s]; step is the 1-based iteration used for the bias corrections (computed on the host in double). */
int spv_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                  const long long *seg_end_host, const float *seg_lr_host, float beta1, float beta2, float eps, int step,
                  void *stream) {
    if (n <= 0) return 0;
    if (nseg < 1 || nseg > kMaxSeg) { spv::set_error(cudaErrorInvalidValue, "spv_adam_step: 1..16 segments"); return (int)cudaErrorInvalidValue; }
    Segs segs;
    for (int i = 0; i < kMaxSeg; ++i) { segs.end[i] = i < nseg ? seg_end_host[i] : n; segs.lr[i] = i < nseg ? seg_lr_host[i] : 0.f; }
    segs.n = nseg;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const long long n4 = (n + 3) / 4;
    adam_kernel<<<spv::cdiv(n4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n4, n, (float4 *)param, (const float4 *)grad, (float4 *)exp_avg, (float4 *)exp_avg_sq, segs, beta1, beta2, eps,
        (float)bc1, (float)sqrt(bc2), nullptr, nullptr);
    return spv::check_launch("spv_adam_step");
}

/* The same update with the optimizer clock and the learning rates in DEVICE memory, so a captured CUDA graph advances the bias
 * corrections on every replay: state_dev = float[4] {step, 1 - b1^step, sqrt(1 - b2^step), -} (zero-initialised once; every call
 * first advances it by one step), lr_dev = float[nseg] (a scheduler updates it with a plain copy outside the graph). */
int spv_adam_step_device(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                         const long long *seg_end_host, const float *lr_dev, float beta1, float beta2, float eps,
                         float *state_dev, void *stream) {
    if (n <= 0) return 0;
    if (nseg < 1 || nseg > kMaxSeg || !lr_dev || !state_dev) { spv::set_error(cudaErrorInvalidValue, "spv_adam_step_device: 1..16 segments, device lr and state"); return (int)cudaErrorInvalidValue; }
    Segs segs;
    for (int i = 0; i < kMaxSeg; ++i) { segs.end[i] = i < nseg ? seg_end_host[i] : n; segs.lr[i] = 0.f; }
    segs.n = nseg;
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev, beta1, beta2);
    const long long n4 = (n + 3) / 4;
    adam_kernel<<<spv::cdiv(n4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n4, n, (float4 *)param, (const float4 *)grad, (float4 *)exp_avg, (float4 *)exp_avg_sq, segs, beta1, beta2, eps, 1.f, 1.f,
        state_dev, lr_dev);
    return spv::check_launch("spv_adam_step_device", 2);
}
}  // extern "C"
