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

__global__ void __launch_bounds__(kThreads)
adam_kernel(long long n4, long long n, float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
            float4 *__restrict__ v, Segs segs, float b1, float b2, float eps, float bc1, float bc2_sqrt,
            const float *__restrict__ state, const float *__restrict__ lr_dev) {
    const long long k = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (k >= n4) return;
    if (state) { bc1 = state[1]; bc2_sqrt = state[2]; }   // device clock (graph replay): advanced by adam_tick_kernel
    float4 P = p[k], M = m[k], V = v[k];
    const float4 G = g[k];
    float *pp = &P.x, *mm = &M.x, *vv = &V.x;
    const float *gg = &G.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const long long e = 4 * k + c;
        if (e >= n) break;
        float lr = lr_dev ? lr_dev[segs.n - 1] : segs.lr[segs.n - 1];
#pragma unroll
        for (int s = kMaxSeg - 1; s >= 0; --s)
            if (s < segs.n && e < segs.end[s]) lr = lr_dev ? lr_dev[s] : segs.lr[s];
        mm[c] = mm[c] + (gg[c] - mm[c]) * (1.f - b1);                 // lerp_
        vv[c] = vv[c] * b2 + (1.f - b2) * gg[c] * gg[c];              // mul_().addcmul_()
        const float denom = sqrtf(vv[c]) / bc2_sqrt + eps;
        pp[c] = pp[c] - (lr / bc1) * (mm[c] / denom);                 // addcdiv_(value=-step_size)
    }
    p[k] = P; m[k] = M; v[k] = V;
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
