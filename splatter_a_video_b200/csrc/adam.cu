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

// Device-resident optimizer clock for the graph-capturable entries: state = float[8], zero-initialised once:
// [step, 1 - b1^step, sqrt(1 - b2^step), unused | b1^step and b2^step as two doubles].
// `ring` (optional, kRing x 2 floats): the step size lr / (1 - b1^t) of the lazily updated segment and sqrt(1 - b2^t) of the last
// kRing steps, so a catch-up over missed steps replays exactly the constants the dense kernel used.
constexpr int kRing = 4096;
__global__ void adam_tick_kernel(float *__restrict__ state, float b1, float b2, const float *__restrict__ lr_dev, int lazy_seg,
                                 float2 *__restrict__ ring) {
    // state[4..7] hold b1^t and b2^t as two doubles, advanced by one multiplication per step (a double pow() here would put
    // ~5 us of single-thread latency on the step's critical path)
    double *pw = reinterpret_cast<double *>(state + 4);
    const float t0 = state[0];
    double p1 = t0 == 0.f ? 1.0 : pw[0], p2 = t0 == 0.f ? 1.0 : pw[1];
    p1 *= (double)b1; p2 *= (double)b2;
    pw[0] = p1; pw[1] = p2;
    const float t = t0 + 1.f;
    const float bc1 = (float)(1.0 - p1), bc2s = (float)sqrt(1.0 - p2);
    state[0] = t;
    state[1] = bc1;
    state[2] = bc2s;
    if (ring) ring[(long long)t % kRing] = make_float2(lr_dev[lazy_seg] / bc1, bc2s);
}

// Explicitly rounded operations (no compiler-chosen contraction): the dense kernel and the lazy catch-up below must produce the
// same bits for the same (p, g, m, v) whatever code surrounds the call.  Square root and the two divisions use the SFU
// approximations (sqrt.approx, rcp-based division: <= 2 ulp, deterministic): the update is accurate to ~1e-6 of ITS OWN size
// (i.e. ~1e-6 * lr on the parameter; tests hold the optimizer to torch.optim.Adam at rtol 2e-6), and a replayed zero-gradient
// step costs ~12 instructions instead of ~45 with IEEE division / square root.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, float lr_over_bc1, float b1, float b2, float eps,
                                            float bc2_sqrt) {
    m = __fmaf_rn(__fsub_rn(g, m), __fsub_rn(1.f, b1), m);                               // lerp_(g, 1 - b1)
    v = __fmaf_rn(__fmul_rn(__fsub_rn(1.f, b2), g), g, __fmul_rn(v, b2));                // mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = __fadd_rn(__fdividef(sqrt_approx(v), bc2_sqrt), eps);
    p = __fmaf_rn(-lr_over_bc1, __fdividef(m, denom), p);                                // addcdiv_(m, denom, value=-step_size)
}

// One float4 of (param, grad, exp_avg, exp_avg_sq) per thread: 16 B read + 12 B written per element, nothing else.  The
// per-segment learning rate is looked up ONCE per float4 (segment boundaries almost never fall inside one); streaming loads /
// stores keep the 1 GB pass out of the way of the L2-resident Gaussian records.
__global__ void __launch_bounds__(kThreads)
adam_kernel(long long n4, long long n, float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
            float4 *__restrict__ v, Segs segs, float b1, float b2, float eps, float bc1, float bc2_sqrt,
            const float *__restrict__ state, const float *__restrict__ lr_dev, long long skip_lo4, long long skip_hi4) {
    long long k = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (k >= skip_lo4) k += skip_hi4 - skip_lo4;          // [skip_lo4, skip_hi4): the lazily updated segment (float4 units)
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
// ---- interval-lazy Adam for the spline coefficients -----------------------------------------------------------------------------
// A step only produces gradient in the <= 2 intervals per rank that its frames fall into (the `dirty` list of deform.cu); every
// other interval still moves under dense Adam (m, v decay, p follows m), which is why the reference's optimizer streams all
// 4*NI*3 coefficients of every Gaussian every step (1 GB of the 1.0 GB pass at 200 k Gaussians x 50 frames; 18 of 19.7 GB at
// 2 M x 120).  Those zero-gradient updates depend on nothing but the element's own (p, m, v) and the step constants, so they can
// be REPLAYED later: an interval is brought up to date (the same adam_update calls with g = 0, the same per-step constants from
// the ring) right before it is read by a forward pass (`prepare`), before it receives gradient (`step`), or all at once
// (`flush`: before densification, checkpoints, rendering other frames).  Same arithmetic in the same order per element as the
// dense kernel -- identical parameters whenever they are observed -- at 2 x 12 instead of 4 * NI * 3 coefficients per Gaussian
// and step.  `last[b]` = optimizer step interval b is current through (global per interval: all Gaussians share the schedule).
struct LazyGeom { long long off; int P, NI, layout; };

__device__ __forceinline__ long long lazy_elem(const LazyGeom &q, int i, int b, int slot) {
    // slot = coefficient k (0..3) * 3 + xyz component
    const long long row = q.off + (long long)i * 12 * q.NI;
    return q.layout ? row + (long long)b * 12 + slot : row + (long long)(slot / 3) * (q.NI * 3) + (long long)b * 3 + (slot % 3);
}

__device__ __forceinline__ float2 step_consts(const float2 *__restrict__ ring, int j, int now, float lr, float b1, float b2) {
    if (now - j < kRing) return ring[j % kRing];
    // older than the ring (an interval idle for thousands of steps): rebuild the constants (current learning rate)
    return make_float2(lr / (float)(1.0 - pow((double)b1, (double)j)), (float)sqrt(1.0 - pow((double)b2, (double)j)));
}

// MODE 0: prepare -- bring the intervals idx1 / idx2 (device scalars) up to the current step before a forward pass reads them
// MODE 1: step    -- the dirty intervals: catch up to step now-1, then the real update of step `now` with their gradient
// MODE 2: flush   -- every interval up to the current step
// One thread per (Gaussian, float4 of the interval's 12 coefficients); the CTA's interval list (deduplicated, stale entries
// only) is built once in shared memory.  Interval-major storage: one 16-byte load / store per array and interval.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 3)      // <= 85 registers, no spills: latency-bound streaming, occupancy over unrolling
adam_lazy_kernel(LazyGeom q, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 const int *__restrict__ idx1_dev, const int *__restrict__ idx2_dev, const int *__restrict__ dirty,
                 const int *__restrict__ last, const float2 *__restrict__ ring, const float *__restrict__ state,
                 const float *__restrict__ lr_dev, int lazy_seg, float b1, float b2, float eps) {
    __shared__ int s_list[16], s_from[16], s_n;
    const int now = (int)state[0];
    if (threadIdx.x == 0) {
        int n = 0;
        auto add = [&](int b) {
            if (b < 0 || b >= q.NI) return;
            for (int w = 0; w < n; ++w) if (s_list[w] == b) return;
            const int from = last[b];
            if (MODE != 1 && from >= now) return;          // already current: nothing to read or write
            s_list[n] = b; s_from[n] = from; ++n;
        };
        if (MODE == 0) { add(idx1_dev[0]); add(idx2_dev[0]); }
        else if (MODE == 1) { const int cnt = min(dirty[0], 16); for (int u = 0; u < cnt; ++u) add(dirty[1 + u]); }
        s_n = n;
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (long long)q.P * 3) return;
    const int i = (int)(t / 3), quad = (int)(t % 3);
    const float lr = lr_dev[lazy_seg];
    const int count = MODE == 2 ? q.NI : s_n;
    const int to = MODE == 1 ? now - 1 : now;
    for (int u = 0; u < count; ++u) {
        const int b = MODE == 2 ? u : s_list[u];
        const int from = MODE == 2 ? last[b] : s_from[u];
        if (MODE == 2 && from >= to) continue;
        float P_[4], M_[4], V_[4], G_[4] = {0.f, 0.f, 0.f, 0.f};
        const long long e0 = lazy_elem(q, i, b, 4 * quad);
        if (q.layout) {      // the four slots are 16 contiguous, aligned bytes
            const float4 a = *reinterpret_cast<const float4 *>(p + e0), bq = *reinterpret_cast<const float4 *>(m + e0),
                         cq = *reinterpret_cast<const float4 *>(v + e0);
            P_[0] = a.x; P_[1] = a.y; P_[2] = a.z; P_[3] = a.w; M_[0] = bq.x; M_[1] = bq.y; M_[2] = bq.z; M_[3] = bq.w;
            V_[0] = cq.x; V_[1] = cq.y; V_[2] = cq.z; V_[3] = cq.w;
            if (MODE == 1) { const float4 d = *reinterpret_cast<const float4 *>(g + e0); G_[0] = d.x; G_[1] = d.y; G_[2] = d.z; G_[3] = d.w; }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long e = lazy_elem(q, i, b, 4 * quad + c);
                P_[c] = p[e]; M_[c] = m[e]; V_[c] = v[e];
                if (MODE == 1) G_[c] = g[e];
            }
        }
        for (int j = from + 1; j <= to; ++j) {          // replay the zero-gradient steps (constants fetched once per step)
            const float2 cs = step_consts(ring, j, now, lr, b1, b2);
#pragma unroll
            for (int c = 0; c < 4; ++c) adam_update(P_[c], 0.f, M_[c], V_[c], cs.x, b1, b2, eps, cs.y);
        }
        if (MODE == 1) {
            const float2 cs = ring[now % kRing];
#pragma unroll
            for (int c = 0; c < 4; ++c) adam_update(P_[c], G_[c], M_[c], V_[c], cs.x, b1, b2, eps, cs.y);
        }
        if (q.layout) {
            *reinterpret_cast<float4 *>(p + e0) = make_float4(P_[0], P_[1], P_[2], P_[3]);
            *reinterpret_cast<float4 *>(m + e0) = make_float4(M_[0], M_[1], M_[2], M_[3]);
            *reinterpret_cast<float4 *>(v + e0) = make_float4(V_[0], V_[1], V_[2], V_[3]);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) { const long long e = lazy_elem(q, i, b, 4 * quad + c); p[e] = P_[c]; m[e] = M_[c]; v[e] = V_[c]; }
        }
    }
}

template <int MODE>
__global__ void adam_lazy_mark_kernel(int NI, const int *__restrict__ idx1_dev, const int *__restrict__ idx2_dev,
                                      const int *__restrict__ dirty, int *__restrict__ last, const float *__restrict__ state) {
    const int now = (int)state[0];
    if (MODE == 0) {
        if (threadIdx.x == 0) {
            const int a = idx1_dev[0], b = idx2_dev[0];
            if (a >= 0 && a < NI) last[a] = now;
            if (b >= 0 && b < NI) last[b] = now;
        }
    } else if (MODE == 1) {
        const int cnt = min(dirty[0], 16);
        if ((int)threadIdx.x < cnt) { const int b = dirty[1 + threadIdx.x]; if (b >= 0 && b < NI) last[b] = now; }
    } else {
        for (int b = threadIdx.x; b < NI; b += blockDim.x) last[b] = now;
    }
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
        (float)bc1, (float)sqrt(bc2), nullptr, nullptr, n4, n4);
    return spv::check_launch("spv_adam_step");
}

/* The same update with the optimizer clock and the learning rates in DEVICE memory, so a captured CUDA graph advances the bias
 * corrections on every replay: state_dev = float[8] {step, 1 - b1^step, sqrt(1 - b2^step), -, b1^step, b2^step as doubles} (zero-initialised once; every call
 * first advances it by one step), lr_dev = float[nseg] (a scheduler updates it with a plain copy outside the graph). */
int spv_adam_step_device(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                         const long long *seg_end_host, const float *lr_dev, float beta1, float beta2, float eps,
                         float *state_dev, void *stream) {
    if (n <= 0) return 0;
    if (nseg < 1 || nseg > kMaxSeg || !lr_dev || !state_dev) { spv::set_error(cudaErrorInvalidValue, "spv_adam_step_device: 1..16 segments, device lr and state"); return (int)cudaErrorInvalidValue; }
    Segs segs;
    for (int i = 0; i < kMaxSeg; ++i) { segs.end[i] = i < nseg ? seg_end_host[i] : n; segs.lr[i] = 0.f; }
    segs.n = nseg;
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state_dev, beta1, beta2, lr_dev, 0, nullptr);
    const long long n4 = (n + 3) / 4;
    adam_kernel<<<spv::cdiv(n4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n4, n, (float4 *)param, (const float4 *)grad, (float4 *)exp_avg, (float4 *)exp_avg_sq, segs, beta1, beta2, eps, 1.f, 1.f,
        state_dev, lr_dev, n4, n4);
    return spv::check_launch("spv_adam_step_device", 2);
}

/* ---- interval-lazy variant for ONE segment of spline coefficients ([P, 4*NI*3], layout as in spv_deform_spline_*) ----------------
 * lazy_seg: index of that segment; last_dev = int[NI] (zero-initialised), ring_dev = float[2 * 4096] (zero-initialised).
 * Identical parameters to spv_adam_step_device whenever an interval is observed after `prepare` / `flush` (see adam.cu). */
static int lazy_geom(LazyGeom &q, long long n, int nseg, const long long *seg_end_host, int lazy_seg, int P, int NI, int layout, const char *where) {
    if (nseg < 1 || nseg > kMaxSeg || lazy_seg < 0 || lazy_seg >= nseg || P <= 0 || NI <= 0) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
    q.off = lazy_seg ? seg_end_host[lazy_seg - 1] : 0;
    q.P = P; q.NI = NI; q.layout = layout;
    const long long hi = lazy_seg == nseg - 1 ? n : seg_end_host[lazy_seg];
    if (hi - q.off < (long long)P * 12 * NI || (q.off & 3) || ((q.off + (long long)P * 12 * NI) & 3)) {
        spv::set_error(cudaErrorInvalidValue, where);      // the segment must hold [P, 4*NI*3] and start / end on float4 boundaries
        return (int)cudaErrorInvalidValue;
    }
    return 0;
}

int spv_adam_lazy_prepare(long long n, int nseg, const long long *seg_end_host, int lazy_seg, int P, int NI, int layout, float *param,
                          float *exp_avg, float *exp_avg_sq, const int *idx1_dev, const int *idx2_dev, int *last_dev, const float *ring_dev,
                          const float *state_dev, const float *lr_dev, float beta1, float beta2, float eps, void *stream) {
    LazyGeom q;
    if (int rc = lazy_geom(q, n, nseg, seg_end_host, lazy_seg, P, NI, layout, "spv_adam_lazy_prepare: bad segment")) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    adam_lazy_kernel<0><<<spv::cdiv((long long)P * 3, kThreads), kThreads, 0, s>>>(q, param, nullptr, exp_avg, exp_avg_sq, idx1_dev, idx2_dev,
                                                                                  nullptr, last_dev, (const float2 *)ring_dev, state_dev, lr_dev,
                                                                                  lazy_seg, beta1, beta2, eps);
    adam_lazy_mark_kernel<0><<<1, 32, 0, s>>>(NI, idx1_dev, idx2_dev, nullptr, last_dev, state_dev);
    return spv::check_launch("spv_adam_lazy_prepare", 2);
}

int spv_adam_lazy_flush(long long n, int nseg, const long long *seg_end_host, int lazy_seg, int P, int NI, int layout, float *param,
                        float *exp_avg, float *exp_avg_sq, int *last_dev, const float *ring_dev, const float *state_dev,
                        const float *lr_dev, float beta1, float beta2, float eps, void *stream) {
    LazyGeom q;
    if (int rc = lazy_geom(q, n, nseg, seg_end_host, lazy_seg, P, NI, layout, "spv_adam_lazy_flush: bad segment")) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    adam_lazy_kernel<2><<<spv::cdiv((long long)P * 3, kThreads), kThreads, 0, s>>>(q, param, nullptr, exp_avg, exp_avg_sq, nullptr, nullptr,
                                                                                  nullptr, last_dev, (const float2 *)ring_dev, state_dev, lr_dev,
                                                                                  lazy_seg, beta1, beta2, eps);
    adam_lazy_mark_kernel<2><<<1, 256, 0, s>>>(NI, nullptr, nullptr, nullptr, last_dev, state_dev);
    return spv::check_launch("spv_adam_lazy_flush", 2);
}

/* One optimizer step: clock tick (+ ring entry), the dense kernel over everything except the lazy segment, the lazy segment's
 * dirty intervals (dirty_dev = int[17] of deform.cu: [0] = count, then interval indices holding this step's gradient). */
int spv_adam_step_lazy(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int nseg,
                       const long long *seg_end_host, const float *lr_dev, float beta1, float beta2, float eps, float *state_dev,
                       int lazy_seg, int P, int NI, int layout, const int *dirty_dev, int *last_dev, float *ring_dev, void *stream) {
    if (n <= 0) return 0;
    if (!lr_dev || !state_dev || !dirty_dev || !last_dev || !ring_dev) { spv::set_error(cudaErrorInvalidValue, "spv_adam_step_lazy: device buffers"); return (int)cudaErrorInvalidValue; }
    LazyGeom q;
    if (int rc = lazy_geom(q, n, nseg, seg_end_host, lazy_seg, P, NI, layout, "spv_adam_step_lazy: bad segment")) return rc;
    Segs segs;
    for (int i = 0; i < kMaxSeg; ++i) { segs.end[i] = i < nseg ? seg_end_host[i] : n; segs.lr[i] = 0.f; }
    segs.n = nseg;
    cudaStream_t s = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, s>>>(state_dev, beta1, beta2, lr_dev, lazy_seg, (float2 *)ring_dev);
    const long long n4 = (n + 3) / 4, lo4 = q.off / 4, hi4 = (q.off + (long long)P * 12 * NI) / 4;
    const long long dense4 = n4 - (hi4 - lo4);
    // the dense parameters on the side stream, the spline intervals on the caller's: two independent streaming passes
    spv::SideLane *lane = dense4 > 0 ? spv::side_lane() : nullptr;
    if (dense4 > 0) {
        cudaStream_t ds = s;
        if (lane) {
            SPV_CUDA_TRY(cudaEventRecord(lane->fork, s), "spv_adam_step_lazy/fork");
            SPV_CUDA_TRY(cudaStreamWaitEvent(lane->stream, lane->fork, 0), "spv_adam_step_lazy/fork");
            ds = lane->stream;
        }
        adam_kernel<<<spv::cdiv(dense4, kThreads), kThreads, 0, ds>>>(n4, n, (float4 *)param, (const float4 *)grad, (float4 *)exp_avg,
                                                                     (float4 *)exp_avg_sq, segs, beta1, beta2, eps, 1.f, 1.f, state_dev, lr_dev,
                                                                     lo4, hi4);
        if (lane) SPV_CUDA_TRY(cudaEventRecord(lane->join, lane->stream), "spv_adam_step_lazy/join");
    }
    adam_lazy_kernel<1><<<spv::cdiv((long long)P * 3, kThreads), kThreads, 0, s>>>(q, param, grad, exp_avg, exp_avg_sq, nullptr, nullptr, dirty_dev,
                                                                                  last_dev, (const float2 *)ring_dev, state_dev, lr_dev, lazy_seg,
                                                                                  beta1, beta2, eps);
    adam_lazy_mark_kernel<1><<<1, 32, 0, s>>>(NI, nullptr, nullptr, dirty_dev, last_dev, state_dev);
    if (lane) SPV_CUDA_TRY(cudaStreamWaitEvent(s, lane->join, 0), "spv_adam_step_lazy/join");
    return spv::check_launch("spv_adam_step_lazy", 4);
}
}  // extern "C"
