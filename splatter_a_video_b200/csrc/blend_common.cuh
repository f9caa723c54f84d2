// Device helpers shared by the blend kernels (blend.cu: separate-array staged/grouped kernels; blend_rec.cu: the
// record-staged kernels of the fused frame path).  The hit decision and alpha come from the SAME inline functions with
// explicit rounding intrinsics everywhere, so every kernel agrees on which Gaussians a pixel applied.
#pragma once
#include "common.cuh"

namespace spv_blend {

constexpr int kBlock = 256;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.99f;
constexpr float kTmin = 0.0001f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLog2AlphaMin = -7.994353436858858f;  // log2(1/255)
constexpr unsigned kFull = 0xffffffffu;

// warp w, lane l -> pixel inside the 16x16 tile: 8x4 footprint per warp, 2x4 warps per tile.
__device__ __forceinline__ void thread_pixel(int tile_x, int tile_y, int &px, int &py) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    px = tile_x * SPV_TILE + ((warp & 1) << 3) + (lane & 7);
    py = tile_y * SPV_TILE + ((warp >> 1) << 2) + (lane >> 3);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Staged splat: g0 = {x, y, a2, b2}, g1 = {c2, log2(opacity), opacity, bias} with
// a2 = -0.5*log2e*a, b2 = -log2e*b, c2 = -0.5*log2e*c  ->  p2 = log2e * power.
__device__ __forceinline__ void stage_splat(float2 xy, float a, float b, float c, float o, float bias, float4 &g0,
                                            float4 &g1) {
    g0 = make_float4(xy.x, xy.y, __fmul_rn(-0.5f * kLog2e, a), __fmul_rn(-kLog2e, b));
    g1 = make_float4(__fmul_rn(-0.5f * kLog2e, c), __log2f(o), o, bias);
}

__device__ __forceinline__ float splat_p2(const float4 g0, float c2, float pxf, float pyf, float &dx, float &dy) {
    dx = __fsub_rn(g0.x, pxf);
    dy = __fsub_rn(g0.y, pyf);
    return __fmaf_rn(g0.z, __fmul_rn(dx, dx), __fmaf_rn(c2, __fmul_rn(dy, dy), __fmul_rn(g0.w, __fmul_rn(dx, dy))));
}

// Does this pixel take this Gaussian?  (alpha_blending.cu:82-88: power <= 0 and alpha >= 1/255)
template <bool HAS_BIAS>
__device__ __forceinline__ bool splat_hits(float p2, const float4 g1) {
    if (HAS_BIAS) {
        const float alpha = fminf(kAlphaMax, __fmaf_rn(g1.z, ex2_approx(p2), g1.w));
        return p2 <= 0.f && alpha >= kAlphaMin;
    }
    return p2 <= 0.f && __fadd_rn(p2, g1.y) >= kLog2AlphaMin;
}

template <bool HAS_BIAS>
__device__ __forceinline__ float splat_alpha(float p2, const float4 g1, float &G) {
    G = ex2_approx(p2);
    return fminf(kAlphaMax, HAS_BIAS ? __fmaf_rn(g1.z, G, g1.w) : __fmul_rn(g1.z, G));
}

// Per-warp pre-filter: can ANY pixel of the warp's 8x4 block take this splat?  Same hit condition as splat_hits, evaluated
// conservatively over the block rectangle: -p2 = A dx^2 + B dx dy + C dy^2 must reach lo - log2(1/255) somewhere on it.
template <bool HAS_BIAS>
__device__ __forceinline__ bool block_may_hit(const float4 g0, const float4 g1, float bx0, float by0) {
    if (HAS_BIAS) return true;   // alpha = o*G + bias: no closed-form bound; the per-pixel test decides
    const float tau2 = g1.y - kLog2AlphaMin;
    if (!(tau2 >= 0.f)) return false;   // opacity below 1/255 (or NaN): never taken
    return spv::tile_may_hit(g0.x, g0.y, -2.f * g0.z, -g0.w, -2.f * g1.x, tau2, bx0, by0, bx0 + 7.f, by0 + 3.f);
}

// Recursive-halving multi-value warp reduction: on return lane l holds, in v[OFF], the warp-wide sum of the value with
// index OFF + (l % N).  N-1 shuffles (+ log2(32/N) butterfly steps for N < 32) instead of 5*N.
template <int N, int OFF, int TOT>
__device__ __forceinline__ void halving_reduce(float (&v)[TOT], int lane) {
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[OFF + i] : v[OFF + i + h];
            const float keep = up ? v[OFF + i + h] : v[OFF + i];
            v[OFF + i] = keep + __shfl_xor_sync(kFull, send, h);
        }
    }
    // N < 32: the sums are still split over the 32/N lane groups
#pragma unroll
    for (int o = N; o < 32; o <<= 1) v[OFF] += __shfl_xor_sync(kFull, v[OFF], o);
}

}  // namespace spv_blend
