// Per-tile front-to-back alpha compositing ("renderCUDA") forward and backward for sm_100a.
//
// Semantics restate the reference kernels
//   /root/reference/src/submodules/dptr/dptr/gs/src/alpha_blending.cu:16-249          (plain)
//   /root/reference/src/submodules/dptr/dptr/gs/src/alpha_blending_enhanced.cu:16-273 (first-K ids, truncation)
//   /root/reference/src/submodules/dptr/dptr/gs/src/alpha_blending_with_bias.cu       (per-Gaussian alpha bias)
// (skip power>0, alpha=min(0.99, o*G [+bias]), skip alpha<1/255, stop before T*(1-alpha)<1e-4; `ncontrib` = 1-based
// list position of the last applied Gaussian; bg added as T*bg; fast exp like the reference's --use_fast_math build).
// The machine mapping is new:
//
//  * one CTA per 16x16 tile, each warp owns a compact 8x4 pixel footprint;
//  * every chunk of the tile's list is staged in shared memory INCLUDING the features, read from the caller's [P,C]
//    layout (the reference re-reads features from global per pixel x Gaussian and needs a [C,P] transpose pass);
//  * TWO-PHASE inner loop.  The reference's per-pixel loop is a serial dependent chain per Gaussian (LDS -> quadratic
//    form -> exp -> compare -> branch), which on B200 is latency-bound at ~27 cycles per (warp, Gaussian).  Here
//    phase 1 evaluates, for 32 Gaussians at once and with no exp, WHETHER each pixel is hit -- the quadratic form is
//    pre-scaled by log2(e) at staging time so "alpha >= 1/255" becomes "p2 + log2(opacity) >= log2(1/255)" -- and packs
//    the answers into one 32-bit mask per lane: 12 independent instructions per (pixel, Gaussian), fully pipelined.
//    One REDUX.OR gives the warp's union; phase 2 visits only the Gaussians that hit at least one pixel of the warp.
//  * backward: the reference issues C+8 global float atomics per (pixel, Gaussian) hit.  Here a warp reduces its C+8
//    partial sums with a recursive-halving shuffle network (NV-1 shuffles for NV values instead of 5*NV), parks the
//    per-warp totals in shared memory, the CTA folds its 8 warps and issues ONE coalesced atomic row per
//    (tile, Gaussian) into a packed [P,NV] buffer that a streaming pass unpacks into the reference's output tensors.
//    The per-channel `accum_rec/last_feature` recurrences collapse into one scalar recurrence on
//    S = <colour behind, dL_dpixel>.
//  * the hit decision and alpha are produced by the same inline functions with explicit rounding intrinsics in the
//    forward and the backward kernel, so both always agree on which Gaussians were applied.
#include <stdlib.h>

#include "blend_common.cuh"
#include "../../include/spv_b200.h"

namespace {
using namespace spv_blend;

// ------------------------------------------------------------------------------------------------ forward
template <int CH, bool HAS_IDX, bool HAS_BIAS>
__global__ void __launch_bounds__(kBlock)
blend_fwd_kernel(int C, int Cstride, int c0, int W, int H, int gx, int K, int trunc,
                 const float2 *__restrict__ uv, const float *__restrict__ conic, const float *__restrict__ opacity,
                 const float *__restrict__ feature, const float *__restrict__ bias,
                 const int *__restrict__ idx_sorted, const int2 *__restrict__ tile_range, float bg, float bgB, float bgC,
                 int cA, int cB, float *__restrict__ rendered, float *__restrict__ final_T, int *__restrict__ ncontrib,
                 int *__restrict__ gs_idx) {
    __shared__ float4 s_g0[kBlock];
    __shared__ float4 s_g1[kBlock];
    __shared__ int s_id[kBlock];
    __shared__ __align__(16) float s_feat[kBlock * CH];

    const int tile = blockIdx.x;
    const int tile_x = tile % gx, tile_y = tile / gx;
    int px, py;
    thread_pixel(tile_x, tile_y, px, py);
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bx0 = (float)(tile_x * SPV_TILE + ((warp & 1) << 3)), by0 = (float)(tile_y * SPV_TILE + ((warp >> 1) << 2));

    const int2 range = tile_range[tile];
    const int n = range.y - range.x;

    float T = 1.0f;
    float F[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) F[c] = 0.f;
    int last = 0, layer = 0;
    bool done = !inside;

    for (int base = 0; base < n; base += kBlock) {
        if (__syncthreads_count(done) == kBlock) break;  // also fences reuse of the staging buffers
        const int m = min(kBlock, n - base);
        if ((int)threadIdx.x < m) {
            const int id = idx_sorted[range.x + base + threadIdx.x];
            s_id[threadIdx.x] = id;
            float4 g0, g1;
            stage_splat(uv[id], conic[3 * id], conic[3 * id + 1], conic[3 * id + 2], opacity[id],
                        HAS_BIAS ? bias[id] : 0.f, g0, g1);
            s_g0[threadIdx.x] = g0;
            s_g1[threadIdx.x] = g1;
        }
        __syncthreads();
        // feature rows: one coalesced row read per Gaussian, 32 Gaussians per warp
        if (lane < C) {
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj) {
                const int j = warp * 32 + jj;
                if (j < m) s_feat[j * CH + lane] = feature[(size_t)s_id[j] * Cstride + c0 + lane];
            }
        }
        __syncthreads();

        for (int j0 = 0; j0 < m; j0 += 32) {
            if (__all_sync(kFull, done)) break;
            // stage 1: one list entry per lane -- can the warp's pixel block take it at all?
            bool maybe = false;
            if (j0 + lane < m) maybe = block_may_hit<HAS_BIAS>(s_g0[j0 + lane], s_g1[j0 + lane], bx0, by0);
            unsigned wm = __ballot_sync(kFull, maybe);
            // stage 2: visit the survivors in list order; every lane decides for its own pixel
            while (wm) {
                const int j = __ffs(wm) - 1;
                wm &= wm - 1;
                const float4 g0 = s_g0[j0 + j];
                const float4 g1 = s_g1[j0 + j];
                float dx, dy, G;
                const float p2 = splat_p2(g0, g1.x, pxf, pyf, dx, dy);
                const bool hit = !done && splat_hits<HAS_BIAS>(p2, g1);
                if (!__any_sync(kFull, hit)) continue;
                if (hit) {
                    const float alpha = splat_alpha<HAS_BIAS>(p2, g1, G);
                    const float next_T = T * (1.f - alpha);
                    if (next_T < kTmin) { done = true; continue; }
                    const float w = alpha * T;
                    const float *fr = s_feat + (j0 + j) * CH;
#pragma unroll
                    for (int c = 0; c < CH; ++c) F[c] = fmaf(fr[c], w, F[c]);
                    T = next_T;
                    last = base + j0 + j + 1;
                    if (HAS_IDX) {
                        if (trunc) {
                            gs_idx[pix * K + layer] = s_id[j0 + j];
                            if (++layer >= K) done = true;
                        } else if (layer < K) {
                            gs_idx[pix * K + layer] = s_id[j0 + j];
                            ++layer;
                        }
                    }
                }
            }
        }
    }

    if (inside) {
        final_T[pix] = T;
        ncontrib[pix] = last;
        const size_t HW = (size_t)H * W;
#pragma unroll
        for (int c = 0; c < CH; ++c)
            if (c < C) rendered[c * HW + pix] = F[c] + T * (c < cA ? bg : (c < cB ? bgB : bgC));
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Packed gradient row layouts.
//  plain / bias (row = NV floats): 0,1 dL_duv  2,3 dL_dabs_uv  4,5,6 dL_dconic  7 dL_dopacity  8.. dL_dfeature  NV-1 dL_dbias
//    NV = 16 (C <= 8), 32 (C <= 24) or 64 (C <= 32).  One launch covers up to 32 channels -- the reference's channel
//    chunking (alpha_blending.cu:440-576), which matters for dL_dabs_uv: |.| is taken of the per-chunk uv gradient.
//  groups (row = 36 floats, NV = 32 + 1 butterfly): the trainer's three passes over [rgb(3)|depth(1)|attrs] in one
//    traversal (dptr_ortho_enhanced.py:342-376): uv/conic from every channel, opacity from rgb+depth only (the
//    attribute pass gets opacity.detach()), 2,3 = |.| and 31,32 = value of the RGB-pass uv gradient (-> abs_ndc / ndc).
enum BwdMode { kPlain = 0, kBias = 1, kGroups = 2 };
constexpr int kRowG = spv::kPackedRowGroups;

// CG = number of leading feature channels whose dL_dfeature is wanted (the caller orders gradient-free attribute groups
// last): only 8 + CG sums go through the reduction network.
template <int NV, int CH, int MODE, int CG>
__global__ void __launch_bounds__(kBlock, NV == 64 ? 1 : ((CH > 8) ? 3 : 2))
blend_bwd_kernel(int C, int Cstride, int c0, int W, int H, int gx,
                 const float2 *__restrict__ uv, const float *__restrict__ conic, const float *__restrict__ opacity,
                 const float *__restrict__ feature, const float *__restrict__ bias,
                 const int *__restrict__ idx_sorted, const int2 *__restrict__ tile_range, float bg, float bgB, float bgC,
                 const float *__restrict__ final_T, const int *__restrict__ ncontrib,
                 const spv::ChanPlanes planes, float *__restrict__ packed) {
    constexpr bool HAS_BIAS = MODE == kBias;
    constexpr bool GROUPS = MODE == kGroups;
    static_assert(GROUPS ? (CH >= 4 && CH <= 23 && 8 + CG <= NV && NV <= 32) : (CG == CH && 8 + CH + (HAS_BIAS ? 1 : 0) <= NV),
                  "row too small");
    constexpr int kG = (NV == 64) ? 16 : 32;       // Gaussians per chunk (keeps s_part at <= 34 KB)
    constexpr int kRow = GROUPS ? kRowG : NV;       // packed row stride in global memory
    constexpr int kSP = GROUPS ? NV + 2 : NV;       // per-warp parking row (+2: the RGB-pass uv gradient of the groups mode)
    constexpr int FS = (CH + 3) & ~3;               // feature row pitch: 16-byte rows -> LDS.128 broadcasts
    // NV == 32 variants keep each pixel's dL_dpixel row in (dynamic) shared memory instead of 23 registers: 113 -> ~85
    // registers, 3 instead of 2 resident CTAs per SM.  Row pitch 20 / 28 words: conflict-free 16-byte row reads.
    constexpr bool D_SMEM = (CH > 8) && NV <= 32;
    constexpr int DS = (FS <= 20) ? 20 : 28;
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ float4 s_g0[kG];
    __shared__ float4 s_g1[kG];
    __shared__ float4 s_con[kG];                    // unscaled conic a,b,c for the gradient formulas
    __shared__ int s_id[kG];
    __shared__ __align__(16) float s_feat[kG * FS];
    __shared__ float s_part[8][kG][kSP];
    __shared__ unsigned s_hit[8];                   // per warp: which chunk entries it parked a row for
    __shared__ int s_max;

    const int tile = blockIdx.x;
    const int tile_x = tile % gx, tile_y = tile / gx;
    int px, py;
    thread_pixel(tile_x, tile_y, px, py);
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)W * py + px;
    const float pxf = (float)px, pyf = (float)py;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float bx0 = (float)(tile_x * SPV_TILE + ((warp & 1) << 3)), by0 = (float)(tile_y * SPV_TILE + ((warp >> 1) << 2));

    const int2 range = tile_range[tile];
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last_contrib = inside ? ncontrib[pix] : 0;

    float d[D_SMEM ? 1 : FS];
    float *dq = s_dyn + threadIdx.x * DS;           // this pixel's dL_dpixel row (D_SMEM)
    // <bg, dL_dpixel> per gradient group (plain: one group)
    float bgdA = 0.f, bgdB = 0.f, bgdC = 0.f;
#pragma unroll
    for (int c = 0; c < FS; ++c) {
        const float dv = (inside && c < C && planes.p[c]) ? planes.p[c][pix] : 0.f;
        if (D_SMEM) dq[c] = dv; else d[c] = dv;
        if (GROUPS && c >= 4) bgdC += dv;
        else if (GROUPS && c == 3) bgdB += dv;
        else bgdA += dv;
    }
    bgdA *= bg; bgdB *= bgB; bgdC *= bgC;

    // positions >= max(last_contrib) are skipped by every pixel of the tile: do not even stage them
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int wmax = __reduce_max_sync(kFull, last_contrib);
    if (lane == 0 && wmax > 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int n_eff = min(range.y - range.x, s_max);

    float last_alpha = 0.f, lfA = 0.f, lfB = 0.f, lfC = 0.f, SA = 0.f, SB = 0.f, SC = 0.f;

    for (int p_hi = n_eff; p_hi > 0; p_hi -= kG) {
        const int m = min(kG, p_hi);
        __syncthreads();  // previous chunk's epilogue is done with s_id / s_part
        if ((int)threadIdx.x < m) {
            const int id = idx_sorted[range.x + p_hi - 1 - threadIdx.x];  // j = 0 is the back-most entry
            s_id[threadIdx.x] = id;
            const float a = conic[3 * id], b = conic[3 * id + 1], c = conic[3 * id + 2];
            float4 g0, g1;
            stage_splat(uv[id], a, b, c, opacity[id], HAS_BIAS ? bias[id] : 0.f, g0, g1);
            s_g0[threadIdx.x] = g0;
            s_g1[threadIdx.x] = g1;
            s_con[threadIdx.x] = make_float4(a, b, c, 0.f);
        }
        __syncthreads();
        if (lane < FS) {  // padding channels (C <= c < FS) are zero-filled: they meet d[c] == 0 but must stay finite
#pragma unroll
            for (int jj = 0; jj < kG / 8; ++jj) {
                const int j = warp * (kG / 8) + jj;
                if (j < m) s_feat[j * FS + lane] = lane < C ? feature[(size_t)s_id[j] * Cstride + c0 + lane] : 0.f;
            }
        }
        __syncthreads();

        // stage 1: one chunk entry per lane -- can the warp's 8x4 pixel block have taken it at all?
        bool maybe = false;
        if (lane < m) maybe = block_may_hit<HAS_BIAS>(s_g0[lane], s_g1[lane], bx0, by0);
        const unsigned wmask = __ballot_sync(kFull, maybe);

        // phase 2: reduce the per-pixel partial gradients of every Gaussian that touched this warp.  The body is
        // branch-free: a lane that did not take the Gaussian runs it with p2 = -inf, i.e. G = alpha = w = 0, so every
        // partial sum it contributes is an exact zero and only the recurrence state needs selects.
        // only the entries that passed the block test are visited (in chunk order: the recurrence needs back-to-front)
        unsigned todo = wmask, hmask = 0u;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1u;
            float dx = 0.f, dy = 0.f;
            const float4 g0 = s_g0[j];
            const float4 g1 = s_g1[j];
            float p2 = splat_p2(g0, g1.x, pxf, pyf, dx, dy);
            // did this pixel apply the Gaussian in the forward pass?  (same test, and before the last contributor)
            const bool hit = splat_hits<HAS_BIAS>(p2, g1) && (p_hi - 1 - j) < last_contrib;
            if (!__any_sync(kFull, hit)) continue;
            hmask |= 1u << j;
            float v[NV];
            float n0 = 0.f, n1 = 0.f;
            {
                const float4 con = s_con[j];
                float Gv;
                p2 = hit ? p2 : -INFINITY;
                float alpha = splat_alpha<HAS_BIAS>(p2, g1, Gv);
                if (HAS_BIAS) alpha = hit ? alpha : 0.f;
                const float rinv = __fdividef(1.f, 1.f - alpha);
                T = T * rinv;  // transmittance in front of this Gaussian (unchanged when alpha == 0)
                const float w = alpha * T;
                const float *fr = s_feat + j * FS;
                const float tb = -T_final * rinv;
                const float om = 1.f - last_alpha;
                float da_all, da_op, da_ndc;
                // <feature, dL_dpixel> per gradient group and the per-channel feature-gradient partials, 4 channels per
                // 16-byte read of the feature row (and of the pixel's gradient row when it lives in shared memory)
                float fdA = 0.f, fdB = 0.f, fdC = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < FS / 4; ++c4) {
                    const float4 ff = *reinterpret_cast<const float4 *>(fr + 4 * c4);
                    float4 dd;
                    if (D_SMEM) dd = *reinterpret_cast<const float4 *>(dq + 4 * c4);
                    else dd = make_float4(d[4 * c4], d[4 * c4 + 1], d[4 * c4 + 2], d[4 * c4 + 3]);
                    const float fv[4] = {ff.x, ff.y, ff.z, ff.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 4 * c4 + k;
                        if (c < CH) {
                            if (GROUPS && c >= 4) fdC = fmaf(fv[k], dv[k], fdC);
                            else if (GROUPS && c == 3) fdB = fmaf(fv[k], dv[k], fdB);
                            else fdA = fmaf(fv[k], dv[k], fdA);
                            if (c < CG) v[8 + c] = w * dv[k];
                        }
                    }
                }
                if (GROUPS) {
                    const float nSA = last_alpha * lfA + om * SA;
                    const float nSB = last_alpha * lfB + om * SB;
                    const float nSC = last_alpha * lfC + om * SC;
                    da_ndc = (fdA - nSA) * T + tb * bgdA;
                    da_op = da_ndc + ((fdB - nSB) * T + tb * bgdB);
                    da_all = da_op + ((fdC - nSC) * T + tb * bgdC);
                    SA = hit ? nSA : SA; SB = hit ? nSB : SB; SC = hit ? nSC : SC;
                    lfA = hit ? fdA : lfA; lfB = hit ? fdB : lfB; lfC = hit ? fdC : lfC;
#pragma unroll
                    for (int c = 8 + CG; c < NV; ++c) v[c] = 0.f;
                } else {
                    const float nSA = last_alpha * lfA + om * SA;
                    da_all = da_op = da_ndc = (fdA - nSA) * T + tb * bgdA;
                    SA = hit ? nSA : SA;
                    lfA = hit ? fdA : lfA;
#pragma unroll
                    for (int c = 8 + CH; c < NV; ++c) v[c] = 0.f;
                }
                last_alpha = hit ? alpha : last_alpha;
                const float dL_dG = g1.z * da_all;
                const float dGx = -Gv * dx * con.x - Gv * dy * con.y;
                const float dGy = -Gv * dy * con.z - Gv * dx * con.y;
                const float g0x = dL_dG * dGx, g0y = dL_dG * dGy;
                v[0] = g0x; v[1] = g0y;
                v[4] = -0.5f * Gv * dx * dx * dL_dG;
                v[5] = -Gv * dx * dy * dL_dG;
                v[6] = -0.5f * Gv * dy * dy * dL_dG;
                v[7] = Gv * da_op;
                if (GROUPS) {
                    const float dL_dG_ndc = g1.z * da_ndc;
                    n0 = dL_dG_ndc * dGx; n1 = dL_dG_ndc * dGy;
                    v[2] = fabsf(n0); v[3] = fabsf(n1);
                } else {
                    v[2] = fabsf(g0x); v[3] = fabsf(g0y);
                    if (HAS_BIAS) v[NV - 1] = hit ? da_all : 0.f;
                }
            }
            if constexpr (NV == 64) {
                halving_reduce<32, 0, NV>(v, lane);
                halving_reduce<32, 32, NV>(v, lane);
                s_part[warp][j][lane] = v[0];
                s_part[warp][j][32 + lane] = v[32];
            } else {
                halving_reduce<NV, 0, NV>(v, lane);
                if (lane < NV) s_part[warp][j][lane] = v[0];
                if constexpr (GROUPS) {
                    // the two RGB-pass sums: one halving step (odd lanes take n1, even lanes n0), then 4 butterfly steps
                    const bool up = (lane & 1) != 0;
                    float r = (up ? n1 : n0) + __shfl_xor_sync(kFull, up ? n0 : n1, 1);
#pragma unroll
                    for (int o = 2; o <= 16; o <<= 1) r += __shfl_xor_sync(kFull, r, o);
                    if (lane < 2) s_part[warp][j][NV + lane] = r;
                }
            }
        }
        if (lane == 0) s_hit[warp] = hmask;
        __syncthreads();
        // fold the warps that parked a row and push one packed row per (tile, Gaussian)
        unsigned hm[8], hany = 0u;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) { hm[w8] = s_hit[w8]; hany |= hm[w8]; }
        if constexpr (NV >= 32) {
#pragma unroll
            for (int jj = 0; jj < kG / 8; ++jj) {
                const int j = warp * (kG / 8) + jj;
                if (j < m && ((hany >> j) & 1u)) {
                    float *row = packed + (size_t)s_id[j] * kRow;
#pragma unroll
                    for (int half = 0; half < NV / 32; ++half) {
                        float s = 0.f;
#pragma unroll
                        for (int w8 = 0; w8 < 8; ++w8)
                            if ((hm[w8] >> j) & 1u) s += s_part[w8][j][half * 32 + lane];
                        if (s != 0.f) atomicAdd(row + half * 32 + lane, s);
                    }
                    if constexpr (GROUPS) {
                        if (lane < 2) {
                            float e = 0.f;
#pragma unroll
                            for (int w8 = 0; w8 < 8; ++w8)
                                if ((hm[w8] >> j) & 1u) e += s_part[w8][j][NV + lane];
                            if (e != 0.f) atomicAdd(row + 31 + lane, e);
                        }
                    }
                }
            }
        } else {  // NV == 16: two Gaussians per warp pass
            const int half = lane >> 4, l16 = lane & 15;
#pragma unroll
            for (int jj = 0; jj < kG / 16; ++jj) {
                const int j = warp * (kG / 8) + jj * 2 + half;
                if (j < m && ((hany >> j) & 1u)) {
                    float *row = packed + (size_t)s_id[j] * kRow;
                    float s = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8)
                        if ((hm[w8] >> j) & 1u) s += s_part[w8][j][l16];
                    if (s != 0.f) atomicAdd(row + l16, s);
                    if constexpr (GROUPS) {
                        if (l16 < 2) {
                            float e = 0.f;
#pragma unroll
                            for (int w8 = 0; w8 < 8; ++w8)
                                if ((hm[w8] >> j) & 1u) e += s_part[w8][j][NV + l16];
                            if (e != 0.f) atomicAdd(row + 31 + l16, e);
                        }
                    }
                }
            }
        }
    }
}

// packed [P,NV] -> the reference's separate gradient tensors.  accumulate: later channel chunks add their share
// of the channel-summed gradients (uv, conic, opacity), exactly like the reference's per-chunk launches do.
template <int NV>
__global__ void __launch_bounds__(kBlock)
unpack_kernel(int P, int C, int Cstride, int c0, int has_bias, int accumulate, const float *__restrict__ packed,
              float2 *__restrict__ dL_duv, float2 *__restrict__ dL_dabs_uv, float *__restrict__ dL_dconic,
              float *__restrict__ dL_dopacity, float *__restrict__ dL_dfeature, float *__restrict__ dL_dbias) {
    const int g = blockIdx.x * kBlock + threadIdx.x;
    if (g >= P) return;
    float r[NV];
    const float4 *row = reinterpret_cast<const float4 *>(packed + (size_t)g * NV);
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) {
        const float4 q = row[k];
        r[4 * k] = q.x; r[4 * k + 1] = q.y; r[4 * k + 2] = q.z; r[4 * k + 3] = q.w;
    }
    if (accumulate) {
        const float2 a = dL_duv[g], b = dL_dabs_uv[g];
        dL_duv[g] = make_float2(a.x + r[0], a.y + r[1]);
        dL_dabs_uv[g] = make_float2(b.x + r[2], b.y + r[3]);
        dL_dconic[3 * g] += r[4]; dL_dconic[3 * g + 1] += r[5]; dL_dconic[3 * g + 2] += r[6];
        dL_dopacity[g] += r[7];
        if (has_bias) dL_dbias[g] += r[NV - 1];
    } else {
        dL_duv[g] = make_float2(r[0], r[1]);
        dL_dabs_uv[g] = make_float2(r[2], r[3]);
        dL_dconic[3 * g] = r[4]; dL_dconic[3 * g + 1] = r[5]; dL_dconic[3 * g + 2] = r[6];
        dL_dopacity[g] = r[7];
        if (has_bias) dL_dbias[g] = r[NV - 1];
    }
#pragma unroll
    for (int c = 0; c < (NV == 64 ? 32 : NV - 8); ++c)
        if (c < C) dL_dfeature[(size_t)g * Cstride + c0 + c] = r[8 + c];
}

__global__ void __launch_bounds__(kBlock)
unpack_groups_kernel(int P, int C, const float *__restrict__ packed, float2 *__restrict__ dL_duv,
                     float2 *__restrict__ dL_duv_ndc, float2 *__restrict__ dL_dabs_uv, float *__restrict__ dL_dconic,
                     float *__restrict__ dL_dopacity, float *__restrict__ dL_dfeature) {
    const int g = blockIdx.x * kBlock + threadIdx.x;
    if (g >= P) return;
    float r[kRowG];
    const float4 *row = reinterpret_cast<const float4 *>(packed + (size_t)g * kRowG);
#pragma unroll
    for (int k = 0; k < kRowG / 4; ++k) {
        const float4 q = row[k];
        r[4 * k] = q.x; r[4 * k + 1] = q.y; r[4 * k + 2] = q.z; r[4 * k + 3] = q.w;
    }
    dL_duv[g] = make_float2(r[0], r[1]);
    dL_dabs_uv[g] = make_float2(r[2], r[3]);
    dL_duv_ndc[g] = make_float2(r[31], r[32]);
    dL_dconic[3 * g] = r[4]; dL_dconic[3 * g + 1] = r[5]; dL_dconic[3 * g + 2] = r[6];
    dL_dopacity[g] = r[7];
#pragma unroll
    for (int c = 0; c < 23; ++c)
        if (c < C) dL_dfeature[(size_t)g * C + c] = r[8 + c];
}

// ------------------------------------------------------------------------------------------------ dispatch
struct FwdArgs {
    int C, Cstride, c0, W, H, gx, K, trunc;
    const float2 *uv; const float *conic, *opacity, *feature, *bias;
    const int *idx_sorted; const int2 *tile_range; float bg, bgB, bgC; int cA, cB;
    float *rendered, *final_T; int *ncontrib, *gs_idx;
};

template <int CH, bool IDX, bool BIAS>
void launch_fwd(const FwdArgs &a, int ntiles, cudaStream_t s) {
    spv::timer_mark(0, 0, s);
    blend_fwd_kernel<CH, IDX, BIAS><<<ntiles, kBlock, 0, s>>>(a.C, a.Cstride, a.c0, a.W, a.H, a.gx, a.K, a.trunc, a.uv,
                                                             a.conic, a.opacity, a.feature, a.bias, a.idx_sorted,
                                                             a.tile_range, a.bg, a.bgB, a.bgC, a.cA, a.cB, a.rendered,
                                                             a.final_T, a.ncontrib, a.gs_idx);
    spv::timer_mark(0, 1, s);
}

template <bool IDX, bool BIAS>
void dispatch_fwd(const FwdArgs &a, int ntiles, cudaStream_t s) {
    const int C = a.C;
    if (C <= 1) launch_fwd<1, IDX, BIAS>(a, ntiles, s);
    else if (C <= 2) launch_fwd<2, IDX, BIAS>(a, ntiles, s);
    else if (C <= 3) launch_fwd<3, IDX, BIAS>(a, ntiles, s);
    else if (C <= 4) launch_fwd<4, IDX, BIAS>(a, ntiles, s);
    else if (C <= 8) launch_fwd<8, IDX, BIAS>(a, ntiles, s);
    else if (C <= 12) launch_fwd<12, IDX, BIAS>(a, ntiles, s);
    else if (C <= 16) launch_fwd<16, IDX, BIAS>(a, ntiles, s);
    else if (C <= 20) launch_fwd<20, IDX, BIAS>(a, ntiles, s);
    else if (C <= 24) launch_fwd<24, IDX, BIAS>(a, ntiles, s);
    else launch_fwd<32, IDX, BIAS>(a, ntiles, s);
}

struct BwdArgs {
    int C, Cstride, c0, W, H, gx;
    const float2 *uv; const float *conic, *opacity, *feature, *bias;
    const int *idx_sorted; const int2 *tile_range; float bg, bgB, bgC;
    const float *final_T; const int *ncontrib; spv::ChanPlanes planes; float *packed;
    int n_grad_channels = 1 << 30;   // groups mode: leading channels whose dL_dfeature is wanted
};

inline spv::ChanPlanes contiguous_planes(const float *base, int C, int W, int H) {
    spv::ChanPlanes pl;
    for (int c = 0; c < 32; ++c) pl.p[c] = (base && c < C) ? base + (size_t)c * H * W : nullptr;
    return pl;
}

template <int NV, int CH, int MODE, int CG = CH>
void launch_bwd(const BwdArgs &a, int ntiles, cudaStream_t s) {
    constexpr int FS = (CH + 3) & ~3;
    constexpr size_t dyn = (CH > 8 && NV <= 32) ? sizeof(float) * kBlock * ((FS <= 20) ? 20 : 28) : 0;
    static std::atomic<unsigned long long> configured{0};   // static (<= 38 KB) + dynamic (<= 28 KB) shared memory exceeds the 48 KB default
    if (dyn) spv::opt_in_dynamic_smem(blend_bwd_kernel<NV, CH, MODE, CG>, dyn, configured);
    spv::timer_mark(1, 0, s);
    blend_bwd_kernel<NV, CH, MODE, CG><<<ntiles, kBlock, dyn, s>>>(a.C, a.Cstride, a.c0, a.W, a.H, a.gx, a.uv, a.conic,
                                                            a.opacity, a.feature, a.bias, a.idx_sorted, a.tile_range,
                                                            a.bg, a.bgB, a.bgC, a.final_T, a.ncontrib, a.planes,
                                                            a.packed);
    spv::timer_mark(1, 1, s);
}

inline int bwd_nv(int C, bool bias) { return C <= (bias ? 7 : 8) ? 16 : (C <= (bias ? 23 : 24) ? 32 : 64); }

template <int MODE>
void dispatch_bwd(const BwdArgs &a, int ntiles, cudaStream_t s) {
    const int C = a.C;
    constexpr bool BIAS = MODE == kBias;
    constexpr int cap16 = BIAS ? 7 : 8, cap32 = BIAS ? 23 : 24;
    if (C <= cap16) {
        if (C <= 1) launch_bwd<16, 1, MODE>(a, ntiles, s);
        else if (C <= 3) launch_bwd<16, 3, MODE>(a, ntiles, s);
        else if (C <= 4) launch_bwd<16, 4, MODE>(a, ntiles, s);
        else launch_bwd<16, cap16, MODE>(a, ntiles, s);
    } else if (C <= cap32) {
        if (C <= 12) launch_bwd<32, 12, MODE>(a, ntiles, s);
        else if (C <= 16) launch_bwd<32, 16, MODE>(a, ntiles, s);
        else if (C <= 20) launch_bwd<32, 20, MODE>(a, ntiles, s);
        else launch_bwd<32, cap32, MODE>(a, ntiles, s);
    } else {
        if (C <= 28) launch_bwd<64, 28, MODE>(a, ntiles, s);
        else launch_bwd<64, 32, MODE>(a, ntiles, s);
    }
}

void dispatch_bwd_groups(const BwdArgs &a, int ntiles, cudaStream_t s) {
    const int C = a.C;
    if (a.n_grad_channels <= 8) {   // <= 8 feature-gradient channels: 16-wide network + the two RGB-pass butterflies
        if (C <= 4) launch_bwd<16, 4, kGroups, 4>(a, ntiles, s);
        else if (C <= 8) launch_bwd<16, 8, kGroups, 8>(a, ntiles, s);
        else if (C <= 12) launch_bwd<16, 12, kGroups, 8>(a, ntiles, s);
        else if (C <= 16) launch_bwd<16, 16, kGroups, 8>(a, ntiles, s);
        else if (C <= 20) launch_bwd<16, 20, kGroups, 8>(a, ntiles, s);
        else launch_bwd<16, 23, kGroups, 8>(a, ntiles, s);
        return;
    }
    if (C <= 12) launch_bwd<32, 12, kGroups>(a, ntiles, s);
    else if (C <= 16) launch_bwd<32, 16, kGroups>(a, ntiles, s);
    else if (C <= 20) launch_bwd<32, 20, kGroups>(a, ntiles, s);
    else launch_bwd<32, 23, kGroups>(a, ntiles, s);
}

}  // namespace

namespace spv {
// Grouped forward; fill_idx = false when the caller has already filled gs_idx with -1 (frame.cu does it on its side stream).
int blend_groups_forward(int P, int C, int W, int H, int K, const float *uv, const float *conic, const float *opacity,
                         const float *feature, const int *idx_sorted, const int *tile_range, float bg_rgb, float bg_depth,
                         float bg_attr, float *rendered, float *final_T, int *ncontrib, int *gs_idx, bool fill_idx, void *stream) {
    (void)P;
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) return 0;
    if (C < 4 || C > 32) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_groups_forward: need 4 <= C <= 32"); return (int)cudaErrorInvalidValue; }
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    const bool has_idx = gs_idx != nullptr && K > 0;
    if (has_idx && fill_idx) SPV_CUDA_TRY(cudaMemsetAsync(gs_idx, 0xFF, sizeof(int) * (size_t)H * W * K, s), "spv_alpha_blend_groups_forward");
    FwdArgs a;
    a.C = C; a.Cstride = C; a.c0 = 0; a.W = W; a.H = H; a.gx = gx; a.K = K; a.trunc = 0;
    a.uv = (const float2 *)uv; a.conic = conic; a.opacity = opacity; a.feature = feature; a.bias = nullptr;
    a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range;
    a.bg = bg_rgb; a.bgB = bg_depth; a.bgC = bg_attr; a.cA = 3; a.cB = 4;
    a.rendered = rendered; a.final_T = final_T; a.ncontrib = ncontrib; a.gs_idx = gs_idx;
    if (has_idx) dispatch_fwd<true, false>(a, ntiles, s); else dispatch_fwd<false, false>(a, ntiles, s);
    return spv::check_launch("spv_alpha_blend_groups_forward");
}
}  // namespace spv

extern "C" {

int spv_alpha_blend_forward(int P, int C, int W, int H, int K, int enable_truncation, const float *uv,
                            const float *conic, const float *opacity, const float *feature, const float *opacity_bias,
                            const int *idx_sorted, const int *tile_range, float bg, float *rendered, float *final_T,
                            int *ncontrib, int *gs_idx, void *stream) {
    (void)P;
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    if (W <= 0 || H <= 0) return 0;
    const bool has_idx = gs_idx != nullptr && K > 0;
    if (has_idx && opacity_bias) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_forward: gs_idx and opacity_bias are exclusive"); return (int)cudaErrorInvalidValue; }
    if (has_idx) SPV_CUDA_TRY(cudaMemsetAsync(gs_idx, 0xFF, sizeof(int) * (size_t)H * W * K, s), "spv_alpha_blend_forward");
    if (C <= 0) {  // the reference launches nothing for C == 0
        SPV_CUDA_TRY(cudaMemsetAsync(final_T, 0, sizeof(float) * (size_t)H * W, s), "spv_alpha_blend_forward");
        SPV_CUDA_TRY(cudaMemsetAsync(ncontrib, 0, sizeof(int) * (size_t)H * W, s), "spv_alpha_blend_forward");
        return 0;
    }
    // channels beyond 32 are processed in chunks; every chunk rewrites final_T / ncontrib with identical values
    // (alpha_blending.cu:287-393).
    for (int c0 = 0; c0 < C; c0 += 32) {
        FwdArgs a;
        a.C = (C - c0 < 32) ? (C - c0) : 32; a.Cstride = C; a.c0 = c0; a.W = W; a.H = H; a.gx = gx; a.K = K;
        a.trunc = enable_truncation;
        a.uv = (const float2 *)uv; a.conic = conic; a.opacity = opacity; a.feature = feature; a.bias = opacity_bias;
        a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range; a.bg = bg; a.bgB = bg; a.bgC = bg;
        a.cA = a.C; a.cB = a.C;
        a.rendered = rendered + (size_t)c0 * H * W; a.final_T = final_T; a.ncontrib = ncontrib; a.gs_idx = gs_idx;
        if (has_idx) dispatch_fwd<true, false>(a, ntiles, s);
        else if (opacity_bias) dispatch_fwd<false, true>(a, ntiles, s);
        else dispatch_fwd<false, false>(a, ntiles, s);
        int rc = spv::check_launch("spv_alpha_blend_forward");
        if (rc) return rc;
    }
    return 0;
}

size_t spv_alpha_blend_backward_workspace_bytes(int P, int C) {
    (void)C;
    return (size_t)(P > 0 ? P : 1) * 64 * sizeof(float);
}

int spv_alpha_blend_backward(int P, int C, int W, int H, const float *uv, const float *conic, const float *opacity,
                             const float *feature, const float *opacity_bias, const int *idx_sorted,
                             const int *tile_range, float bg, const float *final_T, const int *ncontrib,
                             const float *dL_drendered, float *dL_duv, float *dL_dabs_uv, float *dL_dconic,
                             float *dL_dopacity, float *dL_dfeature, float *dL_dopacity_bias, void *workspace,
                             size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return 0;
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    if (ws_bytes < spv_alpha_blend_backward_workspace_bytes(P, C)) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_backward: workspace too small"); return (int)cudaErrorInvalidValue; }
    const bool has_bias = opacity_bias != nullptr;
    if (has_bias && !dL_dopacity_bias) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_backward: dL_dopacity_bias is NULL"); return (int)cudaErrorInvalidValue; }
    float *packed = (float *)workspace;
    const int cap = 32;  // channels per launch == the reference's chunking
    if (C <= 0 || W <= 0 || H <= 0) {
        SPV_CUDA_TRY(cudaMemsetAsync(dL_duv, 0, sizeof(float) * 2 * (size_t)P, s), "spv_alpha_blend_backward");
        SPV_CUDA_TRY(cudaMemsetAsync(dL_dabs_uv, 0, sizeof(float) * 2 * (size_t)P, s), "spv_alpha_blend_backward");
        SPV_CUDA_TRY(cudaMemsetAsync(dL_dconic, 0, sizeof(float) * 3 * (size_t)P, s), "spv_alpha_blend_backward");
        SPV_CUDA_TRY(cudaMemsetAsync(dL_dopacity, 0, sizeof(float) * (size_t)P, s), "spv_alpha_blend_backward");
        if (has_bias) SPV_CUDA_TRY(cudaMemsetAsync(dL_dopacity_bias, 0, sizeof(float) * (size_t)P, s), "spv_alpha_blend_backward");
        return 0;
    }
    for (int c0 = 0; c0 < C; c0 += cap) {
        BwdArgs a;
        a.C = (C - c0 < cap) ? (C - c0) : cap; a.Cstride = C; a.c0 = c0; a.W = W; a.H = H; a.gx = gx;
        a.uv = (const float2 *)uv; a.conic = conic; a.opacity = opacity; a.feature = feature; a.bias = opacity_bias;
        a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range; a.bg = bg; a.bgB = bg; a.bgC = bg;
        a.final_T = final_T; a.ncontrib = ncontrib; a.planes = contiguous_planes(dL_drendered + (size_t)c0 * H * W, a.C, W, H);
        a.packed = packed;
        const int nv = bwd_nv(a.C, has_bias);
        SPV_CUDA_TRY(cudaMemsetAsync(packed, 0, sizeof(float) * (size_t)nv * P, s), "spv_alpha_blend_backward");
        if (has_bias) dispatch_bwd<kBias>(a, ntiles, s); else dispatch_bwd<kPlain>(a, ntiles, s);
        int rc = spv::check_launch("spv_alpha_blend_backward/blend");
        if (rc) return rc;
        const unsigned g = spv::cdiv(P, kBlock);
        if (nv == 16)
            unpack_kernel<16><<<g, kBlock, 0, s>>>(P, a.C, C, c0, has_bias, c0 > 0, packed, (float2 *)dL_duv,
                                                   (float2 *)dL_dabs_uv, dL_dconic, dL_dopacity, dL_dfeature,
                                                   dL_dopacity_bias);
        else if (nv == 64)
            unpack_kernel<64><<<g, kBlock, 0, s>>>(P, a.C, C, c0, has_bias, c0 > 0, packed, (float2 *)dL_duv,
                                                   (float2 *)dL_dabs_uv, dL_dconic, dL_dopacity, dL_dfeature,
                                                   dL_dopacity_bias);
        else
            unpack_kernel<32><<<g, kBlock, 0, s>>>(P, a.C, C, c0, has_bias, c0 > 0, packed, (float2 *)dL_duv,
                                                   (float2 *)dL_dabs_uv, dL_dconic, dL_dopacity, dL_dfeature,
                                                   dL_dopacity_bias);
        rc = spv::check_launch("spv_alpha_blend_backward/unpack");
        if (rc) return rc;
    }
    return 0;
}

// ---- grouped (single-traversal) blending of [rgb(3) | depth(1) | attributes] -------------------------------------
int spv_alpha_blend_groups_forward(int P, int C, int W, int H, int K, const float *uv, const float *conic,
                                   const float *opacity, const float *feature, const int *idx_sorted,
                                   const int *tile_range, float bg_rgb, float bg_depth, float bg_attr, float *rendered,
                                   float *final_T, int *ncontrib, int *gs_idx, void *stream) {
    return spv::blend_groups_forward(P, C, W, H, K, uv, conic, opacity, feature, idx_sorted, tile_range, bg_rgb, bg_depth, bg_attr,
                                     rendered, final_T, ncontrib, gs_idx, /*fill_idx=*/true, stream);
}

size_t spv_alpha_blend_groups_backward_workspace_bytes(int P) { return (size_t)(P > 0 ? P : 1) * kRowG * sizeof(float); }

int spv_alpha_blend_groups_backward(int P, int C, int W, int H, const float *uv, const float *conic,
                                    const float *opacity, const float *feature, const int *idx_sorted,
                                    const int *tile_range, float bg_rgb, float bg_depth, float bg_attr,
                                    const float *final_T, const int *ncontrib, const float *dL_drendered,
                                    float *dL_duv, float *dL_duv_rgb, float *dL_dabs_uv_rgb, float *dL_dconic,
                                    float *dL_dopacity, float *dL_dfeature, void *workspace, size_t ws_bytes,
                                    void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return 0;
    if (C < 4 || C > 23) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_groups_backward: need 4 <= C <= 23"); return (int)cudaErrorInvalidValue; }
    if (ws_bytes < spv_alpha_blend_groups_backward_workspace_bytes(P)) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_groups_backward: workspace too small"); return (int)cudaErrorInvalidValue; }
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    float *packed = (float *)workspace;
    SPV_CUDA_TRY(cudaMemsetAsync(packed, 0, sizeof(float) * (size_t)kRowG * P, s), "spv_alpha_blend_groups_backward");
    if (W > 0 && H > 0) {
        BwdArgs a;
        a.C = C; a.Cstride = C; a.c0 = 0; a.W = W; a.H = H; a.gx = gx;
        a.uv = (const float2 *)uv; a.conic = conic; a.opacity = opacity; a.feature = feature; a.bias = nullptr;
        a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range; a.bg = bg_rgb; a.bgB = bg_depth; a.bgC = bg_attr;
        a.final_T = final_T; a.ncontrib = ncontrib; a.planes = contiguous_planes(dL_drendered, C, W, H); a.packed = packed;
        dispatch_bwd_groups(a, ntiles, s);
        int rc = spv::check_launch("spv_alpha_blend_groups_backward/blend");
        if (rc) return rc;
    }
    unpack_groups_kernel<<<spv::cdiv(P, kBlock), kBlock, 0, s>>>(P, C, packed, (float2 *)dL_duv, (float2 *)dL_duv_rgb,
                                                                 (float2 *)dL_dabs_uv_rgb, dL_dconic, dL_dopacity,
                                                                 dL_dfeature);
    return spv::check_launch("spv_alpha_blend_groups_backward/unpack");
}

/* Grouped backward, blend stage only: upstream gradients as per-channel planes (host array of C device pointers, NULL
 * entries allowed), result left as packed rows (spv::kPackedRowGroups floats per Gaussian) in `packed` for a caller-side
 * unpack (frame.cu).  `packed` must hold P*36 floats. */
int spv_alpha_blend_groups_backward_packed(int P, int C, int W, int H, const float *uv, const float *conic,
                                           const float *opacity, const float *feature, const int *idx_sorted,
                                           const int *tile_range, float bg_rgb, float bg_depth, float bg_attr,
                                           const float *final_T, const int *ncontrib, const float *const *planes_host,
                                           int n_grad_channels, float *packed, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return 0;
    if (C < 4 || C > 23) { spv::set_error(cudaErrorInvalidValue, "spv_alpha_blend_groups_backward_packed: need 4 <= C <= 23"); return (int)cudaErrorInvalidValue; }
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    SPV_CUDA_TRY(cudaMemsetAsync(packed, 0, sizeof(float) * (size_t)kRowG * P, s), "spv_alpha_blend_groups_backward_packed");
    if (W <= 0 || H <= 0) return 0;
    BwdArgs a;
    a.C = C; a.Cstride = C; a.c0 = 0; a.W = W; a.H = H; a.gx = gx;
    a.uv = (const float2 *)uv; a.conic = conic; a.opacity = opacity; a.feature = feature; a.bias = nullptr;
    a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range; a.bg = bg_rgb; a.bgB = bg_depth; a.bgC = bg_attr;
    a.final_T = final_T; a.ncontrib = ncontrib; a.packed = packed;
    for (int c = 0; c < 32; ++c) a.planes.p[c] = c < C ? planes_host[c] : nullptr;
    a.n_grad_channels = n_grad_channels;
    dispatch_bwd_groups(a, ntiles, s);
    return spv::check_launch("spv_alpha_blend_groups_backward_packed");
}

}  // extern "C"
