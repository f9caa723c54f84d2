// Record-staged alpha compositing for the fused frame path (spv_frame_ortho_forward/backward): the single-traversal
// [rgb(3) | depth(1) | attributes] blend of DPTROrthoEnhancedRender.render_iter
// (/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py:330-376; kernels restated:
//  /root/reference/src/submodules/dptr/dptr/gs/src/alpha_blending_enhanced.cu:16-273 and alpha_blending.cu:16-249).
// Same arithmetic as blend.cu (shared inline functions, blend_common.cuh); what changes is how a tile's list reaches the SM:
//
//  * every Gaussian gets ONE 144-byte record per frame (pack_records_kernel, on the frame's side stream):
//      [ x  y  a2 b2 | c2 log2(o) o id | feature[0..23] | a b c 0 ]        (a2,b2,c2 = conic pre-scaled by log2(e))
//    so staging a list entry is a pure copy -- no per-(tile, entry) log2 / scaling, no separate uv / conic / opacity /
//    feature gathers -- and the copy is done by the TMA engine: each thread of the CTA issues one `cp.async.bulk`
//    (global -> shared, 128 or 144 bytes) for the entry it owns, all of them completing on one mbarrier;
//  * the chunks are double-buffered: chunk k+1 is in flight while chunk k is composited, so the warps never wait on a
//    gather (the staged kernels spend 12-18 % of their stall samples there) and one CTA barrier per 128 entries is left
//    (was four per 32 in the backward);
//  * backward: each warp pushes its reduced row straight to the packed gradient buffer with one predicated RED
//    (18 or 32 consecutive floats = at most 4 sectors) instead of parking it in shared memory for a CTA-level fold: the
//    barrier-stall share of the staged kernel (25 % of all stall samples, warps waiting for the slowest warp of the CTA
//    every 32 entries) disappears and 18 KB of shared memory per CTA are freed.
#include "blend_common.cuh"
#include "../../include/spv_b200.h"

namespace {
using namespace spv_blend;

constexpr int kRec = spv::kRecordFloats;   // 36 floats = 144 B per Gaussian
constexpr int kChunk = 128;                // list entries per staging buffer
constexpr int kRowG = spv::kPackedRowGroups;

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SPV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SPV_DONE;\n"
        "bra SPV_WAIT;\n"
        "SPV_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA bulk copy global -> this CTA's shared memory; completion is signalled as `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Row pitch (floats) of a staged record in shared memory: pitch/4 odd, so the per-lane 16-byte reads of 8 consecutive
// records (the block test: one entry per lane) fall into 8 different bank groups.
__host__ __device__ constexpr int rec_pitch(int words) { return ((words / 4) & 1) ? words : words + 4; }

// ---- per-Gaussian records -----------------------------------------------------------------------------------------------
constexpr int kMaxGroups = 8;
struct RecGroups { const float *in[kMaxGroups]; int ch[kMaxGroups]; int start[kMaxGroups]; int n; };

__global__ void __launch_bounds__(kBlock)
pack_records_kernel(int P, int A, const float2 *__restrict__ uv, const float *__restrict__ conic,
                    const float *__restrict__ opacity, const int *__restrict__ radius, const float *__restrict__ rgb,
                    const float *__restrict__ depth, const RecGroups gr, float4 *__restrict__ rec) {
    const long long k = (long long)blockIdx.x * kBlock + threadIdx.x;   // one thread per 16-byte slot of a record
    if (k >= (long long)P * (kRec / 4)) return;
    const int i = (int)(k / (kRec / 4)), s = (int)(k % (kRec / 4));
    if (radius[i] <= 0) return;   // never listed: its record is never read
    float4 o;
    if (s < 2) {
        float4 g0, g1;
        stage_splat(uv[i], conic[3 * i], conic[3 * i + 1], conic[3 * i + 2], opacity[i], 0.f, g0, g1);
        g1.w = __int_as_float(i);
        o = s == 0 ? g0 : g1;
    } else if (s == kRec / 4 - 1) {
        o = make_float4(conic[3 * i], conic[3 * i + 1], conic[3 * i + 2], 0.f);
    } else {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 4 * (s - 2) + q;
            float x = 0.f;
            if (c < 3) x = rgb[3 * i + c];
            else if (c == 3) x = depth[i];
            else if (c < 4 + A) {
                int gi = 0;
#pragma unroll
                for (int t = 1; t < kMaxGroups; ++t) if (t < gr.n && c - 4 >= gr.start[t]) gi = t;
                x = gr.in[gi][(size_t)i * gr.ch[gi] + (c - 4 - gr.start[gi])];
            }
            v[q] = x;
        }
        o = make_float4(v[0], v[1], v[2], v[3]);
    }
    rec[k] = o;
}

// ---- forward ------------------------------------------------------------------------------------------------------------
// CH: feature slots composited (multiple of 4 >= C).  Staged per entry: 8 + CH floats.
template <int CH>
__global__ void __launch_bounds__(kBlock)   // 80 registers, 3 CTAs/SM; capping at 64 registers for 4 CTAs/SM measured slower (167 vs 161 us)
blend_rec_fwd_kernel(int C, int W, int H, int gx, int K, const float *__restrict__ rec, const int *__restrict__ idx_sorted,
                     const int2 *__restrict__ tile_range, const int *__restrict__ tile_order, float bgA, float bgB, float bgC, float *__restrict__ rendered,
                     float *__restrict__ final_T, int *__restrict__ ncontrib, int *__restrict__ gs_idx) {
    constexpr int RW = 8 + CH, RP = rec_pitch(RW);
    __shared__ __align__(128) float s_rec[2][kChunk * RP];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;   // longest tile lists first (sort.cu: tile_scan_kernel)
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
    const int nchunks = (n + kChunk - 1) / kChunk;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int c) {
        const int b = c & 1, m = min(kChunk, n - c * kChunk);
        if (threadIdx.x == 0) mbar_expect_tx(&s_bar[b], (uint32_t)(m * RW * 4));
        if ((int)threadIdx.x < m) {
            const int id = idx_sorted[range.x + c * kChunk + threadIdx.x];
            bulk_g2s(&s_rec[b][threadIdx.x * RP], rec + (size_t)id * kRec, RW * 4, &s_bar[b]);
        }
    };
    if (nchunks > 0) issue(0);

    float T = 1.0f;
    float F[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) F[c] = 0.f;
    int last = 0;
    int *gp = gs_idx + pix * K, *const gend = gp + K;   // next free slot of this pixel's id row
    bool done = !inside;

    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) issue(c + 1);   // its buffer was released by the barrier that ended chunk c-1
        mbar_wait(&s_bar[c & 1], (c >> 1) & 1);
        const float *sr = s_rec[c & 1];
        const int m = min(kChunk, n - c * kChunk), base = c * kChunk;
        for (int j0 = 0; j0 < m; j0 += 32) {
            if (__all_sync(kFull, done)) break;
            // stage 1: one list entry per lane -- can the warp's pixel block take it at all?
            bool maybe = false;
            if (j0 + lane < m) {
                const float4 *r = reinterpret_cast<const float4 *>(sr + (j0 + lane) * RP);
                maybe = block_may_hit<false>(r[0], r[1], bx0, by0);
            }
            unsigned wm = __ballot_sync(kFull, maybe);
            // stage 2: visit the survivors in list order; every lane decides for its own pixel
            while (wm) {
                const int j = j0 + __ffs(wm) - 1;
                wm &= wm - 1;
                const float4 *r = reinterpret_cast<const float4 *>(sr + j * RP);
                const float4 g0 = r[0], g1 = r[1];
                float dx, dy, G;
                const float p2 = splat_p2(g0, g1.x, pxf, pyf, dx, dy);
                const bool hit = !done && splat_hits<false>(p2, g1);
                if (!__any_sync(kFull, hit)) continue;
                if (hit) {
                    const float alpha = splat_alpha<false>(p2, g1, G);
                    const float next_T = T * (1.f - alpha);
                    if (next_T < kTmin) { done = true; continue; }
                    const float w = alpha * T;
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; ++c4) {
                        const float4 f = r[2 + c4];
                        F[4 * c4] = fmaf(f.x, w, F[4 * c4]); F[4 * c4 + 1] = fmaf(f.y, w, F[4 * c4 + 1]);
                        F[4 * c4 + 2] = fmaf(f.z, w, F[4 * c4 + 2]); F[4 * c4 + 3] = fmaf(f.w, w, F[4 * c4 + 3]);
                    }
                    T = next_T;
                    last = base + j + 1;
                    if (gp < gend) *gp++ = __float_as_int(g1.w);
                }
            }
        }
        if (__syncthreads_count(done) == kBlock) {   // also releases buffer c&1 for chunk c+2
            if (c + 1 < nchunks) mbar_wait(&s_bar[(c + 1) & 1], ((c + 1) >> 1) & 1);   // never exit under an in-flight copy
            break;
        }
    }

    if (inside) {
        final_T[pix] = T;
        ncontrib[pix] = last;
        const size_t HW = (size_t)H * W;
#pragma unroll
        for (int c = 0; c < CH; ++c)
            if (c < C) rendered[c * HW + pix] = F[c] + T * (c < 3 ? bgA : (c < 4 ? bgB : bgC));
    }
}

// ---- backward -----------------------------------------------------------------------------------------------------------
// Packed gradient row (kRowG = 36 floats per Gaussian): 0,1 dL_duv (all channels)  2,3 |RGB-pass dL_duv|  4,5,6 dL_dconic
// 7 dL_dopacity (rgb + depth passes: the attribute pass gets opacity.detach(), dptr_ortho_enhanced.py:362)
// 8.. dL_dfeature  31,32 RGB-pass dL_duv.  CG = leading feature channels whose gradient is wanted.  Reduction networks:
// CG <= 8: 16-wide (8 geometric sums + 8 features) + the RGB-pass pair on a 2-wide one (21 shuffles);
// CG <= 14: 16-wide + an 8-wide one carrying features 8..13 and the RGB-pass pair (25 shuffles); else 32-wide + the pair (36).
// Ring staging.  Round 1's kernel synchronised its 8 warps at every 128-entry chunk; ncu (profiles/r01_ncu_blend_rec_bwd.csv)
// showed the CTA barrier as its top stall reason (2.2 warps per issue slot) -- the 8x4-pixel blocks of a tile take a different
// number of list entries, and every chunk ended when the slowest block was done.  Here the warps of a tile are decoupled
// (measured: -2.4 %, profiles/r02_bwd_variants.txt -- the barrier was where warps waited, not what bounds the kernel):
//   * the list is staged into a ring of kRing buffers of kRingChunk entries (same 36 KB as the two 128-entry buffers), all of
//     them requested up front -- tile lists average ~290 entries, so most tiles never wait again;
//   * a warp that finishes a chunk bumps that buffer's counter; the LAST warp to do so (by definition nobody is left reading
//     it) re-arms the buffer's mbarrier and issues the bulk copies of the chunk kRing further down the list.  Nobody ever
//     waits for a slower warp, only for data;
//   * the per-pixel recurrences run on R = sum over the Gaussians behind of <f, dL_dpixel> alpha T (one FMA per group and
//     entry, no state selects: a lane that did not take the Gaussian has w = 0) instead of the normalised accum_rec form:
//         dL_dalpha = T <f, d> - (R + T_final <bg, d>) / (1 - alpha)         (alpha_blending.cu:180-246, same sum)
constexpr int kRing = 4, kRingChunk = 64;

__device__ __forceinline__ int atom_add_acq_rel_shared(int *addr, int v) {
    int old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
    return old;
}

// ABS: the |RGB-pass dL_duv| pair (packed columns 2,3) is wanted.  The renderer only reads it under densify_abs_grad_enable
// (dptr_ortho_enhanced.py:378-383 returns ONE of ndc / abs_ndc as viewspace_points); without it the RGB-pass pair takes its two
// slots in the 16-wide network and the extra network shrinks or disappears:
//                    ABS                                          !ABS
//   CG <= 8          16-wide + pair butterfly       21 shuffles   16-wide                        16
//   CG <= 12         16-wide + 8-wide (feat + pair) 25            16-wide + 4-wide (feat 8..11)  22
//   CG <= 14         16-wide + 8-wide               25            16-wide + 8-wide (feat 8..13)  25
//   else             32-wide + pair butterfly       36            32-wide                        31
template <int CH, int CG, bool ABS>
__global__ void __launch_bounds__(kBlock, 3)
blend_rec_bwd_kernel(int C, int W, int H, int gx, const float *__restrict__ rec, const int *__restrict__ idx_sorted,
                     const int2 *__restrict__ tile_range, const int *__restrict__ tile_order, float bgA, float bgB, float bgC,
                     const float *__restrict__ final_T, const int *__restrict__ ncontrib, const spv::ChanPlanes planes,
                     float *__restrict__ packed) {
    constexpr int NV = (CG <= 14) ? 16 : 32;
    constexpr int NU = (CG <= 8 || NV == 32) ? 0 : ((!ABS && CG <= 12) ? 4 : 8);   // second network: features 8.. (+ the pair if ABS)
    static_assert(CH % 4 == 0 && CH >= 4 && CH <= 24 && 8 + CG <= 32 && CG <= CH, "unsupported channel configuration");
    static_assert(kRing * kRingChunk == kBlock, "one bulk copy per thread fills the whole ring");
    constexpr int RP = kRec;                        // 36: pitch/4 = 9 is odd
    constexpr int DS = rec_pitch(CH);               // dL_dpixel row pitch, pitch/4 odd
    extern __shared__ __align__(128) float s_dyn[];
    float *s_rec0 = s_dyn;                                            // [kRing][kRingChunk][RP]
    float *dq = s_dyn + kRing * kRingChunk * RP + threadIdx.x * DS;   // this pixel's dL_dpixel row
    __shared__ __align__(8) uint64_t s_full[kRing];
    __shared__ int s_done[kRing];
    __shared__ int s_max;

    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;   // longest tile lists first (sort.cu: tile_scan_kernel)
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

    if (threadIdx.x == 0) {
        s_max = 0;
#pragma unroll
        for (int b = 0; b < kRing; ++b) { mbar_init(&s_full[b], 1); s_done[b] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int wmax = __reduce_max_sync(kFull, last_contrib);   // positions >= wmax were applied by no pixel of this warp
    if (lane == 0 && wmax > 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int n_eff = min(range.y - range.x, s_max);   // ... and positions >= s_max by no pixel of the tile: never staged
    const int nchunks = (n_eff + kRingChunk - 1) / kRingChunk;

    // chunk c covers list positions [p_hi - m, p_hi), p_hi = n_eff - c*kRingChunk; slot t holds position p_hi - 1 - t
    {   // the first kRing chunks: one bulk copy per thread
        const int c = threadIdx.x / kRingChunk, t = threadIdx.x % kRingChunk;
        if (c < nchunks) {
            const int p_hi = n_eff - c * kRingChunk, m = min(kRingChunk, p_hi);
            if (t == 0) mbar_expect_tx(&s_full[c], (uint32_t)(m * kRec * 4));
            if (t < m) {
                const int id = idx_sorted[range.x + p_hi - 1 - t];
                bulk_g2s(s_rec0 + (c * kRingChunk + t) * RP, rec + (size_t)id * kRec, kRec * 4, &s_full[c]);
            }
        }
    }

    // <bg, dL_dpixel> per gradient group, pre-multiplied by T_final
    float tfA = 0.f, tfB = 0.f, tfC = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float dv = (inside && c < C && planes.p[c]) ? planes.p[c][pix] : 0.f;
        dq[c] = dv;
        if (c >= 4) tfC += dv;
        else if (c == 3) tfB += dv;
        else tfA += dv;
    }
    tfA *= bgA * T_final; tfB *= bgB * T_final; tfC *= bgC * T_final;

    float RA = 0.f, RB = 0.f, RC = 0.f;

    for (int c = 0; c < nchunks; ++c) {
        const int b = c % kRing;
        mbar_wait(&s_full[b], (c / kRing) & 1);   // every warp waits for every chunk, also one it will skip: a refill is only
                                                  // ever issued into a buffer whose previous copies have landed
        const float *sr = s_rec0 + b * kRingChunk * RP;
        const int p_hi = n_eff - c * kRingChunk, m = min(kRingChunk, p_hi);
        for (int j0 = 0; j0 < m; j0 += 32) {
            if (p_hi - 1 - (j0 + 31) >= wmax) continue;   // the whole sub-batch lies behind this warp's last contributor
            // stage 1: one chunk entry per lane -- can the warp's 8x4 pixel block have taken it at all?
            bool maybe = false;
            if (j0 + lane < m && p_hi - 1 - (j0 + lane) < wmax) {
                const float4 *r = reinterpret_cast<const float4 *>(sr + (j0 + lane) * RP);
                maybe = block_may_hit<false>(r[0], r[1], bx0, by0);
            }
            unsigned todo = __ballot_sync(kFull, maybe);
            // stage 2 (chunk order = back to front).  Branch-free body: a lane that did not take the Gaussian runs it with
            // p2 = -inf, i.e. G = alpha = w = 0 and 1/(1-alpha) = 1 -- every partial sum it contributes is an exact zero and
            // its T / R state does not move.
            while (todo) {
                const int j = j0 + __ffs(todo) - 1;
                todo &= todo - 1u;
                const float4 *r = reinterpret_cast<const float4 *>(sr + j * RP);
                float dx = 0.f, dy = 0.f;
                const float4 g0 = r[0], g1 = r[1];
                float p2 = splat_p2(g0, g1.x, pxf, pyf, dx, dy);
                // did this pixel apply the Gaussian in the forward pass?  (same test, and before its last contributor)
                const bool hit = splat_hits<false>(p2, g1) && (p_hi - 1 - j) < last_contrib;
                if (!__any_sync(kFull, hit)) continue;
                float v[NV];
                float u[NU > 0 ? NU : 1];   // second network: features 8.. (| RGB-pass pair when ABS)
                float n0, n1;
                {
                    const float4 con = r[kRec / 4 - 1];
                    float Gv;
                    p2 = hit ? p2 : -INFINITY;
                    const float alpha = splat_alpha<false>(p2, g1, Gv);
                    const float rinv = __fdividef(1.f, 1.f - alpha);
                    T = T * rinv;  // transmittance in front of this Gaussian (unchanged when alpha == 0)
                    const float w = alpha * T;
                    float fdA = 0.f, fdB = 0.f, fdC = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; ++c4) {
                        const float4 ff = r[2 + c4];
                        const float4 dd = *reinterpret_cast<const float4 *>(dq + 4 * c4);
                        const float fv[4] = {ff.x, ff.y, ff.z, ff.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int ch = 4 * c4 + k;
                            if (ch >= 4) fdC = fmaf(fv[k], dv[k], fdC);
                            else if (ch == 3) fdB = fmaf(fv[k], dv[k], fdB);
                            else fdA = fmaf(fv[k], dv[k], fdA);
                            if (ch < CG) {
                                if (NU > 0 && ch >= 8) u[ch - 8] = w * dv[k];
                                else v[8 + ch] = w * dv[k];
                            }
                        }
                    }
                    const float da_ndc = fmaf(T, fdA, -rinv * (RA + tfA));
                    const float da_op = da_ndc + fmaf(T, fdB, -rinv * (RB + tfB));
                    const float da_all = da_op + fmaf(T, fdC, -rinv * (RC + tfC));
                    RA = fmaf(fdA, w, RA); RB = fmaf(fdB, w, RB); RC = fmaf(fdC, w, RC);
#pragma unroll
                    for (int q = 8 + (NU > 0 ? 8 : CG); q < NV; ++q) v[q] = 0.f;
#pragma unroll
                    for (int q = CG - 8; q < NU - (ABS ? 2 : 0); ++q) u[q] = 0.f;
                    const float dL_dG = g1.z * da_all;
                    const float dGx = -Gv * dx * con.x - Gv * dy * con.y;
                    const float dGy = -Gv * dy * con.z - Gv * dx * con.y;
                    v[0] = dL_dG * dGx; v[1] = dL_dG * dGy;
                    v[4] = -0.5f * Gv * dx * dx * dL_dG;
                    v[5] = -Gv * dx * dy * dL_dG;
                    v[6] = -0.5f * Gv * dy * dy * dL_dG;
                    v[7] = Gv * da_op;
                    const float dL_dG_ndc = g1.z * da_ndc;
                    n0 = dL_dG_ndc * dGx; n1 = dL_dG_ndc * dGy;
                    if constexpr (ABS) { v[2] = fabsf(n0); v[3] = fabsf(n1); }
                    else { v[2] = n0; v[3] = n1; }        // the RGB-pass pair rides in the |.| slots
                }
                halving_reduce<NV, 0, NV>(v, lane);   // lane l (< NV) now holds the warp-wide sum of value l
                float *row = packed + (size_t)__float_as_int(g1.w) * kRowG;
                // network slot -> packed column: 2,3 hold the RGB-pass pair (columns 31,32) when !ABS
                const int vcol = (!ABS && (lane == 2 || lane == 3)) ? 29 + lane : lane;
                if constexpr (NU > 0) {
                    if constexpr (ABS) { u[NU - 2] = n0; u[NU - 1] = n1; }
                    halving_reduce<NU, 0, NU>(u, lane);     // lane l holds the sum of u[l % NU]
                    const int t = lane - 16;                // one RED: lanes 0..15 <- v, lanes 16..16+NU-1 <- u
                    const float val = lane < 16 ? v[0] : u[0];
                    const int nf = ABS ? NU - 2 : NU;       // feature slots of the second network
                    const int col = lane < 16 ? vcol : (t < nf ? 16 + t : 31 + (t - nf));
                    const bool live = lane < 16 || (lane < 16 + NU && (t >= nf || t < CG - 8));
                    if (live && val != 0.f) atomicAdd(row + col, val);
                } else if constexpr (ABS) {
                    // the two RGB-pass sums: one halving step (odd lanes take n1, even lanes n0), then 4 butterfly steps
                    const bool up = (lane & 1) != 0;
                    float e = (up ? n1 : n0) + __shfl_xor_sync(kFull, up ? n0 : n1, 1);
#pragma unroll
                    for (int o = 2; o <= 16; o <<= 1) e += __shfl_xor_sync(kFull, e, o);
                    if constexpr (NV == 16) {   // one RED: lanes 0..15 the network's sums, lanes 16,17 the RGB-pass pair
                        const float val = lane < 16 ? v[0] : e;
                        const int col = lane < 16 ? lane : 31 + (lane & 1);
                        if (lane < 18 && val != 0.f) atomicAdd(row + col, val);
                    } else {
                        if (lane < 8 + CG && v[0] != 0.f) atomicAdd(row + lane, v[0]);
                        if (lane < 2 && e != 0.f) atomicAdd(row + 31 + lane, e);
                    }
                } else {
                    if (lane < 8 + CG && v[0] != 0.f) atomicAdd(row + vcol, v[0]);
                }
            }
        }
        // release buffer b; the last of the 8 warps to get here refills it with chunk c + kRing
        if (c + kRing < nchunks) {      // (uniform over the CTA: buffers that will not be refilled need no bookkeeping)
            __syncwarp();
            int old = 0;
            if (lane == 0) old = atom_add_acq_rel_shared(&s_done[b], 1);
            old = __shfl_sync(kFull, old, 0);
            if (old == kBlock / 32 - 1) {
                const int cn = c + kRing, q_hi = n_eff - cn * kRingChunk, mn = min(kRingChunk, q_hi);
                if (lane == 0) {
                    s_done[b] = 0;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warps' reads of the buffer before the async writes
                    mbar_expect_tx(&s_full[b], (uint32_t)(mn * kRec * 4));
                }
                __syncwarp();
#pragma unroll
                for (int t = lane; t < kRingChunk; t += 32)
                    if (t < mn) {
                        const int id = idx_sorted[range.x + q_hi - 1 - t];
                        bulk_g2s(s_rec0 + (b * kRingChunk + t) * RP, rec + (size_t)id * kRec, kRec * 4, &s_full[b]);
                    }
            }
        }
    }
}

struct RecFwdArgs {
    int C, W, H, gx, K;
    const float *rec; const int *idx_sorted; const int2 *tile_range; const int *tile_order; float bgA, bgB, bgC;
    float *rendered, *final_T; int *ncontrib, *gs_idx;
};

template <int CH>
void launch_rec_fwd(const RecFwdArgs &a, int ntiles, cudaStream_t s) {
    spv::timer_mark(0, 0, s);
    blend_rec_fwd_kernel<CH><<<ntiles, kBlock, 0, s>>>(a.C, a.W, a.H, a.gx, a.K, a.rec, a.idx_sorted, a.tile_range, a.tile_order, a.bgA, a.bgB,
                                                      a.bgC, a.rendered, a.final_T, a.ncontrib, a.gs_idx);
    spv::timer_mark(0, 1, s);
}

struct RecBwdArgs {
    int C, W, H, gx;
    const float *rec; const int *idx_sorted; const int2 *tile_range; const int *tile_order; float bgA, bgB, bgC;
    const float *final_T; const int *ncontrib; spv::ChanPlanes planes; float *packed;
};

template <int CH, int CG, bool ABS>
void launch_rec_bwd(const RecBwdArgs &a, int ntiles, cudaStream_t s) {
    constexpr size_t dyn = sizeof(float) * (kRing * kRingChunk * kRec + kBlock * rec_pitch(CH));
    static std::atomic<unsigned long long> configured{0};   // up to 64.5 KB of dynamic shared memory: above the 48 KB default
    spv::opt_in_dynamic_smem(blend_rec_bwd_kernel<CH, CG, ABS>, dyn, configured);
    spv::timer_mark(1, 0, s);
    blend_rec_bwd_kernel<CH, CG, ABS><<<ntiles, kBlock, dyn, s>>>(a.C, a.W, a.H, a.gx, a.rec, a.idx_sorted, a.tile_range, a.tile_order, a.bgA, a.bgB,
                                                                 a.bgC, a.final_T, a.ncontrib, a.planes, a.packed);
    spv::timer_mark(1, 1, s);
}

template <int CH, bool ABS>
void dispatch_rec_bwd_cg(const RecBwdArgs &a, int n_grad, int ntiles, cudaStream_t s) {
    // feature-gradient channels reduced: 4 (rgb + depth only), 8, 12, 14, or all CH
    if (n_grad <= 4) launch_rec_bwd<CH, 4, ABS>(a, ntiles, s);
    else if (n_grad <= 8 && CH >= 8) launch_rec_bwd<CH, (CH >= 8 ? 8 : CH), ABS>(a, ntiles, s);
    else if (n_grad <= 12 && CH >= 16) launch_rec_bwd<CH, (CH >= 16 ? 12 : CH), ABS>(a, ntiles, s);
    else if (n_grad <= 14 && CH >= 16) launch_rec_bwd<CH, (CH >= 16 ? 14 : CH), ABS>(a, ntiles, s);
    else launch_rec_bwd<CH, (CH > 23 ? 23 : CH), ABS>(a, ntiles, s);
}

template <int CH>
void dispatch_rec_bwd(const RecBwdArgs &a, int n_grad, bool want_abs, int ntiles, cudaStream_t s) {
    if (want_abs) dispatch_rec_bwd_cg<CH, true>(a, n_grad, ntiles, s);
    else dispatch_rec_bwd_cg<CH, false>(a, n_grad, ntiles, s);
}

}  // namespace

namespace spv {

int pack_records(int P, int A, const float *uv, const float *conic, const float *opacity, const int *radius, const float *rgb,
                 const float *depth, int n_groups, const float *const *attr_ptrs, const int *attr_channels, float *rec,
                 void *stream) {
    if (P <= 0) return 0;
    RecGroups gr;
    gr.n = n_groups;
    int start = 0;
    for (int q = 0; q < kMaxGroups; ++q) {
        gr.in[q] = q < n_groups ? attr_ptrs[q] : nullptr;
        gr.ch[q] = q < n_groups ? attr_channels[q] : 0;
        gr.start[q] = start;
        start += gr.ch[q];
    }
    pack_records_kernel<<<spv::cdiv((long long)P * (kRec / 4), kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        P, A, (const float2 *)uv, conic, opacity, radius, rgb, depth, gr, (float4 *)rec);
    return spv::check_launch("spv_frame/pack_records");
}

int blend_records_forward(int C, int W, int H, int K, const float *rec, const int *idx_sorted, const int *tile_range,
                          const int *tile_order, float bg_rgb, float bg_depth, float bg_attr, float *rendered, float *final_T,
                          int *ncontrib, int *gs_idx, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) return 0;
    if (C < 4 || C > 23 || K <= 0 || !gs_idx) { spv::set_error(cudaErrorInvalidValue, "blend_records_forward: need 4 <= C <= 23 and K > 0"); return (int)cudaErrorInvalidValue; }
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    RecFwdArgs a;
    a.C = C; a.W = W; a.H = H; a.gx = gx; a.K = K; a.rec = rec; a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range;
    a.tile_order = tile_order;
    a.bgA = bg_rgb; a.bgB = bg_depth; a.bgC = bg_attr; a.rendered = rendered; a.final_T = final_T; a.ncontrib = ncontrib;
    a.gs_idx = gs_idx;
    if (C <= 4) launch_rec_fwd<4>(a, ntiles, s);
    else if (C <= 8) launch_rec_fwd<8>(a, ntiles, s);
    else if (C <= 12) launch_rec_fwd<12>(a, ntiles, s);
    else if (C <= 16) launch_rec_fwd<16>(a, ntiles, s);
    else if (C <= 20) launch_rec_fwd<20>(a, ntiles, s);
    else launch_rec_fwd<24>(a, ntiles, s);
    return spv::check_launch("spv_frame/blend_records_forward");
}

int blend_records_backward(int P, int C, int W, int H, const float *rec, const int *idx_sorted, const int *tile_range,
                           const int *tile_order, float bg_rgb, float bg_depth, float bg_attr, const float *final_T, const int *ncontrib,
                           const float *const *planes_host, int n_grad_channels, bool want_abs, float *packed, bool packed_is_zero,
                           void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return 0;
    if (C < 4 || C > 23) { spv::set_error(cudaErrorInvalidValue, "blend_records_backward: need 4 <= C <= 23"); return (int)cudaErrorInvalidValue; }
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), ntiles = gx * gy;
    if (!packed_is_zero) SPV_CUDA_TRY(cudaMemsetAsync(packed, 0, sizeof(float) * (size_t)kRowG * P, s), "blend_records_backward");
    if (W <= 0 || H <= 0) return 0;
    // Trailing image channels without an upstream gradient (e.g. the trainer renders mask / dino / pos_poly_feat images it puts
    // no loss on, trainer_fragGS.py:600-640) contribute exact zeros to every sum: they are not traversed at all.
    int C_live = 4;
    for (int c = 4; c < C; ++c) if (planes_host[c]) C_live = c + 1;
    if (n_grad_channels > C_live) n_grad_channels = C_live;   // their feature gradients are exact zeros (the rows were cleared)
    C = C_live;
    RecBwdArgs a;
    a.C = C; a.W = W; a.H = H; a.gx = gx; a.rec = rec; a.idx_sorted = idx_sorted; a.tile_range = (const int2 *)tile_range;
    a.tile_order = tile_order;
    a.bgA = bg_rgb; a.bgB = bg_depth; a.bgC = bg_attr; a.final_T = final_T; a.ncontrib = ncontrib; a.packed = packed;
    for (int c = 0; c < 32; ++c) a.planes.p[c] = c < C ? planes_host[c] : nullptr;
    const int ng = n_grad_channels < 4 ? 4 : (n_grad_channels > C ? C : n_grad_channels);
    if (C <= 4) dispatch_rec_bwd<4>(a, ng, want_abs, ntiles, s);
    else if (C <= 8) dispatch_rec_bwd<8>(a, ng, want_abs, ntiles, s);
    else if (C <= 12) dispatch_rec_bwd<12>(a, ng, want_abs, ntiles, s);
    else if (C <= 16) dispatch_rec_bwd<16>(a, ng, want_abs, ntiles, s);
    else if (C <= 20) dispatch_rec_bwd<20>(a, ng, want_abs, ntiles, s);
    else dispatch_rec_bwd<24>(a, ng, want_abs, ntiles, s);
    return spv::check_launch("spv_frame/blend_records_backward");
}

}  // namespace spv
