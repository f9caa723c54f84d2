// Gradient exchange packing for frame-parallel training.  The reference has no gradient collective
// (`--distributed` only barriers, /root/reference/src/train.py:19-31,210-213); this is the data path of the new one.
//
// The flat gradient buffer holds 180 floats per Gaussian at 50 frames, but a step only produces non-zero gradient in
//   * the dense parameters (scaling, rotation, opacity, attributes),
//   * a fixed SUBSET of slices of some parameters (SH bases 0, 2, 6, 12 under the renderer's constant view direction (0,0,1),
//     dptr_ortho_enhanced.py:270-271),
//   * one or two SPARSE slices of the spline coefficients -- the intervals of this rank's two frame times, different per
//     rank (dynamic_gaussian_with_base_point_cloud.py:239-247).
// pack: flat gradient -> [all-reduce buffer | all-gather buffer (+ the interval indices)], scaled by 1/world; the collectives
// are NCCL's (one all-reduce, one all-gather); unpack: all-reduce result back into the flat gradient, every rank's sparse
// slices added at their own intervals in rank order (bit-identical on every rank), and the intervals recorded in the
// `dirty` list the deformation backward uses to keep the coefficient gradient clean (deform.cu).  Two streaming kernels
// instead of ~25 torch index_select / cat / index_add launches.
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {
constexpr int kThreads = 256;
constexpr int kMaxSeg = SPV_EXCHANGE_MAX_SEGMENTS;
constexpr int kMaxSel = SPV_EXCHANGE_MAX_SELECT;

constexpr int kMaxWidth = 48;   // floats per Gaussian a subset / sparse segment may contribute

struct Seg {
    long long flat_off;   // first float of the parameter in the flat buffer
    long long comm_off;   // first float of its block in the all-reduce (dense / subset) or all-gather (sparse) buffer
    int A, B, Cn;         // the parameter's per-Gaussian row viewed as [A, B, Cn]; slices are taken along B
    int mode;             // 0 dense (all of B), 1 subset (sel[] host-known), 2 sparse (indices read from the device)
    int nsel;
    int w;                // floats per Gaussian in the comm buffer = A * nsel * Cn
    int cta0, ncta;       // this segment's CTAs inside the 1-D grid
    int vec4;             // dense segment whose offsets and length are multiples of 4 floats: moved as float4
    short off[kMaxWidth];         // gather modes: comm column k -> offset inside the Gaussian's [A,B,Cn] row (sparse: slice 0)
    unsigned char slot[kMaxWidth];   // sparse: which of the rank's slices column k belongs to
};
struct Plan {
    Seg seg[kMaxSeg];
    int nseg, ncta;
    const int *sparse_idx[2];     // device scalars: this rank's active slices (sparse mode)
    long long n_ar, n_ag;         // floats in the all-reduce part / the all-gather payload (without the index tail)
};

// One 1-D grid; every CTA first finds its segment (uniform, <= 16 compares).  Dense segments are a scaled linear copy
// (their comm block has the parameter's own layout); gather segments run 8 Gaussians per CTA with the lanes over the
// Gaussian's comm columns and a host-built column -> row-offset table: no integer division anywhere.
__global__ void __launch_bounds__(kThreads)
exchange_pack_kernel(int P, Plan plan, const float *__restrict__ flat_grad, float scale, float *__restrict__ comm_ar,
                     float *__restrict__ comm_ag) {
    int q = 0;
    for (int t = 1; t < plan.nseg; ++t) if ((int)blockIdx.x >= plan.seg[t].cta0) q = t;
    const Seg &s = plan.seg[q];
    const int cta = blockIdx.x - s.cta0;
    if (s.mode == 0) {
        const long long r = (long long)cta * kThreads + threadIdx.x;
        if (s.vec4) {
            if (r < (long long)P * s.w / 4) {
                const float4 v = reinterpret_cast<const float4 *>(flat_grad + s.flat_off)[r];
                reinterpret_cast<float4 *>(comm_ar + s.comm_off)[r] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
            }
        } else if (r < (long long)P * s.w) comm_ar[s.comm_off + r] = flat_grad[s.flat_off + r] * scale;
        return;
    }
    const int lane = threadIdx.x & 31, i = cta * 8 + (threadIdx.x >> 5);
    if (s.mode == 2 && cta == 0 && threadIdx.x < kMaxSel)   // index tail of the all-gather payload (int bits)
        reinterpret_cast<int *>(comm_ag)[plan.n_ag + threadIdx.x] = (int)threadIdx.x < s.nsel ? plan.sparse_idx[threadIdx.x][0] : -1;
    if (i >= P) return;
    const float *row = flat_grad + s.flat_off + (long long)i * (s.A * s.B * s.Cn);
    float *dst = (s.mode == 2 ? comm_ag : comm_ar) + s.comm_off + (long long)i * s.w;
    int b0 = 0, b1 = 0;
    if (s.mode == 2) { b0 = plan.sparse_idx[0][0]; b1 = s.nsel > 1 ? plan.sparse_idx[1][0] : b0; }
    for (int k = lane; k < s.w; k += 32) {
        float v;
        if (s.mode == 2) {
            const int sl = s.slot[k];
            v = (sl == 1 && b1 == b0) ? 0.f : row[s.off[k] + (sl ? b1 : b0) * s.Cn] * scale;   // same interval twice: sent once
        } else v = row[s.off[k]] * scale;
        dst[k] = v;
    }
}

__global__ void __launch_bounds__(kThreads)
exchange_unpack_kernel(int P, Plan plan, int world, const float *__restrict__ comm_ar, const float *__restrict__ gathered,
                       long long ag_stride, float *__restrict__ flat_grad, int *__restrict__ dirty) {
    int q = 0;
    for (int t = 1; t < plan.nseg; ++t) if ((int)blockIdx.x >= plan.seg[t].cta0) q = t;
    const Seg &s = plan.seg[q];
    const int cta = blockIdx.x - s.cta0;
    if (s.mode == 0) {
        const long long r = (long long)cta * kThreads + threadIdx.x;
        if (s.vec4) {
            if (r < (long long)P * s.w / 4) reinterpret_cast<float4 *>(flat_grad + s.flat_off)[r] = reinterpret_cast<const float4 *>(comm_ar + s.comm_off)[r];
        } else if (r < (long long)P * s.w) flat_grad[s.flat_off + r] = comm_ar[s.comm_off + r];
        return;
    }
    const int lane = threadIdx.x & 31, i = cta * 8 + (threadIdx.x >> 5);
    if (s.mode == 1) {
        if (i >= P) return;
        float *row = flat_grad + s.flat_off + (long long)i * (s.A * s.B * s.Cn);
        const float *src = comm_ar + s.comm_off + (long long)i * s.w;
        for (int k = lane; k < s.w; k += 32) row[s.off[k]] = src[k];
        return;
    }
    // sparse: every rank's slices are summed per interval in (rank, slot) order -- bit-identical on every rank
    const int nsel = s.nsel, pairs = world * nsel;   // <= 16 contributions
    int idx[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
        idx[j] = j < pairs ? reinterpret_cast<const int *>(gathered + (j / nsel) * ag_stride)[plan.n_ag + (j % nsel)] : -1;
    if (cta == 0 && threadIdx.x == 0 && dirty) {
        for (int j = 0; j < 16; ++j) if (j < pairs) dirty[1 + j] = idx[j];
        dirty[0] = pairs;
    }
    if (i >= P) return;
    float *row = flat_grad + s.flat_off + (long long)i * (s.A * s.B * s.Cn);
    const float *src = gathered + s.comm_off + (long long)i * s.w;
    for (int k = lane; k < s.w; k += 32) {
        if (s.slot[k] != 0) continue;                 // one lane per (a, c): the columns of slot 0
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j >= pairs) break;
            const int bj = idx[j];
            bool first = bj >= 0 && bj < s.B;
#pragma unroll
            for (int j2 = 0; j2 < 16; ++j2) if (j2 < j) first &= idx[j2] != bj;
            if (!first) continue;
            float sum = 0.f;
#pragma unroll
            for (int j2 = 0; j2 < 16; ++j2)
                if (j2 >= j && j2 < pairs && idx[j2] == bj) sum += src[(j2 / nsel) * ag_stride + k + (j2 % nsel) * s.Cn];
            row[s.off[k] + bj * s.Cn] = sum;
        }
    }
}

// out[e] = scale * sum over ranks (in rank order: bit-identical on every rank) of gathered[r * stride + e]
__global__ void __launch_bounds__(kThreads)
exchange_reduce_kernel(long long n4, int world, const float4 *__restrict__ gathered, long long stride4, float scale,
                       float4 *__restrict__ out) {
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= n4) return;
    float4 acc = gathered[e];
    for (int r = 1; r < world; ++r) {
        const float4 v = gathered[r * stride4 + e];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[e] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
}

// The same reduction straight out of the peers' memory (NVLink P2P loads from symmetric buffers; no NCCL collective):
// [0, n_red) is summed over ranks in rank order, [n_red, n_row) is copied rank by rank into the local `rows` (the gathered
// position gradients + frame scalars the deferred spline backward reads).  One pass: every remote byte crosses NVLink once.
struct PeerPtrs { const float4 *p[8]; };
__global__ void __launch_bounds__(kThreads)
exchange_reduce_peers_kernel(long long n_red4, long long n_row4, int world, PeerPtrs peers, float scale, float4 *__restrict__ reduced,
                             float4 *__restrict__ rows, long long row_stride4) {
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= n_row4) return;
    if (e < n_red4) {
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < world) v[r] = peers.p[r][e];   // all loads in flight before the adds
        float4 acc = v[0];
#pragma unroll
        for (int r = 1; r < 8; ++r) if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        reduced[e] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    } else if (rows) {
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < world) rows[r * row_stride4 + e] = peers.p[r][e];
    }
}

// The gathered tails alone ([n_red, n_row) of every rank's row -> local `rows`): lets the caller run the gather and what
// depends on it (the deferred spline backward) on a second stream next to the reduction.
constexpr int kPeerUnroll = 4;   // float4 per thread and peer: world x 4 x 16 B of NVLink reads in flight per thread
__global__ void __launch_bounds__(kThreads)
exchange_gather_peers_kernel(long long n_red4, long long n_row4, int world, PeerPtrs peers, float4 *__restrict__ rows,
                             long long row_stride4) {
    const long long base = n_red4 + (long long)blockIdx.x * (kThreads * kPeerUnroll) + threadIdx.x;
    for (int r0 = 0; r0 < world; r0 += 2) {          // two peers x kPeerUnroll loads in flight, then their stores
        float4 v[2][kPeerUnroll];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                const long long e = base + (long long)u * kThreads;
                if (r0 + rr < world && e < n_row4) v[rr][u] = peers.p[r0 + rr][e];
            }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                const long long e = base + (long long)u * kThreads;
                if (r0 + rr < world && e < n_row4) rows[(r0 + rr) * row_stride4 + e] = v[rr][u];
            }
    }
}

// Two-phase variant for larger groups (inbound volume 2(N-1)/N instead of N-1 times the summed block): phase 1 -- every
// rank sums ITS 1/N slice of the block over all peers and publishes it in its symmetric `red` area (and copies the gathered
// tails); a barrier; phase 2 -- every rank fetches the other slices from their owners.  Same rank-order sums, so the result is
// still bit-identical on every rank.
__global__ void __launch_bounds__(kThreads)
exchange_reduce_scatter_kernel(long long n_red4, long long n_row4, long long chunk4, int rank, int world, PeerPtrs peers, float scale,
                               float4 *__restrict__ red_pub, float4 *__restrict__ reduced, float4 *__restrict__ rows,
                               long long row_stride4) {
    const long long lo = (long long)rank * chunk4, hi = min(n_red4, lo + chunk4), mine = max(hi - lo, 0ll);
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t < mine) {
        const long long e = lo + t;
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < world) v[r] = peers.p[r][e];
        float4 acc = v[0];
#pragma unroll
        for (int r = 1; r < 8; ++r) if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        acc = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
        red_pub[e] = acc;
        reduced[e] = acc;
    } else {
        const long long e = n_red4 + (t - mine);
        if (e >= n_row4 || !rows) return;
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < world) rows[r * row_stride4 + e] = peers.p[r][e];
    }
}

__global__ void __launch_bounds__(kThreads)
exchange_fetch_reduced_kernel(long long n_red4, long long chunk4, int rank, PeerPtrs red_peers, float4 *__restrict__ reduced) {
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= n_red4) return;
    const int owner = (int)(e / chunk4);
    if (owner != rank) reduced[e] = red_peers.p[owner][e];
}

// NVLS form: the NVSwitch does the sum.  Every rank reads ITS 1/N slice of the block through the MULTICAST address with
// multimem.ld_reduce (the switch returns the fp32 sum over all ranks' buffers: one inbound copy instead of N-1), scales it and
// multimem.st's it to the `red` area of EVERY rank at once.  Per rank the summed block costs ~1x its size on the wire for any
// N, and every rank receives the very same bits.  The gathered tails still travel as peer loads.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Publish by PUSH (NVLS mode): the summed block of the staging row is copied into this rank's half of the symmetric buffer (where
// the peers' multimem.ld_reduce will read it) and the gathered tail -- position gradients + frame scalars, which every rank needs
// from every rank -- is written with ONE multicast store per 16 bytes into slot `rank` of EVERY rank's gathered area.  Stores are
// fire-and-forget: after the barrier that follows, every rank holds all tails locally and the spline backward can start at once,
// next to the in-switch reduction, instead of behind a pull of (N-1) x 24 B per Gaussian through peer LOADS (measured at N = 4:
// 74 us of gather + 33 us of spline tail behind it, the longest branch of the exchange).
__global__ void __launch_bounds__(kThreads)
exchange_publish_kernel(long long n_red4, long long n_row4, const float4 *__restrict__ row, float4 *__restrict__ sym_row,
                        float4 *__restrict__ mc_gathered_slot) {
    const long long e0 = (long long)blockIdx.x * (kThreads * kPeerUnroll) + threadIdx.x;
    float4 v[kPeerUnroll];
#pragma unroll
    for (int u = 0; u < kPeerUnroll; ++u) {
        const long long e = e0 + (long long)u * kThreads;
        if (e < n_row4) v[u] = row[e];
    }
#pragma unroll
    for (int u = 0; u < kPeerUnroll; ++u) {
        const long long e = e0 + (long long)u * kThreads;
        if (e >= n_row4) continue;
        if (e < n_red4) sym_row[e] = v[u];
        else multimem_st(mc_gathered_slot + (e - n_red4), v[u]);
    }
}

__global__ void __launch_bounds__(kThreads)
exchange_nvls_kernel(long long n_red4, long long n_row4, long long chunk4, int rank, int world, const float4 *__restrict__ mc_row,
                     float4 *__restrict__ mc_red, PeerPtrs peers, float scale, float4 *__restrict__ rows, long long row_stride4) {
    const long long lo = (long long)rank * chunk4, hi = min(n_red4, lo + chunk4), mine = max(hi - lo, 0ll);
    const long long mine_blocks = (mine + kThreads * kPeerUnroll - 1) / (kThreads * kPeerUnroll);
    if ((long long)blockIdx.x < mine_blocks) {
        // kPeerUnroll in-switch reductions in flight per thread, then their multicast stores
        const long long base = lo + (long long)blockIdx.x * (kThreads * kPeerUnroll) + threadIdx.x;
        float4 v[kPeerUnroll];
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const long long e = base + (long long)u * kThreads;
            if (e < hi) v[u] = multimem_ld_reduce_add(mc_row + e);
        }
#pragma unroll
        for (int u = 0; u < kPeerUnroll; ++u) {
            const long long e = base + (long long)u * kThreads;
            if (e < hi) multimem_st(mc_red + e, make_float4(v[u].x * scale, v[u].y * scale, v[u].z * scale, v[u].w * scale));
        }
    } else {
        const long long e = n_red4 + ((long long)blockIdx.x - mine_blocks) * kThreads + threadIdx.x;
        if (e >= n_row4 || !rows) return;
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < world) rows[r * row_stride4 + e] = peers.p[r][e];
    }
}

int build_plan(Plan &plan, int P, int nseg, const spv_exchange_segment *segs, const int *const *sparse_idx_dev, const char *where) {
    if (nseg < 1 || nseg > kMaxSeg) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
    plan.nseg = nseg;
    long long ar = 0, ag = 0;
    int n_sparse = 0, cta = 0;
    plan.sparse_idx[0] = plan.sparse_idx[1] = nullptr;
    for (int q = 0; q < nseg; ++q) {
        const spv_exchange_segment &in = segs[q];
        Seg &s = plan.seg[q];
        s.flat_off = in.flat_offset; s.A = in.A; s.B = in.B; s.Cn = in.C; s.mode = in.mode;
        s.nsel = in.mode == 0 ? in.B : in.nsel;
        bool bad = in.A < 1 || in.B < 1 || in.C < 1 || in.mode < 0 || in.mode > 2 || s.nsel < 1 || s.nsel > kMaxSel;
        s.w = bad ? 0 : s.A * s.nsel * s.Cn;
        bad = bad || (in.mode != 0 && (s.w > kMaxWidth || (long long)s.A * s.B * s.Cn > 32767)) ||
              (in.mode == 2 && (s.nsel > 2 || ++n_sparse > 1 || !sparse_idx_dev));
        if (bad) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
        for (int k = 0; k < kMaxWidth; ++k) { s.off[k] = 0; s.slot[k] = 0; }
        if (in.mode != 0)
            for (int k = 0; k < s.w; ++k) {   // comm column k = (a, si, c)
                const int a = k / (s.nsel * s.Cn), si = (k / s.Cn) % s.nsel, c = k % s.Cn;
                const int b = in.mode == 1 ? in.sel[si] : 0;
                if (b < 0 || b >= in.B) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
                s.off[k] = (short)((a * s.B + b) * s.Cn + c);
                s.slot[k] = (unsigned char)si;
            }
        const long long n = (long long)P * s.w;
        const long long comm_next = in.mode == 2 ? ag : ar;
        s.vec4 = (in.mode == 0 && (n & 3) == 0 && (s.flat_off & 3) == 0 && (comm_next & 3) == 0) ? 1 : 0;
        const long long nc = in.mode == 0 ? ((s.vec4 ? n / 4 : n) + kThreads - 1) / kThreads : ((long long)P + 7) / 8;
        if (cta + nc >= (1ll << 31)) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
        s.cta0 = cta; s.ncta = (int)nc; cta += (int)nc;
        if (in.mode == 2) { s.comm_off = ag; ag += n; for (int t = 0; t < s.nsel; ++t) plan.sparse_idx[t] = sparse_idx_dev[t]; }
        else { s.comm_off = ar; ar += n; }
    }
    plan.ncta = cta;
    plan.n_ar = ar; plan.n_ag = ag;
    return 0;
}
}  // namespace

extern "C" {

int spv_exchange_sizes(int P, int nseg, const spv_exchange_segment *segs, long long *n_allreduce, long long *n_allgather) {
    Plan plan;
    static const int *dummy[kMaxSel] = {nullptr};
    int rc = build_plan(plan, P, nseg, segs, dummy, "spv_exchange_sizes: bad segment table");
    if (rc) return rc;
    *n_allreduce = plan.n_ar;
    *n_allgather = plan.n_ag + (plan.n_ag ? kMaxSel : 0);   // payload + index tail
    return 0;
}

int spv_exchange_pack(int P, int nseg, const spv_exchange_segment *segs, const int *const *sparse_idx_dev, const float *flat_grad,
                      float scale, float *comm_allreduce, float *comm_allgather, void *stream) {
    if (P <= 0) return 0;
    Plan plan;
    int rc = build_plan(plan, P, nseg, segs, sparse_idx_dev, "spv_exchange_pack: bad segment table");
    if (rc) return rc;
    exchange_pack_kernel<<<plan.ncta, kThreads, 0, (cudaStream_t)stream>>>(P, plan, flat_grad, scale, comm_allreduce, comm_allgather);
    return spv::check_launch("spv_exchange_pack");
}

int spv_exchange_unpack(int P, int nseg, const spv_exchange_segment *segs, int world, const float *comm_allreduce,
                        const float *gathered, float *flat_grad, int *dirty, void *stream) {
    if (P <= 0) return 0;
    if (world < 1 || world > 8) { spv::set_error(cudaErrorInvalidValue, "spv_exchange_unpack: 1..8 ranks"); return (int)cudaErrorInvalidValue; }
    Plan plan;
    static const int *dummy[kMaxSel] = {nullptr};
    int rc = build_plan(plan, P, nseg, segs, dummy, "spv_exchange_unpack: bad segment table");
    if (rc) return rc;
    const long long stride = plan.n_ag + (plan.n_ag ? kMaxSel : 0);
    exchange_unpack_kernel<<<plan.ncta, kThreads, 0, (cudaStream_t)stream>>>(P, plan, world, comm_allreduce, gathered, stride, flat_grad,
                                                                             dirty);
    return spv::check_launch("spv_exchange_unpack");
}

/* out[0..n) = scale * sum_r gathered[r*stride + 0..n): the local reduction behind a single all-gather (n and stride multiples
 * of 4, 16-byte aligned buffers). */
int spv_exchange_reduce(long long n, int world, const float *gathered, long long stride, float scale, float *out, void *stream) {
    if (n <= 0) return 0;
    if ((n & 3) || (stride & 3) || world < 1) { spv::set_error(cudaErrorInvalidValue, "spv_exchange_reduce: n and stride must be multiples of 4"); return (int)cudaErrorInvalidValue; }
    exchange_reduce_kernel<<<spv::cdiv(n / 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(n / 4, world, (const float4 *)gathered,
                                                                                             stride / 4, scale, (float4 *)out);
    return spv::check_launch("spv_exchange_reduce");
}

/* peer_rows: host array of `world` device pointers (this rank's own row included, in rank order) to rows of n_row floats laid
 * out as [n_red summed floats | gathered floats]; the caller has synchronised the ranks (all rows written) before this launch. */
int spv_exchange_reduce_peers(long long n_red, long long n_row, int world, const float *const *peer_rows, float scale, float *reduced,
                              float *rows, long long row_stride, void *stream) {
    if (n_row <= 0) return 0;
    if ((n_red & 3) || (n_row & 3) || (row_stride & 3) || world < 1 || world > 8 || n_red > n_row) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_reduce_peers: sizes must be multiples of 4, 1..8 ranks");
        return (int)cudaErrorInvalidValue;
    }
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (const float4 *)peer_rows[r < world ? r : 0];
    exchange_reduce_peers_kernel<<<spv::cdiv((rows ? n_row : (n_red > 0 ? n_red : 4)) / 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n_red / 4, n_row / 4, world, pp, scale, (float4 *)reduced, (float4 *)rows, row_stride / 4);
    return spv::check_launch("spv_exchange_reduce_peers");
}

/* Two-phase form (world >= 4): phase 1 sums this rank's 1/world slice of [0, n_red) over the peers' rows into red_pub (this
 * rank's symmetric area, read by the peers in phase 2) and `reduced`, and copies the gathered tails; the caller places a
 * cross-rank barrier; phase 2 fetches the other slices from peer_red[owner]. */
int spv_exchange_reduce_scatter_peers(long long n_red, long long n_row, int rank, int world, const float *const *peer_rows, float scale,
                                      float *red_pub, float *reduced, float *rows, long long row_stride, void *stream) {
    if (n_row <= 0) return 0;
    if ((n_red & 3) || (n_row & 3) || (row_stride & 3) || world < 1 || world > 8 || n_red > n_row || rank < 0 || rank >= world) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_reduce_scatter_peers: sizes must be multiples of 4, 1..8 ranks");
        return (int)cudaErrorInvalidValue;
    }
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (const float4 *)peer_rows[r < world ? r : 0];
    const long long n_red4 = n_red / 4, n_row4 = n_row / 4, chunk4 = (n_red4 + world - 1) / world;
    const long long lo = (long long)rank * chunk4, hi = lo + chunk4 < n_red4 ? lo + chunk4 : n_red4;
    const long long work = (hi > lo ? hi - lo : 0) + (rows ? n_row4 - n_red4 : 0);
    exchange_reduce_scatter_kernel<<<spv::cdiv(work > 0 ? work : 1, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        n_red4, n_row4, chunk4 > 0 ? chunk4 : 1, rank, world, pp, scale, (float4 *)red_pub, (float4 *)reduced, (float4 *)rows, row_stride / 4);
    return spv::check_launch("spv_exchange_reduce_scatter_peers");
}

/* rows[r*row_stride + e] = peer_rows[r][e] for e in [n_red, n_row): the gather part alone (the reduction entry points skip it
 * when called with rows == NULL). */
int spv_exchange_gather_peers(long long n_red, long long n_row, int world, const float *const *peer_rows, float *rows,
                              long long row_stride, void *stream) {
    if (n_row <= n_red) return 0;
    if ((n_red & 3) || (n_row & 3) || (row_stride & 3) || world < 1 || world > 8) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_gather_peers: sizes must be multiples of 4, 1..8 ranks");
        return (int)cudaErrorInvalidValue;
    }
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (const float4 *)peer_rows[r < world ? r : 0];
    exchange_gather_peers_kernel<<<spv::cdiv((n_row - n_red) / 4, kThreads * kPeerUnroll), kThreads, 0, (cudaStream_t)stream>>>(
        n_red / 4, n_row / 4, world, pp, (float4 *)rows, row_stride / 4);
    return spv::check_launch("spv_exchange_gather_peers");
}

/* NVLS form: mc_row / mc_red = MULTICAST addresses of the symmetric row / red areas; after the caller's second barrier the
 * summed block (scaled) is in every rank's own red area. */
int spv_exchange_nvls(long long n_red, long long n_row, int rank, int world, const float *mc_row, float *mc_red,
                      const float *const *peer_rows, float scale, float *rows, long long row_stride, void *stream) {
    if (n_row <= 0) return 0;
    if ((n_red & 3) || (n_row & 3) || (row_stride & 3) || world < 1 || world > 8 || n_red > n_row || rank < 0 || rank >= world || !mc_row || !mc_red) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_nvls: sizes must be multiples of 4, 1..8 ranks, multicast pointers set");
        return (int)cudaErrorInvalidValue;
    }
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (const float4 *)peer_rows[r < world ? r : 0];
    const long long n_red4 = n_red / 4, n_row4 = n_row / 4, chunk4 = (n_red4 + world - 1) / world;
    const long long lo = (long long)rank * chunk4, hi = lo + chunk4 < n_red4 ? lo + chunk4 : n_red4;
    const long long mine = hi > lo ? hi - lo : 0;
    const long long blocks = (mine + kThreads * kPeerUnroll - 1) / (kThreads * kPeerUnroll) + (rows ? (n_row4 - n_red4 + kThreads - 1) / kThreads : 0);
    exchange_nvls_kernel<<<(unsigned)(blocks > 0 ? blocks : 1), kThreads, 0, (cudaStream_t)stream>>>(
        n_red4, n_row4, chunk4 > 0 ? chunk4 : 1, rank, world, (const float4 *)mc_row, (float4 *)mc_red, pp, scale, (float4 *)rows,
        row_stride / 4);
    return spv::check_launch("spv_exchange_nvls");
}

/* NVLS publish: row[0, n_red) -> sym_row (this rank's symmetric staging row, local address); row[n_red, n_row) -> multicast store
 * into mc_gathered_slot = the multicast address of slot `rank` of the gathered area (n_row - n_red floats per slot). */
int spv_exchange_publish(long long n_red, long long n_row, const float *row, float *sym_row, float *mc_gathered_slot, void *stream) {
    if (n_row <= 0) return 0;
    if ((n_red & 3) || (n_row & 3) || n_red > n_row || !row || !sym_row || !mc_gathered_slot) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_publish: sizes must be multiples of 4, pointers set");
        return (int)cudaErrorInvalidValue;
    }
    exchange_publish_kernel<<<spv::cdiv(n_row / 4, kThreads * kPeerUnroll), kThreads, 0, (cudaStream_t)stream>>>(
        n_red / 4, n_row / 4, (const float4 *)row, (float4 *)sym_row, (float4 *)mc_gathered_slot);
    return spv::check_launch("spv_exchange_publish");
}

int spv_exchange_fetch_reduced(long long n_red, int rank, int world, const float *const *peer_red, float *reduced, void *stream) {
    if (n_red <= 0) return 0;
    if ((n_red & 3) || world < 1 || world > 8 || rank < 0 || rank >= world) {
        spv::set_error(cudaErrorInvalidValue, "spv_exchange_fetch_reduced: n_red must be a multiple of 4, 1..8 ranks");
        return (int)cudaErrorInvalidValue;
    }
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (const float4 *)peer_red[r < world ? r : 0];
    const long long n_red4 = n_red / 4, chunk4 = (n_red4 + world - 1) / world;
    exchange_fetch_reduced_kernel<<<spv::cdiv(n_red4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(n_red4, chunk4 > 0 ? chunk4 : 1, rank, pp,
                                                                                                  (float4 *)reduced);
    return spv::check_launch("spv_exchange_fetch_reduced");
}

}  // extern "C"
