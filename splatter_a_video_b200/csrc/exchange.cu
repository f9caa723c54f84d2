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

struct Seg {
    long long flat_off;   // first float of the parameter in the flat buffer
    long long comm_off;   // first float of its block in the all-reduce (dense / subset) or all-gather (sparse) buffer
    int A, B, Cn;         // the parameter's per-Gaussian row viewed as [A, B, Cn]; slices are taken along B
    int mode;             // 0 dense (all of B), 1 subset (sel[] host-known), 2 sparse (indices read from the device)
    int nsel;
    int sel[kMaxSel];
};
struct Plan {
    Seg seg[kMaxSeg];
    int nseg;
    const int *sparse_idx[kMaxSel];   // device scalars: this rank's active slices (sparse mode)
    long long n_ar, n_ag;             // floats in the all-reduce part / the all-gather payload (without the index tail)
};

__device__ __forceinline__ int seg_width(const Seg &s) { return s.A * s.nsel * s.Cn; }

__global__ void __launch_bounds__(kThreads)
exchange_pack_kernel(int P, Plan plan, const float *__restrict__ flat_grad, float scale, float *__restrict__ comm_ar,
                     float *__restrict__ comm_ag) {
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long total = plan.n_ar + plan.n_ag;
    if (e >= total + kMaxSel) return;
    if (e >= total) {   // index tail of the all-gather payload (int bits)
        const int t = (int)(e - total);
        int v = -1;
        for (int q = 0; q < plan.nseg; ++q)
            if (plan.seg[q].mode == 2 && t < plan.seg[q].nsel) v = plan.sparse_idx[t][0];
        reinterpret_cast<int *>(comm_ag)[plan.n_ag + t] = v;
        return;
    }
    const bool ag = e >= plan.n_ar;
    const long long local = ag ? e - plan.n_ar : e;
    int q = 0;
#pragma unroll 1
    for (int t = 0; t < plan.nseg; ++t) {
        const Seg &s = plan.seg[t];
        if ((s.mode == 2) == ag && local >= s.comm_off && local < s.comm_off + (long long)P * seg_width(s)) q = t;
    }
    const Seg &s = plan.seg[q];
    const int w = seg_width(s);
    const long long r = local - s.comm_off;
    const int i = (int)(r / w), k = (int)(r % w);
    const int a = k / (s.nsel * s.Cn), si = (k / s.Cn) % s.nsel, c = k % s.Cn;
    int b;
    bool dup = false;
    if (s.mode == 2) {
        b = plan.sparse_idx[si][0];
        for (int t = 0; t < si; ++t) dup |= plan.sparse_idx[t][0] == b;   // the same interval twice: sent once
    } else b = s.sel[si];
    const float v = flat_grad[s.flat_off + (long long)i * s.A * s.B * s.Cn + ((long long)a * s.B + b) * s.Cn + c];
    (ag ? comm_ag : comm_ar)[local] = dup ? 0.f : v * scale;
}

__global__ void __launch_bounds__(kThreads)
exchange_unpack_kernel(int P, Plan plan, int world, const float *__restrict__ comm_ar, const float *__restrict__ gathered,
                       long long ag_stride, float *__restrict__ flat_grad, int *__restrict__ dirty) {
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    // the sparse part is walked once per (Gaussian, a, c): n_ag / nsel work items
    int sq = -1;
    for (int t = 0; t < plan.nseg; ++t) if (plan.seg[t].mode == 2) sq = t;
    const int nsel = sq >= 0 ? plan.seg[sq].nsel : 1;
    const long long n_sp = plan.n_ag / nsel;
    if (e == 0 && dirty && sq >= 0) {
        int n = 0;
        for (int r = 0; r < world; ++r)
            for (int t = 0; t < nsel; ++t)
                if (n < 16) dirty[1 + n++] = reinterpret_cast<const int *>(gathered + r * ag_stride)[plan.n_ag + t];
        dirty[0] = n;
    }
    if (e < plan.n_ar) {
        int q = 0;
#pragma unroll 1
        for (int t = 0; t < plan.nseg; ++t) {
            const Seg &s = plan.seg[t];
            if (s.mode != 2 && e >= s.comm_off && e < s.comm_off + (long long)P * seg_width(s)) q = t;
        }
        const Seg &s = plan.seg[q];
        const int w = seg_width(s);
        const long long r = e - s.comm_off;
        const int i = (int)(r / w), k = (int)(r % w);
        const int a = k / (s.nsel * s.Cn), si = (k / s.Cn) % s.nsel, c = k % s.Cn;
        flat_grad[s.flat_off + (long long)i * s.A * s.B * s.Cn + ((long long)a * s.B + s.sel[si]) * s.Cn + c] = comm_ar[e];
        return;
    }
    const long long u = e - plan.n_ar;
    if (u >= n_sp || sq < 0) return;
    const Seg &s = plan.seg[sq];
    const int per = s.A * s.Cn;                       // work items per Gaussian
    const int i = (int)(u / per), k = (int)(u % per);
    const int a = k / s.Cn, c = k % s.Cn;
    float *row = flat_grad + s.flat_off + (long long)i * s.A * s.B * s.Cn + (long long)a * s.B * s.Cn + c;
    const int pairs = world * nsel;                   // <= 16 (rank, slot) contributions, summed per interval in rank order
    for (int j = 0; j < pairs; ++j) {
        const int rj = j / nsel, tj = j % nsel;
        const int bj = reinterpret_cast<const int *>(gathered + rj * ag_stride)[plan.n_ag + tj];
        bool first = true;
        for (int j2 = 0; j2 < j; ++j2)
            first &= reinterpret_cast<const int *>(gathered + (j2 / nsel) * ag_stride)[plan.n_ag + (j2 % nsel)] != bj;
        if (!first || bj < 0 || bj >= s.B) continue;
        float sum = 0.f;
        for (int j2 = j; j2 < pairs; ++j2) {
            const int r2 = j2 / nsel, t2 = j2 % nsel;
            if (reinterpret_cast<const int *>(gathered + r2 * ag_stride)[plan.n_ag + t2] != bj) continue;
            sum += gathered[r2 * ag_stride + s.comm_off + ((long long)i * s.A * nsel + (long long)a * nsel + t2) * s.Cn + c];
        }
        row[(long long)bj * s.Cn] = sum;
    }
}

int build_plan(Plan &plan, int P, int nseg, const spv_exchange_segment *segs, const int *const *sparse_idx_dev, const char *where) {
    if (nseg < 1 || nseg > kMaxSeg) { spv::set_error(cudaErrorInvalidValue, where); return (int)cudaErrorInvalidValue; }
    plan.nseg = nseg;
    long long ar = 0, ag = 0;
    int n_sparse = 0;
    for (int t = 0; t < kMaxSel; ++t) plan.sparse_idx[t] = nullptr;
    for (int q = 0; q < nseg; ++q) {
        const spv_exchange_segment &in = segs[q];
        Seg &s = plan.seg[q];
        s.flat_off = in.flat_offset; s.A = in.A; s.B = in.B; s.Cn = in.C; s.mode = in.mode;
        s.nsel = in.mode == 0 ? in.B : in.nsel;
        if (in.A < 1 || in.B < 1 || in.C < 1 || in.mode < 0 || in.mode > 2 || s.nsel < 1 || s.nsel > kMaxSel ||
            (in.mode == 2 && (s.nsel > 2 || ++n_sparse > 1 || !sparse_idx_dev))) {
            spv::set_error(cudaErrorInvalidValue, where);
            return (int)cudaErrorInvalidValue;
        }
        for (int t = 0; t < kMaxSel; ++t) s.sel[t] = in.mode == 0 ? t : (t < s.nsel && in.mode == 1 ? in.sel[t] : 0);
        const long long n = (long long)P * s.A * s.nsel * s.Cn;
        if (in.mode == 2) { s.comm_off = ag; ag += n; for (int t = 0; t < s.nsel; ++t) plan.sparse_idx[t] = sparse_idx_dev[t]; }
        else { s.comm_off = ar; ar += n; }
    }
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
    const long long total = plan.n_ar + plan.n_ag + (plan.n_ag ? kMaxSel : 0);
    exchange_pack_kernel<<<spv::cdiv(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, plan, flat_grad, scale, comm_allreduce,
                                                                                           comm_allgather);
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
    int nsel = 1;
    for (int q = 0; q < nseg; ++q) if (plan.seg[q].mode == 2) nsel = plan.seg[q].nsel;
    const long long total = plan.n_ar + plan.n_ag / nsel;
    const long long stride = plan.n_ag + (plan.n_ag ? kMaxSel : 0);
    exchange_unpack_kernel<<<spv::cdiv(total > 0 ? total : 1, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P, plan, world, comm_allreduce, gathered, stride, flat_grad, dirty);
    return spv::check_launch("spv_exchange_unpack");
}

}  // extern "C"
