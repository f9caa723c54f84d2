// Per-frame deformation of the active model: cubic-spline position
//   pos(t) = base + c3 + c2*d + c1*d^2 + c0*d^3,  coeff = pos_cubic_node[P, 4, NI, 3], (interval idx, d) from the frame time
// restating /root/reference/src/dynamic_gaussian_with_base_point_cloud.py:236-250 (get_position), and its backward
// (gradient to the 4 coefficients of the active interval; `position` itself is frozen there, :97-99).
// The interval index and the in-interval distance are read from DEVICE memory so a captured CUDA graph can be replayed
// for any frame by updating two scalars.  One coalesced streaming pass: 12 B read + 48 B coefficients + 12 B written per
// Gaussian forward; the reference evaluates it with ~10 torch kernels over [P,3].
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
deform_fwd_kernel(int P, int NI, int layout, const float *__restrict__ base, const float *__restrict__ coeff,
                  const int *__restrict__ idx_dev, const float *__restrict__ dist_dev, float *__restrict__ pos) {
    const int k = blockIdx.x * kThreads + threadIdx.x;   // one thread per (Gaussian, xyz component)
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const int idx = idx_dev[0];
    const float d = dist_dev[0];
    const size_t s = layout ? 3 : (size_t)NI * 3, si = layout ? 12 : 3;      // plane / interval strides: [4][NI][3] or [NI][4][3]
    const float *q = coeff + (size_t)i * 4 * NI * 3 + (size_t)idx * si + c;
    // same association as the reference: c3 + c2*d + c1*d**2 + c0*d**3, then + position
    const float v = q[3 * s] + q[2 * s] * d + q[1 * s] * (d * d) + q[0] * (d * d * d);
    pos[k] = v + base[k];
}

__global__ void __launch_bounds__(kThreads)
deform_bwd_kernel(int P, int NI, int layout, const int *__restrict__ idx_dev, const float *__restrict__ dist_dev,
                  const float *__restrict__ dL_dpos, float *__restrict__ dL_dcoeff, int accumulate) {
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const int idx = idx_dev[0];
    const float d = dist_dev[0];
    const float g = dL_dpos[k];
    const size_t s = layout ? 3 : (size_t)NI * 3, si = layout ? 12 : 3;
    float *q = dL_dcoeff + (size_t)i * 4 * NI * 3 + (size_t)idx * si + c;
    if (accumulate) { q[3 * s] += g; q[2 * s] += g * d; q[1 * s] += g * (d * d); q[0] += g * (d * d * d); }
    else { q[3 * s] = g; q[2 * s] = g * d; q[1 * s] = g * (d * d); q[0] = g * (d * d * d); }
}
// Two frame times in one pass (the trainer evaluates the model at ids1 and ids2 every step, trainer_fragGS.py:486-487: the
// position of frame ids1 is rendered, the position of ids2 travels as the `track_gs` attribute, :506-508).  The two active
// intervals are equal or adjacent for neighbouring frames, so the second evaluation re-uses the sectors of the first.
__global__ void __launch_bounds__(kThreads)
deform_fwd2_kernel(int P, int NI, int layout, const float *__restrict__ base, const float *__restrict__ coeff,
                   const int *__restrict__ idx1_dev, const float *__restrict__ dist1_dev, const int *__restrict__ idx2_dev,
                   const float *__restrict__ dist2_dev, float *__restrict__ pos1, float *__restrict__ pos2) {
    const int k = blockIdx.x * kThreads + threadIdx.x;   // one thread per (Gaussian, xyz component)
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const int i1 = idx1_dev[0], i2 = idx2_dev[0];
    const float d1 = dist1_dev[0], d2 = dist2_dev[0];
    const size_t s = layout ? 3 : (size_t)NI * 3, si = layout ? 12 : 3;
    const float *row = coeff + (size_t)i * 4 * NI * 3 + c;
    const float *q = row + (size_t)i1 * si;
    const float a0 = q[0], a1 = q[s], a2 = q[2 * s], a3 = q[3 * s];
    float b0 = a0, b1 = a1, b2 = a2, b3 = a3;
    if (i2 != i1) { const float *r = row + (size_t)i2 * si; b0 = r[0]; b1 = r[s]; b2 = r[2 * s]; b3 = r[3 * s]; }
    const float bs = base[k];
    pos1[k] = (a3 + a2 * d1 + a1 * (d1 * d1) + a0 * (d1 * d1 * d1)) + bs;
    pos2[k] = (b3 + b2 * d2 + b1 * (d2 * d2) + b0 * (d2 * d2 * d2)) + bs;
}

// Backward of the pair into a gradient SINK (e.g. the parameter's slice of a flat gradient buffer) that is kept clean
// incrementally: `dirty` (device, [0] = count, [1..16] = interval indices) lists the intervals that hold non-zero gradient
// from earlier calls / gradient exchanges; they are zeroed here unless re-written, instead of clearing all 4*NI*3 floats per
// Gaussian every step (96 MB at 200 k Gaussians x 50 frames).  g2 may be NULL (ids2 position used without gradient).
__global__ void __launch_bounds__(kThreads)
deform_bwd2_kernel(int P, int NI, int layout, const int *__restrict__ idx1_dev, const float *__restrict__ dist1_dev,
                   const int *__restrict__ idx2_dev, const float *__restrict__ dist2_dev, const float *__restrict__ g1,
                   const float *__restrict__ g2, const int *__restrict__ dirty, float *__restrict__ dL_dcoeff) {
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const int i1 = idx1_dev[0], i2 = g2 ? idx2_dev[0] : i1;
    const float d1 = dist1_dev[0], d2 = dist2_dev[0];
    const size_t s = layout ? 3 : (size_t)NI * 3, si = layout ? 12 : 3;
    float *row = dL_dcoeff + (size_t)i * 4 * NI * 3 + c;
    const int nd = min(dirty[0], 16);
    for (int t = 0; t < nd; ++t) {
        const int z = dirty[1 + t];
        bool skip = z == i1 || z == i2 || z < 0 || z >= NI;
        for (int t2 = 0; t2 < t; ++t2) skip |= dirty[1 + t2] == z;       // listed twice: cleared once
        if (skip) continue;
        float *q = row + (size_t)z * si;
        q[0] = 0.f; q[s] = 0.f; q[2 * s] = 0.f; q[3 * s] = 0.f;
    }
    const float ga = g1[k], gb = g2 ? g2[k] : 0.f;
    float *q = row + (size_t)i1 * si;
    if (i2 == i1) {
        q[3 * s] = ga + gb; q[2 * s] = ga * d1 + gb * d2; q[1 * s] = ga * (d1 * d1) + gb * (d2 * d2);
        q[0] = ga * (d1 * d1 * d1) + gb * (d2 * d2 * d2);
    } else {
        q[3 * s] = ga; q[2 * s] = ga * d1; q[1 * s] = ga * (d1 * d1); q[0] = ga * (d1 * d1 * d1);
        float *r = row + (size_t)i2 * si;
        r[3 * s] = gb; r[2 * s] = gb * d2; r[1 * s] = gb * (d2 * d2); r[0] = gb * (d2 * d2 * d2);
    }
}

__global__ void deform_dirty_set_kernel(const int *__restrict__ idx1_dev, const int *__restrict__ idx2_dev, int *__restrict__ dirty) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int a = idx1_dev[0], b = idx2_dev[0];
        dirty[0] = a == b ? 1 : 2; dirty[1] = a; dirty[2] = b;
    }
}

__global__ void __launch_bounds__(kThreads)
deform_defer_kernel(int P, const float *__restrict__ g1, const float *__restrict__ g2, const int *__restrict__ idx1_dev,
                    const float *__restrict__ dist1_dev, const int *__restrict__ idx2_dev, const float *__restrict__ dist2_dev,
                    float *__restrict__ payload) {
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k < 3 * P) { payload[k] = g1[k]; payload[3 * P + k] = g2 ? g2[k] : 0.f; }
    if (k == 0) {
        float *tail = payload + 6 * (size_t)P;
        tail[0] = __int_as_float(idx1_dev[0]); tail[1] = dist1_dev[0];
        tail[2] = __int_as_float(idx2_dev[0]); tail[3] = dist2_dev[0];
    }
}

// Entries e = 2*rank + slot (<= 16).  Thread 0 of every CTA groups them by interval once (shared memory); each thread then
// owns one (Gaussian, xyz component): clears the dirty intervals nobody re-writes and accumulates g * {d^3, d^2, d, 1} per
// distinct interval in entry order.
__global__ void __launch_bounds__(kThreads)
deform_bwd_gathered_kernel(int P, int NI, int layout, int world, const float *__restrict__ gathered, long long stride, float scale,
                           int *__restrict__ dirty, float *__restrict__ dL_dcoeff) {
    __shared__ int s_idx[16], s_uniq[16], s_nu, s_nd, s_dirty[16];
    __shared__ unsigned s_members[16];
    __shared__ float s_dist[16];
    const int ne = 2 * world;
    __shared__ int s_draw[17];
    // the <= 16 frame scalars and the dirty list are fetched by 33 threads in parallel (one global round trip), ...
    if ((int)threadIdx.x < ne) {
        const int e = threadIdx.x;
        const float *tail = gathered + (e >> 1) * stride + 6 * (size_t)P + 2 * (e & 1);
        s_idx[e] = __float_as_int(tail[0]); s_dist[e] = tail[1];
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + 17) s_draw[threadIdx.x - 32] = dirty[threadIdx.x - 32];
    __syncthreads();
    if (threadIdx.x == 0) {   // ... grouped by interval from shared memory
        int nu = 0;
        for (int e = 0; e < ne; ++e) {
            const int b = s_idx[e];
            if (b < 0 || b >= NI) continue;
            int u = 0;
            while (u < nu && s_uniq[u] != b) ++u;
            if (u == nu) { s_uniq[nu] = b; s_members[nu] = 0u; ++nu; }
            s_members[u] |= 1u << e;
        }
        s_nu = nu;
        int nd = 0;
        const int cnt = min(s_draw[0], 16);
        for (int t = 0; t < cnt; ++t) {
            const int z = s_draw[1 + t];
            bool skip = z < 0 || z >= NI;                                 // invalid, re-written below, or already listed
            for (int u = 0; u < nu; ++u) skip |= s_uniq[u] == z;
            for (int u = 0; u < nd; ++u) skip |= s_dirty[u] == z;
            if (!skip) s_dirty[nd++] = z;
        }
        s_nd = nd;
    }
    __syncthreads();
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const size_t s = layout ? 3 : (size_t)NI * 3, si = layout ? 12 : 3;
    float *row = dL_dcoeff + (size_t)i * 4 * NI * 3 + c;
    for (int t = 0; t < s_nd; ++t) {
        float *q = row + (size_t)s_dirty[t] * si;
        q[0] = 0.f; q[s] = 0.f; q[2 * s] = 0.f; q[3 * s] = 0.f;
    }
    for (int u = 0; u < s_nu; ++u) {
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        unsigned m = s_members[u];
        while (m) {
            const int e = __ffs(m) - 1;
            m &= m - 1u;
            const float g = gathered[(e >> 1) * stride + (size_t)(e & 1) * 3 * P + k] * scale, d = s_dist[e];
            q3 += g; q2 += g * d; q1 += g * (d * d); q0 += g * (d * d * d);
        }
        float *q = row + (size_t)s_uniq[u] * si;
        q[3 * s] = q3; q[2 * s] = q2; q[1 * s] = q1; q[0] = q0;
    }
}

__global__ void deform_dirty_gathered_kernel(int P, int world, const float *__restrict__ gathered, long long stride,
                                             int *__restrict__ dirty) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int ne = 2 * world;
        int n = 0;
        for (int e = 0; e < ne; ++e) {
            const int b = __float_as_int(gathered[(e >> 1) * stride + 6 * (size_t)P + 2 * (e & 1)]);
            bool seen = false;
            for (int t = 0; t < n; ++t) seen |= dirty[1 + t] == b;
            if (!seen) dirty[1 + n++] = b;
        }
        dirty[0] = n;
    }
}

// Rotation of the active model at frame time t (dynamic_gaussian_with_base_point_cloud.py:184-198, get_rotation):
//   q = rotation + sum_k rot_poly_feat[:,k,:] * t^k + sum_l rot_fourier_feat[:,l,:] * basis_l(t)   (both feature sums are
//   .detach()ed there), returned through F.normalize.  basis = [t^0..t^3 | cos(t*pi*(1..4)) | sin(t*pi*(1..4))] is read from
//   DEVICE memory (12 floats) so a captured graph can be replayed for any frame.
__global__ void __launch_bounds__(kThreads)
deform_rot_fwd_kernel(int P, const float4 *__restrict__ rot, const float4 *__restrict__ poly /*[P,4]*/,
                      const float4 *__restrict__ fourier /*[P,8]*/, const float *__restrict__ basis, float4 *__restrict__ out,
                      float *__restrict__ inv_norm) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float4 q = rot[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 f = poly[4 * (size_t)i + k]; const float b = basis[k];
        q.x += f.x * b; q.y += f.y * b; q.z += f.z * b; q.w += f.w * b;
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        const float4 f = fourier[8 * (size_t)i + l]; const float b = basis[4 + l];
        q.x += f.x * b; q.y += f.y * b; q.z += f.z * b; q.w += f.w * b;
    }
    const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);   // F.normalize eps
    const float r = 1.0f / n;
    out[i] = make_float4(q.x * r, q.y * r, q.z * r, q.w * r);
    inv_norm[i] = r;
}

__global__ void __launch_bounds__(kThreads)
deform_rot_bwd_kernel(int P, const float4 *__restrict__ qhat, const float *__restrict__ inv_norm,
                      const float4 *__restrict__ g, float4 *__restrict__ g_rot) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float4 q = qhat[i], d = g[i];
    const float dot = q.x * d.x + q.y * d.y + q.z * d.z + q.w * d.w, r = inv_norm[i];
    g_rot[i] = make_float4((d.x - q.x * dot) * r, (d.y - q.y * dot) * r, (d.z - q.z * dot) * r, (d.w - q.w * dot) * r);
}
// ---- polynomial + Fourier position of the alternative model (dynamic_gaussian_points.py:170-186) ---------------------------------
//   pos(t) = position + sum_k pos_poly_feat[:, k, :] * t^k + sum_k pos_fourier_feat[:, k, :] * {cos, sin}(t * pi * (k mod 4 + 1))
// = position + feat[P, 12, 3] . basis[12]: the only dense contraction on the path (K = 12) -- a bandwidth-bound
// stream (156 B read, 12 B written per Gaussian), one thread per (Gaussian, component).  `basis` (t^0..t^3 | cos | sin) is read from
// device memory so a captured graph replays for any frame.  Sum order = the reference's: position + (poly 0..3) + (fourier 0..7).
__global__ void __launch_bounds__(kThreads)
deform_pf_fwd_kernel(int P, const float *__restrict__ position, const float *__restrict__ poly /*[P,4,3]*/,
                     const float *__restrict__ fourier /*[P,8,3]*/, const float *__restrict__ basis, float *__restrict__ pos) {
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const float *pf = poly + (size_t)i * 12 + c, *ff = fourier + (size_t)i * 24 + c;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) a += pf[3 * q] * basis[q];        // torch.sum(feat * basis, dim=1): ascending k
#pragma unroll
    for (int q = 0; q < 8; ++q) b += ff[3 * q] * basis[4 + q];
    pos[k] = (position[k] + a) + b;
}

__global__ void __launch_bounds__(kThreads)
deform_pf_bwd_kernel(int P, const float *__restrict__ basis, const float *__restrict__ dL_dpos, float *__restrict__ dL_dposition,
                     float *__restrict__ dL_dpoly, float *__restrict__ dL_dfourier) {
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= 3 * P) return;
    const int i = k / 3, c = k % 3;
    const float g = dL_dpos[k];
    if (dL_dposition) dL_dposition[k] = g;
    if (dL_dpoly) {
        float *o = dL_dpoly + (size_t)i * 12 + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) o[3 * q] = g * basis[q];
    }
    if (dL_dfourier) {
        float *o = dL_dfourier + (size_t)i * 24 + c;
#pragma unroll
        for (int q = 0; q < 8; ++q) o[3 * q] = g * basis[4 + q];
    }
}

}  // namespace

extern "C" {

int spv_deform_spline_forward(int P, int NI, int layout, const float *base, const float *coeff, const int *idx_dev,
                              const float *dist_dev, float *pos, void *stream) {
    if (P <= 0) return 0;
    deform_fwd_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, NI, layout, base, coeff, idx_dev, dist_dev, pos);
    return spv::check_launch("spv_deform_spline_forward");
}

/* accumulate == 0: dL_dcoeff is cleared first (only the active interval's 12 floats per Gaussian are non-zero). */
int spv_deform_spline_backward(int P, int NI, int layout, const int *idx_dev, const float *dist_dev, const float *dL_dpos,
                               float *dL_dcoeff, int accumulate, void *stream) {
    if (P <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) SPV_CUDA_TRY(cudaMemsetAsync(dL_dcoeff, 0, sizeof(float) * 12 * (size_t)NI * P, s), "spv_deform_spline_backward");
    deform_bwd_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, s>>>(P, NI, layout, idx_dev, dist_dev, dL_dpos, dL_dcoeff, accumulate);
    return spv::check_launch("spv_deform_spline_backward");
}

int spv_deform_spline_forward2(int P, int NI, int layout, const float *base, const float *coeff, const int *idx1_dev, const float *dist1_dev,
                               const int *idx2_dev, const float *dist2_dev, float *pos1, float *pos2, void *stream) {
    if (P <= 0) return 0;
    deform_fwd2_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, NI, layout, base, coeff, idx1_dev, dist1_dev,
                                                                                           idx2_dev, dist2_dev, pos1, pos2);
    return spv::check_launch("spv_deform_spline_forward2");
}

/* dL_dcoeff is a sink that this call keeps clean through `dirty` (int[17] on the device, zero-initialised by the caller
 * together with the sink): intervals listed there are zeroed unless re-written, then the list becomes {idx1, idx2}. */
int spv_deform_spline_backward2(int P, int NI, int layout, const int *idx1_dev, const float *dist1_dev, const int *idx2_dev,
                                const float *dist2_dev, const float *dL_dpos1, const float *dL_dpos2 /*or NULL*/, int *dirty,
                                float *dL_dcoeff, void *stream) {
    if (P <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    deform_bwd2_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, s>>>(P, NI, layout, idx1_dev, dist1_dev, idx2_dev, dist2_dev, dL_dpos1,
                                                                        dL_dpos2, dirty, dL_dcoeff);
    deform_dirty_set_kernel<<<1, 32, 0, s>>>(idx1_dev, dL_dpos2 ? idx2_dev : idx1_dev, dirty);
    return spv::check_launch("spv_deform_spline_backward2", 2);
}

int spv_deform_defer(int P, const float *dL_dpos1, const float *dL_dpos2, const int *idx1_dev, const float *dist1_dev,
                     const int *idx2_dev, const float *dist2_dev, float *payload, void *stream) {
    if (P <= 0) return 0;
    deform_defer_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, dL_dpos1, dL_dpos2, idx1_dev, dist1_dev,
                                                                                            idx2_dev, dist2_dev, payload);
    return spv::check_launch("spv_deform_defer");
}

int spv_deform_spline_backward_gathered(int P, int NI, int layout, int world, const float *gathered, long long stride, float scale,
                                        int *dirty, float *dL_dcoeff, void *stream) {
    if (P <= 0) return 0;
    if (world < 1 || world > 8) { spv::set_error(cudaErrorInvalidValue, "spv_deform_spline_backward_gathered: 1..8 ranks"); return (int)cudaErrorInvalidValue; }
    cudaStream_t s = (cudaStream_t)stream;
    deform_bwd_gathered_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, s>>>(P, NI, layout, world, gathered, stride, scale, dirty, dL_dcoeff);
    deform_dirty_gathered_kernel<<<1, 32, 0, s>>>(P, world, gathered, stride, dirty);
    return spv::check_launch("spv_deform_spline_backward_gathered", 2);
}

int spv_deform_rotation_forward(int P, const float *rotation, const float *rot_poly_feat, const float *rot_fourier_feat,
                                const float *basis_dev /*[12]*/, float *out /*[P,4]*/, float *inv_norm /*[P]*/, void *stream) {
    if (P <= 0) return 0;
    deform_rot_fwd_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        P, (const float4 *)rotation, (const float4 *)rot_poly_feat, (const float4 *)rot_fourier_feat, basis_dev, (float4 *)out, inv_norm);
    return spv::check_launch("spv_deform_rotation_forward");
}

int spv_deform_rotation_backward(int P, const float *out, const float *inv_norm, const float *dL_dout, float *dL_drotation,
                                 void *stream) {
    if (P <= 0) return 0;
    deform_rot_bwd_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, (const float4 *)out, inv_norm,
                                                                                     (const float4 *)dL_dout, (float4 *)dL_drotation);
    return spv::check_launch("spv_deform_rotation_backward");
}

int spv_deform_polyfourier_forward(int P, const float *position, const float *pos_poly_feat, const float *pos_fourier_feat,
                                   const float *basis_dev /*[12]*/, float *pos /*[P,3]*/, void *stream) {
    if (P <= 0) return 0;
    deform_pf_fwd_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, position, pos_poly_feat, pos_fourier_feat,
                                                                                             basis_dev, pos);
    return spv::check_launch("spv_deform_polyfourier_forward");
}

/* Gradients are WRITTEN (not accumulated); any of the three outputs may be NULL (e.g. dL_dposition under detach_pos). */
int spv_deform_polyfourier_backward(int P, const float *basis_dev, const float *dL_dpos, float *dL_dposition, float *dL_dpos_poly_feat,
                                    float *dL_dpos_fourier_feat, void *stream) {
    if (P <= 0) return 0;
    deform_pf_bwd_kernel<<<spv::cdiv(3ll * P, kThreads), kThreads, 0, (cudaStream_t)stream>>>(P, basis_dev, dL_dpos, dL_dposition,
                                                                                             dL_dpos_poly_feat, dL_dpos_fourier_feat);
    return spv::check_launch("spv_deform_polyfourier_backward");
}

}  // extern "C"
