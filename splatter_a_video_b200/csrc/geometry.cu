// Per-Gaussian stages of the rasterizer for sm_100a: projection (perspective + orthographic), 3D covariance,
// EWA splat (perspective + orthographic), spherical harmonics -- forward and backward.
//
// These kernels decide DISCRETE outputs (radius, tile rectangles, cull masks), which must be bit-exact against
// the oracle, so this translation unit is compiled with -fmad=false and uses only IEEE fp32 operations
// (div.rn / sqrt.rn); the arithmetic order restates the reference kernels cited per function
// (/root/reference/src/submodules/dptr/dptr/gs/src/*.cu).  They are pure HBM streaming passes
// (280 B/Gaussian forward), one thread per Gaussian, 256-thread CTAs, grid rounded up
// to whole waves of the 148 SMs by the launcher.
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {

constexpr int kThreads = 256;

struct M3 { float m[3][3]; };  // m[col][row]: column-major like the reference's glm::mat3

__device__ __forceinline__ M3 mul(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w)
            r.m[c][w] = a.m[0][w] * b.m[c][0] + a.m[1][w] * b.m[c][1] + a.m[2][w] * b.m[c][2];
    return r;
}
__device__ __forceinline__ M3 transpose(const M3 &a) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w) r.m[c][w] = a.m[w][c];
    return r;
}
__device__ __forceinline__ M3 quat_to_R(const float4 q) {  // compute_cov3d.cu:24-40, q = (r,x,y,z)
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}
__device__ __forceinline__ M3 scale_to_S(float sx, float sy, float sz) {
    M3 S;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w) S.m[c][w] = 0.f;
    S.m[0][0] = sx; S.m[1][1] = sy; S.m[2][2] = sz;
    return S;
}
__device__ __forceinline__ M3 cov6_to_M3(const float *c) {
    M3 V;
    V.m[0][0] = c[0]; V.m[0][1] = c[1]; V.m[0][2] = c[2];
    V.m[1][0] = c[1]; V.m[1][1] = c[3]; V.m[1][2] = c[4];
    V.m[2][0] = c[2]; V.m[2][1] = c[4]; V.m[2][2] = c[5];
    return V;
}

// ------------------------------------------------------------------------------------------------ K1/K2
__global__ void __launch_bounds__(kThreads)
project_point_fwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ intr,
                         const float *__restrict__ extr, int W, int H, float nearest, float extent,
                         float2 *__restrict__ uv, float *__restrict__ depth) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    const float tx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];
    const float ty = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
    const float tz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
    // project_point.cu:31,34-35: the reciprocal and the -0.5 are double-precision expressions there.
    const float norm1 = (float)(1.0 / ((double)tz + 1e-7));
    const float u = (float)((double)(intr[0] * tx * norm1 + intr[2]) - 0.5);
    const float v = (float)((double)(intr[1] * ty * norm1 + intr[3]) - 0.5);
    bool cull = false;
    if (nearest > 0) cull = tz <= nearest;
    if (extent > 0) {
        const float xmin = (float)((double)((1 - extent) * W) * 0.5), xmax = (float)((double)((1 + extent) * W) * 0.5);
        const float ymin = (float)((double)((1 - extent) * H) * 0.5), ymax = (float)((double)((1 + extent) * H) * 0.5);
        cull = cull || u < xmin || u > xmax || v < ymin || v > ymax;
    }
    uv[i] = cull ? make_float2(0.f, 0.f) : make_float2(u, v);
    depth[i] = cull ? 0.f : tz;
}

__global__ void __launch_bounds__(kThreads)
project_point_bwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ intr,
                         const float *__restrict__ extr, const float *__restrict__ depth,
                         const float2 *__restrict__ dL_duv, const float *__restrict__ dL_ddepth,
                         float *__restrict__ dL_dxyz, float *__restrict__ dL_dintr, float *__restrict__ dL_dextr) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const bool live = i < P && depth[i] != 0.f;
    float gi[4] = {0.f, 0.f, 0.f, 0.f};
    float ge[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) ge[k] = 0.f;
    if (i < P) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (live) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            const float tx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];
            const float ty = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
            const float tz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
            const float norm1 = (float)(1.0 / (double)tz);
            const float norm2 = (float)(1.0 / (double)(tz * tz));
            const float gu = dL_duv[i].x, gv = dL_duv[i].y, gd = dL_ddepth[i];
            gx += (intr[0] * (extr[0] * tz - tx * extr[8]) * norm2) * gu;
            gx += (intr[1] * (extr[4] * tz - ty * extr[8]) * norm2) * gv;
            gx += extr[8] * gd;
            gy += (intr[0] * (extr[1] * tz - tx * extr[9]) * norm2) * gu;
            gy += (intr[1] * (extr[5] * tz - ty * extr[9]) * norm2) * gv;
            gy += extr[9] * gd;
            gz += (intr[0] * (extr[2] * tz - tx * extr[10]) * norm2) * gu;
            gz += (intr[1] * (extr[6] * tz - ty * extr[10]) * norm2) * gv;
            gz += extr[10] * gd;
            if (dL_dintr) { gi[0] = tx * norm1 * gu; gi[1] = ty * norm1 * gv; gi[2] = gu; gi[3] = gv; }
            if (dL_dextr) {
                ge[0] = intr[0] * px * norm1 * gu; ge[1] = intr[0] * py * norm1 * gu;
                ge[2] = intr[0] * pz * norm1 * gu; ge[3] = intr[0] * norm1 * gu;
                ge[4] = intr[1] * px * norm1 * gv; ge[5] = intr[1] * py * norm1 * gv;
                ge[6] = intr[1] * pz * norm1 * gv; ge[7] = intr[1] * norm1 * gv;
                ge[8] = -intr[0] * px * tx * norm2 * gu + -intr[1] * px * ty * norm2 * gv + px * gd;
                ge[9] = -intr[0] * py * tx * norm2 * gu + -intr[1] * py * ty * norm2 * gv + py * gd;
                ge[10] = -intr[0] * pz * tx * norm2 * gu + -intr[1] * pz * ty * norm2 * gv + pz * gd;
                ge[11] = -intr[0] * tx * norm2 * gu + -intr[1] * ty * norm2 * gv + gd;
            }
        }
        dL_dxyz[3 * i] = gx; dL_dxyz[3 * i + 1] = gy; dL_dxyz[3 * i + 2] = gz;
    }
    // Camera gradients: warp-shuffle reduction, one atomic per warp (the reference issues one per thread).
    if (dL_dintr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = gi[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dL_dintr + k, v);
        }
    }
    if (dL_dextr) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float v = ge[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dL_dextr + k, v);
        }
    }
}

// Orthographic projection: pointrix/renderer/dptr_ortho_enhanced.py:177-202.  The bodies are device functions so the fused
// per-frame kernels at the end of this file execute exactly the same operations as the staged kernels.
struct OrthoBounds { float nearest, xmin, xmax, ymin, ymax; };

__device__ __forceinline__ void project_ortho_body(float px, float py, float pz, const float *__restrict__ extr, int W, int H,
                                                   const OrthoBounds &bd, float2 &uv, float &depth) {
    const float cx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];
    const float cy = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
    const float cz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
    const float u = (cx + 1.0f) * (float)W / 2.0f - 0.5f;
    const float v = (cy + 1.0f) * (float)H / 2.0f - 0.5f;
    float d = cz;
    if (isnan(d)) d = 0.0f;                                              // nan_to_num
    else if (isinf(d)) d = d > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    const bool mask = (d <= bd.nearest) || (u < bd.xmin) || (u > bd.xmax) || (v < bd.ymin) || (v > bd.ymax);
    uv = mask ? make_float2(0.f, 0.f) : make_float2(u, v);
    depth = mask ? 0.f : d;
}

__device__ __forceinline__ void project_ortho_bwd_body(const float *__restrict__ extr, float hw, float hh, float depth, float2 guv,
                                                       float gdepth, float &gx, float &gy, float &gz) {
    gx = 0.f; gy = 0.f; gz = 0.f;
    if (depth != 0.f) {  // masked rows were overwritten with 0 in the forward: no gradient
        const float a = guv.x * hw, b = guv.y * hh, c = gdepth;
        gx = extr[0] * a + extr[4] * b + extr[8] * c;
        gy = extr[1] * a + extr[5] * b + extr[9] * c;
        gz = extr[2] * a + extr[6] * b + extr[10] * c;
    }
}

__global__ void __launch_bounds__(kThreads)
project_point_ortho_fwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ extr, int W, int H,
                               float nearest, float xmin, float xmax, float ymin, float ymax,
                               float2 *__restrict__ uv, float *__restrict__ depth) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const OrthoBounds bd{nearest, xmin, xmax, ymin, ymax};
    float2 o;
    float d;
    project_ortho_body(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], extr, W, H, bd, o, d);
    uv[i] = o;
    depth[i] = d;
}

__global__ void __launch_bounds__(kThreads)
project_point_ortho_bwd_kernel(int P, const float *__restrict__ extr, float hw, float hh,
                               const float *__restrict__ depth, const float2 *__restrict__ dL_duv,
                               const float *__restrict__ dL_ddepth, float *__restrict__ dL_dxyz) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float gx, gy, gz;
    project_ortho_bwd_body(extr, hw, hh, depth[i], dL_duv[i], dL_ddepth[i], gx, gy, gz);
    dL_dxyz[3 * i] = gx; dL_dxyz[3 * i + 1] = gy; dL_dxyz[3 * i + 2] = gz;
}

// ------------------------------------------------------------------------------------------------ K3/K4
__device__ __forceinline__ void cov3d_body(float sx, float sy, float sz, float4 q, bool visible, float (&c)[6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = 0.f;
    if (visible) {
        const M3 M = mul(scale_to_S(sx, sy, sz), quat_to_R(q));
        const M3 Sg = mul(transpose(M), M);
        c[0] = Sg.m[0][0]; c[1] = Sg.m[0][1]; c[2] = Sg.m[0][2]; c[3] = Sg.m[1][1]; c[4] = Sg.m[1][2]; c[5] = Sg.m[2][2];
    }
}

__device__ __forceinline__ void cov3d_bwd_body(const float (&s)[3], float4 q, bool visible, const float *g, float (&gs)[3], float4 &gq) {
    gs[0] = gs[1] = gs[2] = 0.f;
    gq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (visible) {
        const M3 R = quat_to_R(q);
        const M3 M = mul(scale_to_S(s[0], s[1], s[2]), R);
        M3 dS;  // compute_cov3d.cu:69-77
        dS.m[0][0] = g[0]; dS.m[0][1] = 0.5f * g[1]; dS.m[0][2] = 0.5f * g[2];
        dS.m[1][0] = 0.5f * g[1]; dS.m[1][1] = g[3]; dS.m[1][2] = 0.5f * g[4];
        dS.m[2][0] = 0.5f * g[2]; dS.m[2][1] = 0.5f * g[4]; dS.m[2][2] = g[5];
        M3 M2;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int w = 0; w < 3; ++w) M2.m[c][w] = 2.0f * M.m[c][w];
        const M3 dM = mul(M2, dS);
        const M3 Rt = transpose(R);
        M3 dMt = transpose(dM);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            gs[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int w = 0; w < 3; ++w) dMt.m[k][w] *= s[k];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
#define D(a, b) dMt.m[a][b]
        gq.x = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
        gq.y = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) -
               4 * x * (D(2, 2) + D(1, 1));
        gq.z = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) -
               4 * y * (D(2, 2) + D(0, 0));
        gq.w = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) -
               4 * z * (D(1, 1) + D(0, 0));
#undef D
    }
}

__global__ void __launch_bounds__(kThreads)
cov3d_fwd_kernel(int P, const float *__restrict__ scales, const float4 *__restrict__ uquats,
                 const uint8_t *__restrict__ visible, float *__restrict__ cov3d) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float c[6];
    cov3d_body(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], uquats[i], visible[i] != 0, c);
    float2 *o = reinterpret_cast<float2 *>(cov3d + 6 * (size_t)i);
    o[0] = make_float2(c[0], c[1]); o[1] = make_float2(c[2], c[3]); o[2] = make_float2(c[4], c[5]);
}

__global__ void __launch_bounds__(kThreads)
cov3d_bwd_kernel(int P, const float *__restrict__ scales, const float4 *__restrict__ uquats,
                 const uint8_t *__restrict__ visible, const float *__restrict__ dL_dcov3d,
                 float *__restrict__ dL_dscales, float4 *__restrict__ dL_duquats) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    float gs[3];
    float4 gq;
    cov3d_bwd_body(s, uquats[i], visible[i] != 0, dL_dcov3d + 6 * (size_t)i, gs, gq);
    dL_dscales[3 * i] = gs[0]; dL_dscales[3 * i + 1] = gs[1]; dL_dscales[3 * i + 2] = gs[2];
    dL_duquats[i] = gq;
}

// ------------------------------------------------------------------------------------------------ K5/K6
struct EwaFrame { M3 T, J, Wm; float t[3]; };

__device__ __forceinline__ EwaFrame ewa_frame(const float *p, const float *intr, const float *extr) {
    EwaFrame f;
    const float fx = intr[0], fy = intr[1];
    f.t[0] = extr[0] * p[0] + extr[1] * p[1] + extr[2] * p[2] + extr[3];
    f.t[1] = extr[4] * p[0] + extr[5] * p[1] + extr[6] * p[2] + extr[7];
    f.t[2] = extr[8] * p[0] + extr[9] * p[1] + extr[10] * p[2] + extr[11];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int w = 0; w < 3; ++w) f.J.m[c][w] = 0.f;
    f.J.m[0][0] = fx / f.t[2]; f.J.m[1][1] = fy / f.t[2];
    f.J.m[2][0] = -(fx * f.t[0]) / (f.t[2] * f.t[2]); f.J.m[2][1] = -(fy * f.t[1]) / (f.t[2] * f.t[2]);
    f.Wm.m[0][0] = extr[0]; f.Wm.m[0][1] = extr[4]; f.Wm.m[0][2] = extr[8];
    f.Wm.m[1][0] = extr[1]; f.Wm.m[1][1] = extr[5]; f.Wm.m[1][2] = extr[9];
    f.Wm.m[2][0] = extr[2]; f.Wm.m[2][1] = extr[6]; f.Wm.m[2][2] = extr[10];
    f.T = mul(f.J, f.Wm);
    return f;
}

__global__ void __launch_bounds__(kThreads)
ewa_fwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ cov3d, const float *__restrict__ intr,
               const float *__restrict__ extr, const float2 *__restrict__ uv, int gx, int gy,
               const uint8_t *__restrict__ visible, float *__restrict__ conic, int *__restrict__ radius,
               int *__restrict__ tiles) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float k0 = 0.f, k1 = 0.f, k2 = 0.f;
    int rad = 0, nt = 0;
    if (visible[i]) {
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const EwaFrame f = ewa_frame(p, intr, extr);
        const M3 cov2 = mul(mul(f.T, cov6_to_M3(cov3d + 6 * (size_t)i)), transpose(f.T));
        const float cx = cov2.m[0][0] + 0.3f, cy = cov2.m[0][1], cz = cov2.m[1][1] + 0.3f;
        const float det = cx * cz - cy * cy;
        if (det != 0.0f) {
            const float mid = 0.5f * (cx + cz);
            const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
            const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
            const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
            int x0, y0, x1, y1;
            const float2 c = uv[i];
            spv::tile_rect(c.x, c.y, (int)my_radius, gx, gy, x0, y0, x1, y1);
            if ((x1 - x0) * (y1 - y0) != 0) {
                const float det_inv = 1.f / det;
                k0 = cz * det_inv; k1 = -cy * det_inv; k2 = cx * det_inv;
                rad = (int)my_radius;
                nt = (y1 - y0) * (x1 - x0);
            }
        }
    }
    conic[3 * i] = k0; conic[3 * i + 1] = k1; conic[3 * i + 2] = k2;
    radius[i] = rad;
    tiles[i] = nt;
}

__device__ __forceinline__ void conic_grad_to_cov2d(float cx, float cy, float cz, float det, const float *g,
                                                    float &dcx, float &dcy, float &dcz) {
    const float nom = 1.0f / (det * det);  // ewa_project.cu:135-143
    dcx = nom * (-cz * cz * g[0] + cy * cz * g[1] + (det - cx * cz) * g[2]);
    dcy = nom * (2 * cy * cz * g[0] - (det + 2 * cy * cy) * g[1] + 2 * cx * cy * g[2]);
    dcz = nom * ((det - cx * cz) * g[0] + cx * cy * g[1] - cx * cx * g[2]);
}

__global__ void __launch_bounds__(kThreads)
ewa_bwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ cov3d, const float *__restrict__ intr,
               const float *__restrict__ extr, const int *__restrict__ radius, const float *__restrict__ dL_dconic,
               float *__restrict__ dL_dxyz, float *__restrict__ dL_dcov3d, float *__restrict__ dL_dintr,
               float *__restrict__ dL_dextr) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float gp[3] = {0.f, 0.f, 0.f};
    float gi[2] = {0.f, 0.f};
    float ge[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) ge[k] = 0.f;
    if (i < P && radius[i] > 0) {
        const float fx = intr[0], fy = intr[1];
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const float *c3 = cov3d + 6 * (size_t)i;
        const float *g = dL_dconic + 3 * (size_t)i;
        const EwaFrame f = ewa_frame(p, intr, extr);
        const M3 &T = f.T;
        const M3 cov2 = mul(mul(T, cov6_to_M3(c3)), transpose(T));
        const float cx = cov2.m[0][0] + 0.3f, cy = cov2.m[0][1], cz = cov2.m[1][1] + 0.3f;
        const float det = cx * cz - cy * cy;
        if (det != 0.0f) {
            float dcx, dcy, dcz;
            conic_grad_to_cov2d(cx, cy, cz, det, g, dcx, dcy, dcz);
#define TT(a, b) T.m[a][b]
            o[0] += TT(0, 0) * TT(0, 0) * dcx; o[0] += TT(0, 0) * TT(0, 1) * dcy; o[0] += TT(0, 1) * TT(0, 1) * dcz;
            o[1] += 2 * TT(0, 0) * TT(1, 0) * dcx; o[1] += (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dcy; o[1] += 2 * TT(0, 1) * TT(1, 1) * dcz;
            o[2] += 2 * TT(0, 0) * TT(2, 0) * dcx; o[2] += (TT(0, 0) * TT(2, 1) + TT(0, 1) * TT(2, 0)) * dcy; o[2] += 2 * TT(0, 1) * TT(2, 1) * dcz;
            o[3] += TT(1, 0) * TT(1, 0) * dcx; o[3] += TT(1, 0) * TT(1, 1) * dcy; o[3] += TT(1, 1) * TT(1, 1) * dcz;
            o[4] += 2 * TT(1, 0) * TT(2, 0) * dcx; o[4] += (TT(1, 0) * TT(2, 1) + TT(1, 1) * TT(2, 0)) * dcy; o[4] += 2 * TT(1, 1) * TT(2, 1) * dcz;
            o[5] += TT(2, 0) * TT(2, 0) * dcx; o[5] += TT(2, 0) * TT(2, 1) * dcy; o[5] += TT(2, 1) * TT(2, 1) * dcz;
            float dT00 = 0, dT01 = 0, dT10 = 0, dT11 = 0, dT20 = 0, dT21 = 0;
            dT00 += 2 * (TT(0, 0) * c3[0] + TT(1, 0) * c3[1] + TT(2, 0) * c3[2]) * dcx;
            dT00 += (TT(0, 1) * c3[0] + TT(1, 1) * c3[1] + TT(2, 1) * c3[2]) * dcy;
            dT01 += (TT(0, 0) * c3[0] + TT(1, 0) * c3[1] + TT(2, 0) * c3[2]) * dcy;
            dT01 += 2 * (TT(0, 1) * c3[0] + TT(1, 1) * c3[1] + TT(2, 1) * c3[2]) * dcz;
            dT10 += 2 * (TT(0, 0) * c3[1] + TT(1, 0) * c3[3] + TT(2, 0) * c3[4]) * dcx;
            dT10 += (TT(0, 1) * c3[1] + TT(1, 1) * c3[3] + TT(2, 1) * c3[4]) * dcy;
            dT11 += (TT(0, 0) * c3[1] + TT(1, 0) * c3[3] + TT(2, 0) * c3[4]) * dcy;
            dT11 += 2 * (TT(0, 1) * c3[1] + TT(1, 1) * c3[3] + TT(2, 1) * c3[4]) * dcz;
            dT20 += 2 * (TT(0, 0) * c3[2] + TT(1, 0) * c3[4] + TT(2, 0) * c3[5]) * dcx;
            dT20 += (TT(0, 1) * c3[2] + TT(1, 1) * c3[4] + TT(2, 1) * c3[5]) * dcy;
            dT21 += (TT(0, 0) * c3[2] + TT(1, 0) * c3[4] + TT(2, 0) * c3[5]) * dcy;
            dT21 += 2 * (TT(0, 1) * c3[2] + TT(1, 1) * c3[4] + TT(2, 1) * c3[5]) * dcz;
#undef TT
#define WW(a, b) f.Wm.m[a][b]
            const float dJ00 = WW(0, 0) * dT00 + WW(1, 0) * dT10 + WW(2, 0) * dT20;
            const float dJ20 = WW(0, 2) * dT00 + WW(1, 2) * dT10 + WW(2, 2) * dT20;
            const float dJ11 = WW(0, 1) * dT01 + WW(1, 1) * dT11 + WW(2, 1) * dT21;
            const float dJ21 = WW(0, 2) * dT01 + WW(1, 2) * dT11 + WW(2, 2) * dT21;
#undef WW
            const float tz = 1.f / f.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
            const float dtx = -fx * tz2 * dJ20;
            const float dty = -fy * tz2 * dJ21;
            const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * f.t[0]) * tz3 * dJ20 +
                              (2 * fy * f.t[1]) * tz3 * dJ21;
            if (dL_dintr) {
                gi[0] = tz * dJ00 + -f.t[0] * tz2 * dJ20;
                gi[1] = tz * dJ11 + -f.t[1] * tz2 * dJ21;
            }
            if (dL_dextr) {
                ge[0] = f.J.m[0][0] * dT00 + p[0] * dtx; ge[1] = f.J.m[0][0] * dT10 + p[1] * dtx;
                ge[2] = f.J.m[0][0] * dT20 + p[2] * dtx; ge[3] = dtx;
                ge[4] = f.J.m[1][1] * dT01 + p[0] * dty; ge[5] = f.J.m[1][1] * dT11 + p[1] * dty;
                ge[6] = f.J.m[1][1] * dT21 + p[2] * dty; ge[7] = dty;
                ge[8] = (f.J.m[2][0] * dT00 + f.J.m[2][1] * dT01) + p[0] * dtz;
                ge[9] = (f.J.m[2][0] * dT10 + f.J.m[2][1] * dT11) + p[1] * dtz;
                ge[10] = (f.J.m[2][0] * dT20 + f.J.m[2][1] * dT21) + p[2] * dtz;
                ge[11] = dtz;
            }
            gp[0] = extr[0] * dtx + extr[4] * dty + extr[8] * dtz;
            gp[1] = extr[1] * dtx + extr[5] * dty + extr[9] * dtz;
            gp[2] = extr[2] * dtx + extr[6] * dty + extr[10] * dtz;
        }
    }
    if (i < P) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dL_dcov3d[6 * (size_t)i + k] = o[k];
        dL_dxyz[3 * i] = gp[0]; dL_dxyz[3 * i + 1] = gp[1]; dL_dxyz[3 * i + 2] = gp[2];
    }
    if (dL_dintr) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float v = gi[k];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dL_dintr + k, v);
        }
    }
    if (dL_dextr) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            float v = ge[k];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dL_dextr + k, v);
        }
    }
}

// Orthographic EWA (dptr_ortho_enhanced.py:26-111): T = [[W/2,0,0],[0,H/2,0]] @ R is per-frame constant.
struct OrthoT { float t[2][3]; };

__device__ __forceinline__ OrthoT ortho_T(const float *extr, float jx, float jy) {
    OrthoT o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.t[0][k] = jx * extr[k] + 0.0f * extr[4 + k] + 0.0f * extr[8 + k];
        o.t[1][k] = 0.0f * extr[k] + jy * extr[4 + k] + 0.0f * extr[8 + k];
    }
    return o;
}

__device__ __forceinline__ void ortho_cov2d(const OrthoT &T, const float *c, float &c00, float &c01, float &c11) {
    const float S[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    float M[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) M[a][k] = T.t[a][0] * S[0][k] + T.t[a][1] * S[1][k] + T.t[a][2] * S[2][k];
    c00 = (M[0][0] * T.t[0][0] + M[0][1] * T.t[0][1] + M[0][2] * T.t[0][2]) + 0.3f;
    c01 = M[0][0] * T.t[1][0] + M[0][1] * T.t[1][1] + M[0][2] * T.t[1][2];
    c11 = (M[1][0] * T.t[1][0] + M[1][1] * T.t[1][1] + M[1][2] * T.t[1][2]) + 0.3f;
}

__device__ __forceinline__ void ewa_ortho_body(const float *c6, const float *__restrict__ extr, float jx, float jy, float2 c, int gx,
                                               int gy, bool visible, float (&conic)[3], int &radius, int &tiles) {
    const OrthoT T = ortho_T(extr, jx, jy);
    float c00, c01, c11;
    ortho_cov2d(T, c6, c00, c01, c11);
    const float det = c00 * c11 - c01 * c01;
    const float k0 = c11 / det, k1 = -c01 / det, k2 = c00 / det;
    const float b = (c00 + c11) / 2.0f;
    float disc = b * b - det;
    if (disc < 0.1f) disc = 0.1f;
    const float v1 = b + sqrtf(disc), v2 = b - sqrtf(disc);
    const float rad = ceilf(3.0f * sqrtf(v1 > v2 ? v1 : v2));
    const float f0 = (c.x - rad) / 16.0f, f1 = (c.y - rad) / 16.0f;
    const float f2 = (c.x + rad + 16.0f - 1.0f) / 16.0f, f3 = (c.y + rad + 16.0f - 1.0f) / 16.0f;
    const bool finite = isfinite(f0) && isfinite(f1) && isfinite(f2) && isfinite(f3) && isfinite(k0) && isfinite(k1) &&
                        isfinite(k2) && isfinite(rad) && fabsf(f0) < 2.0e9f && fabsf(f1) < 2.0e9f &&
                        fabsf(f2) < 2.0e9f && fabsf(f3) < 2.0e9f;
    int nt = 0;
    if (finite) {
        const int x0 = min(max((int)f0, 0), gx), y0 = min(max((int)f1, 0), gy);
        const int x1 = min(max((int)f2, 0), gx), y1 = min(max((int)f3, 0), gy);
        nt = (x1 - x0) * (y1 - y0);
    }
    const bool mask = finite && nt != 0 && det != 0.0f && visible;
    conic[0] = mask ? k0 : 0.f; conic[1] = mask ? k1 : 0.f; conic[2] = mask ? k2 : 0.f;
    radius = mask ? (int)rad : 0;
    tiles = mask ? nt : 0;
}

// Gradient of the ortho EWA w.r.t. cov3d only (J is constant, so xyz receives nothing): what torch autograd
// produces for dptr_ortho_enhanced.py:42-63,107.
__device__ __forceinline__ void ewa_ortho_bwd_body(const float *c6, const float *__restrict__ extr, float jx, float jy, int radius,
                                                   const float *g_conic, float (&o)[6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = 0.f;
    if (radius > 0) {
        const OrthoT T = ortho_T(extr, jx, jy);
        float c00, c01, c11;
        ortho_cov2d(T, c6, c00, c01, c11);
        const float det = c00 * c11 - c01 * c01;
        float d00, d01, d11;
        conic_grad_to_cov2d(c00, c01, c11, det, g_conic, d00, d01, d11);
        // dL/dS_jk = d00 T0j T0k + d01 T0j T1k + d11 T1j T1k ; symmetric entries of the 6-vector add up.
#define A(j) T.t[0][j]
#define B(j) T.t[1][j]
        o[0] = d00 * A(0) * A(0) + d01 * A(0) * B(0) + d11 * B(0) * B(0);
        o[1] = 2 * d00 * A(0) * A(1) + d01 * (A(0) * B(1) + A(1) * B(0)) + 2 * d11 * B(0) * B(1);
        o[2] = 2 * d00 * A(0) * A(2) + d01 * (A(0) * B(2) + A(2) * B(0)) + 2 * d11 * B(0) * B(2);
        o[3] = d00 * A(1) * A(1) + d01 * A(1) * B(1) + d11 * B(1) * B(1);
        o[4] = 2 * d00 * A(1) * A(2) + d01 * (A(1) * B(2) + A(2) * B(1)) + 2 * d11 * B(1) * B(2);
        o[5] = d00 * A(2) * A(2) + d01 * A(2) * B(2) + d11 * B(2) * B(2);
#undef A
#undef B
    }
}

__global__ void __launch_bounds__(kThreads)
ewa_ortho_fwd_kernel(int P, const float *__restrict__ cov3d, const float *__restrict__ extr, float jx, float jy,
                     const float2 *__restrict__ uv, int gx, int gy, const uint8_t *__restrict__ visible,
                     float *__restrict__ conic, int *__restrict__ radius, int *__restrict__ tiles) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float k[3];
    int rad, nt;
    ewa_ortho_body(cov3d + 6 * (size_t)i, extr, jx, jy, uv[i], gx, gy, visible[i] != 0, k, rad, nt);
    conic[3 * i] = k[0]; conic[3 * i + 1] = k[1]; conic[3 * i + 2] = k[2];
    radius[i] = rad;
    tiles[i] = nt;
}

__global__ void __launch_bounds__(kThreads)
ewa_ortho_bwd_kernel(int P, const float *__restrict__ cov3d, const float *__restrict__ extr, float jx, float jy,
                     const int *__restrict__ radius, const float *__restrict__ dL_dconic,
                     float *__restrict__ dL_dcov3d) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float o[6];
    ewa_ortho_bwd_body(cov3d + 6 * (size_t)i, extr, jx, jy, radius[i], dL_dconic + 3 * (size_t)i, o);
#pragma unroll
    for (int k = 0; k < 6; ++k) dL_dcov3d[6 * (size_t)i + k] = o[k];
}

// ---- fused per-frame geometry (spv_frame_ortho_forward/backward): the same bodies back to back in one pass -------------
// forward: projection -> visibility -> covariance -> EWA; one launch and one read of the inputs instead of four launches with
// uv / depth / vis / cov3d round trips.  radii_out = the caller-facing copy of `radius`; dirs = the renderer's constant view
// direction (0,0,1) for the SH kernels.
__global__ void __launch_bounds__(kThreads)
frame_geometry_fwd_kernel(int P, const float *__restrict__ xyz, const float *__restrict__ scales, const float4 *__restrict__ uquats,
                          const float *__restrict__ extr, int W, int H, OrthoBounds bd, float jx, float jy, int gx, int gy,
                          float2 *__restrict__ uv, float *__restrict__ depth, uint8_t *__restrict__ vis, float *__restrict__ cov3d,
                          float *__restrict__ conic, int *__restrict__ radius, int *__restrict__ tiles, int *__restrict__ radii_out) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    float2 o;
    float d;
    project_ortho_body(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], extr, W, H, bd, o, d);
    const bool visible = d != 0.f;                                   // dptr_ortho_enhanced.py:295
    float c[6];
    cov3d_body(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], uquats[i], visible, c);
    float k[3];
    int rad, nt;
    ewa_ortho_body(c, extr, jx, jy, o, gx, gy, visible, k, rad, nt);
    uv[i] = o; depth[i] = d; vis[i] = visible;
    float2 *oc = reinterpret_cast<float2 *>(cov3d + 6 * (size_t)i);
    oc[0] = make_float2(c[0], c[1]); oc[1] = make_float2(c[2], c[3]); oc[2] = make_float2(c[4], c[5]);
    conic[3 * i] = k[0]; conic[3 * i + 1] = k[1]; conic[3 * i + 2] = k[2];
    radius[i] = rad; tiles[i] = nt; radii_out[i] = rad;
}

// backward: (dL_duv, dL_ddepth) -> position ; dL_dconic -> cov3d -> scaling, rotation, with the gradients read straight from
// the packed rows of the blend backward (row layout: spv::kPackedRowGroups).
__global__ void __launch_bounds__(kThreads)
frame_geometry_bwd_kernel(int P, const float *__restrict__ packed, const float *__restrict__ scales, const float4 *__restrict__ uquats,
                          const float *__restrict__ extr, float jx, float jy, const float *__restrict__ depth,
                          const uint8_t *__restrict__ vis, const float *__restrict__ cov3d, const int *__restrict__ radius,
                          float *__restrict__ dL_dxyz, float *__restrict__ dL_dscales, float4 *__restrict__ dL_duquats) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float4 *row = reinterpret_cast<const float4 *>(packed + (size_t)i * spv::kPackedRowGroups);
    const float4 r0 = row[0], r1 = row[1], r2 = row[2];   // 0,1 dL_duv | 4,5,6 dL_dconic | 11 dL_ddepth
    float gx, gy, gz;
    project_ortho_bwd_body(extr, jx, jy, depth[i], make_float2(r0.x, r0.y), r2.w, gx, gy, gz);
    dL_dxyz[3 * i] = gx; dL_dxyz[3 * i + 1] = gy; dL_dxyz[3 * i + 2] = gz;
    const float gcon[3] = {r1.x, r1.y, r1.z};
    float gcov[6];
    ewa_ortho_bwd_body(cov3d + 6 * (size_t)i, extr, jx, jy, radius[i], gcon, gcov);
    const float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    float gs[3];
    float4 gq;
    cov3d_bwd_body(s, uquats[i], vis[i] != 0, gcov, gs, gq);
    dL_dscales[3 * i] = gs[0]; dL_dscales[3 * i + 1] = gs[1]; dL_dscales[3 * i + 2] = gs[2];
    dL_duquats[i] = gq;
}

// ------------------------------------------------------------------------------------------------ K7-K10
__constant__ float SH_C0 = 0.28209479177387814f;
__constant__ float SH_C1 = 0.4886025119029199f;
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// SH kernels move 192 B (deg 3) per Gaussian each way -- the fattest per-Gaussian stream.  A
// thread-per-Gaussian access pattern strides by 192 B, so the coefficient rows go through shared memory: the CTA's
// contiguous [128 x NB*3] slab is loaded / stored with coalesced 16-byte accesses and each thread works on its own row
// (row pitch NB*3+1 words: conflict-free).  The arithmetic is unchanged (bit-exact against the oracle).
constexpr int kShThreads = 128;     // forward: one 25 KB slab
constexpr int kShBwdThreads = 64;   // backward: coefficient slab + gradient slab

template <int ROW, int NT>
__device__ __forceinline__ void sh_slab_load(const float *__restrict__ g, float *s, int rows_valid) {
    const int nfl = rows_valid * ROW;                       // floats in this CTA's slab
    const float4 *g4 = reinterpret_cast<const float4 *>(g); // slab base is 16-byte aligned (128*ROW*4 bytes per CTA)
    for (int q = threadIdx.x; q * 4 < nfl; q += NT) {
        float v[4];
        if (q * 4 + 3 < nfl) { const float4 t = g4[q]; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { for (int k = 0; k < 4; ++k) v[k] = (q * 4 + k < nfl) ? g[q * 4 + k] : 0.f; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = q * 4 + k;
            if (e < nfl) s[(e / ROW) * (ROW + 1) + e % ROW] = v[k];
        }
    }
}

template <int ROW, int NT>
__device__ __forceinline__ void sh_slab_store(float *__restrict__ g, const float *s, int rows_valid) {
    const int nfl = rows_valid * ROW;
    float4 *g4 = reinterpret_cast<float4 *>(g);
    for (int q = threadIdx.x; q * 4 < nfl; q += NT) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = q * 4 + k;
            v[k] = e < nfl ? s[(e / ROW) * (ROW + 1) + e % ROW] : 0.f;
        }
        if (q * 4 + 3 < nfl) g4[q] = make_float4(v[0], v[1], v[2], v[3]);
        else { for (int k = 0; k < 4; ++k) if (q * 4 + k < nfl) g[q * 4 + k] = v[k]; }
    }
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_fwd_kernel(int P, const float *__restrict__ shs, const float *__restrict__ dirs,
              const uint8_t *__restrict__ visible, int free_variant, float *__restrict__ colors,
              uint8_t *__restrict__ clamped) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int ROW = NB * 3;
    __shared__ float s_sh[kShThreads * (ROW + 1)];
    const int base = blockIdx.x * kShThreads;
    const int rows = min(kShThreads, P - base);
    sh_slab_load<ROW, kShThreads>(shs + (size_t)base * ROW, s_sh, rows);
    __syncthreads();
    const int i = base + threadIdx.x;
    if (i >= P) return;
    float out[3] = {0.f, 0.f, 0.f};
    uint8_t cl[3] = {1, 1, 1};  // torch::ones for rows the kernel skips (compute_sh.cu:245)
    if (visible == nullptr || visible[i]) {   // NULL mask = every point visible
        const float *sh = s_sh + threadIdx.x * (ROW + 1);
        const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
#define S(k) sh[(k) * 3 + ch]
            float r = SH_C0 * S(0);
            if (DEG > 0) {
                r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
                if (DEG > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                        SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                    if (DEG > 2) {
                        r = r + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                            SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                            SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                            SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
                            SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                    }
                }
            }
#undef S
            if (!free_variant) {
                r += 0.5f;
                cl[ch] = (r < 0);
                out[ch] = r < 0.0f ? 0.0f : r;
            } else {
                out[ch] = r;
            }
        }
    }
    colors[3 * i] = out[0]; colors[3 * i + 1] = out[1]; colors[3 * i + 2] = out[2];
    if (clamped) { clamped[3 * i] = cl[0]; clamped[3 * i + 1] = cl[1]; clamped[3 * i + 2] = cl[2]; }
}

// DIRS = false (dL_ddirs == NULL, e.g. the renderers' constant view direction): the SH coefficients are not even read --
// the gradient of the coefficients only needs the direction's basis values and the (clamp-masked) colour gradient.
template <int DEG, bool DIRS>
__global__ void __launch_bounds__(kShBwdThreads)
sh_bwd_kernel(int P, const float *__restrict__ shs, const float *__restrict__ dirs,
              const uint8_t *__restrict__ visible, const uint8_t *__restrict__ clamped,
              const float *__restrict__ dL_dcolors, float *__restrict__ dL_dshs, float *__restrict__ dL_ddirs) {
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int ROW = NB * 3;
    __shared__ float s_sh[DIRS ? kShBwdThreads * (ROW + 1) : 1];
    __shared__ float s_g[kShBwdThreads * (ROW + 1)];
    const int base = blockIdx.x * kShBwdThreads;
    const int rows = min(kShBwdThreads, P - base);
    if (DIRS) {
        sh_slab_load<ROW, kShBwdThreads>(shs + (size_t)base * ROW, s_sh, rows);
        __syncthreads();
    }
    const int i = base + threadIdx.x;
    float gdir[3] = {0.f, 0.f, 0.f};
    if (i < P) {
        float *dsh = s_g + threadIdx.x * (ROW + 1);
#pragma unroll
        for (int k = 0; k < ROW; ++k) dsh[k] = 0.f;     // invisible rows stay zero (torch::zeros in the reference)
        if (visible == nullptr || visible[i]) {   // NULL mask = every point visible
            const float *sh = s_sh + (DIRS ? threadIdx.x * (ROW + 1) : 0);
            const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float g = dL_dcolors[3 * i + ch];
                if (clamped) g *= clamped[3 * i + ch] ? 0.0f : 1.0f;
#define S(k) (DIRS ? sh[(k) * 3 + ch] : 0.f)
#define DS(k) dsh[(k) * 3 + ch]
                float dx = 0, dy = 0, dz = 0;
                DS(0) = SH_C0 * g;
                if (DEG > 0) {
                    DS(1) = (-SH_C1 * y) * g; DS(2) = (SH_C1 * z) * g; DS(3) = (-SH_C1 * x) * g;
                    dx = -SH_C1 * S(3); dy = -SH_C1 * S(1); dz = SH_C1 * S(2);
                    if (DEG > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        DS(4) = (SH_C2[0] * xy) * g; DS(5) = (SH_C2[1] * yz) * g;
                        DS(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * g;
                        DS(7) = (SH_C2[3] * xz) * g; DS(8) = (SH_C2[4] * (xx - yy)) * g;
                        dx += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
                        dy += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
                        dz += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
                        if (DEG > 2) {
                            DS(9) = (SH_C3[0] * y * (3.f * xx - yy)) * g; DS(10) = (SH_C3[1] * xy * z) * g;
                            DS(11) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * g;
                            DS(12) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * g;
                            DS(13) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * g;
                            DS(14) = (SH_C3[5] * z * (xx - yy)) * g; DS(15) = (SH_C3[6] * x * (xx - 3.f * yy)) * g;
                            dx += (SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy +
                                   SH_C3[3] * S(12) * -3.f * 2.f * xz + SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) +
                                   SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy));
                            dy += (SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz +
                                   SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * S(12) * -3.f * 2.f * yz +
                                   SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz +
                                   SH_C3[6] * S(15) * -3.f * 2.f * xy);
                            dz += (SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz +
                                   SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * S(13) * 4.f * 2.f * xz +
                                   SH_C3[5] * S(14) * (xx - yy));
                        }
                    }
                }
#undef S
#undef DS
                gdir[0] += dx * g; gdir[1] += dy * g; gdir[2] += dz * g;
            }
        }
        if (DIRS) { dL_ddirs[3 * i] = gdir[0]; dL_ddirs[3 * i + 1] = gdir[1]; dL_ddirs[3 * i + 2] = gdir[2]; }
    }
    __syncthreads();
    sh_slab_store<ROW, kShBwdThreads>(dL_dshs + (size_t)base * ROW, s_g, rows);
}

// ---- SH along the constant view direction (0,0,1) -----------------------------------------------------------------------------
// The ortho renderers evaluate the colour along direction (0,0,1) for every Gaussian (dptr_ortho_enhanced.py:270-271).  There only
// the bases 0, 2, 6 and 12 are non-zero; every other basis multiplies its coefficient by 0 on the way in and its gradient by 0 on
// the way out.  These two kernels take the coefficients of those four bases alone, [P,4,3] in that order: 48 instead of 192 bytes
// per Gaussian.  Same operations in the same order as sh_fwd_kernel<3> / sh_bwd_kernel<3,false> with x = y = 0, z = 1 (the
// skipped terms are exact zeros), so colours and gradients are bit-identical to the 16-basis kernels'.
__global__ void __launch_bounds__(kThreads)
sh_z_fwd_kernel(int P, const float4 *__restrict__ shs_z, float *__restrict__ colors, uint8_t *__restrict__ clamped) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float4 a = shs_z[3 * i], b = shs_z[3 * i + 1], c = shs_z[3 * i + 2];
    const float s[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    const float k2 = SH_C2[2] * 2.0f, k3 = SH_C3[3] * 2.0f;    // C2[2] (2 zz - xx - yy), C3[3] z (2 zz - 3 xx - 3 yy) at z = 1
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float r = SH_C0 * s[ch];
        r = r + SH_C1 * s[3 + ch];
        r = r + k2 * s[6 + ch];
        r = r + k3 * s[9 + ch];
        r += 0.5f;
        if (clamped) clamped[3 * i + ch] = (r < 0);
        colors[3 * i + ch] = r < 0.0f ? 0.0f : r;
    }
}

__global__ void __launch_bounds__(kThreads)
sh_z_bwd_kernel(int P, const uint8_t *__restrict__ clamped, const float *__restrict__ dL_dcolors, float4 *__restrict__ dL_dshs_z) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const float k2 = SH_C2[2] * 2.0f, k3 = SH_C3[3] * 2.0f;
    float d[12];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float g = dL_dcolors[3 * i + ch];
        if (clamped) g *= clamped[3 * i + ch] ? 0.0f : 1.0f;
        d[ch] = SH_C0 * g; d[3 + ch] = SH_C1 * g; d[6 + ch] = k2 * g; d[9 + ch] = k3 * g;
    }
    dL_dshs_z[3 * i] = make_float4(d[0], d[1], d[2], d[3]);
    dL_dshs_z[3 * i + 1] = make_float4(d[4], d[5], d[6], d[7]);
    dL_dshs_z[3 * i + 2] = make_float4(d[8], d[9], d[10], d[11]);
}

inline dim3 grid_for(int P) { return dim3(spv::cdiv(P, kThreads)); }

}  // namespace

// ---- internal entry points of the fused frame path (frame.cu) ----------------------------------------------------------------
namespace spv {
int frame_geometry_forward(int P, const float *xyz, const float *scales, const float *uquats, const float *extr, int W, int H,
                           float nearest, float extent, float *uv, float *depth, uint8_t *vis, float *cov3d, float *conic, int *radius,
                           int *tiles, int *radii_out, void *stream) {
    if (P <= 0) return 0;
    OrthoBounds bd;
    bd.nearest = nearest;   // same narrowing as spv_project_point_ortho_forward
    bd.xmin = (float)((1.0 - (double)extent) * W * 0.5); bd.xmax = (float)((1.0 + (double)extent) * W * 0.5);
    bd.ymin = (float)((1.0 - (double)extent) * H * 0.5); bd.ymax = (float)((1.0 + (double)extent) * H * 0.5);
    frame_geometry_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, xyz, scales, (const float4 *)uquats, extr, W, H, bd, (float)((double)W / 2.0), (float)((double)H / 2.0), spv::tiles_x(W),
        spv::tiles_y(H), (float2 *)uv, depth, vis, cov3d, conic, radius, tiles, radii_out);
    return spv::check_launch("spv_frame_ortho_forward/geometry");
}

int frame_geometry_backward(int P, const float *packed, const float *scales, const float *uquats, const float *extr, int W, int H,
                            const float *depth, const uint8_t *vis, const float *cov3d, const int *radius, float *dL_dxyz,
                            float *dL_dscales, float *dL_duquats, void *stream) {
    if (P <= 0) return 0;
    frame_geometry_bwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, packed, scales, (const float4 *)uquats, extr, (float)((double)W / 2.0), (float)((double)H / 2.0), depth, vis, cov3d, radius,
        dL_dxyz, dL_dscales, (float4 *)dL_duquats);
    return spv::check_launch("spv_frame_ortho_backward/geometry");
}
}  // namespace spv

// ================================================================================================ C ABI
extern "C" {

int spv_project_point_forward(int P, const float *xyz, const float *intr, const float *extr, int W, int H,
                              float nearest, float extent, float *uv, float *depth, void *stream) {
    if (P <= 0) return 0;
    project_point_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, xyz, intr, extr, W, H, nearest, extent, (float2 *)uv, depth);
    return spv::check_launch("spv_project_point_forward");
}

int spv_project_point_backward(int P, const float *xyz, const float *intr, const float *extr, const float *depth,
                               const float *dL_duv, const float *dL_ddepth, float *dL_dxyz, float *dL_dintr,
                               float *dL_dextr, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (dL_dintr) SPV_CUDA_TRY(cudaMemsetAsync(dL_dintr, 0, 4 * sizeof(float), s), "spv_project_point_backward");
    if (dL_dextr) SPV_CUDA_TRY(cudaMemsetAsync(dL_dextr, 0, 12 * sizeof(float), s), "spv_project_point_backward");
    if (P <= 0) return 0;
    project_point_bwd_kernel<<<grid_for(P), kThreads, 0, s>>>(P, xyz, intr, extr, depth, (const float2 *)dL_duv,
                                                             dL_ddepth, dL_dxyz, dL_dintr, dL_dextr);
    return spv::check_launch("spv_project_point_backward");
}

int spv_project_point_ortho_forward(int P, const float *xyz, const float *extr, int W, int H, float nearest,
                                    float extent, float *uv, float *depth, void *stream) {
    if (P <= 0) return 0;
    // python doubles narrowed to fp32 when compared against an fp32 tensor (dptr_ortho_enhanced.py:189-192)
    const float xmin = (float)((1.0 - (double)extent) * W * 0.5), xmax = (float)((1.0 + (double)extent) * W * 0.5);
    const float ymin = (float)((1.0 - (double)extent) * H * 0.5), ymax = (float)((1.0 + (double)extent) * H * 0.5);
    project_point_ortho_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, xyz, extr, W, H, nearest, xmin, xmax, ymin, ymax, (float2 *)uv, depth);
    return spv::check_launch("spv_project_point_ortho_forward");
}

int spv_project_point_ortho_backward(int P, const float *extr, int W, int H, const float *depth, const float *dL_duv,
                                     const float *dL_ddepth, float *dL_dxyz, void *stream) {
    if (P <= 0) return 0;
    project_point_ortho_bwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, extr, (float)W / 2.0f, (float)H / 2.0f, depth, (const float2 *)dL_duv, dL_ddepth, dL_dxyz);
    return spv::check_launch("spv_project_point_ortho_backward");
}

int spv_compute_cov3d_forward(int P, const float *scales, const float *uquats, const uint8_t *visible, float *cov3d,
                              void *stream) {
    if (P <= 0) return 0;
    cov3d_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, scales, (const float4 *)uquats, visible,
                                                                         cov3d);
    return spv::check_launch("spv_compute_cov3d_forward");
}

int spv_compute_cov3d_backward(int P, const float *scales, const float *uquats, const uint8_t *visible,
                               const float *dL_dcov3d, float *dL_dscales, float *dL_duquats, void *stream) {
    if (P <= 0) return 0;
    cov3d_bwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, scales, (const float4 *)uquats, visible,
                                                                         dL_dcov3d, dL_dscales, (float4 *)dL_duquats);
    return spv::check_launch("spv_compute_cov3d_backward");
}

int spv_ewa_project_forward(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                            const float *uv, int W, int H, const uint8_t *visible, float *conic, int *radius,
                            int *tiles, void *stream) {
    if (P <= 0) return 0;
    ewa_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, xyz, cov3d, intr, extr, (const float2 *)uv, spv::tiles_x(W), spv::tiles_y(H), visible, conic, radius, tiles);
    return spv::check_launch("spv_ewa_project_forward");
}

int spv_ewa_project_backward(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                             const int *radius, const float *dL_dconic, float *dL_dxyz, float *dL_dcov3d,
                             float *dL_dintr, float *dL_dextr, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (dL_dintr) SPV_CUDA_TRY(cudaMemsetAsync(dL_dintr, 0, 4 * sizeof(float), s), "spv_ewa_project_backward");
    if (dL_dextr) SPV_CUDA_TRY(cudaMemsetAsync(dL_dextr, 0, 12 * sizeof(float), s), "spv_ewa_project_backward");
    if (P <= 0) return 0;
    ewa_bwd_kernel<<<grid_for(P), kThreads, 0, s>>>(P, xyz, cov3d, intr, extr, radius, dL_dconic, dL_dxyz, dL_dcov3d,
                                                   dL_dintr, dL_dextr);
    return spv::check_launch("spv_ewa_project_backward");
}

int spv_ewa_project_ortho_forward(int P, const float *cov3d, const float *extr, const float *uv, int W, int H,
                                  const uint8_t *visible, float *conic, int *radius, int *tiles, void *stream) {
    if (P <= 0) return 0;
    ewa_ortho_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, cov3d, extr, (float)((double)W / 2.0), (float)((double)H / 2.0), (const float2 *)uv, spv::tiles_x(W),
        spv::tiles_y(H), visible, conic, radius, tiles);
    return spv::check_launch("spv_ewa_project_ortho_forward");
}

int spv_ewa_project_ortho_backward(int P, const float *cov3d, const float *extr, int W, int H, const int *radius,
                                   const float *dL_dconic, float *dL_dcov3d, void *stream) {
    if (P <= 0) return 0;
    ewa_ortho_bwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, cov3d, extr, (float)((double)W / 2.0), (float)((double)H / 2.0), radius, dL_dconic, dL_dcov3d);
    return spv::check_launch("spv_ewa_project_ortho_backward");
}

/* SH colour along the constant direction (0,0,1): shs_z = [P,4,3], the coefficients of the bases 0, 2, 6, 12 (the only ones that
 * direction reaches).  Bit-identical to spv_compute_sh_forward / _backward(deg 3, dirs = (0,0,1)) on the full [P,16,3] tensor. */
int spv_compute_sh_z_forward(int P, const float *shs_z, float *colors, uint8_t *clamped, void *stream) {
    if (P <= 0) return 0;
    sh_z_fwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, (const float4 *)shs_z, colors, clamped);
    return spv::check_launch("spv_compute_sh_z_forward");
}

int spv_compute_sh_z_backward(int P, const uint8_t *clamped, const float *dL_dcolors, float *dL_dshs_z, void *stream) {
    if (P <= 0) return 0;
    sh_z_bwd_kernel<<<grid_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, clamped, dL_dcolors, (float4 *)dL_dshs_z);
    return spv::check_launch("spv_compute_sh_z_backward");
}

int spv_compute_sh_forward(int P, const float *shs, int deg, const float *dirs, const uint8_t *visible,
                           int free_variant, float *colors, uint8_t *clamped, void *stream) {
    if (P <= 0) return 0;
    if (deg < 0 || deg > 3) { spv::set_error(cudaErrorInvalidValue, "spv_compute_sh_forward: deg must be 0..3"); return (int)cudaErrorInvalidValue; }
    cudaStream_t s = (cudaStream_t)stream;
    dim3 g(spv::cdiv(P, kShThreads));
    switch (deg) {
        case 0: sh_fwd_kernel<0><<<g, kShThreads, 0, s>>>(P, shs, dirs, visible, free_variant, colors, clamped); break;
        case 1: sh_fwd_kernel<1><<<g, kShThreads, 0, s>>>(P, shs, dirs, visible, free_variant, colors, clamped); break;
        case 2: sh_fwd_kernel<2><<<g, kShThreads, 0, s>>>(P, shs, dirs, visible, free_variant, colors, clamped); break;
        default: sh_fwd_kernel<3><<<g, kShThreads, 0, s>>>(P, shs, dirs, visible, free_variant, colors, clamped); break;
    }
    return spv::check_launch("spv_compute_sh_forward");
}

int spv_compute_sh_backward(int P, const float *shs, int deg, const float *dirs, const uint8_t *visible,
                            const uint8_t *clamped, const float *dL_dcolors, int S_alloc, float *dL_dshs,
                            float *dL_ddirs, void *stream) {
    if (P <= 0) return 0;
    if (deg < 0 || deg > 3) { spv::set_error(cudaErrorInvalidValue, "spv_compute_sh_backward: deg must be 0..3"); return (int)cudaErrorInvalidValue; }
    cudaStream_t s = (cudaStream_t)stream;
    // the kernel writes every (deg+1)^2-row (zeros for invisible points); only the tail of an over-allocated gradient
    // tensor (S_alloc > (deg+1)^2, the reference's stride quirk) needs the torch::zeros-like clear
    const int nb = (deg + 1) * (deg + 1);
    if (S_alloc > nb)
        SPV_CUDA_TRY(cudaMemsetAsync(dL_dshs + (size_t)P * nb * 3, 0, sizeof(float) * 3 * (size_t)(S_alloc - nb) * (size_t)P, s),
                     "spv_compute_sh_backward");
    dim3 g(spv::cdiv(P, kShBwdThreads));
    switch (deg) {
#define SPV_SH_BWD(D)                                                                                                            \
    if (dL_ddirs) sh_bwd_kernel<D, true><<<g, kShBwdThreads, 0, s>>>(P, shs, dirs, visible, clamped, dL_dcolors, dL_dshs, dL_ddirs);  \
    else sh_bwd_kernel<D, false><<<g, kShBwdThreads, 0, s>>>(P, shs, dirs, visible, clamped, dL_dcolors, dL_dshs, nullptr)
        case 0: SPV_SH_BWD(0); break;
        case 1: SPV_SH_BWD(1); break;
        case 2: SPV_SH_BWD(2); break;
        default: SPV_SH_BWD(3); break;
#undef SPV_SH_BWD
    }
    return spv::check_launch("spv_compute_sh_backward");
}

}  // extern "C"
