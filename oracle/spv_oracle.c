/*
 * spv_oracle.c -- CPU restatement of the reference rasterizer (DPTR) hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in splatter_a_video_b200/ may import, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * CPU baseline.  Plain C, IEEE fp32, no fast-math, no FMA contraction
 * (compile with -ffp-contract=off): every function follows the operation order
 * of the reference file:line it cites (paths relative to
 * /root/reference/src/submodules/dptr/dptr/gs/ unless they start with src/).
 *
 * Parity status: the reference ships no tests / golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference
 * ITSELF: the unmodified reference .cu files are compiled for sm_100a into
 * oracle/_ref/_C.so (oracle/ref_build/Makefile) and tests/test_ref_pin_gpu.py
 * compares every function below with it on the GPU box.
 *
 * Deliberate, documented deviations from the compiled reference:
 *   - IEEE expf/sqrtf/div instead of --use_fast_math approximations
 *     (setup.py:51); differences are ~1e-7 relative and are covered by the
 *     tolerance + "fragile" masks (pixels/Gaussians whose discrete decisions
 *     sit within frag_eps of a threshold).
 *   - backward accumulates the per-Gaussian sums in double (the reference uses
 *     order-nondeterministic float atomics, alpha_blending.cu:219-246).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16 /* include/config.h:7 */
#define BLOCK_Y 16 /* include/config.h:8 */

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* include/utils.h:17-37 get_rect; the int radius is promoted to float. */
static void get_rect(float px, float py, int max_radius, int gx, int gy,
                     int *minx, int *miny, int *maxx, int *maxy) {
    float r = (float)max_radius;
    *minx = imin(gx, imax(0, (int)((px - r) / (float)BLOCK_X)));
    *miny = imin(gy, imax(0, (int)((py - r) / (float)BLOCK_Y)));
    *maxx = imin(gx, imax(0, (int)((px + r + (float)BLOCK_X - 1.0f) / (float)BLOCK_X)));
    *maxy = imin(gy, imax(0, (int)((py + r + (float)BLOCK_Y - 1.0f) / (float)BLOCK_Y)));
}

/* ------------------------------------------------------------------------- */
/* K1: src/project_point.cu:13-57 (perspective projection + culling).         */
void orc_project_point_fwd(int P, const float *xyz, const float *intr, const float *extr,
                           int W, int H, float nearest, float extent,
                           float *uv, float *depth) {
    memset(uv, 0, sizeof(float) * 2 * (size_t)P);
    memset(depth, 0, sizeof(float) * (size_t)P);
    for (int i = 0; i < P; ++i) {
        float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        float tx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];
        float ty = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
        float tz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
        /* :31 `1.0 / (tmp.z + 1e-7)` is evaluated in double, then narrowed. */
        float norm1 = (float)(1.0 / ((double)tz + 1e-7));
        /* :34-35 `... - 0.5` is a double subtraction of a float value. */
        float u = (float)((double)(intr[0] * tx * norm1 + intr[2]) - 0.5);
        float v = (float)((double)(intr[1] * ty * norm1 + intr[3]) - 0.5);
        int near_cull = 0, extent_cull = 0;
        if (nearest > 0) near_cull = tz <= nearest;
        if (extent > 0) {
            float xmin = (float)((double)((1 - extent) * W) * 0.5);
            float xmax = (float)((double)((1 + extent) * W) * 0.5);
            float ymin = (float)((double)((1 - extent) * H) * 0.5);
            float ymax = (float)((double)((1 + extent) * H) * 0.5);
            extent_cull = u < xmin || u > xmax || v < ymin || v > ymax;
        }
        if (near_cull || extent_cull) continue;
        uv[2 * i] = u;
        uv[2 * i + 1] = v;
        depth[i] = tz;
    }
}

/* K2: src/project_point.cu:59-145.  dL_dintr / dL_dextr may be NULL. */
void orc_project_point_bwd(int P, const float *xyz, const float *intr, const float *extr,
                           const float *depth, const float *dL_duv, const float *dL_ddepth,
                           float *dL_dxyz, float *dL_dintr, float *dL_dextr) {
    double gi[4] = {0, 0, 0, 0}, ge[12] = {0};
    memset(dL_dxyz, 0, sizeof(float) * 3 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (depth[i] == 0) continue;
        float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        float tx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];
        float ty = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
        float tz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
        float norm1 = (float)(1.0 / (double)tz);
        float norm2 = (float)(1.0 / (double)(tz * tz));
        float gu = dL_duv[2 * i], gv = dL_duv[2 * i + 1], gd = dL_ddepth[i];
        float gx = 0, gy = 0, gz = 0;
        gx += (intr[0] * (extr[0] * tz - tx * extr[8]) * norm2) * gu;
        gx += (intr[1] * (extr[4] * tz - ty * extr[8]) * norm2) * gv;
        gx += extr[8] * gd;
        gy += (intr[0] * (extr[1] * tz - tx * extr[9]) * norm2) * gu;
        gy += (intr[1] * (extr[5] * tz - ty * extr[9]) * norm2) * gv;
        gy += extr[9] * gd;
        gz += (intr[0] * (extr[2] * tz - tx * extr[10]) * norm2) * gu;
        gz += (intr[1] * (extr[6] * tz - ty * extr[10]) * norm2) * gv;
        gz += extr[10] * gd;
        dL_dxyz[3 * i] = gx; dL_dxyz[3 * i + 1] = gy; dL_dxyz[3 * i + 2] = gz;
        if (dL_dintr) {
            gi[0] += tx * norm1 * gu; gi[1] += ty * norm1 * gv; gi[2] += gu; gi[3] += gv;
        }
        if (dL_dextr) {
            ge[0] += intr[0] * px * norm1 * gu; ge[1] += intr[0] * py * norm1 * gu;
            ge[2] += intr[0] * pz * norm1 * gu; ge[3] += intr[0] * norm1 * gu;
            ge[4] += intr[1] * px * norm1 * gv; ge[5] += intr[1] * py * norm1 * gv;
            ge[6] += intr[1] * pz * norm1 * gv; ge[7] += intr[1] * norm1 * gv;
            ge[8] += -intr[0] * px * tx * norm2 * gu; ge[8] += -intr[1] * px * ty * norm2 * gv; ge[8] += px * gd;
            ge[9] += -intr[0] * py * tx * norm2 * gu; ge[9] += -intr[1] * py * ty * norm2 * gv; ge[9] += py * gd;
            ge[10] += -intr[0] * pz * tx * norm2 * gu; ge[10] += -intr[1] * pz * ty * norm2 * gv; ge[10] += pz * gd;
            ge[11] += -intr[0] * tx * norm2 * gu; ge[11] += -intr[1] * ty * norm2 * gv; ge[11] += gd;
        }
    }
    if (dL_dintr) for (int k = 0; k < 4; ++k) dL_dintr[k] = (float)gi[k];
    if (dL_dextr) for (int k = 0; k < 12; ++k) dL_dextr[k] = (float)ge[k];
}

/* Orthographic projection used by the trainer:
 * src/pointrix/renderer/dptr_ortho_enhanced.py:177-202 (torch ops restated). */
void orc_project_point_ortho_fwd(int P, const float *xyz, const float *extr /* row stride 4 */,
                                 int W, int H, float nearest, float extent,
                                 float *uv, float *depth) {
    /* python scalars are doubles, cast to float32 when compared with a float32 tensor */
    float xmin = (float)((1.0 - (double)extent) * W * 0.5), xmax = (float)((1.0 + (double)extent) * W * 0.5);
    float ymin = (float)((1.0 - (double)extent) * H * 0.5), ymax = (float)((1.0 + (double)extent) * H * 0.5);
    for (int i = 0; i < P; ++i) {
        float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        float cx = extr[0] * px + extr[1] * py + extr[2] * pz + extr[3];   /* :180 matmul + t */
        float cy = extr[4] * px + extr[5] * py + extr[6] * pz + extr[7];
        float cz = extr[8] * px + extr[9] * py + extr[10] * pz + extr[11];
        float u = (cx + 1.0f) * (float)W / 2.0f - 0.5f;                     /* :183,185 */
        float v = (cy + 1.0f) * (float)H / 2.0f - 0.5f;
        float d = cz;
        if (isnan(d)) d = 0.0f;                                             /* :187 nan_to_num */
        else if (isinf(d)) d = d > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        int mask = (d <= nearest) || (u < xmin) || (u > xmax) || (v < ymin) || (v > ymax);
        uv[2 * i] = mask ? 0.0f : u;
        uv[2 * i + 1] = mask ? 0.0f : v;
        depth[i] = mask ? 0.0f : d;
    }
}

/* ------------------------------------------------------------------------- */
/* K3: src/compute_cov3d.cu:14-58,119-129.  glm is column-major: m[c][r].     */
typedef struct { float m[3][3]; } m3; /* m[col][row] */

static m3 m3_mul(m3 a, m3 b) {
    m3 r;
    for (int c = 0; c < 3; ++c)
        for (int w = 0; w < 3; ++w)
            r.m[c][w] = a.m[0][w] * b.m[c][0] + a.m[1][w] * b.m[c][1] + a.m[2][w] * b.m[c][2];
    return r;
}
static m3 m3_t(m3 a) {
    m3 r;
    for (int c = 0; c < 3; ++c) for (int w = 0; w < 3; ++w) r.m[c][w] = a.m[w][c];
    return r;
}
static m3 quat_to_R(const float *q) { /* :24-40, (r,x,y,z); glm ctor fills columns */
    float r = q[0], x = q[1], y = q[2], z = q[3];
    m3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}
static m3 scale_to_S(const float *s) {
    m3 S; memset(&S, 0, sizeof S);
    S.m[0][0] = s[0]; S.m[1][1] = s[1]; S.m[2][2] = s[2];
    return S;
}

void orc_compute_cov3d_fwd(int P, const float *scales, const float *uquats,
                           const unsigned char *visible, float *cov3d) {
    memset(cov3d, 0, sizeof(float) * 6 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (!visible[i]) continue;
        m3 M = m3_mul(scale_to_S(scales + 3 * i), quat_to_R(uquats + 4 * i)); /* :49 M = S*R */
        m3 Sg = m3_mul(m3_t(M), M);                                           /* :50 */
        float *c = cov3d + 6 * i;
        c[0] = Sg.m[0][0]; c[1] = Sg.m[0][1]; c[2] = Sg.m[0][2];
        c[3] = Sg.m[1][1]; c[4] = Sg.m[1][2]; c[5] = Sg.m[2][2];
    }
}

/* K4: src/compute_cov3d.cu:60-117,131-147. */
void orc_compute_cov3d_bwd(int P, const float *scales, const float *uquats,
                           const unsigned char *visible, const float *dL_dcov3d,
                           float *dL_dscales, float *dL_duquats) {
    memset(dL_dscales, 0, sizeof(float) * 3 * (size_t)P);
    memset(dL_duquats, 0, sizeof(float) * 4 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (!visible[i]) continue;
        const float *s = scales + 3 * i, *q = uquats + 4 * i, *g = dL_dcov3d + 6 * i;
        m3 R = quat_to_R(q);
        m3 M = m3_mul(scale_to_S(s), R);
        m3 dS; /* :69-77, columns */
        dS.m[0][0] = g[0]; dS.m[0][1] = 0.5f * g[1]; dS.m[0][2] = 0.5f * g[2];
        dS.m[1][0] = 0.5f * g[1]; dS.m[1][1] = g[3]; dS.m[1][2] = 0.5f * g[4];
        dS.m[2][0] = 0.5f * g[2]; dS.m[2][1] = 0.5f * g[4]; dS.m[2][2] = g[5];
        m3 M2;
        for (int c = 0; c < 3; ++c) for (int w = 0; w < 3; ++w) M2.m[c][w] = 2.0f * M.m[c][w];
        m3 dM = m3_mul(M2, dS); /* :81 */
        m3 Rt = m3_t(R), dMt = m3_t(dM);
        float *gs = dL_dscales + 3 * i, *gq = dL_duquats + 4 * i;
        for (int k = 0; k < 3; ++k)
            gs[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
        for (int k = 0; k < 3; ++k) for (int w = 0; w < 3; ++w) dMt.m[k][w] *= s[k];
        float r = q[0], x = q[1], y = q[2], z = q[3];
#define D(a, b) dMt.m[a][b]
        gq[0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
        gq[1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) -
                4 * x * (D(2, 2) + D(1, 1));
        gq[2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) -
                4 * y * (D(2, 2) + D(0, 0));
        gq[3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) -
                4 * z * (D(1, 1) + D(0, 0));
#undef D
    }
}

/* ------------------------------------------------------------------------- */
/* K5: src/ewa_project.cu:16-83 (perspective EWA, radius, tile rect).         */
static void ewa_T(const float *p, const float *intr, const float *extr, m3 *T, m3 *J, m3 *Wm, float *t) {
    float fx = intr[0], fy = intr[1];
    t[0] = extr[0] * p[0] + extr[1] * p[1] + extr[2] * p[2] + extr[3];
    t[1] = extr[4] * p[0] + extr[5] * p[1] + extr[6] * p[2] + extr[7];
    t[2] = extr[8] * p[0] + extr[9] * p[1] + extr[10] * p[2] + extr[11];
    memset(J, 0, sizeof *J);
    J->m[0][0] = fx / t[2]; J->m[1][1] = fy / t[2];
    J->m[2][0] = -(fx * t[0]) / (t[2] * t[2]); J->m[2][1] = -(fy * t[1]) / (t[2] * t[2]);
    Wm->m[0][0] = extr[0]; Wm->m[0][1] = extr[4]; Wm->m[0][2] = extr[8];
    Wm->m[1][0] = extr[1]; Wm->m[1][1] = extr[5]; Wm->m[1][2] = extr[9];
    Wm->m[2][0] = extr[2]; Wm->m[2][1] = extr[6]; Wm->m[2][2] = extr[10];
    *T = m3_mul(*J, *Wm);
}
static m3 cov3d_to_m3(const float *c) {
    m3 V;
    V.m[0][0] = c[0]; V.m[0][1] = c[1]; V.m[0][2] = c[2];
    V.m[1][0] = c[1]; V.m[1][1] = c[3]; V.m[1][2] = c[4];
    V.m[2][0] = c[2]; V.m[2][1] = c[4]; V.m[2][2] = c[5];
    return V;
}

void orc_ewa_project_fwd(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                         const float *uv, int W, int H, const unsigned char *visible,
                         float *conic, int *radius, int *tiles) {
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    memset(conic, 0, sizeof(float) * 3 * (size_t)P);
    memset(radius, 0, sizeof(int) * (size_t)P);
    memset(tiles, 0, sizeof(int) * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (!visible[i]) continue;
        m3 T, J, Wm; float t[3];
        ewa_T(xyz + 3 * i, intr, extr, &T, &J, &Wm, t);
        m3 cov2 = m3_mul(m3_mul(T, cov3d_to_m3(cov3d + 6 * i)), m3_t(T)); /* :55 */
        float cx = cov2.m[0][0] + 0.3f, cy = cov2.m[0][1], cz = cov2.m[1][1] + 0.3f;
        float det = cx * cz - cy * cy;
        if (det == 0.0f) continue;
        float mid = 0.5f * (cx + cz);
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        int x0, y0, x1, y1;
        get_rect(uv[2 * i], uv[2 * i + 1], (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        float det_inv = 1.f / det;
        conic[3 * i] = cz * det_inv; conic[3 * i + 1] = -cy * det_inv; conic[3 * i + 2] = cx * det_inv;
        radius[i] = (int)my_radius;
        tiles[i] = (y1 - y0) * (x1 - x0);
    }
}

/* K6: src/ewa_project.cu:85-252.  dL_dintr / dL_dextr may be NULL. */
void orc_ewa_project_bwd(int P, const float *xyz, const float *cov3d, const float *intr, const float *extr,
                         const int *radius, const float *dL_dconic,
                         float *dL_dxyz, float *dL_dcov3d, float *dL_dintr, float *dL_dextr) {
    double gi[4] = {0}, ge[12] = {0};
    memset(dL_dxyz, 0, sizeof(float) * 3 * (size_t)P);
    memset(dL_dcov3d, 0, sizeof(float) * 6 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (!(radius[i] > 0)) continue;
        float fx = intr[0], fy = intr[1];
        const float *p = xyz + 3 * i, *c3 = cov3d + 6 * i, *g = dL_dconic + 3 * i;
        m3 T, J, Wm; float t[3];
        ewa_T(p, intr, extr, &T, &J, &Wm, t);
        m3 cov2 = m3_mul(m3_mul(T, cov3d_to_m3(c3)), m3_t(T));
        float cx = cov2.m[0][0] + 0.3f, cy = cov2.m[0][1], cz = cov2.m[1][1] + 0.3f;
        float det = cx * cz - cy * cy;
        if (det == 0.0f) continue;
        float nom = 1.0f / (det * det);
        float dcx = nom * (-cz * cz * g[0] + cy * cz * g[1] + (det - cx * cz) * g[2]);
        float dcy = nom * (2 * cy * cz * g[0] - (det + 2 * cy * cy) * g[1] + 2 * cx * cy * g[2]);
        float dcz = nom * ((det - cx * cz) * g[0] + cx * cy * g[1] - cx * cx * g[2]);
#define TT(a, b) T.m[a][b]
        float *o = dL_dcov3d + 6 * i;
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
#define WW(a, b) Wm.m[a][b]
        float dJ00 = WW(0, 0) * dT00 + WW(1, 0) * dT10 + WW(2, 0) * dT20;
        float dJ20 = WW(0, 2) * dT00 + WW(1, 2) * dT10 + WW(2, 2) * dT20;
        float dJ11 = WW(0, 1) * dT01 + WW(1, 1) * dT11 + WW(2, 1) * dT21;
        float dJ21 = WW(0, 2) * dT01 + WW(1, 2) * dT11 + WW(2, 2) * dT21;
#undef WW
        float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        float dtx = -fx * tz2 * dJ20;
        float dty = -fy * tz2 * dJ21;
        float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ20 + (2 * fy * t[1]) * tz3 * dJ21;
        if (dL_dintr) {
            gi[0] += tz * dJ00; gi[0] += -t[0] * tz2 * dJ20;
            gi[1] += tz * dJ11; gi[1] += -t[1] * tz2 * dJ21;
        }
        if (dL_dextr) {
            ge[0] += J.m[0][0] * dT00; ge[1] += J.m[0][0] * dT10; ge[2] += J.m[0][0] * dT20;
            ge[4] += J.m[1][1] * dT01; ge[5] += J.m[1][1] * dT11; ge[6] += J.m[1][1] * dT21;
            ge[8] += J.m[2][0] * dT00 + J.m[2][1] * dT01;
            ge[9] += J.m[2][0] * dT10 + J.m[2][1] * dT11;
            ge[10] += J.m[2][0] * dT20 + J.m[2][1] * dT21;
            ge[0] += p[0] * dtx; ge[1] += p[1] * dtx; ge[2] += p[2] * dtx; ge[3] += dtx;
            ge[4] += p[0] * dty; ge[5] += p[1] * dty; ge[6] += p[2] * dty; ge[7] += dty;
            ge[8] += p[0] * dtz; ge[9] += p[1] * dtz; ge[10] += p[2] * dtz; ge[11] += dtz;
        }
        dL_dxyz[3 * i] = extr[0] * dtx + extr[4] * dty + extr[8] * dtz;
        dL_dxyz[3 * i + 1] = extr[1] * dtx + extr[5] * dty + extr[9] * dtz;
        dL_dxyz[3 * i + 2] = extr[2] * dtx + extr[6] * dty + extr[10] * dtz;
    }
    if (dL_dintr) for (int k = 0; k < 4; ++k) dL_dintr[k] = (float)gi[k];
    if (dL_dextr) for (int k = 0; k < 12; ++k) dL_dextr[k] = (float)ge[k];
}

/* Orthographic EWA used by the trainer:
 * src/pointrix/renderer/dptr_ortho_enhanced.py:18-111 (torch ops restated;
 * matmul accumulation taken left to right).  visible = depth != 0.           */
void orc_ewa_project_ortho_fwd(int P, const float *cov3d, const float *extr /* row stride 4 */,
                               const float *uv, int W, int H, const unsigned char *visible,
                               float *conic, int *radius, int *tiles) {
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    float jx = (float)((double)W / 2.0), jy = (float)((double)H / 2.0); /* :36-37 */
    float T[2][3];
    for (int k = 0; k < 3; ++k) {           /* :42 T = Jmat @ Wmat, J = [[jx,0,0],[0,jy,0]] */
        T[0][k] = jx * extr[k] + 0.0f * extr[4 + k] + 0.0f * extr[8 + k];
        T[1][k] = 0.0f * extr[k] + jy * extr[4 + k] + 0.0f * extr[8 + k];
    }
    for (int i = 0; i < P; ++i) {
        const float *c = cov3d + 6 * i;
        float S[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
        float M[2][3], c2[2][2];
        for (int a = 0; a < 2; ++a) for (int k = 0; k < 3; ++k)
            M[a][k] = T[a][0] * S[0][k] + T[a][1] * S[1][k] + T[a][2] * S[2][k];
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b)
            c2[a][b] = M[a][0] * T[b][0] + M[a][1] * T[b][1] + M[a][2] * T[b][2];
        float c00 = c2[0][0] + 0.3f, c11 = c2[1][1] + 0.3f, c01 = c2[0][1];
        float det = c00 * c11 - c01 * c01;
        float k0 = c11 / det, k1 = -c01 / det, k2 = c00 / det;            /* :56-63 */
        float b = (c00 + c11) / 2.0f;                                       /* :65 */
        float disc = b * b - det;
        if (disc < 0.1f) disc = 0.1f;                                       /* :66 clamp(min=0.1) */
        float v1 = b + sqrtf(disc), v2 = b - sqrtf(disc);
        float rad = ceilf(3.0f * sqrtf(v1 > v2 ? v1 : v2));                 /* :68 */
        float u = uv[2 * i], v = uv[2 * i + 1];
        /* :73-76 float -> int32 assignment truncates toward zero */
        float f0 = (u - rad) / (float)BLOCK_X, f1 = (v - rad) / (float)BLOCK_Y;
        float f2 = (u + rad + (float)BLOCK_X - 1.0f) / (float)BLOCK_X;
        float f3 = (v + rad + (float)BLOCK_Y - 1.0f) / (float)BLOCK_Y;
        /* Non-finite intermediates (NaN/Inf covariances) make the torch float->int
         * conversion undefined; we define such rows as masked out (all zero).   */
        int finite = isfinite(f0) && isfinite(f1) && isfinite(f2) && isfinite(f3) &&
                     isfinite(k0) && isfinite(k1) && isfinite(k2) && isfinite(rad) &&
                     fabsf(f0) < 2.0e9f && fabsf(f1) < 2.0e9f && fabsf(f2) < 2.0e9f && fabsf(f3) < 2.0e9f;
        int nt = 0;
        if (finite) {
            int x0 = imin(imax((int)f0, 0), gx), y0 = imin(imax((int)f1, 0), gy);      /* :82-95 */
            int x1 = imin(imax((int)f2, 0), gx), y1 = imin(imax((int)f3, 0), gy);
            nt = (x1 - x0) * (y1 - y0);
        }
        int mask = finite && (nt != 0) && (det != 0.0f) && visible[i];      /* :100-101 */
        conic[3 * i] = mask ? k0 : 0.0f; conic[3 * i + 1] = mask ? k1 : 0.0f; conic[3 * i + 2] = mask ? k2 : 0.0f;
        radius[i] = mask ? (int)rad : 0;
        tiles[i] = mask ? nt : 0;
    }
}

/* ------------------------------------------------------------------------- */
/* K7-K10: src/compute_sh.cu:32-195 / src/compute_sh_free.cu (no +0.5, no clamp). */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
static const int NUM_SH_BASES[4] = {1, 4, 9, 16};

/* Per-point stride of shs is num_sh_bases[deg] (compute_sh.cu:45), NOT shs.size(1): preserved. */
void orc_compute_sh_fwd(int P, const float *shs, int deg, const float *dirs, const unsigned char *visible,
                        int free_variant, float *colors, unsigned char *clamped) {
    memset(colors, 0, sizeof(float) * 3 * (size_t)P);
    if (clamped) memset(clamped, 1, 3 * (size_t)P); /* torch::ones (compute_sh.cu:245) */
    int nb = NUM_SH_BASES[deg];
    for (int i = 0; i < P; ++i) {
        if (!visible[i]) continue;
        const float *sh = shs + (size_t)i * nb * 3;
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        for (int ch = 0; ch < 3; ++ch) {
#define S(k) sh[(k) * 3 + ch]
            float r = SH_C0 * S(0);
            if (deg > 0) {
                r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
                if (deg > 1) {
                    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                        SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                    if (deg > 2) {
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
                if (clamped) clamped[3 * i + ch] = (r < 0);
                colors[3 * i + ch] = r < 0.0f ? 0.0f : r;
            } else {
                colors[3 * i + ch] = r;
            }
        }
    }
}

void orc_compute_sh_bwd(int P, const float *shs, int deg, const float *dirs, const unsigned char *visible,
                        const unsigned char *clamped /* NULL for the free variant */, const float *dL_dcolors,
                        int S_alloc /* shs.size(1): dL_dshs is [P,S_alloc,3] zero-init */,
                        float *dL_dshs, float *dL_ddirs) {
    memset(dL_dshs, 0, sizeof(float) * 3 * (size_t)S_alloc * (size_t)P);
    memset(dL_ddirs, 0, sizeof(float) * 3 * (size_t)P);
    int nb = NUM_SH_BASES[deg];
    for (int i = 0; i < P; ++i) {
        if (!visible[i]) continue;
        const float *sh = shs + (size_t)i * nb * 3;
        float *dsh = dL_dshs + (size_t)i * nb * 3;
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float gdir[3] = {0, 0, 0};
        for (int ch = 0; ch < 3; ++ch) {
            float g = dL_dcolors[3 * i + ch];
            if (clamped) g *= clamped[3 * i + ch] ? 0.0f : 1.0f;
#define S(k) sh[(k) * 3 + ch]
#define DS(k) dsh[(k) * 3 + ch]
            float dx = 0, dy = 0, dz = 0;
            DS(0) = SH_C0 * g;
            if (deg > 0) {
                DS(1) = (-SH_C1 * y) * g; DS(2) = (SH_C1 * z) * g; DS(3) = (-SH_C1 * x) * g;
                dx = -SH_C1 * S(3); dy = -SH_C1 * S(1); dz = SH_C1 * S(2);
                if (deg > 1) {
                    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DS(4) = (SH_C2[0] * xy) * g; DS(5) = (SH_C2[1] * yz) * g;
                    DS(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * g;
                    DS(7) = (SH_C2[3] * xz) * g; DS(8) = (SH_C2[4] * (xx - yy)) * g;
                    dx += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
                    dy += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
                    dz += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
                    if (deg > 2) {
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
        dL_ddirs[3 * i] = gdir[0]; dL_ddirs[3 * i + 1] = gdir[1]; dL_ddirs[3 * i + 2] = gdir[2];
    }
}

/* ------------------------------------------------------------------------- */
/* sort_gaussian: gs/sort_gaussian.py:41-54 + src/sort_gaussian.cu:15-69.     */
typedef struct { int64_t key; int32_t seq; int32_t idx; } isect_t;
static int isect_cmp(const void *a, const void *b) {
    const isect_t *x = (const isect_t *)a, *y = (const isect_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq); /* emulates the stable sort */
}

/* returns I = number of intersections (= cumsum(tiles)[P-1]). */
int64_t orc_count_intersections(int P, const int *tiles) {
    int64_t I = 0;
    for (int i = 0; i < P; ++i) I += tiles[i];
    return I;
}

/* key_sorted may be NULL.  tile_range is [ntiles,2] zero-initialised here. */
void orc_sort_gaussian(int P, const float *uv, const float *depth, int W, int H, const int *radius,
                       const int *tiles, int64_t I, int *idx_sorted, int64_t *key_sorted, int *tile_range) {
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    memset(tile_range, 0, sizeof(int) * 2 * (size_t)gx * gy);
    if (I <= 0) return;
    isect_t *buf = (isect_t *)calloc((size_t)I, sizeof(isect_t));
    int64_t cur = 0; /* cumsum offset of this Gaussian (sort_gaussian.cu:30) */
    for (int i = 0; i < P; ++i) {
        if (radius[i] > 0) {
            int x0, y0, x1, y1;
            get_rect(uv[2 * i], uv[2 * i + 1], radius[i], gx, gy, &x0, &y0, &x1, &y1);
            int32_t dbits; memcpy(&dbits, depth + i, 4);
            int64_t depth_id = (int64_t)dbits; /* sign-extended (:32) */
            int64_t w = cur;
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x) {
                    if (w >= I) break;
                    int64_t tile_id = (int64_t)y * gx + x;
                    buf[w].key = (tile_id << 32) | depth_id;
                    buf[w].idx = i; buf[w].seq = (int32_t)w;
                    ++w;
                }
        }
        cur += tiles[i];
    }
    /* slots never written keep key 0 / idx 0, exactly like the zero-initialised reference buffers */
    for (int64_t k = 0; k < I; ++k) if (buf[k].seq == 0 && k != 0) buf[k].seq = (int32_t)k;
    qsort(buf, (size_t)I, sizeof(isect_t), isect_cmp);
    for (int64_t k = 0; k < I; ++k) {
        idx_sorted[k] = buf[k].idx;
        if (key_sorted) key_sorted[k] = buf[k].key;
        int cur_t = (int)(buf[k].key >> 32);
        if (k == 0) tile_range[2 * cur_t] = 0;
        if (k == I - 1) tile_range[2 * cur_t + 1] = (int)I;
        if (k > 0) {
            int prev_t = (int)(buf[k - 1].key >> 32);
            if (prev_t != cur_t) { tile_range[2 * prev_t + 1] = (int)k; tile_range[2 * cur_t] = (int)k; }
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------------- */
/* K15/K17/K19 forward: src/alpha_blending.cu:16-110,
 * src/alpha_blending_enhanced.cu:16-134, src/alpha_blending_with_bias.cu.
 * feature is [P,C] row-major (the reference transposes to [C,P] itself, :282).
 * gs_idx ([H,W,K], -1 padded) and opacity_bias are optional (NULL).
 * fragile ([H,W], optional): set to 1 where any discrete decision of that pixel
 * (power>0, alpha<1/255, next_T<1e-4) lies within relative frag_eps of its
 * threshold, i.e. where an implementation with ~1e-7 different exp() may
 * legitimately take the other branch.                                        */
void orc_alpha_blend_fwd(int P, int C, int W, int H, int K, int enable_truncation,
                         const float *uv, const float *conic, const float *opacity, const float *feature,
                         const float *opacity_bias, const int *idx_sorted, const int *tile_range, float bg,
                         float *rendered, float *final_T, int *ncontrib, int *gs_idx,
                         unsigned char *fragile, float frag_eps) {
    (void)P;
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const double thr = 1.0 / (double)255.0f; /* `alpha < 1.0 / 255.0f` is a double compare (:87) */
    if (gs_idx) for (size_t k = 0; k < (size_t)H * W * K; ++k) gs_idx[k] = -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        int r0 = tile_range[2 * tile], r1 = tile_range[2 * tile + 1];
        float *F = (float *)malloc(sizeof(float) * (size_t)(C > 0 ? C : 1));
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                size_t pix = (size_t)W * py + px;
                float pxf = (float)px, pyf = (float)py;
                float T = 1.0f;
                unsigned contributor = 0, last_contributor = 0;
                int layer_cnt = 0, frag = 0, done = 0;
                for (int c = 0; c < C; ++c) F[c] = 0.0f;
                for (int k = r0; k < r1 && !done; ++k) {
                    int g = idx_sorted[k];
                    contributor++;
                    float vx = uv[2 * g] - pxf, vy = uv[2 * g + 1] - pyf;
                    float ca = conic[3 * g], cb = conic[3 * g + 1], cc = conic[3 * g + 2];
                    float power = -0.5f * (ca * vx * vx + cc * vy * vy) - cb * vx * vy;
                    if (fabsf(power) < frag_eps) frag = 1;
                    if (power > 0) continue;
                    float alpha = opacity[g] * expf(power);
                    if (opacity_bias) alpha = alpha + opacity_bias[g];
                    alpha = fminf(0.99f, alpha);
                    if (fabs((double)alpha - thr) < (double)frag_eps * thr) frag = 1;
                    if ((double)alpha < thr) continue;
                    float next_T = T * (1 - alpha);
                    if (fabsf(next_T - 0.0001f) < frag_eps * 0.0001f) frag = 1;
                    if (next_T < 0.0001f) { done = 1; continue; }
                    for (int c = 0; c < C; ++c) F[c] += feature[(size_t)g * C + c] * alpha * T;
                    T = next_T;
                    last_contributor = contributor;
                    if (gs_idx) {
                        if (enable_truncation) {
                            gs_idx[pix * K + layer_cnt] = g;
                            layer_cnt++;
                            if (layer_cnt >= K) { done = 1; continue; }
                        } else if (layer_cnt < K) {
                            gs_idx[pix * K + layer_cnt] = g;
                            layer_cnt++;
                        }
                    }
                }
                final_T[pix] = T;
                ncontrib[pix] = (int)last_contributor;
                if (fragile) fragile[pix] = (unsigned char)frag;
                for (int c = 0; c < C; ++c) rendered[(size_t)c * H * W + pix] = F[c] + T * bg;
            }
        free(F);
    }
}

/* K16/K18/K20 backward: src/alpha_blending.cu:112-249 (+ with_bias.cu).
 * Per-Gaussian sums are accumulated in double and narrowed at the end.
 * dL_dopacity_bias may be NULL (no-bias variants).                           */
void orc_alpha_blend_bwd(int P, int C, int W, int H,
                         const float *uv, const float *conic, const float *opacity, const float *feature,
                         const float *opacity_bias, const int *idx_sorted, const int *tile_range, float bg,
                         const float *final_T, const int *ncontrib, const float *dL_drendered,
                         float *dL_duv, float *dL_dabs_uv, float *dL_dconic, float *dL_dopacity,
                         float *dL_dfeature /* [P,C] */, float *dL_dopacity_bias) {
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    size_t nacc = (size_t)P * (size_t)(9 + C);
    double *acc = (double *)calloc(nacc ? nacc : 1, sizeof(double));
    /* layout per Gaussian: duv(2) dabs(2) dconic(3) dop(1) dbias(1) dfeat(C) */
    int S = 9 + C;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        int r0 = tile_range[2 * tile], r1 = tile_range[2 * tile + 1];
        float *buf = (float *)malloc(sizeof(float) * 3 * (size_t)(C > 0 ? C : 1));
        float *accum_rec = buf, *last_feature = buf + C, *dpix = buf + 2 * C;
        for (int ly = 0; ly < BLOCK_Y; ++ly)
            for (int lx = 0; lx < BLOCK_X; ++lx) {
                int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                size_t pix = (size_t)W * py + px;
                float pxf = (float)px, pyf = (float)py;
                const float T_final = final_T[pix];
                float T = T_final;
                unsigned contributor = (unsigned)(r1 - r0);
                const int last_contributor = ncontrib[pix];
                float last_alpha = 0;
                for (int c = 0; c < C; ++c) {
                    accum_rec[c] = 0; last_feature[c] = 0;
                    dpix[c] = dL_drendered[(size_t)c * H * W + pix];
                }
                for (int k = r1 - 1; k >= r0; --k) { /* back to front (:198 idx_sorted[range.y - progress - 1]) */
                    int g = idx_sorted[k];
                    contributor--;
                    if (contributor >= (unsigned)last_contributor) continue;
                    float vx = uv[2 * g] - pxf, vy = uv[2 * g + 1] - pyf;
                    float ca = conic[3 * g], cb = conic[3 * g + 1], cc = conic[3 * g + 2];
                    float power = -0.5f * (ca * vx * vx + cc * vy * vy) - cb * vx * vy;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float opac = opacity[g];
                    float a0 = opac * G;
                    if (opacity_bias) a0 = a0 + opacity_bias[g];
                    const float alpha = fminf(0.99f, a0);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    double *A = acc + (size_t)g * S;
                    const float dGx = -G * vx * ca - G * vy * cb;
                    const float dGy = -G * vy * cc - G * vx * cb;
                    /* The reference launches one kernel per chunk of <=32 channels (alpha_blending.cu:440-576):
                     * dL_dalpha, and therefore the |.| that feeds dL_dabs_uv, is formed per chunk and the chunk
                     * results are summed by the atomics.  Same here. */
                    float dL_dalpha = 0.0f, abs0 = 0.0f, abs1 = 0.0f;
                    for (int cb0 = 0; cb0 < C; cb0 += 32) {
                        const int cb1 = cb0 + 32 < C ? cb0 + 32 : C;
                        float da = 0.0f, bgd = 0.0f;
                        for (int c = cb0; c < cb1; ++c) {
                            const float cf = feature[(size_t)g * C + c];
                            accum_rec[c] = last_alpha * last_feature[c] + (1.f - last_alpha) * accum_rec[c];
                            last_feature[c] = cf;
                            da += (cf - accum_rec[c]) * dpix[c];
                            float v = dchannel_dcolor * dpix[c];
#pragma omp atomic
                            A[9 + c] += (double)v;
                        }
                        da *= T;
                        for (int c = cb0; c < cb1; ++c) bgd += bg * dpix[c];
                        da += (-T_final / (1.f - alpha)) * bgd;
                        abs0 += fabsf(opac * da * dGx);
                        abs1 += fabsf(opac * da * dGy);
                        dL_dalpha += da;
                    }
                    last_alpha = alpha;
                    const float dL_dG = opac * dL_dalpha;
                    float g0 = dL_dG * dGx, g1 = dL_dG * dGy;
                    float g2 = -0.5f * G * vx * vx * dL_dG, g3 = -G * vx * vy * dL_dG, g4 = -0.5f * G * vy * vy * dL_dG;
                    float g5 = G * dL_dalpha;
#pragma omp atomic
                    A[0] += (double)g0;
#pragma omp atomic
                    A[1] += (double)g1;
#pragma omp atomic
                    A[2] += (double)abs0;
#pragma omp atomic
                    A[3] += (double)abs1;
#pragma omp atomic
                    A[4] += (double)g2;
#pragma omp atomic
                    A[5] += (double)g3;
#pragma omp atomic
                    A[6] += (double)g4;
#pragma omp atomic
                    A[7] += (double)g5;
#pragma omp atomic
                    A[8] += (double)dL_dalpha;
                }
            }
        free(buf);
    }
    for (int g = 0; g < P; ++g) {
        const double *A = acc + (size_t)g * S;
        dL_duv[2 * g] = (float)A[0]; dL_duv[2 * g + 1] = (float)A[1];
        dL_dabs_uv[2 * g] = (float)A[2]; dL_dabs_uv[2 * g + 1] = (float)A[3];
        dL_dconic[3 * g] = (float)A[4]; dL_dconic[3 * g + 1] = (float)A[5]; dL_dconic[3 * g + 2] = (float)A[6];
        dL_dopacity[g] = (float)A[7];
        if (dL_dopacity_bias) dL_dopacity_bias[g] = (float)A[8];
        for (int c = 0; c < C; ++c) dL_dfeature[(size_t)g * C + c] = (float)A[9 + c];
    }
    free(acc);
}

int orc_version(void) { return 1; }

/* Diagnostics (not part of the reference): for every list entry, how many pixels of its tile pass the hit test
 * (power <= 0 and alpha >= 1/255), ignoring transmittance termination, and how many of the tile's eight 8x4 pixel
 * blocks contain at least one such pixel. */
void orc_count_entry_hits(int W, int H, const float *uv, const float *conic, const float *opacity,
                          const int *idx_sorted, const int *tile_range, int *hits, int *warp_hits) {
    int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        int tx = tile % gx, ty = tile / gx;
        for (int k = tile_range[2 * tile]; k < tile_range[2 * tile + 1]; ++k) {
            int g = idx_sorted[k], n = 0, wmask = 0;
            for (int ly = 0; ly < BLOCK_Y; ++ly)
                for (int lx = 0; lx < BLOCK_X; ++lx) {
                    int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                    if (px >= W || py >= H) continue;
                    float vx = uv[2 * g] - (float)px, vy = uv[2 * g + 1] - (float)py;
                    float power = -0.5f * (conic[3 * g] * vx * vx + conic[3 * g + 2] * vy * vy) - conic[3 * g + 1] * vx * vy;
                    if (power > 0) continue;
                    if (fminf(0.99f, opacity[g] * expf(power)) < 1.0f / 255.0f) continue;
                    ++n;
                    wmask |= 1 << ((ly / 4) * 2 + (lx / 8));
                }
            hits[k] = n;
            warp_hits[k] = __builtin_popcount(wmask);
        }
    }
}
