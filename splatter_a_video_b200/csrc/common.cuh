// Shared helpers for the sm_100a rasterizer kernels (no torch, no third-party headers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#define SPV_TILE 16
#define SPV_TILE_PIX 256

namespace spv {

// Error plumbing: every extern "C" entry returns this after its launches.
void set_error(cudaError_t e, const char *where);
void timer_mark(int slot, int edge, cudaStream_t s);  // no-op unless spv_kernel_timer_enable(1)
int get_option(const char *name);  // runtime switches of experimental kernel variants (runtime.cu); 0 = validated default
void count_launches(int n);  // bookkeeping for spv_launch_count(): kernels this library has launched
inline int check_launch(const char *where, int launches = 1) {
    count_launches(launches);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error(e, where);
        (void)cudaGetLastError();  // clear the sticky-less error so later calls can proceed
        return (int)e;
    }
    return 0;
}
#define SPV_CUDA_TRY(expr, where)                 \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) {                  \
            spv::set_error(_e, where);            \
            return (int)_e;                       \
        }                                         \
    } while (0)

// cudaFuncSetAttribute applies to the CURRENT device only: opt-ins to more than 48 KB of dynamic shared memory are made once
// per (kernel, device).  `done` is a per-call-site bit mask over device ordinals (devices >= 64 simply repeat the cheap call).
template <typename Kernel>
inline void opt_in_dynamic_smem(Kernel kernel, size_t bytes, std::atomic<unsigned long long> &done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    // ... and the maximal carve-out: an SM keeps its shared-memory / L1 split while CTAs are resident, so kernels meant to run
    // side by side (the three image losses; a frame's blend next to another frame's) must ask for the same split
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (bit) done.fetch_or(bit, std::memory_order_release);
}
// the same carve-out preference for a kernel with static shared memory only
template <typename Kernel>
inline void prefer_max_carveout(Kernel kernel, std::atomic<unsigned long long> &done) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (bit) done.fetch_or(bit, std::memory_order_release);
}
inline int sm_count() {   // of the current device (the runtime caches device attributes: this is a table lookup)
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

// A second stream for the branches of a call that do not depend on each other (frame forward: SH colours + record packing + the
// id-image fill next to the projection / binning / sort chain; frame backward: the SH gradient next to the covariance chain; lazy
// Adam: the dense parameters next to the spline intervals).  Forked from and joined back into the caller's stream with events, so
// the caller still sees one in-order stream -- and a stream capture records the branches as parallel graph nodes.  One lane per
// (host thread, device): streams and events belong to the device that was current when they were created.  runtime.cu.
// `stream`: side work next to the caller's stream.  `hi`: a stream of the greatest priority for serial chains of small kernels
// (the binning chain): the block scheduler hands free SM slots to its kernels before the backlog of a big side-lane grid.
struct SideLane {
    cudaStream_t stream = nullptr, hi = nullptr;
    cudaEvent_t fork = nullptr, mid = nullptr, join = nullptr, hi_join = nullptr;
    bool ok = false;
};
SideLane *side_lane();

inline int tiles_x(int W) { return (W + SPV_TILE - 1) / SPV_TILE; }
inline int tiles_y(int H) { return (H + SPV_TILE - 1) / SPV_TILE; }
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Tile rectangle of a splat (semantics of the reference's get_rect, include/utils.h:17-37):
// truncating float->int, clamped to [0, grid].
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy,
                                          int &x0, int &y0, int &x1, int &y1) {
    const float r = (float)radius;
    x0 = min(gx, max(0, (int)((px - r) / 16.0f)));
    y0 = min(gy, max(0, (int)((py - r) / 16.0f)));
    x1 = min(gx, max(0, (int)((px + r + 16.0f - 1.0f) / 16.0f)));
    y1 = min(gy, max(0, (int)((py + r + 16.0f - 1.0f) / 16.0f)));
}

// Conservative ellipse-vs-pixel-rectangle test shared by the tile culling (sort.cu) and the per-warp pre-filter of the blend
// kernels: can q(d) = 0.5 a dx^2 + b dx dy + 0.5 c dy^2 (d = centre - pixel) drop to <= tau somewhere on the rectangle of
// pixel centres [x0,x1] x [y0,y1]?  q is a convex quadratic: zero if the centre is inside, else its minimum is on an edge.
__device__ __forceinline__ float quad_min_on_segment(float A, float B, float C0, float lo, float hi) {
    // min over t in [lo,hi] of A t^2 + B t + C0, A >= 0
    // fast division: any t in [lo,hi] gives a value >= the true minimum, and a 2-ulp error of the minimiser moves the value by
    // O(1e-13) of it -- far inside the safety margin of tile_may_hit
    float t = (A > 0.f) ? fminf(fmaxf(__fdividef(-B, 2.f * A), lo), hi) : ((B > 0.f) ? lo : hi);
    return (A * t + B) * t + C0;
}

__device__ __forceinline__ bool tile_may_hit(float cx, float cy, float a, float b, float c, float tau, float x0, float y0,
                                             float x1, float y1) {
    // rectangle of pixel centres [x0,x1] x [y0,y1]; d = centre - pixel
    if (cx >= x0 && cx <= x1 && cy >= y0 && cy <= y1) return true;
    const float dxl = cx - x0, dxh = cx - x1, dyl = cy - y0, dyh = cy - y1;   // dx in [dxh, dxl], dy in [dyh, dyl]
    // q(dx,dy) = 0.5 a dx^2 + b dx dy + 0.5 c dy^2 ; the minimum over the box lies on its boundary here
    float m = quad_min_on_segment(0.5f * a, b * dyl, 0.5f * c * dyl * dyl, dxh, dxl);
    m = fminf(m, quad_min_on_segment(0.5f * a, b * dyh, 0.5f * c * dyh * dyh, dxh, dxl));
    m = fminf(m, quad_min_on_segment(0.5f * c, b * dxl, 0.5f * a * dxl * dxl, dyh, dyl));
    m = fminf(m, quad_min_on_segment(0.5f * c, b * dxh, 0.5f * a * dxh * dxh, dyh, dyl));
    return m <= tau * 1.0005f + 1e-4f;
}


// Upstream image gradients given as one [H,W] plane per channel (NULL = that channel has no gradient).  Lets the fused
// frame path hand the autograd engine's separate per-image gradients to the kernel without concatenating them.
struct ChanPlanes { const float *p[32]; };

// blend.cu: grouped forward with the -1 fill of the id image optional (frame.cu fills it on its side stream)
int blend_groups_forward(int P, int C, int W, int H, int K, const float *uv, const float *conic, const float *opacity,
                         const float *feature, const int *idx_sorted, const int *tile_range, float bg_rgb, float bg_depth,
                         float bg_attr, float *rendered, float *final_T, int *ncontrib, int *gs_idx, bool fill_idx, void *stream);

// Packed per-Gaussian gradient row of the grouped blend backward (stride kPackedRowGroups floats):
//   0,1 dL_duv(all)  2,3 |RGB-pass dL_duv|  4,5,6 dL_dconic  7 dL_dopacity  8..8+C-1 dL_dfeature  31,32 RGB-pass dL_duv
constexpr int kPackedRowGroups = 36;

// geometry.cu: the per-Gaussian geometry of one orthographic frame in one pass each way (frame.cu)
int frame_geometry_forward(int P, const float *xyz, const float *scales, const float *uquats, const float *extr, int W, int H,
                           float nearest, float extent, float *uv, float *depth, uint8_t *vis, float *cov3d, float *conic, int *radius,
                           int *tiles, int *radii_out, void *stream);
int frame_geometry_backward(int P, const float *packed, const float *scales, const float *uquats, const float *extr, int W, int H,
                            const float *depth, const uint8_t *vis, const float *cov3d, const int *radius, float *dL_dxyz,
                            float *dL_dscales, float *dL_duquats, void *stream);

// sort.cu: spv_bin_tiles + the longest-list-first tile order the blend kernels launch in (tile_order = int[T] or NULL)
int bin_tiles_ordered(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                      const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status, int *tile_order,
                      bool cleared, void *workspace, size_t ws_bytes, void *stream);
// clears tile_range / status / the workspace's counters in one launch (bin_tiles_ordered(cleared = true) then skips it)
int bin_tiles_clear(int P, int64_t I_cap, int W, int H, int *tile_range, int *status, void *workspace, void *stream);

// blend_rec.cu: record-staged blending of the fused frame path.  One record of kRecordFloats floats per Gaussian:
//   [x y a2 b2 | c2 log2(o) o id | feature[0..23] = rgb(3) depth(1) attributes, zero padded | a b c 0]
constexpr int kRecordFloats = 36;
int pack_records(int P, int A, const float *uv, const float *conic, const float *opacity, const int *radius, const float *rgb,
                 const float *depth, int n_groups, const float *const *attr_ptrs, const int *attr_channels, float *rec,
                 void *stream);
int blend_records_forward(int C, int W, int H, int K, const float *rec, const int *idx_sorted, const int *tile_range,
                          const int *tile_order, float bg_rgb, float bg_depth, float bg_attr, float *rendered, float *final_T,
                          int *ncontrib, int *gs_idx, void *stream);
int blend_records_backward(int P, int C, int W, int H, const float *rec, const int *idx_sorted, const int *tile_range,
                           const int *tile_order, float bg_rgb, float bg_depth, float bg_attr, const float *final_T, const int *ncontrib,
                           const float *const *planes_host, int n_grad_channels, bool want_abs, float *packed, bool packed_is_zero,
                           void *stream);

}  // namespace spv
