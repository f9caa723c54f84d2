// Tile binning for sm_100a: inclusive scan of tiles-touched, (tile|depth) key emission, radix sort of the
// SIGNIFICANT key bits only, per-tile ranges.  Integer work: results must be bit-exact.
//
// Reference: gs/sort_gaussian.py:41-54 (torch.cumsum, torch.sort over all 64 bits carrying int64 indices,
// torch.gather) + src/sort_gaussian.cu:15-69.  Differences that keep the result identical:
//   * key = (tile << 32) | fp32 bits of depth, as there, but only bits [0, 32 + ceil(log2(#tiles))) are sorted
//     (43 bits at 480p instead of 64 -> 6 instead of 8 radix passes) and the payload is the 4-byte Gaussian id
//     itself, so there is no int64 index array and no gather pass (12 B/intersection/pass instead of 16 + gather).
//   * LSD radix sort is stable, so equal (tile, depth) keys keep emission order = ascending Gaussian id, the
//     order torch.sort(stable) over the reference's emission order yields.
//   * depth must be positive (it always is after near culling); the reference's sign extension of negative
//     depths corrupts its own tile bits (sort_gaussian.cu:32,36), which we do not reproduce.
// The radix sort itself is CUB's DeviceRadixSort (header-only CCCL 12.9 templates instantiated in this TU);
// everything around it is hand-written.
#include <cub/cub.cuh>

#include "common.cuh"
#include "../../include/spv_b200.h"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
emit_keys_kernel(int P, const float2 *__restrict__ uv, const float *__restrict__ depth, const int *__restrict__ radius,
                 const int *__restrict__ offsets, int gx, int gy, long long I, unsigned long long *__restrict__ keys,
                 int *__restrict__ vals) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const int r = radius[i];
    if (r <= 0) return;
    int x0, y0, x1, y1;
    const float2 c = uv[i];
    spv::tile_rect(c.x, c.y, r, gx, gy, x0, y0, x1, y1);
    long long cur = (i == 0) ? 0 : offsets[i - 1];
    const unsigned long long dbits = (unsigned long long)__float_as_uint(depth[i]);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            if (cur >= I) return;  // inconsistent tiles[] (never with our own ewa_project outputs)
            keys[cur] = ((unsigned long long)(y * gx + x) << 32) | dbits;
            vals[cur] = i;
            ++cur;
        }
}

__global__ void __launch_bounds__(kThreads)
tile_range_kernel(long long I, const unsigned long long *__restrict__ keys_sorted, int2 *__restrict__ tile_range) {
    const long long k = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (k >= I) return;
    const int cur = (int)(keys_sorted[k] >> 32);
    if (k == 0) tile_range[cur].x = 0;
    else {
        const int prev = (int)(keys_sorted[k - 1] >> 32);
        if (prev != cur) { tile_range[prev].y = (int)k; tile_range[cur].x = (int)k; }
    }
    if (k == I - 1) tile_range[cur].y = (int)I;
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

inline int key_end_bit(int ntiles) {
    int b = 0;
    while ((1 << b) < ntiles) ++b;
    return 32 + (b > 0 ? b : 1);
}

size_t sort_temp_bytes(long long I) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, bytes, (const unsigned long long *)nullptr,
                                    (unsigned long long *)nullptr, (const int *)nullptr, (int *)nullptr, (int)I, 0, 64);
    return bytes;
}

}  // namespace

extern "C" {

size_t spv_sort_scan_workspace_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, bytes, (const int *)nullptr, (int *)nullptr, P > 0 ? P : 1);
    return align_up(bytes);
}

int spv_sort_scan(int P, const int *tiles, int *offsets, void *workspace, size_t ws_bytes, void *stream) {
    if (P <= 0) return 0;
    size_t need = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, need, tiles, offsets, P);
    if (ws_bytes < need) { spv::set_error(cudaErrorInvalidValue, "spv_sort_scan: workspace too small"); return (int)cudaErrorInvalidValue; }
    SPV_CUDA_TRY(cub::DeviceScan::InclusiveSum(workspace, need, tiles, offsets, P, (cudaStream_t)stream), "spv_sort_scan");
    return spv::check_launch("spv_sort_scan", 2);
}

size_t spv_sort_workspace_bytes(int P, int64_t I) {
    (void)P;
    if (I <= 0) return 256;
    return align_up(8 * (size_t)I) * 2 + align_up(4 * (size_t)I) + align_up(sort_temp_bytes(I));
}

int spv_sort_gaussian(int P, int64_t I, const float *uv, const float *depth, const int *radius, const int *offsets,
                      int W, int H, int *idx_sorted, int *tile_range, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H);
    SPV_CUDA_TRY(cudaMemsetAsync(tile_range, 0, sizeof(int) * 2 * (size_t)gx * gy, s), "spv_sort_gaussian");
    if (P <= 0 || I <= 0) return 0;  // the reference launches <<<0,256>>> here (invalid, unchecked)
    if (I >= (1ll << 31)) { spv::set_error(cudaErrorInvalidValue, "spv_sort_gaussian: more than 2^31-1 intersections"); return (int)cudaErrorInvalidValue; }
    if (ws_bytes < spv_sort_workspace_bytes(P, I)) { spv::set_error(cudaErrorInvalidValue, "spv_sort_gaussian: workspace too small"); return (int)cudaErrorInvalidValue; }
    char *w = (char *)workspace;
    unsigned long long *keys_in = (unsigned long long *)w; w += align_up(8 * (size_t)I);
    unsigned long long *keys_out = (unsigned long long *)w; w += align_up(8 * (size_t)I);
    int *vals_in = (int *)w; w += align_up(4 * (size_t)I);
    size_t temp = sort_temp_bytes(I);
    // zero-fill like the reference's torch::zeros so slots an inconsistent caller leaves unwritten are defined
    SPV_CUDA_TRY(cudaMemsetAsync(keys_in, 0, 8 * (size_t)I, s), "spv_sort_gaussian");
    SPV_CUDA_TRY(cudaMemsetAsync(vals_in, 0, 4 * (size_t)I, s), "spv_sort_gaussian");
    emit_keys_kernel<<<spv::cdiv(P, kThreads), kThreads, 0, s>>>(P, (const float2 *)uv, depth, radius, offsets, gx, gy,
                                                                 (long long)I, keys_in, vals_in);
    int rc = spv::check_launch("spv_sort_gaussian/emit");
    if (rc) return rc;
    SPV_CUDA_TRY(cub::DeviceRadixSort::SortPairs((void *)w, temp, (const unsigned long long *)keys_in, keys_out,
                                                 (const int *)vals_in, idx_sorted, (int)I, 0, key_end_bit(gx * gy), s),
                 "spv_sort_gaussian/sort");
    tile_range_kernel<<<spv::cdiv(I, kThreads), kThreads, 0, s>>>((long long)I, keys_out, (int2 *)tile_range);
    // CUB onesweep: histogram + exclusive-sum + one kernel per 8-bit digit
    return spv::check_launch("spv_sort_gaussian/range", 1 + 2 + (key_end_bit(gx * gy) + 7) / 8);
}

}  // extern "C"
