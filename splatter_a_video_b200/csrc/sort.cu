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

// ---- exact tile culling + capacity-bounded emission (fused frame path; no host synchronisation) -----------------------
// The reference bins a splat into every tile of the bounding square of radius ceil(3*sigma_max) (utils.h:17-37); 31 % of
// those (tile, splat) pairs contain no pixel with alpha >= 1/255 on the DAVIS-shaped workload.  Dropping them cannot change
// any pixel: a pixel takes a splat only if power >= -tau with tau = ln(255*opacity), i.e. only inside the ellipse
// q(d) = 0.5*(a dx^2 + c dy^2) + b dx dy <= tau, and the test below keeps every tile whose pixel-centre rectangle
// intersects that ellipse (minimum of the convex quadratic over the rectangle, with a relative safety margin).
// Pass 1 (EMIT = false): per-splat count of surviving tiles + a 64-bit survival mask over its candidate rectangle (row-major;
// rectangles with more than 64 tiles -- rare -- are simply re-tested in pass 2).  Pass 2 (EMIT = true): write (key, id) at
// the scanned offsets.  EIGHT lanes share one splat and stride over its candidate tiles: tile rectangles are heavy-tailed
// (2 % of the splats cover 36+ tiles), and with one thread per splat half of all warps wait on such a lane.
constexpr int kLanesPerSplat = 8;

// MODE 0: count (per-splat counts + masks)   1: emit at the scanned per-splat offsets (radix path)
// MODE 2: count + per-tile histogram          3: emit (depth|id) keys into the per-tile segments (tile-sort path)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
cull_count_emit_kernel(int P, const float2 *__restrict__ uv, const float *__restrict__ depth, const int *__restrict__ radius,
                       const float *__restrict__ conic, const float *__restrict__ opacity, int cull, int W, int H, int gx,
                       int gy, int *__restrict__ counts, unsigned long long *__restrict__ masks,
                       const int *__restrict__ offsets, long long cap, unsigned long long *__restrict__ keys,
                       int *__restrict__ vals, int *__restrict__ status, int *__restrict__ tile_count,
                       const int2 *__restrict__ tile_seg) {
    constexpr bool EMIT = (MODE & 1) != 0, TILES = MODE >= 2;
    const int t = blockIdx.x * kThreads + threadIdx.x;
    const int i = t / kLanesPerSplat, sub = t % kLanesPerSplat;          // splat, lane within the splat's group
    const int lane = threadIdx.x & 31, gshift = lane & ~(kLanesPerSplat - 1);
    const bool live = i < P;
    const int r = live ? radius[i] : 0;
    int n = 0;                       // tiles kept so far (identical on the 8 lanes of a group)
    unsigned long long m = 0ull;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, area = 0;
    float2 c = make_float2(0.f, 0.f);
    float ka = 0.f, kb = 0.f, kc = 0.f, tau = 0.f;
    bool any = false, retest = cull != 0;
    long long cur = 0;
    unsigned long long dbits = 0;
    if (r > 0) {
        c = uv[i];
        spv::tile_rect(c.x, c.y, r, gx, gy, x0, y0, x1, y1);
        area = (x1 - x0) * (y1 - y0);
        const float o = opacity[i];
        any = (!cull || (o * 255.f >= 0.999f)) && area > 0;
        ka = conic[3 * i]; kb = conic[3 * i + 1]; kc = conic[3 * i + 2];
        tau = __logf(255.f * o);    // alpha = min(.99, o G) >= 1/255  <=>  o G >= 1/255  <=>  power >= -ln(255 o)
        if (EMIT) {
            if (!TILES) cur = (i == 0) ? 0 : offsets[i - 1];
            dbits = (unsigned long long)__float_as_uint(depth[i]);
            if (cull && area <= 64) { m = masks[i]; retest = false; }
        }
    }
    // every lane of the warp runs the same number of rounds (the ballots below are warp-wide)
    int rounds = any ? (area + kLanesPerSplat - 1) / kLanesPerSplat : 0;
    rounds = __reduce_max_sync(0xffffffffu, rounds);
    const int w = max(x1 - x0, 1);
    const float w_inv = __fdividef(1.f, (float)w);   // k / w through a float reciprocal: exact for k < 2^20 ((k + 0.5) / w is never
                                                     // within 0.5 / w of an integer, the reciprocal is good to 2^-21)
    for (int it = 0; it < rounds; ++it) {
        const int k = it * kLanesPerSplat + sub;
        bool keep = any && k < area;
        int x = 0, y = 0;
        if (keep) {
            const int q = __float2int_rz(((float)k + 0.5f) * w_inv);
            y = y0 + q; x = x0 + (k - q * w);
            if (EMIT && cull && !retest) keep = (m >> k) & 1ull;
            else if (retest) {
                const float px0 = (float)(x * 16), py0 = (float)(y * 16);
                const float px1 = fminf(px0 + 15.f, (float)(W - 1)), py1 = fminf(py0 + 15.f, (float)(H - 1));
                keep = spv::tile_may_hit(c.x, c.y, ka, kb, kc, tau, px0, py0, px1, py1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const unsigned grp = (bal >> gshift) & ((1u << kLanesPerSplat) - 1u);
        if (EMIT && TILES) {
            if (keep) {   // slot order inside a tile's segment is arbitrary: the tile sort orders by (depth, id)
                const int tl = y * gx + x;
                const int2 seg = tile_seg[tl];
                // capacity overflow: the scan clipped this tile's segment (possibly to (0,0)) -- entries beyond it are dropped
                // here instead of spilling into the neighbouring (or, for an empty range, the first) tiles' slots
                if (seg.y > seg.x) {
                    const int pos = seg.x + atomicAdd(&tile_count[tl], 1);
                    if (pos < seg.y) keys[pos] = (dbits << 32) | (unsigned long long)(unsigned)i;
                }
            }
        } else if (EMIT) {
            if (keep) {
                const long long pos = cur + n + __popc(grp & ((1u << sub) - 1u));
                if (pos < cap) {
                    keys[pos] = ((unsigned long long)(y * gx + x) << 32) | dbits;
                    vals[pos] = i;
                }
            }
        } else {
            if (it < 64 / kLanesPerSplat) m |= (unsigned long long)grp << (it * kLanesPerSplat);
            if (TILES && keep) atomicAdd(&tile_count[y * gx + x], 1);
        }
        n += __popc(grp);
    }
    if (!EMIT && live && sub == 0) { if (!TILES) counts[i] = n; masks[i] = m; }
    if (EMIT && !TILES && live && i == P - 1 && sub == 0) {
        const long long total = (long long)offsets[P - 1];
        status[0] = (int)(total < cap ? total : cap);
        status[1] = total > cap ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kThreads)
tile_range_dev_kernel(long long cap, const int *__restrict__ status, const unsigned long long *__restrict__ keys_sorted,
                      int2 *__restrict__ tile_range) {
    const long long k = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long I = status[0];
    if (k >= I || k >= cap) return;
    const int cur = (int)(keys_sorted[k] >> 32);
    if (k == 0) tile_range[cur].x = 0;
    else {
        const int prev = (int)(keys_sorted[k - 1] >> 32);
        if (prev != cur) { tile_range[prev].y = (int)k; tile_range[cur].x = (int)k; }
    }
    if (k == I - 1) tile_range[cur].y = (int)I;
}

// ---- tile-segment binning: per-tile histogram -> scan over tiles -> scatter -> per-tile sort in shared memory --------
// The global radix sort orders (tile | depth) keys although the tile of every entry is known when it is emitted.  Here the
// count pass also histograms the tiles, one CTA scans the T tile counts into segment ranges, the emit pass scatters
// (depth bits << 32 | Gaussian id) keys into the segments (slot order arbitrary), and every tile sorts its own segment in
// shared memory (bitonic network over 64-bit keys; ids are unique, so the result is the unique (depth, id) order -- exactly
// what the stable radix sort over emission order yields).  12 B/intersection of global traffic instead of 6 radix passes.
constexpr int kScanThreads = 1024;
constexpr int kOrderBuckets = 64;         // tile_order: tiles bucketed by list length (32 entries per bucket), longest first
constexpr int kSortSmall = 1024;          // keys a 256-thread CTA sorts in static shared memory (8 KB); the heavy tail of the
                                          // tile-list distribution goes to the 1024-thread CTAs of tile_sort_big_kernel
constexpr int kSortBig = 25600;           // keys a 1024-thread CTA sorts in 200 KB of dynamic shared memory

__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int T, long long cap, int *__restrict__ tile_count, int2 *__restrict__ tile_range, int *__restrict__ status,
                 int *__restrict__ big_queue /*[0] = count, [1] = head, [2..] = tile ids*/, int *__restrict__ tile_order /*or NULL*/) {
    __shared__ long long s_warp[kScanThreads / 32];
    __shared__ long long s_total;
    __shared__ int s_bucket[kOrderBuckets], s_cursor[kOrderBuckets];
    if (tile_order && threadIdx.x < kOrderBuckets) s_bucket[threadIdx.x] = 0;
    const int per = (T + kScanThreads - 1) / kScanThreads;
    const int lo = min(T, (int)threadIdx.x * per), hi = min(T, lo + per);
    long long sum = 0;
    for (int t = lo; t < hi; ++t) sum += tile_count[t];
    // block-wide exclusive scan of the per-thread sums
    long long incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        s_warp[lane] = w;
        if (lane == 31) s_total = w;
    }
    __syncthreads();
    long long run = (incl - sum) + (warp > 0 ? s_warp[warp - 1] : 0);
    for (int t = lo; t < hi; ++t) {
        const int n = tile_count[t];
        const long long a = run < cap ? run : cap, b = (run + n) < cap ? (run + n) : cap;
        tile_range[t] = (n > 0 && b > a) ? make_int2((int)a, (int)b) : make_int2(0, 0);
        if (b - a > kSortSmall) big_queue[2 + atomicAdd(&big_queue[0], 1)] = t;
        if (tile_order) atomicAdd(&s_bucket[min(kOrderBuckets - 1, (int)((b - a) >> 5))], 1);
        tile_count[t] = 0;   // becomes the emit pass's cursor
        run += n;
    }
    if (threadIdx.x == 0) {
        status[0] = (int)(s_total < cap ? s_total : cap);
        status[1] = s_total > cap ? 1 : 0;
    }
    if (!tile_order) return;
    // Launch order of the blend kernels: CTAs are dispatched in blockIdx order as SM slots free up, so listing the tiles longest
    // list first turns the launch into a longest-processing-time-first work queue -- the heavy tail of the list-length
    // distribution (a few tiles under large splats) starts first instead of deciding when the kernel ends.  Counting sort over
    // kOrderBuckets length buckets; the order inside a bucket is irrelevant (tiles are independent).
    __syncthreads();
    if (threadIdx.x == 0) {
        int at = 0;
        for (int q = kOrderBuckets - 1; q >= 0; --q) { s_cursor[q] = at; at += s_bucket[q]; }
    }
    __syncthreads();
    for (int t = lo; t < hi; ++t) {
        const int2 r = tile_range[t];
        tile_order[atomicAdd(&s_cursor[min(kOrderBuckets - 1, (r.y - r.x) >> 5)], 1)] = t;
    }
}

// Bitonic network in its "flip" form (every compare-exchange moves the minimum to the lower index), so positions >= n
// behave as +inf padding without being stored.  keys: shared or global memory.
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_u64(unsigned long long *keys, int n) {
    int lg = 0;                                   // np2 = 1 << lg = next power of two >= n
    while ((1 << lg) < n) ++lg;
    const int half = (1 << lg) >> 1;
    for (int lk = 1; lk <= lg; ++lk) {            // k = 1 << lk
        const int k = 1 << lk, lh = lk - 1;
        for (int i = threadIdx.x; i < half; i += THREADS) {   // flip step: partner = mirror inside the k-block
            const int off = i & ((1 << lh) - 1), base = (i >> lh) << lk;
            const int a = base + off, b = base + (k - 1 - off);
            if (b < n) {
                const unsigned long long x = keys[a], y = keys[b];
                if (x > y) { keys[a] = y; keys[b] = x; }
            }
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; --lj) {    // j = 1 << lj
            for (int i = threadIdx.x; i < half; i += THREADS) {
                const int a = ((i >> lj) << (lj + 1)) + (i & ((1 << lj) - 1)), b = a + (1 << lj);
                if (b < n) {
                    const unsigned long long x = keys[a], y = keys[b];
                    if (x > y) { keys[a] = y; keys[b] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// Register-blocked bitonic network for one tile segment: E consecutive keys per thread (256 E >= n, padded with all-ones keys).
// Compare-exchange distances below E stay inside a thread, distances below 32 E are lane shuffles, only the rest (6 of the 36 / 45 /
// 55 stages for E = 1 / 2 / 4) go through shared memory with a CTA barrier -- the all-shared-memory network above pays a barrier
// per stage (measured 34 us for the 1620 tiles of a 854x480 frame, average list 286 keys).  Keys are unique (Gaussian id in the
// low word), so the result is the same permutation whatever the network.
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int mask) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, mask), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), mask);
    return ((unsigned long long)hi << 32) | lo;
}

template <int E>
__device__ __forceinline__ void tile_sort_regs(const int2 r, const unsigned long long *__restrict__ keys, int *__restrict__ idx_sorted,
                                               unsigned long long *s_keys) {
    constexpr int N = kThreads * E;
    const int n = r.y - r.x, t = threadIdx.x;
    unsigned long long v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = E * t + e < n ? keys[r.x + E * t + e] : ~0ull;
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j >= 1; j >>= 1) {
            if (j < E) {                              // both keys in this thread (jj: compile-time copy of j, registers stay registers)
#pragma unroll
                for (int jj = E >> 1; jj >= 1; jj >>= 1) {
                    if (j != jj) continue;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        if ((e & jj) == 0) {
                            const bool asc = ((E * t + e) & k) == 0;
                            const unsigned long long a = v[e], b = v[e | jj];
                            if ((a > b) == asc) { v[e] = b; v[e | jj] = a; }
                        }
                    }
                }
            } else if (j < 32 * E) {                  // partner in this warp
                const bool lower = (t & (j / E)) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long o = shfl_xor_u64(v[e], j / E);
                    const bool keep_min = lower == (((E * t + e) & k) == 0);
                    v[e] = (keep_min == (v[e] < o)) ? v[e] : o;
                }
            } else {                                  // partner in another warp
#pragma unroll
                for (int e = 0; e < E; ++e) s_keys[E * t + e] = v[e];
                __syncthreads();
                const bool lower = (t & (j / E)) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long o = s_keys[(E * t + e) ^ j];
                    const bool keep_min = lower == (((E * t + e) & k) == 0);
                    v[e] = (keep_min == (v[e] < o)) ? v[e] : o;
                }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (E * t + e < n) idx_sorted[r.x + E * t + e] = (int)(unsigned)(v[e] & 0xffffffffull);
}

__global__ void __launch_bounds__(kThreads)
tile_sort_small_kernel(const int2 *__restrict__ tile_range, const unsigned long long *__restrict__ keys,
                       int *__restrict__ idx_sorted) {
    __shared__ unsigned long long s_keys[kSortSmall];
    const int2 r = tile_range[blockIdx.x];
    const int n = r.y - r.x;
    if (n <= 0 || n > kSortSmall) return;   // larger segments: tile_sort_big_kernel
    if (n <= kThreads) tile_sort_regs<1>(r, keys, idx_sorted, s_keys);
    else if (n <= 2 * kThreads) tile_sort_regs<2>(r, keys, idx_sorted, s_keys);
    else tile_sort_regs<4>(r, keys, idx_sorted, s_keys);
}

// Persistent CTAs over the queue of segments with more than kSortSmall keys (a few dozen heavy tiles on the DAVIS-shaped
// workload -- they would otherwise set the duration of the whole sort; every tile at 5 M Gaussians x 480p).  Up to kSortBig keys in shared memory, beyond that in place in global memory (L2).
__global__ void __launch_bounds__(kScanThreads)
tile_sort_big_kernel(const int2 *__restrict__ tile_range, unsigned long long *__restrict__ keys, int *__restrict__ idx_sorted,
                     int *__restrict__ big_queue) {
    extern __shared__ __align__(16) unsigned long long s_big[];
    __shared__ int s_item;
    const int count = big_queue[0];
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(&big_queue[1], 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= count) return;
        const int2 r = tile_range[big_queue[2 + item]];
        const int n = r.y - r.x;
        if (n <= kSortBig) {
            for (int i = threadIdx.x; i < n; i += kScanThreads) s_big[i] = keys[r.x + i];
            __syncthreads();
            bitonic_sort_u64<kScanThreads>(s_big, n);
            for (int i = threadIdx.x; i < n; i += kScanThreads) idx_sorted[r.x + i] = (int)(unsigned)(s_big[i] & 0xffffffffull);
        } else {
            __syncthreads();
            bitonic_sort_u64<kScanThreads>(keys + r.x, n);   // __syncthreads orders this CTA's global accesses
            for (int i = threadIdx.x; i < n; i += kScanThreads) idx_sorted[r.x + i] = (int)(unsigned)(keys[r.x + i] & 0xffffffffull);
        }
        __syncthreads();
    }
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

// ---- capacity-bounded binning without host synchronisation (fused frame path) -------------------------------------
size_t spv_bin_capacity_workspace_bytes(int P, int64_t I_cap) {
    if (I_cap <= 0) I_cap = 1;
    size_t scan = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, scan, (const int *)nullptr, (int *)nullptr, P > 0 ? P : 1);
    size_t temp = sort_temp_bytes(I_cap);
    if (scan > temp) temp = scan;
    return align_up(8 * (size_t)I_cap) * 2 + align_up(4 * (size_t)I_cap) + align_up(4 * (size_t)(P > 0 ? P : 1)) * 2 +
           align_up(8 * (size_t)(P > 0 ? P : 1)) + align_up(temp);
}

/* Bins the splats of one frame with at most I_cap intersections.  status[0] = number of intersections kept,
 * status[1] = 1 if I_cap was too small (device memory, read it asynchronously).  cull != 0 enables exact tile culling. */
int spv_bin_capacity(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                     const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status,
                     void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H);
    SPV_CUDA_TRY(cudaMemsetAsync(tile_range, 0, sizeof(int) * 2 * (size_t)gx * gy, s), "spv_bin_capacity");
    SPV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * 2, s), "spv_bin_capacity");
    if (P <= 0 || I_cap <= 0) return 0;
    if (I_cap >= (1ll << 31)) { spv::set_error(cudaErrorInvalidValue, "spv_bin_capacity: capacity too large"); return (int)cudaErrorInvalidValue; }
    if (ws_bytes < spv_bin_capacity_workspace_bytes(P, I_cap)) { spv::set_error(cudaErrorInvalidValue, "spv_bin_capacity: workspace too small"); return (int)cudaErrorInvalidValue; }
    char *w = (char *)workspace;
    unsigned long long *keys_in = (unsigned long long *)w; w += align_up(8 * (size_t)I_cap);
    unsigned long long *keys_out = (unsigned long long *)w; w += align_up(8 * (size_t)I_cap);
    int *vals_in = (int *)w; w += align_up(4 * (size_t)I_cap);
    int *counts = (int *)w; w += align_up(4 * (size_t)P);
    int *offsets = (int *)w; w += align_up(4 * (size_t)P);
    unsigned long long *masks = (unsigned long long *)w; w += align_up(8 * (size_t)P);
    size_t scan = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, scan, counts, offsets, P);
    size_t temp = sort_temp_bytes(I_cap);
    const unsigned g = spv::cdiv((long long)P * kLanesPerSplat, kThreads);
    cull_count_emit_kernel<0><<<g, kThreads, 0, s>>>(P, (const float2 *)uv, depth, radius, conic, opacity, cull, W, H, gx,
                                                    gy, counts, masks, nullptr, (long long)I_cap, nullptr, nullptr, nullptr, nullptr, nullptr);
    int rc = spv::check_launch("spv_bin_capacity/count");
    if (rc) return rc;
    SPV_CUDA_TRY(cub::DeviceScan::InclusiveSum((void *)w, scan, counts, offsets, P, s), "spv_bin_capacity/scan");
    // unused slots keep an all-ones key: they sort behind every real (tile, depth) key
    SPV_CUDA_TRY(cudaMemsetAsync(keys_in, 0xFF, 8 * (size_t)I_cap, s), "spv_bin_capacity");
    cull_count_emit_kernel<1><<<g, kThreads, 0, s>>>(P, (const float2 *)uv, depth, radius, conic, opacity, cull, W, H, gx, gy,
                                                    nullptr, masks, offsets, (long long)I_cap, keys_in, vals_in, status, nullptr, nullptr);
    rc = spv::check_launch("spv_bin_capacity/emit", 3);
    if (rc) return rc;
    int tb = 1;
    while ((1 << tb) < gx * gy + 1) ++tb;   // room for the all-ones sentinel above the largest tile id
    SPV_CUDA_TRY(cub::DeviceRadixSort::SortPairs((void *)w, temp, (const unsigned long long *)keys_in, keys_out,
                                                 (const int *)vals_in, idx_sorted, (int)I_cap, 0, 32 + tb, s),
                 "spv_bin_capacity/sort");
    tile_range_dev_kernel<<<spv::cdiv(I_cap, kThreads), kThreads, 0, s>>>((long long)I_cap, status, keys_out, (int2 *)tile_range);
    return spv::check_launch("spv_bin_capacity/range", 1 + 2 + (32 + tb + 7) / 8);
}

// ---- capacity-bounded binning by tile segments + per-tile shared-memory sort (fused frame path) ---------------------
size_t spv_bin_tiles_workspace_bytes(int P, int64_t I_cap, int W, int H) {
    if (I_cap <= 0) I_cap = 1;
    const size_t T = (size_t)spv::tiles_x(W > 0 ? W : 1) * spv::tiles_y(H > 0 ? H : 1);
    return align_up(8 * (size_t)I_cap) + align_up(8 * (size_t)(P > 0 ? P : 1)) + align_up(4 * T) + align_up(4 * (T + 2));
}

/* Same contract as spv_bin_capacity (idx_sorted / tile_range / status are bit-identical when nothing overflows). */
int spv_bin_tiles(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                  const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status,
                  void *workspace, size_t ws_bytes, void *stream) {
    return spv::bin_tiles_ordered(P, I_cap, uv, depth, radius, conic, opacity, cull, W, H, idx_sorted, tile_range, status, nullptr,
                                  /*cleared=*/false, workspace, ws_bytes, stream);
}

}  // extern "C"

namespace {
struct TileBinWs { unsigned long long *keys, *masks; int *tile_count, *big_queue; };
TileBinWs carve_tile_bin(void *workspace, int P, int64_t I_cap, int T) {
    char *w = (char *)workspace;
    TileBinWs b;
    b.keys = (unsigned long long *)w; w += align_up(8 * (size_t)I_cap);
    b.masks = (unsigned long long *)w; w += align_up(8 * (size_t)(P > 0 ? P : 1));
    b.tile_count = (int *)w; w += align_up(4 * (size_t)T);
    b.big_queue = (int *)w;
    return b;
}
// tile_range[2T] = 0, status[2] = 0, tile_count[T] = 0, queue count / head = 0: ONE launch instead of three memset nodes (in a
// captured graph every memset node costs ~5 us of dependency latency in front of the count pass)
__global__ void __launch_bounds__(kThreads)
bin_clear_kernel(int T, int *__restrict__ tile_range, int *__restrict__ status, int *__restrict__ tile_count, int *__restrict__ queue) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < 2 * T) tile_range[i] = 0;
    if (i < T) tile_count[i] = 0;
    if (i < 2) { status[i] = 0; queue[i] = 0; }
}
}  // namespace

int spv::bin_tiles_clear(int P, int64_t I_cap, int W, int H, int *tile_range, int *status, void *workspace, void *stream) {
    const int T = spv::tiles_x(W) * spv::tiles_y(H);
    if (T <= 0) return 0;
    if (I_cap <= 0) I_cap = 1;
    TileBinWs b = carve_tile_bin(workspace, P, I_cap, T);
    bin_clear_kernel<<<spv::cdiv(2 * (long long)T, kThreads), kThreads, 0, (cudaStream_t)stream>>>(T, tile_range, status, b.tile_count, b.big_queue);
    return spv::check_launch("spv_bin_tiles/clear");
}

/* spv_bin_tiles + tile_order (int[T], or NULL): the tiles listed longest list first -- the blend kernels' launch order.
 * cleared: the caller already ran bin_tiles_clear on this workspace / tile_range / status (the frame path does, ahead of its chain). */
int spv::bin_tiles_ordered(int P, int64_t I_cap, const float *uv, const float *depth, const int *radius, const float *conic,
                           const float *opacity, int cull, int W, int H, int *idx_sorted, int *tile_range, int *status,
                           int *tile_order, bool cleared, void *workspace, size_t ws_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = spv::tiles_x(W), gy = spv::tiles_y(H), T = gx * gy;
    if (P <= 0 || I_cap <= 0 || T <= 0) {   // nothing to bin: empty ranges, no workspace needed
        if (!cleared && T > 0) SPV_CUDA_TRY(cudaMemsetAsync(tile_range, 0, sizeof(int) * 2 * (size_t)T, s), "spv_bin_tiles");
        if (!cleared) SPV_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * 2, s), "spv_bin_tiles");
        return 0;
    }
    if (I_cap >= (1ll << 31)) { spv::set_error(cudaErrorInvalidValue, "spv_bin_tiles: capacity too large"); return (int)cudaErrorInvalidValue; }
    if (ws_bytes < spv_bin_tiles_workspace_bytes(P, I_cap, W, H)) { spv::set_error(cudaErrorInvalidValue, "spv_bin_tiles: workspace too small"); return (int)cudaErrorInvalidValue; }
    if (!cleared) { const int rc0 = spv::bin_tiles_clear(P, I_cap, W, H, tile_range, status, workspace, stream); if (rc0) return rc0; }
    TileBinWs b = carve_tile_bin(workspace, P, I_cap, T);
    unsigned long long *keys = b.keys, *masks = b.masks;
    int *tile_count = b.tile_count, *big_queue = b.big_queue;
    const unsigned g = spv::cdiv((long long)P * kLanesPerSplat, kThreads);
    cull_count_emit_kernel<2><<<g, kThreads, 0, s>>>(P, (const float2 *)uv, depth, radius, conic, opacity, cull, W, H, gx, gy,
                                                    nullptr, masks, nullptr, (long long)I_cap, nullptr, nullptr, nullptr, tile_count,
                                                    nullptr);
    tile_scan_kernel<<<1, kScanThreads, 0, s>>>(T, (long long)I_cap, tile_count, (int2 *)tile_range, status, big_queue, tile_order);
    cull_count_emit_kernel<3><<<g, kThreads, 0, s>>>(P, (const float2 *)uv, depth, radius, conic, opacity, cull, W, H, gx, gy,
                                                    nullptr, masks, nullptr, (long long)I_cap, keys, nullptr, nullptr, tile_count,
                                                    (const int2 *)tile_range);
    int rc = spv::check_launch("spv_bin_tiles/emit", 3);
    if (rc) return rc;
    tile_sort_small_kernel<<<T, kThreads, 0, s>>>((const int2 *)tile_range, keys, idx_sorted);
    static std::atomic<unsigned long long> configured{0};
    spv::opt_in_dynamic_smem(tile_sort_big_kernel, kSortBig * 8, configured);
    tile_sort_big_kernel<<<spv::sm_count(), kScanThreads, kSortBig * 8, s>>>((const int2 *)tile_range, keys, idx_sorted, big_queue);
    return spv::check_launch("spv_bin_tiles/sort", 2);
}
