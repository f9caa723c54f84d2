// ABI version + error reporting shared by every extern "C" entry point.
#include "common.cuh"
#include "../../include/spv_b200.h"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace spv {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", where, cudaGetErrorString(e), (int)e);
}

// Optional CUDA-event bracket around the blend kernels (slot 0: forward, 1: backward) so a caller can time the dominant
// kernel INSIDE a real step (also inside a captured graph: external event-record nodes).  Off by default: zero cost.
static std::atomic<int> g_timer_on{0};
static cudaEvent_t g_timer_ev[2][2];
static bool g_timer_made = false;
void timer_mark(int slot, int edge, cudaStream_t s) {
    if (!g_timer_on.load(std::memory_order_relaxed) || slot < 0 || slot > 1) return;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    cudaEventRecordWithFlags(g_timer_ev[slot][edge], s,
                             st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}

constexpr int kMaxDevices = 64;
static thread_local SideLane g_side[kMaxDevices];
SideLane *side_lane() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    SideLane &l = g_side[dev];
    if (!l.ok) {
        if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&l.hi, cudaStreamNonBlocking, greatest) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.hi_join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.mid, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        l.ok = true;
    }
    return &l;
}

// Small runtime switches for kernel variants (default 0).  First read falls back to the environment variable SPV_<NAME>
// (upper case) so a run can be switched without code changes.
//   bwd_variant   reserved for A/B runs of backward blend kernel variants (none selectable at the moment)
//   flat_chain    1: the frame path's geometry + binning chain runs on the caller's stream instead of the high-priority lane
//   depth_staged  1: spv_loss_depth_dpt as six kernels (select passes, statistics, residuals, gradient) instead of the fused one
struct Option { const char *name; const char *env; std::atomic<int> value; };
static Option g_options[] = {{"bwd_variant", "SPV_BWD_VARIANT", {-1}}, {"flat_chain", "SPV_FLAT_CHAIN", {-1}},
                             {"depth_staged", "SPV_DEPTH_STAGED", {-1}}};
int get_option(const char *name) {
    for (Option &o : g_options) {
        if (strcmp(name, o.name) != 0) continue;
        int v = o.value.load(std::memory_order_relaxed);
        if (v < 0) {
            const char *e = getenv(o.env);
            v = e ? atoi(e) : 0;
            if (v < 0) v = 0;
            o.value.store(v, std::memory_order_relaxed);
        }
        return v;
    }
    return 0;
}
int set_option(const char *name, int value) {
    for (Option &o : g_options)
        if (strcmp(name, o.name) == 0) { o.value.store(value < 0 ? 0 : value, std::memory_order_relaxed); return 0; }
    return 1;
}
}  // namespace spv

extern "C" {
int spv_set_option(const char *name, int value) {
    if (!name || spv::set_option(name, value)) { spv::set_error(cudaErrorInvalidValue, "spv_set_option: unknown option"); return (int)cudaErrorInvalidValue; }
    return 0;
}
int spv_kernel_timer_enable(int on) {
    if (on && !spv::g_timer_made) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                cudaError_t e = cudaEventCreate(&spv::g_timer_ev[i][j]);
                if (e != cudaSuccess) { spv::set_error(e, "spv_kernel_timer_enable"); return (int)e; }
            }
        spv::g_timer_made = true;
    }
    spv::g_timer_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return 0;
}
int spv_kernel_timer_read(int slot, float *ms) {
    if (!spv::g_timer_made || slot < 0 || slot > 1 || !ms) { spv::set_error(cudaErrorInvalidValue, "spv_kernel_timer_read: timer not enabled"); return (int)cudaErrorInvalidValue; }
    cudaError_t e = cudaEventSynchronize(spv::g_timer_ev[slot][1]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, spv::g_timer_ev[slot][0], spv::g_timer_ev[slot][1]);
    if (e != cudaSuccess) { spv::set_error(e, "spv_kernel_timer_read"); return (int)e; }
    return 0;
}
int spv_abi_version(void) { return SPV_ABI_VERSION; }
const char *spv_last_error(void) { return spv::g_err; }
long long spv_launch_count(void) { return spv::g_launches.load(std::memory_order_relaxed); }
}
