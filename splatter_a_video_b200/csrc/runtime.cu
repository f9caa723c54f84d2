// ABI version + error reporting shared by every extern "C" entry point.
#include "common.cuh"
#include "../../include/spv_b200.h"

#include <atomic>

namespace spv {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", where, cudaGetErrorString(e), (int)e);
}
}  // namespace spv

extern "C" {
int spv_abi_version(void) { return SPV_ABI_VERSION; }
const char *spv_last_error(void) { return spv::g_err; }
long long spv_launch_count(void) { return spv::g_launches.load(std::memory_order_relaxed); }
}
