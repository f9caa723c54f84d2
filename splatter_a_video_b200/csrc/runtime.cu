// ABI version + error reporting shared by every extern "C" entry point.
#include "common.cuh"
#include "../../include/spv_b200.h"

namespace spv {
static thread_local char g_err[512] = "";
void set_error(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof g_err, "%s: %s (%d)", where, cudaGetErrorString(e), (int)e);
}
}  // namespace spv

extern "C" {
int spv_abi_version(void) { return SPV_ABI_VERSION; }
const char *spv_last_error(void) { return spv::g_err; }
}
