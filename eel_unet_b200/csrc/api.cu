// Error plumbing and small bookkeeping entry points of the C ABI.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace eel {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return EEL_ERR_CUDA;
    }
    return EEL_OK;
}

}  // namespace eel

extern "C" {
const char* eel_last_error(void) { return eel::g_err; }
int eel_version(void) { return 200; }
int eel_num_sms(void) { return eel::kNumSMs; }
long long eel_launch_count(void) { return eel::g_launches.load(); }
}
