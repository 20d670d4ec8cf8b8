// ABI bookkeeping: version, status names, last CUDA error (thread-local, diagnostics only).
#include <atomic>
#include <cstdlib>
#include "common.cuh"

namespace b200w {
constexpr int kLaunchLog = 64;
static thread_local int g_last_cuda_error = 0;
int set_last_cuda_error(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return B200W_ERR_LAUNCH;
}
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200W_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
static std::atomic<unsigned long long> g_launches{0};
static const char* g_log[kLaunchLog] = {nullptr};
void note_launch(const char* kernel) {
    const unsigned long long n = g_launches.fetch_add(1, std::memory_order_relaxed);
    g_log[n % kLaunchLog] = kernel;
}
}  // namespace b200w

extern "C" unsigned long long b200w_kernel_launches(void) { return b200w::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* b200w_kernel_log(int back) {
    const unsigned long long n = b200w::g_launches.load(std::memory_order_relaxed);
    if (back < 0 || back >= b200w::kLaunchLog || (unsigned long long)back >= n) return "";
    const char* s = b200w::g_log[(n - 1 - (unsigned long long)back) % b200w::kLaunchLog];
    return s ? s : "";
}

extern "C" int b200w_abi_version(void) { return B200W_ABI_VERSION; }

extern "C" int b200w_last_cuda_error(void) { return b200w::g_last_cuda_error; }

extern "C" const char* b200w_status_string(int status) {
    switch (status) {
        case B200W_OK: return "ok";
        case B200W_ERR_BAD_MODE: return "Unkown pad type";
        case B200W_ERR_BAD_TAPS: return "bad filter taps (null, empty or more than B200W_MAX_TAPS)";
        case B200W_ERR_NULL_POINTER: return "null pointer";
        case B200W_ERR_BAD_SHAPE: return "bad shape";
        case B200W_ERR_REFLECT_PAD: return "reflect padding must be smaller than the padded dimension";
        case B200W_ERR_PER_TOO_SHORT: return "periodization needs a signal at least as long as the filter";
        case B200W_ERR_LAUNCH: return "CUDA launch failed";
        case B200W_ERR_WORKSPACE: return "workspace missing or too small";
        case B200W_ERR_BAD_WINDOW: return "SSIM window size must be odd and at most 11";
        default: return "unknown status";
    }
}
