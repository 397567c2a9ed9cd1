// ABI bookkeeping: version, thread-local error message, launch counter, cached device attributes.
#include "common.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

namespace lr {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

char* last_error_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

cudaError_t ensure_max_dynamic_smem(const void* func, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> done;          // (device, kernel) -> configured limit
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, func);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done[key] = bytes;
    return e;
}

int sm_count() {
    // Immutable after first use; one value per process is enough (one process drives one GPU).
    static int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        return v;
    }();
    return n;
}

}  // namespace lr

extern "C" int lr_version(void) { return LR_ABI_VERSION; }
extern "C" const char* lr_last_error(void) { return lr::last_error_buf(); }
extern "C" unsigned long long lr_launch_count(void) { return lr::g_launches.load(std::memory_order_relaxed); }

// Zero-fill through the copy engine (a memset node under graph capture, not a kernel launch).
extern "C" int lr_memset(void* ptr, size_t bytes, lr_stream_t stream) {
    if (bytes == 0) return LR_OK;
    LR_CHECK_ARG(ptr, "lr_memset: null pointer");
    cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, stream);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_memset: %s", cudaGetErrorString(e));
    return LR_OK;
}
