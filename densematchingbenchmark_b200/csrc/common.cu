#include "common.cuh"

namespace dmb {

std::atomic<long long> g_launches{0};

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    // cached per DEVICE (the calling thread's current device): a process may drive several GPUs
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace dmb

extern "C" int dmb_b200_abi_version(void) { return 1; }
extern "C" const char* dmb_b200_last_error(void) { return dmb::err_buf(); }
extern "C" int64_t dmb_b200_launch_count(void) { return (int64_t)dmb::g_launches.load(); }
