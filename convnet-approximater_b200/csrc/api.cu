// Library-level C ABI: error string, device check, launch counter.
#include "common.cuh"

#include <atomic>
#include <mutex>

namespace cna {
static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};
void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int dev_info(DevInfo* out) {
    constexpr int kMaxDev = 64;
    static DevInfo cache[kMaxDev];
    static bool have[kMaxDev] = {};
    static std::mutex mu;
    int dev = 0;
    CNA_CUDA(cudaGetDevice(&dev));
    CNA_CHECK(dev >= 0 && dev < kMaxDev, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        DevInfo d{dev, 0, 0};
        CNA_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
        CNA_CUDA(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cache[dev] = d;
        have[dev] = true;
    }
    *out = cache[dev];
    return 0;
}
}  // namespace cna

extern "C" {

int cna_version(void) { return 100; }

const char* cna_last_error(void) { return cna::g_err.c_str(); }

unsigned long long cna_launch_count(void) { return cna::g_launches.load(); }

int cna_device_check(void) {
    int dev = 0;
    CNA_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CNA_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        CNA_FAIL(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                 std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
    }
    return 0;
}

}  // extern "C"
