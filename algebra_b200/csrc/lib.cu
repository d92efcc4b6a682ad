// lib.cu — error string, launch counter, version (shared state of the library)
#include <cstdlib>
#include <mutex>
#include <cstdint>
#include "common.cuh"
namespace ab200 {
static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};
void set_last_error(const std::string &msg) { t_last_error = msg; }
static std::mutex g_init_mutex;
static bool g_inited[64] = {false};
int ensure_device_init() {
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_init_mutex);
    if (dev < 64 && g_inited[dev]) return 0;
    // Freed stream-ordered scratch stays cached up to a BOUNDED amount (default 96 GiB, B200_POOL_RETAIN_GB read once): enough for
    // the scratch of a 2^26 MSM to be reused by the next call without re-allocation, while anything above it goes back to the
    // driver at the next synchronisation instead of starving other allocators of the process (e.g. PyTorch's).
    static const uint64_t retain = [] {
        const char *e = getenv("B200_POOL_RETAIN_GB");
        const double gb = e ? atof(e) : 96.0;
        return (uint64_t)((gb < 0 ? 0 : gb) * (double)(1ull << 30));
    }();
    cudaMemPool_t pool;
    AB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = retain;
    AB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    if (dev < 64) g_inited[dev] = true;
    return 0;
}
}  // namespace ab200
extern "C" const char *b200_version(void) { return "algebra_b200 0.1.0 (sm_100a)"; }
extern "C" const char *b200_last_error(void) { return ab200::t_last_error.c_str(); }
extern "C" unsigned long long b200_launch_count(void) { return ab200::g_launches.load(); }
