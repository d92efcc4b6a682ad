// lib.cu — error string, launch counter, version (shared state of the library)
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
    cudaMemPool_t pool;
    AB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = ~0ull;
    AB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    if (dev < 64) g_inited[dev] = true;
    return 0;
}
}  // namespace ab200
extern "C" const char *b200_version(void) { return "algebra_b200 0.1.0 (sm_100a)"; }
extern "C" const char *b200_last_error(void) { return ab200::t_last_error.c_str(); }
extern "C" unsigned long long b200_launch_count(void) { return ab200::g_launches.load(); }
