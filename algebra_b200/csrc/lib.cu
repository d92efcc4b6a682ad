// lib.cu — error string, launch counter, version (shared state of the library)
#include <mutex>
#include "common.cuh"
namespace ab200 {
static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};
void set_last_error(const std::string &msg) { t_last_error = msg; }
}  // namespace ab200
extern "C" const char *b200_version(void) { return "algebra_b200 0.1.0 (sm_100a)"; }
extern "C" const char *b200_last_error(void) { return ab200::t_last_error.c_str(); }
extern "C" unsigned long long b200_launch_count(void) { return ab200::g_launches.load(); }
