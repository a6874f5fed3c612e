// c_abi.cu -- library-level entry points of libu2pl_b200.so (version, error text, launch counter).
#include <atomic>
#include <cstring>
#include "common.cuh"

namespace u2pl {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void set_error(const char *msg) { strncpy(g_err, msg, sizeof(g_err) - 1); g_err[sizeof(g_err) - 1] = 0; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace u2pl

extern "C" int u2pl_abi_version(void) { return U2PL_ABI_VERSION; }
extern "C" const char *u2pl_last_error(void) { return u2pl::g_err; }
extern "C" int64_t u2pl_launch_count(void) { return u2pl::g_launches.load(std::memory_order_relaxed); }
