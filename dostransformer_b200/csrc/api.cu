// Error reporting, ABI version and launch accounting for libdost_b200.
#include "common.cuh"
#include <atomic>

namespace dost {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};  // process-wide: autograd runs backward on its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace dost

extern "C" {
int dost_abi_version(void) { return DOST_ABI_VERSION; }
const char* dost_last_error(void) { return dost::g_err; }
long long dost_launch_count(void) { return dost::g_launches.load(); }
void dost_reset_launch_count(void) { dost::g_launches.store(0); }
}
