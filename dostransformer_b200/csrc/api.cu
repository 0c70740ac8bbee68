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

// One word per flag in mapped, portable host memory: kernels raise a flag with a plain store (flags are only ever
// set on the device), the host polls without synchronising.
static unsigned int* g_errw_host = nullptr;
static unsigned int* g_errw_dev = nullptr;
unsigned int* device_error_words() { return g_errw_dev; }
static int init_error_words() {
  if (g_errw_host) return DOST_OK;
  void* h = nullptr;
  if (cudaHostAlloc(&h, sizeof(unsigned int) * kDevErrWords, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    set_error("device_errors_init: cudaHostAlloc failed");
    return DOST_ERR_LAUNCH;
  }
  for (int i = 0; i < kDevErrWords; ++i) static_cast<volatile unsigned int*>(h)[i] = 0u;
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeHost(h);
    set_error("device_errors_init: cudaHostGetDevicePointer failed");
    return DOST_ERR_LAUNCH;
  }
  g_errw_host = static_cast<unsigned int*>(h);
  g_errw_dev = static_cast<unsigned int*>(d);
  return DOST_OK;
}
}  // namespace dost

extern "C" {
int dost_abi_version(void) { return DOST_ABI_VERSION; }
const char* dost_last_error(void) { return dost::g_err; }
long long dost_launch_count(void) { return dost::g_launches.load(); }
void dost_reset_launch_count(void) { dost::g_launches.store(0); }
int dost_device_errors_init(void) { return dost::init_error_words(); }
unsigned int dost_device_errors(int clear) {
  if (!dost::g_errw_host) return 0u;
  unsigned int mask = 0u;
  volatile unsigned int* w = dost::g_errw_host;
  for (int i = 0; i < dost::kDevErrWords; ++i)
    if (w[i]) {
      mask |= 1u << i;
      if (clear) w[i] = 0u;
    }
  return mask;
}
}
