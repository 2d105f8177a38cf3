// Error state, launch counter, device info.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace ipavsr {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace ipavsr

extern "C" {
const char* ipavsr_last_error(void) { return ipavsr::g_err; }
int ipavsr_version(void) { return 200; }
#ifndef IPAVSR_SRC_HASH
#define IPAVSR_SRC_HASH "unknown"
#endif
const char* ipavsr_source_hash(void) { return IPAVSR_SRC_HASH; }
uint64_t ipavsr_launch_count(void) { return ipavsr::g_launches.load(); }
void ipavsr_launch_count_add(uint64_t n) { ipavsr::count_launch(n); }

int ipavsr_device_info(int* sm, int* major, int* minor, int* max_smem_optin) {
  int dev = 0;
  IPAVSR_CUDA(cudaGetDevice(&dev));
  if (sm) IPAVSR_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  if (major) IPAVSR_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  if (minor) IPAVSR_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (max_smem_optin) IPAVSR_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return IPAVSR_OK;
}
}
