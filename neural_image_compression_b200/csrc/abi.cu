// Error plumbing + device check for the C ABI (include/nic.h).
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace nic {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NIC_OK;
  return fail(NIC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

static std::atomic<uint64_t> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace nic

extern "C" {

int nic_version(void) { return NIC_ABI_VERSION; }

const char* nic_last_error(void) { return nic::g_err; }

int nic_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return nic::fail(NIC_E_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0)
    return nic::fail(NIC_E_UNSUPPORTED_ARCH,
                     "libnic_b200 is built for sm_100a only; device %d is sm_%d%d (no fallback)", dev, major, minor);
  return NIC_OK;
}

uint64_t nic_launch_count(void) { return nic::launch_count(); }

int32_t nic_partials_per_image(void) { return nic::kPartials; }

}  // extern "C"
