// C-ABI plumbing: version, error reporting, device check.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace bn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BN_OK;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return BN_ERR_CUDA;
}
}  // namespace bn

extern "C" __attribute__((visibility("default"))) int bn_abi_version(void) { return BN_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* bn_last_error(void) { return bn::g_err; }

extern "C" __attribute__((visibility("default"))) int bn_device_check(int device) {
  cudaDeviceProp prop;
  BN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    bn::set_error("device %d is sm_%d%d; this library only runs on sm_100-class (B200) GPUs",
                  device, prop.major, prop.minor);
    return BN_ERR_DEVICE;
  }
  return BN_OK;
}
