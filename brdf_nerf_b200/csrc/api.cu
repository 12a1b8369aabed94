// C-ABI plumbing: version, error reporting, device check.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <vector>

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
static unsigned long long g_launches = 0;
int after_launch(const char* kernel) {
  ++g_launches;
  return check_cuda(cudaGetLastError(), kernel);
}

struct ProfRec { cudaEvent_t a, b; int kind; double work; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
bool prof_enabled() { return g_prof_on; }
void prof_begin(int kind, double work, cudaStream_t s) {
  if (!g_prof_on) return;
  ProfRec r; r.kind = kind; r.work = work;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t s) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, s);
}
}  // namespace bn

extern "C" __attribute__((visibility("default"))) unsigned long long bn_launch_count(void) { return bn::g_launches; }

extern "C" __attribute__((visibility("default"))) int bn_profile_enable(int on) {
  for (auto& r : bn::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  bn::g_prof.clear();
  bn::g_prof_on = on != 0;
  return BN_OK;
}

// Sums the recorded launches per kind (kind < n_kinds): count, milliseconds, work (flops or bytes).
extern "C" __attribute__((visibility("default")))
int bn_profile_collect(int n_kinds, long long* count, double* ms, double* work) {
  BN_CHECK_ARG(count && ms && work && n_kinds > 0, "bad arguments");
  for (int k = 0; k < n_kinds; ++k) { count[k] = 0; ms[k] = 0; work[k] = 0; }
  BN_CUDA(cudaDeviceSynchronize());
  for (auto& r : bn::g_prof) {
    if (r.kind < 0 || r.kind >= n_kinds) continue;
    float t = 0.f;
    BN_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    count[r.kind] += 1; ms[r.kind] += t; work[r.kind] += r.work;
  }
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_abi_version(void) { return BN_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* bn_last_error(void) { return bn::g_err; }

extern "C" __attribute__((visibility("default"))) int bn_device_check(int device) {
  int major = 0, minor = 0;
  BN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  BN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    bn::set_error("device %d is sm_%d%d; this library only runs on sm_100-class (B200) GPUs", device, major, minor);
    return BN_ERR_DEVICE;
  }
  return BN_OK;
}
