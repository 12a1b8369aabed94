// Shared helpers for the sm_100a kernels of the BRDF-NeRF render path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/brdfnerf_b200.h"

namespace bn {

// ---- error plumbing (thread-local message returned by bn_last_error) ----
void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);

#define BN_CHECK_ARG(cond, msg)                                                     \
  do { if (!(cond)) { bn::set_error("%s: %s", __func__, msg); return BN_ERR_ARG; } } while (0)
#define BN_CUDA(call)                                                               \
  do { int _rc = bn::check_cuda((call), #call); if (_rc) return _rc; } while (0)
// every kernel launch of the library goes through after_launch(): error check + launch counter
int after_launch(const char* kernel);
#define BN_LAUNCH_CHECK()                                                            \
  do { int _rc = bn::after_launch(__func__); if (_rc) return _rc; } while (0)

// optional per-launch CUDA-event timing of the GEMM family (bench.py roofline leg)
void prof_begin(int kind, double work, cudaStream_t s);
void prof_end(cudaStream_t s);
bool prof_enabled();

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// inclusive warp scans
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(kFull, v, o); if (lane >= o) v *= t; }
  return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(kFull, v, o); if (lane >= o) v += t; }
  return v;
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// element type helpers shared by the fp32 (SIMT) and bf16 (tcgen05) MLP paths
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace bn
