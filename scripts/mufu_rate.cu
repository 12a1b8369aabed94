// Microbenchmark: MUFU throughput (sin / cos / ex2 / rsq) and the FMA-pipe alternative (polynomial cosine) on this chip, in lane
// results per clock per SM.  The fused trunk's epilogue evaluates sin and cos of every pre-activation; whether that is bound by the
// MUFU pipe decides what a cheaper epilogue has to look like.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mufu_rate.cu -o scripts/_build/mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float poly_cos(float x) {
  // cos(x) = (-1)^n cos(pi f2), n = rint(x / pi), f2 = x / pi - n in [-0.5, 0.5]; even polynomial of degree 8 in f2
  const float r = fmaf(x, 0.318309886f, 12582912.0f);          // magic number rounding: the integer n in the low mantissa bits
  const float nf = r - 12582912.0f;
  const float f2 = fmaf(x, 0.318309886f, -nf);
  const float u = f2 * f2;
  float p = fmaf(u, 0.2353306f, -1.3352627f);
  p = fmaf(p, u, 4.0587121f);
  p = fmaf(p, u, -4.9348022f);
  p = fmaf(p, u, 1.0f);
  return __uint_as_float(__float_as_uint(p) ^ (__float_as_uint(r) << 31));
}

template <int OP>
__global__ void __launch_bounds__(512, 1) mufu_kernel(float* out, int iters, long long* cyc) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.001f * threadIdx.x + 0.37f * j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) x[j] = __sinf(x[j]);
      else if (OP == 1) x[j] = __cosf(x[j]);
      else if (OP == 2) x[j] = exp2f(x[j]) * 0.25f;                      // ex2.approx
      else if (OP == 3) x[j] = rsqrtf(x[j] + 1.5f);
      else if (OP == 4) x[j] = poly_cos(x[j]);
      else if (OP == 5) { float s = __sinf(x[j]); float c = __cosf(x[j]); x[j] = s + c; }
      else if (OP == 6) { float s = __sinf(x[j]); float c = poly_cos(x[j]); x[j] = s + c; }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  mufu_kernel<OP><<<148, 512>>>(out, iters, cyc);
  mufu_kernel<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double results = 512.0 * 8 * iters * ops_per_iter;              // per SM
  printf("%-34s %7.2f results / clk / SM   (%lld cycles)\n", name, results / (double)h, h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("__sinf (FMUL + MUFU.SIN)", 1);
  run<1>("__cosf (FMUL + MUFU.COS)", 1);
  run<2>("exp2f (MUFU.EX2)", 1);
  run<3>("rsqrtf (MUFU.RSQ)", 1);
  run<4>("polynomial cosine (FMA pipe)", 1);
  run<5>("sin + cos, both MUFU", 2);
  run<6>("sin MUFU + cos polynomial", 2);
  // accuracy of the polynomial on [-100, 100]
  return 0;
}
