// Per-ray device code of the sample generator (one warp per ray): shared by sampler.cu (bn_sample_guided / bn_merge_samples)
// and coarse_to_fine.cu (the fused kernel).  Every arithmetic operation is an explicit round-to-nearest intrinsic, so the results
// do not depend on the translation unit's -fmad setting.  See sampler.cu for the reference map and the arithmetic contract.
#pragma once
#include "common.cuh"

namespace bn {

constexpr int kMaxBins = 256;       // n_samples / guided_samples upper bound for the guided kernel
constexpr int kWarpsPerBlock = 4;

// Sum of x[0..n) in ATen's CPU order; x in shared memory, result replicated on all lanes.
__device__ inline float aten_row_sum_warp(const float* x, int n, int lane) {
  const int V = 8, ILP = 4;
  const int nvec = n / V, groups = nvec / ILP;
  const int k = lane >> 3, j = lane & 7;
  float p = 0.0f;
  for (int i = 0; i < groups; ++i) p = __fadd_rn(p, x[(i * ILP + k) * V + j]);
  if (k == 0)
    for (int v = groups * ILP; v < nvec; ++v) p = __fadd_rn(p, x[v * V + j]);
  float p1 = __shfl_sync(kFull, p, j + 8), p2 = __shfl_sync(kFull, p, j + 16), p3 = __shfl_sync(kFull, p, j + 24);
  if (k == 0) p = __fadd_rn(__fadd_rn(__fadd_rn(p, p1), p2), p3);
  float acc = 0.0f;
  for (int s = nvec * V; s < n; ++s) acc = __fadd_rn(acc, x[s]);
  for (int jj = 0; jj < V; ++jj) acc = __fadd_rn(acc, __shfl_sync(kFull, p, jj));
  return acc;
}

// ascending bitonic sort of keys in shared memory, m = power of two, one warp
__device__ inline void bitonic_sort_warp(float* key, int m, int lane) {
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += kWarp) {
        int l = i ^ j;
        if (l > i) {
          float a = key[i], b = key[l];
          bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[l] = a; }
        }
      }
      __syncwarp();
    }
  }
}

struct GuidedArgs {
  const float* z1; const float* depth; const float* weights;
  const float* t_vals; const float* gauss_w; const float* u_pred;
  const float* near0; const float* far0;
  const long long* valid_depth; const float* gt_depth; int gt_depth_stride; const float* gt_std;
  const float* u_gt;
  float* z2; float* std_out;
  float d_range; int N, S1, G;
};

// guided samples of ray r by one warp; A (q / weights / pdf), E (bin edges), C (cdf), Sm (samples, padded to a power of two):
// kMaxBins floats of shared memory each
__device__ __forceinline__ void guided_ray(const GuidedArgs& a, int r, int lane, float* A, float* E, float* C, float* Sm) {
  const int S1 = a.S1, n = a.G;
  const float eps = 1e-5f;
  const float k = a.d_range;

  // 1. sampling std of pass 1 (always evaluated: it is also an output)
  const float d_pred = a.depth[r];
  for (int i = lane; i < S1; i += kWarp) {
    float dz = __fsub_rn(a.z1[(long long)r * S1 + i], d_pred);
    A[i] = __fmul_rn(__fmul_rn(dz, dz), a.weights[(long long)r * S1 + i]);
  }
  __syncwarp();
  const float std_pred = __fsqrt_rn(aten_row_sum_warp(A, S1, lane));
  if (a.std_out && lane == 0) a.std_out[r] = std_pred;
  __syncwarp();

  // 2. window centre / half width: predicted depth, or the ground-truth depth for supervised rays
  float c = d_pred, sd = std_pred;
  const float* u = a.u_pred + (long long)r * n;
  if (a.valid_depth && a.valid_depth[r] > 0) {
    c = a.gt_depth[(long long)r * a.gt_depth_stride];
    sd = a.gt_std[r];
    u = a.u_gt + (long long)r * n;
  }
  const float near0 = *a.near0, far0 = *a.far0;
  float lo = __fsub_rn(c, __fmul_rn(k, sd)), hi = __fadd_rn(c, __fmul_rn(k, sd));
  // 3. clamp to the chunk's [near, far], then make the window symmetric about the centre
  lo = fminf(fmaxf(lo, near0), far0);
  hi = fminf(fmaxf(hi, near0), far0);
  const float half = fminf(fabsf(__fsub_rn(hi, c)), fabsf(__fsub_rn(lo, c)));
  lo = __fsub_rn(c, half);
  hi = __fadd_rn(c, half);
  // 4-5. bin edges
  const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(n - 1));
  for (int j = lane; j < n; j += kWarp) {
    float tj = a.t_vals[j];
    E[j] = __fadd_rn(__fmul_rn(lo, __fsub_rn(1.0f, tj)), __fmul_rn(hi, tj));
  }
  __syncwarp();
  // 6. Gaussian bin weights (+eps)
  const float inv_den = __fadd_rn(step, eps);
  for (int j = lane; j < n - 1; j += kWarp) {
    float factor = __fdiv_rn(__fsub_rn(E[j + 1], E[j]), inv_den);
    A[j] = __fadd_rn(__fmul_rn(factor, a.gauss_w[j]), eps);
  }
  __syncwarp();
  // 7. pdf
  const float tot = aten_row_sum_warp(A, n - 1, lane);
  __syncwarp();
  for (int j = lane; j < n - 1; j += kWarp) A[j] = __fdiv_rn(A[j], tot);
  __syncwarp();
  // 8. cdf: sequential fp64 accumulation, each prefix rounded to fp32
  if (lane == 0) {
    double acc = 0.0;
    C[0] = 0.0f;
    for (int j = 0; j < n - 1; ++j) { acc += (double)A[j]; C[j + 1] = (float)acc; }
  }
  __syncwarp();
  // 9. inverse-CDF lookup
  int m = 1; while (m < n) m <<= 1;
  for (int s = lane; s < m; s += kWarp) {
    float v = __int_as_float(0x7f800000);
    if (s < n) {
      float us = u[s];
      int lo_i = 0, hi_i = n;                    // ATen's cus_upper_bound, step for step (CDF may be non-monotone)
      while (lo_i < hi_i) { int mid = (lo_i + hi_i) >> 1; if (!(C[mid] > us)) lo_i = mid + 1; else hi_i = mid; }
      int below = max(lo_i - 1, 0), above = min(lo_i, n - 1);
      float c0 = C[below], c1 = C[above], b0 = E[below], b1 = E[above];
      float den = __fsub_rn(c1, c0);
      if (den < eps) den = 1.0f;
      v = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(us, c0), den), __fsub_rn(b1, b0)));
    }
    Sm[s] = v;
  }
  __syncwarp();
  // 10. ascending sort
  bitonic_sort_warp(Sm, m, lane);
  for (int s = lane; s < n; s += kWarp) a.z2[(long long)r * n + s] = Sm[s];
}

// Sort the concatenation [z1 | z2] per ray carrying the source index (stable on ties).
// (no __restrict__: in the fused kernel z1 / z2 were written by this very warp)
__device__ __forceinline__ void merge_ray(const float* z1, const float* z2, float* z_out, long long* idx_out, float* unsort_out,
                                          int S1, int G, int r, int lane, float* K, int* I) {
  const int S = S1 + G;
  int m = 1; while (m < S) m <<= 1;
  for (int i = lane; i < m; i += kWarp) {
    float v = __int_as_float(0x7f800000);
    if (i < S1) v = z1[(long long)r * S1 + i];
    else if (i < S) v = z2[(long long)r * G + (i - S1)];
    K[i] = v; I[i] = i;
    if (i < S && unsort_out) unsort_out[(long long)r * S + i] = v;
  }
  __syncwarp();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += kWarp) {
        int l = i ^ j;
        if (l > i) {
          float a = K[i], b = K[l]; int ia = I[i], ib = I[l];
          bool gt = (a > b) || (a == b && ia > ib);
          bool up = (i & k) == 0;
          if (gt == up) { K[i] = b; K[l] = a; I[i] = ib; I[l] = ia; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < S; i += kWarp) {
    z_out[(long long)r * S + i] = K[i];
    if (idx_out) idx_out[(long long)r * S + i] = (long long)I[i];
  }
}

}  // namespace bn
