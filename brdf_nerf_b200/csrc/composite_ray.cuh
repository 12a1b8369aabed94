// Per-ray volume compositing (one warp per ray): shared by composite.cu (bn_composite_sigma / bn_composite_forward) and
// coarse_to_fine.cu (the fused stratified-pass -> guided samples -> merge kernel).  See composite.cu for the reference map.
#pragma once
#include "common.cuh"

namespace bn {

struct CompositeFwd {
  const float* z;        // (N,S)
  const float* packed;   // (N,S,C)  channel `sigma_ch` is the density
  const float* noise;    // (N,S) or null
  const float* irr;      // (N,S) or null : per-sample irradiance scalar (sun visibility)
  float noise_std;
  float *alpha, *trans, *weights;          // (N,S) each, alpha/trans nullable
  float *depth, *wsum, *std;               // (N) each, wsum/std nullable
  float *acc;                              // (N,C)  Σ w·x  (entry sigma_ch holds Σ w·sigma, unused)
  float *acc_irr;                          // (N,4)  Σ w·irr·[x0,x1,x2,1]  (only when irr != null)
  int N, S, sigma_ch;
  // optional: `packed` holds the rows in the MLP's generation order (all stratified samples [N][S1], then all guided
  // samples [N][S - S1]) and sort_idx (N,S) maps depth order to it — the gather bn_permute_samples would do, folded in
  const long long* sort_idx; int S1;
};

// row of `packed` that holds sample i (depth order) of ray r
__device__ __forceinline__ long long packed_row(const long long* sort_idx, int N, int S, int S1, int r, int i) {
  if (sort_idx == nullptr) return (long long)r * S + i;
  const long long j = sort_idx[(long long)r * S + i];
  return j < S1 ? (long long)r * S1 + j : (long long)N * S1 + (long long)r * (S - S1) + (j - S1);
}

template <int C>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&x)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int v = 0; v < C / 4; ++v) {
      float4 q = __ldg(reinterpret_cast<const float4*>(p) + v);
      x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = __ldg(p + c);
  }
}
template <int C>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&x)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int v = 0; v < C / 4; ++v)
      reinterpret_cast<float4*>(p)[v] = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] = x[c];
  }
}

// C == 1 is the sigma-only pass (packed == sigma, nothing but depth accumulated).
// One ray per warp; also the first stage of the fused pass-1 -> pass-2 kernel (coarse_to_fine.cu).
template <int C>
__device__ __forceinline__ void composite_ray(const CompositeFwd& a, int r, int lane) {
  const int S = a.S;
  const long long base = (long long)r * S;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  float acc_i[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int kSig = (C == 1) ? 0 : 3;      // density channel of the packed row (spsbrdfnerf.py:694)
  float depth = 0.f, wsum = 0.f;
  float carry = 1.0f;                         // transmittance in front of the current 32-sample row
  // software pipeline: the loads of 32-sample row k+1 are in flight while row k goes through exp / scan / accumulate
  // (one row per iteration left a warp with ~640 B outstanding: 61 % of the HBM roofline at 65 536 rays)
  float zn = 0.f, nzn = 0.f, irn = 0.f;
  float xn[C];
#pragma unroll
  for (int c = 0; c < C; ++c) xn[c] = 0.f;
  auto fetch = [&](int i) {
    if (i < S) {
      zn = __ldg(a.z + base + i);
      load_row<C>(a.packed + packed_row(a.sort_idx, a.N, S, a.S1, r, i) * C, xn);
      if (a.noise) nzn = __ldg(a.noise + base + i);
      if constexpr (C > 1) { if (a.irr) irn = __ldg(a.irr + base + i); }
    } else {
      zn = 0.f; nzn = 0.f; irn = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) xn[c] = 0.f;
    }
  };
  fetch(lane);
  for (int i0 = 0; i0 < S; i0 += kWarp) {
    const int i = i0 + lane;
    const bool ok = i < S;
    const float zi = zn, nz = nzn, ir = irn;
    float x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = xn[c];
    fetch(i + kWarp);                                          // next row (zeros past the end)
    float znext = __shfl_down_sync(kFull, zi, 1);
    const float zfirst_next = __shfl_sync(kFull, zn, 0);       // z of the next row's first sample, for lane 31
    if (lane == kWarp - 1) znext = zfirst_next;
    float sg = x[kSig];
    if (a.noise) sg += nz * a.noise_std;
    const float delta = (i + 1 < S) ? (znext - zi) : 1e10f;
    // accurate expf: alpha feeds the guided sampler through the weights
    const float al = ok ? 1.0f - expf(-delta * fmaxf(sg, 0.f)) : 0.f;
    const float f = 1.0f - al + 1e-10f;
    float incl = warp_scan_mul(ok ? f : 1.0f, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    const float w = al * T;
    carry *= __shfl_sync(kFull, incl, kWarp - 1);
    if (ok) {
      if (a.alpha) a.alpha[base + i] = al;
      if (a.trans) a.trans[base + i] = T;
      a.weights[base + i] = w;
      depth += w * zi; wsum += w;
      if constexpr (C > 1) {
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += w * x[c];
        if (a.irr) {
          float wi = w * ir;
          acc_i[0] += wi * x[0]; acc_i[1] += wi * x[1]; acc_i[2] += wi * x[2]; acc_i[3] += wi;
        }
      }
    }
  }
  depth = warp_sum(depth); wsum = warp_sum(wsum);
  if (a.std) {
    // Σ w (z-d)² evaluated in the numerically safe two-pass form
    __syncwarp();
    float s2 = 0.f;
    for (int i = lane; i < S; i += kWarp) { float dz = a.z[base + i] - depth; s2 += dz * dz * a.weights[base + i]; }
    s2 = warp_sum(s2);
    if (lane == 0) a.std[r] = sqrtf(s2);
  }
  if constexpr (C > 1) {
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = warp_sum(acc[c]);
    if (a.irr) { for (int c = 0; c < 4; ++c) acc_i[c] = warp_sum(acc_i[c]); }
  }
  if (lane == 0) {
    a.depth[r] = depth;
    if (a.wsum) a.wsum[r] = wsum;
    if constexpr (C > 1) {
#pragma unroll
      for (int c = 0; c < C; ++c) a.acc[(long long)r * C + c] = acc[c];
      if (a.irr && a.acc_irr) { for (int c = 0; c < 4; ++c) a.acc_irr[r * 4 + c] = acc_i[c]; }
    }
  }
}

}  // namespace bn
