// CUDA-core GEMMs with fp32 accumulation.  They are the strict-fp32 mode of the MLP (parity against
// the fp32 oracle within 1e-3) and the on-device checker of the tcgen05 path; the same epilogue
// functors (epilogues.cuh) are used by both mainloops.
//
//   gemm_tn_simt : C[m][n] = sum_k A[m][k] * B[n][k]      (forward Linear, dgrad with W^T packed)
//   gemm_nt_simt : C[n][k] = sum_p A[p][n] * B[p][k]      (wgrad, reduction over points, split-P)
#pragma once
#include "common.cuh"

namespace bn {

constexpr int kSimtBM = 128, kSimtBN = 128, kSimtBK = 16, kSimtPad = 4;

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(w[j] << 16); v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
}

template <typename T, class Epi>
__global__ void __launch_bounds__(256) gemm_tn_simt(const T* __restrict__ A, long long lda,
                                                    const T* __restrict__ B, long long ldb,
                                                    int M, int N, int K, Epi epi) {
  __shared__ float As[kSimtBK][kSimtBM + kSimtPad];
  __shared__ float Bs[kSimtBK][kSimtBN + kSimtPad];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * kSimtBM, n0 = blockIdx.x * kSimtBN;
  const int lrow = tid / 2, lk = (tid % 2) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kSimtBK) {
    float va[8], vb[8];
    if (m0 + lrow < M) load8<T>(A + (long long)(m0 + lrow) * lda + k0 + lk, va);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] = 0.f;
    }
    if (n0 + lrow < N) load8<T>(B + (long long)(n0 + lrow) * ldb + k0 + lk, vb);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { As[lk + j][lrow] = va[j]; Bs[lk + j][lrow] = vb[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) epi.template apply<8>(m0 + ty * 8 + i, n0 + tx * 8, acc[i]);
}

// A: [P, Nn] (row p contiguous in n), B: [P, Kk].  grid = (ceil(Kk/128), ceil(Nn/128), splits)
template <typename T, class Epi>
__global__ void __launch_bounds__(256) gemm_nt_simt(const T* __restrict__ A, long long lda,
                                                    const T* __restrict__ B, long long ldb,
                                                    int Nn, int Kk, long long P, long long p_per_split, Epi epi) {
  __shared__ float As[kSimtBK][kSimtBM + kSimtPad];
  __shared__ float Bs[kSimtBK][kSimtBN + kSimtPad];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int n0 = blockIdx.y * kSimtBM, c0 = blockIdx.x * kSimtBN;
  const long long p_begin = (long long)blockIdx.z * p_per_split;
  const long long p_end = min(P, p_begin + p_per_split);
  const int lp = tid / 16, lo = (tid % 16) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (long long p0 = p_begin; p0 < p_end; p0 += kSimtBK) {
    float va[8], vb[8];
    const long long p = p0 + lp;
    if (p < p_end && n0 + lo < Nn) load8<T>(A + p * lda + n0 + lo, va);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) va[j] = 0.f;
    }
    if (p < p_end && c0 + lo < Kk) load8<T>(B + p * ldb + c0 + lo, vb);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) vb[j] = 0.f;
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lp][lo]) = make_float4(va[0], va[1], va[2], va[3]);
    *reinterpret_cast<float4*>(&As[lp][lo + 4]) = make_float4(va[4], va[5], va[6], va[7]);
    *reinterpret_cast<float4*>(&Bs[lp][lo]) = make_float4(vb[0], vb[1], vb[2], vb[3]);
    *reinterpret_cast<float4*>(&Bs[lp][lo + 4]) = make_float4(vb[4], vb[5], vb[6], vb[7]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) epi.template apply<8>(n0 + ty * 8 + i, c0 + tx * 8, acc[i]);
}

template <typename T, class Epi>
int launch_tn_simt(const T* A, long long lda, const T* B, long long ldb, int M, int N, int K, const Epi& epi, cudaStream_t s) {
  if (K % kSimtBK != 0 || (lda % 8) || (ldb % 8)) { set_error("gemm_tn_simt: K %% 16 / ld %% 8 alignment (K=%d)", K); return BN_ERR_ARG; }
  dim3 grid(ceil_div(N, kSimtBN), ceil_div(M, kSimtBM));
  gemm_tn_simt<T, Epi><<<grid, 256, 0, s>>>(A, lda, B, ldb, M, N, K, epi);
  return after_launch("gemm_tn_simt");
}

template <typename T, class Epi>
int launch_nt_simt(const T* A, long long lda, const T* B, long long ldb, int Nn, int Kk, long long P, const Epi& epi, cudaStream_t s) {
  if ((Nn % 8) || (Kk % 8) || (lda % 8) || (ldb % 8)) { set_error("gemm_nt_simt: dims must be multiples of 8"); return BN_ERR_ARG; }
  const int tiles = ceil_div(Kk, kSimtBN) * ceil_div(Nn, kSimtBM);
  int splits = (int)max(1LL, min((long long)ceil_div(148 * 4, tiles), ceil_div_ll(P, 256)));
  long long per = ceil_div_ll(ceil_div_ll(P, splits), kSimtBK) * kSimtBK;
  splits = (int)ceil_div_ll(P, per);
  dim3 grid(ceil_div(Kk, kSimtBN), ceil_div(Nn, kSimtBM), splits);
  gemm_nt_simt<T, Epi><<<grid, 256, 0, s>>>(A, lda, B, ldb, Nn, Kk, P, per, epi);
  return after_launch("gemm_nt_simt");
}

}  // namespace bn
