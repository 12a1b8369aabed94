// Fused GEMM epilogues of the SIREN MLP, shared by the SIMT (fp32) and tcgen05 (bf16) mainloops.
// Each functor receives one output row fragment: `n` consecutive columns [col0, col0+n) of row `row`
// as fp32 accumulators, and writes the fused result straight to global memory.
//
// Reference ops fused here (paths relative to /root/reference):
//   Siren.forward  sin(w0 * Linear(x))        models/nerf.py:23-33, models/spsbrdfnerf.py:636-646
//   feats_from_xyz (bias only)                models/spsbrdfnerf.py:688
//   autograd of sin / Linear (dgrad, wgrad)   implicit in the reference
#pragma once
#include "common.cuh"

namespace bn {

template <typename T, int n> struct Pack;
template <int n> struct Pack<float, n> {
  static_assert(n % 4 == 0, "");
  __device__ static void store(float* p, const float (&v)[n]) {
#pragma unroll
    for (int i = 0; i < n / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  __device__ static void load(const float* p, float (&v)[n]) {
#pragma unroll
    for (int i = 0; i < n / 4; ++i) {
      float4 q = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  }
};
template <int n> struct Pack<__nv_bfloat16, n> {
  static_assert(n % 8 == 0, "");
  __device__ static void store(__nv_bfloat16* p, const float (&v)[n]) {
#pragma unroll
    for (int i = 0; i < n / 8; ++i) {
      uint4 q;
      __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
      __nv_bfloat162 d = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      q.x = *reinterpret_cast<uint32_t*>(&a); q.y = *reinterpret_cast<uint32_t*>(&b);
      q.z = *reinterpret_cast<uint32_t*>(&c); q.w = *reinterpret_cast<uint32_t*>(&d);
      reinterpret_cast<uint4*>(p)[i] = q;
    }
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[n]) {
#pragma unroll
    for (int i = 0; i < n / 8; ++i) {
      uint4 q = reinterpret_cast<const uint4*>(p)[i];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
        v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  }
};

// H = sin(w0 (acc + b));  C = w0 cos(w0 (acc + b))  (C only when training).
// kFast selects the MUFU sin/cos (bf16 path, hidden layers: |arg| is O(1), error << bf16 ulp).
template <typename T, bool kFast>
struct EpiSin {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  const float* bias; float w0;
  T* H; long long ldh;
  T* Cc; long long ldc;       // nullable
  int M, N;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= M || col0 >= N) return;
    float h[n], c[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float a = w0 * (acc[j] + __ldg(bias + col0 + j));
      if (kFast) { h[j] = __sinf(a); c[j] = w0 * __cosf(a); }
      else { float s, co; sincosf(a, &s, &co); h[j] = s; c[j] = w0 * co; }
    }
    Pack<T, n>::store(H + (long long)row * ldh + col0, h);
    if (Cc) Pack<T, n>::store(Cc + (long long)row * ldc + col0, c);
  }
};

// out = acc + b
template <typename T>
struct EpiBias {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  const float* bias; T* out; long long ld; int M, N;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= M || col0 >= N) return;
    float o[n];
#pragma unroll
    for (int j = 0; j < n; ++j) o[j] = acc[j] + __ldg(bias + col0 + j);
    Pack<T, n>::store(out + (long long)row * ld + col0, o);
  }
};

// dgrad: out = (acc [+ addend]) [* mulc] [+ add2]   (dZ_{l-1} = (dZ_l W_l [+ direct grads]) ⊙ C_{l-1}
// [+ second-order term]); raw_out optionally receives (acc + addend) before the mask.
// colsum (tcgen05 path only, where the 32 lanes of a warp hold 32 consecutive rows of the same
// columns): the bias gradient sum_rows(out) is reduced across the warp with a 31-shuffle butterfly
// and added with one coalesced red.global per 32 columns — no separate pass over dZ.
template <typename T>
struct EpiDgrad {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  const T* addend; long long lda;    // nullable
  const T* mulc; long long ldm;      // nullable
  T* out; long long ld; int M, N;
  T* raw_out = nullptr; long long ldr = 0;
  const T* add2 = nullptr; long long ld2 = 0;
  float* colsum = nullptr;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (col0 >= N) return;                       // warp-uniform
    const bool valid = row < M;
    float o[n];
#pragma unroll
    for (int j = 0; j < n; ++j) o[j] = valid ? acc[j] : 0.f;
    if (valid) {
      if (addend) {
        float t[n]; Pack<T, n>::load(addend + (long long)row * lda + col0, t);
#pragma unroll
        for (int j = 0; j < n; ++j) o[j] += t[j];
      }
      if (raw_out) Pack<T, n>::store(raw_out + (long long)row * ldr + col0, o);
      if (mulc) {
        float t[n]; Pack<T, n>::load(mulc + (long long)row * ldm + col0, t);
#pragma unroll
        for (int j = 0; j < n; ++j) o[j] *= t[j];
      }
      if (add2) {
        float t[n]; Pack<T, n>::load(add2 + (long long)row * ld2 + col0, t);
#pragma unroll
        for (int j = 0; j < n; ++j) o[j] += t[j];
      }
      Pack<T, n>::store(out + (long long)row * ld + col0, o);
    }
    if constexpr (n == 32) {
      if (colsum) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool hi = (lane & s) != 0;
#pragma unroll
          for (int j = 0; j < s; ++j) {
            const float send = hi ? o[j] : o[j + s];
            const float keep = hi ? o[j + s] : o[j];
            o[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        atomicAdd(colsum + col0 + lane, o[0]);     // lane L now holds the sum of column col0 + L
      }
    }
  }
};

// skinny weight gradients through the NT GEMM: output row o (< n_rows) is the gradient of one
// head's second-layer weight row; only columns [c0, c1) of it are real and go to dst[col - c0].
struct EpiSkinnyRow { float* dst; int c0, c1; };
struct EpiSkinny {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  int n_rows; EpiSkinnyRow r[24];
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= n_rows) return;
    const EpiSkinnyRow q = r[row];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const int col = col0 + j;
      if (col >= q.c0 && col < q.c1) atomicAdd(q.dst + (col - q.c0), acc[j]);
    }
  }
};

// second-order sweep of the analytic normals (adjoint of a_{l-1} = (a_l W_l) ⊙ c_{l-1}):
//   acc = abar_l ;  ubar_l = abar_l ⊙ c_l  -> ubar ;  zb_l = (abar_l ⊙ u_l) * (-w0^2 h_l) -> overwrites u_l
template <typename T>
struct EpiSecond {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  const T* Cc; long long ldc;
  T* U; long long ldu;               // in: u_l, out: zb_l
  const T* H; long long ldh;
  T* ubar; long long ldo;
  float neg_w0sq; int M, N;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= M || col0 >= N) return;
    float c[n], u[n], h[n], o[n];
    Pack<T, n>::load(Cc + (long long)row * ldc + col0, c);
    Pack<T, n>::load(U + (long long)row * ldu + col0, u);
    Pack<T, n>::load(H + (long long)row * ldh + col0, h);
#pragma unroll
    for (int j = 0; j < n; ++j) { o[j] = acc[j] * c[j]; u[j] = acc[j] * u[j] * neg_w0sq * h[j]; }
    Pack<T, n>::store(ubar + (long long)row * ldo + col0, o);
    Pack<T, n>::store(U + (long long)row * ldu + col0, u);
  }
};

// wgrad: dW[row][map(col)] += acc, fp32 atomics into the flat gradient bucket.  Columns in
// [pad_lo, pad_hi) are K-padding of the packed operand and are dropped; columns >= pad_hi shift
// down by (pad_hi - pad_lo) — this maps the padded [enc(60)|pad(4)|h(512)] layout back to the
// reference's Linear(572, 512) weight.
struct EpiWgrad {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  float* dW; long long ld; int M, N; int pad_lo, pad_hi;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= M) return;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const int col = col0 + j;
      if (col >= N || (col >= pad_lo && col < pad_hi)) continue;
      const int dst = col >= pad_hi ? col - (pad_hi - pad_lo) : col;
      atomicAdd(dW + (long long)row * ld + dst, acc[j]);
    }
  }
};

// accumulator read back and dropped: isolates the mainloop in bench_gemm.py
struct EpiNull {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;
  float* sink;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (acc[0] == 1.2345e38f && acc[n - 1] == -1.2345e38f) sink[0] = acc[1] + row + col0;      // never true: keeps the loads alive
  }
};

// plain fp32 store (debug / unit-test entry point)
struct EpiStoreF32 {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;   // direct global-memory epilogue (tc::EPI_DIRECT)
  float* out; long long ld; int M, N;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (row >= M) return;
#pragma unroll
    for (int j = 0; j < n; ++j) if (col0 + j < N) out[(long long)row * ld + col0 + j] = acc[j];
  }
};

}  // namespace bn
