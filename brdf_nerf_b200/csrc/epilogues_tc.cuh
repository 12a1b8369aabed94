// TMA-staged epilogue functors of the tcgen05 GEMM (gemm_tc.cuh, modes EPI_TMA_BF16 / EPI_TMA_RED_F32).
// Same math as the direct functors of epilogues.cuh (which the fp32 CUDA-core mode keeps using); here a
// thread owns 32 columns of one row of a 32-row x 64-column unit, element-wise operands arrive as packed bf16x2
// registers read from the warp's swizzled operand boxes, results leave as packed bf16x2 registers.
//
// Reference ops fused here (paths relative to /root/reference):
//   Siren.forward  sin(w0 * Linear(x))        models/nerf.py:23-33, models/spsbrdfnerf.py:636-646
//   feats_from_xyz (bias only)                models/spsbrdfnerf.py:688
//   autograd of sin / Linear (dgrad, wgrad)   implicit in the reference
//   double backward of calc_normals           models/spsbrdfnerf.py:648-660
#pragma once
#include "gemm_tc.cuh"

namespace bn {
namespace tc {

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// boxes of 64 bf16 columns x 32 rows over a row-major [rows, cols] matrix
inline int stream_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld) {
  return make_map_bf16(m, base, rows, cols, ld, 64, 32);
}

// H = sin(w0 (acc + b));  C = w0 cos(w0 (acc + b))  (C only when kSaveC: training / analytic normals).
// MUFU sin/cos: |arg| is O(1) in the hidden layers and <= 30 |lin| in layer 0, where the absolute
// error (~2e-6) is far below the bf16 resolution of the stored activation.
template <bool kSaveC>
struct EpiSinT {
  static constexpr int kMode = EPI_TMA_BF16, kIn = 0, kOut = kSaveC ? 2 : 1;
  CUtensorMap out_map[kOut];
  const float* bias; float w0;
  __device__ __forceinline__ void compute(int, int col0, const float (&acc)[32], const uint32_t (&)[1][16],
                                                           uint32_t (&out)[kOut][16]) const {
    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(bp + j);
      const float a0 = w0 * (acc[4 * j] + b.x), a1 = w0 * (acc[4 * j + 1] + b.y);
      const float a2 = w0 * (acc[4 * j + 2] + b.z), a3 = w0 * (acc[4 * j + 3] + b.w);
      out[0][2 * j] = bf_pack(__sinf(a0), __sinf(a1));
      out[0][2 * j + 1] = bf_pack(__sinf(a2), __sinf(a3));
      if constexpr (kSaveC) {
        out[1][2 * j] = bf_pack(w0 * __cosf(a0), w0 * __cosf(a1));
        out[1][2 * j + 1] = bf_pack(w0 * __cosf(a2), w0 * __cosf(a3));
      }
    }
  }
  __device__ __forceinline__ void load(int, void*, uint64_t*, int, int) const {}
  __device__ __forceinline__ void store(int o, const void* slot, int col0, int row0) const { tma_store_2d(&out_map[o], slot, col0, row0); }
};

// out = acc + b
struct EpiBiasT {
  static constexpr int kMode = EPI_TMA_BF16, kIn = 0, kOut = 1;
  CUtensorMap out_map[1];
  const float* bias;
  __device__ __forceinline__ void compute(int, int col0, const float (&acc)[32], const uint32_t (&)[1][16],
                                                           uint32_t (&out)[1][16]) const {
    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(bp + j);
      out[0][2 * j] = bf_pack(acc[4 * j] + b.x, acc[4 * j + 1] + b.y);
      out[0][2 * j + 1] = bf_pack(acc[4 * j + 2] + b.z, acc[4 * j + 3] + b.w);
    }
  }
  __device__ __forceinline__ void load(int, void*, uint64_t*, int, int) const {}
  __device__ __forceinline__ void store(int, const void* slot, int col0, int row0) const { tma_store_2d(&out_map[0], slot, col0, row0); }
};

// dgrad: out = (acc [+ addend] [+ rank]) [* mulc] [+ add2];  raw = acc + addend (before the mask), optional.
// Operand streams in order [addend][mulc][add2], outputs [out][raw].
// rank (kRank): sum_k r[row][k] * v_k[col], a rank-<=4 term built from per-row scalars (bf16, 4 per row) and
// per-column fp32 vectors — the direct gradients of sigma / the learned normal into h_{L-1}
// (dsigma * w_sigma + sum_k dv_k * Wg_k) without ever materialising that [P, F] matrix.
// colsum (bias gradient of the layer below = column sums of `out`): the 32 lanes hold 32 consecutive rows
// of the same 32 columns; a 31-shuffle butterfly leaves lane L with the sum of column L, added with one
// coalesced red.global.  Rows past M contribute zeros (TMA zero-fills their A rows and operands).
template <bool kAdd, bool kMul, bool kAdd2, bool kRaw, bool kRank = false>
struct EpiDgradT {
  static constexpr int kMode = EPI_TMA_BF16, kIn = (int)kAdd + (int)kMul + (int)kAdd2, kOut = 1 + (int)kRaw;
  CUtensorMap in_map[kIn > 0 ? kIn : 1];
  CUtensorMap out_map[kOut];
  float* colsum;
  const __nv_bfloat16* rank_rows; long long rank_ld;      // r[row][0..3]
  const float* rank_col[4]; int n_rank; int M;
  __device__ __forceinline__ void compute(int row, int col0, const float (&acc)[32],
                                          const uint32_t (&in)[kIn > 0 ? kIn : 1][16], uint32_t (&out)[kOut][16]) const {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = acc[j];
    int s = 0;
    if constexpr (kAdd) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { o[2 * j] += bf_lo(in[s][j]); o[2 * j + 1] += bf_hi(in[s][j]); }
      ++s;
    }
    if constexpr (kRank) {
      float rs[4] = {0.f, 0.f, 0.f, 0.f};
      if (row < M) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(rank_rows + (long long)row * rank_ld));
        rs[0] = bf_lo(r.x); rs[1] = bf_hi(r.x); rs[2] = bf_lo(r.y); rs[3] = bf_hi(r.y);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < n_rank) {
          const float4* wp = reinterpret_cast<const float4*>(rank_col[k] + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = __ldg(wp + j);
            o[4 * j] = fmaf(rs[k], w.x, o[4 * j]); o[4 * j + 1] = fmaf(rs[k], w.y, o[4 * j + 1]);
            o[4 * j + 2] = fmaf(rs[k], w.z, o[4 * j + 2]); o[4 * j + 3] = fmaf(rs[k], w.w, o[4 * j + 3]);
          }
        }
      }
    }
    if constexpr (kRaw) {
#pragma unroll
      for (int j = 0; j < 16; ++j) out[1][j] = bf_pack(o[2 * j], o[2 * j + 1]);
    }
    if constexpr (kMul) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { o[2 * j] *= bf_lo(in[s][j]); o[2 * j + 1] *= bf_hi(in[s][j]); }
      ++s;
    }
    if constexpr (kAdd2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { o[2 * j] += bf_lo(in[s][j]); o[2 * j + 1] += bf_hi(in[s][j]); }
      ++s;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) out[0][j] = bf_pack(o[2 * j], o[2 * j + 1]);
    if (colsum) {
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int st = 16; st >= 1; st >>= 1) {
        const bool hi = (lane & st) != 0;
#pragma unroll
        for (int j = 0; j < st; ++j) {
          const float send = hi ? o[j] : o[j + st];
          const float keep = hi ? o[j + st] : o[j];
          o[j] = keep + __shfl_xor_sync(0xffffffffu, send, st);
        }
      }
      atomicAdd(colsum + col0 + lane, o[0]);
    }
  }
  __device__ __forceinline__ void load(int i, void* slot, uint64_t* bar, int col0, int row0) const { tma_load_2d(slot, &in_map[i], bar, col0, row0); }
  __device__ __forceinline__ void store(int o, const void* slot, int col0, int row0) const { tma_store_2d(&out_map[o], slot, col0, row0); }
};

// second-order sweep of the analytic normals (adjoint of a_{l-1} = (a_l W_l) ⊙ c_{l-1}):
//   acc = abar_l ;  ubar_l = abar_l ⊙ c_l  -> out 0 ;  zb_l = (abar_l ⊙ u_l) * (-w0^2 h_l) -> out 1 (over u_l)
// operand streams [c_l][u_l][h_l]
struct EpiSecondT {
  static constexpr int kMode = EPI_TMA_BF16, kIn = 3, kOut = 2;
  CUtensorMap in_map[3];
  CUtensorMap out_map[2];
  float neg_w0sq;
  __device__ __forceinline__ void compute(int, int, const float (&acc)[32], const uint32_t (&in)[3][16],
                                          uint32_t (&out)[2][16]) const {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t c = in[0][j], u = in[1][j], h = in[2][j];
      const float a0 = acc[2 * j], a1 = acc[2 * j + 1];
      out[0][j] = bf_pack(a0 * bf_lo(c), a1 * bf_hi(c));
      out[1][j] = bf_pack(a0 * bf_lo(u) * neg_w0sq * bf_lo(h), a1 * bf_hi(u) * neg_w0sq * bf_hi(h));
    }
  }
  __device__ __forceinline__ void load(int i, void* slot, uint64_t* bar, int col0, int row0) const { tma_load_2d(slot, &in_map[i], bar, col0, row0); }
  __device__ __forceinline__ void store(int o, const void* slot, int col0, int row0) const { tma_store_2d(&out_map[o], slot, col0, row0); }
};

// wgrad: dW[row][map(col)] += acc as TMA reduce-adds of 32x32 fp32 boxes into the flat gradient bucket.
// Columns of the packed operand in [pad_lo, split) are K-padding and are clipped by map_lo's extent;
// columns >= split land at dW[row][col - (split - pad_lo)] through map_hi — this maps the padded
// [enc(60)|pad(4)|h(512)] layout back to the reference's Linear(572, 512) weight.
template <bool kWithBias>
struct EpiWgradT {
  static constexpr int kMode = EPI_TMA_RED_F32, kIn = 0, kOut = 1;
  static constexpr bool kBias = kWithBias;   // bias gradient sum_p G[p][row] from an extra N=16 MMA against ones
  CUtensorMap map_lo, map_hi;
  int split;                              // first packed column served by map_hi
  float* bias; int M;                     // nullable: bias[row] += sum_p G[p][row], row < M
  __device__ __forceinline__ void compute(int, int, const float (&acc)[32], uint32_t (&out)[32]) const {
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = __float_as_uint(acc[j]);
  }
  __device__ __forceinline__ void load(int, void*, uint64_t*, int, int) const {}
  __device__ __forceinline__ void store(int, const void* slot, int col0, int row0) const {
    if (col0 < split) tma_reduce_add_2d(&map_lo, slot, col0, row0);
    else tma_reduce_add_2d(&map_hi, slot, col0 - split, row0);
  }
};

// dW [M rows, Kreal cols, pitch ld] from a packed operand of N columns with padding [pad_lo, pad_hi)
template <bool kB> inline int make_wgrad(EpiWgradT<kB>* e, float* dW, long long ld, int M, int N, int pad_lo, int pad_hi, float* bias) {
  e->split = pad_hi; e->bias = bias; e->M = M;
  if (int rc = make_map_f32(&e->map_lo, dW, M, pad_lo, ld, 32, 32)) return rc;
  if (N > pad_hi) return make_map_f32(&e->map_hi, dW + pad_lo, M, N - pad_hi, ld, 32, 32);
  e->map_hi = e->map_lo;
  return BN_OK;
}
inline bool wgrad_tma_ok(const float* dW, long long ld, int N, int pad_lo, int pad_hi) {
  return tma_ok(dW, ld, 4) && (N <= pad_hi || tma_ok(dW + pad_lo, ld, 4)) && pad_hi % 32 == 0;
}

}  // namespace tc
}  // namespace bn
