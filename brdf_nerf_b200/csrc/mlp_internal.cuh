#pragma once
// Internal declarations shared by mlp.cu and normals.cu.
// K-B: positional encoding + SIREN MLP (trunk, sigma / feature / colour / BRDF heads), forward and
// backward, as a chain of fused GEMMs.
//
// Replaces (reference, paths relative to /root/reference):
//   xyz = o + d z                        rendering.py:184,216,254,273
//   Mapping.forward                      models/nerf.py:53-70
//   SpSBRDFNeRF.calc_features            models/spsbrdfnerf.py:636-646 (layers :513-524)
//   sigma / feats / rgb / BRDF heads     models/spsbrdfnerf.py:527-535,582-613,682-755
//   autograd backward of all of it       implicit (dgrad + wgrad)
//
// Two precision modes share this orchestration and the epilogue functors:
//   BN_PREC_BF16 : tcgen05.mma (gemm_tc.cuh) — bf16 activations/weights in HBM, fp32 accumulation in
//                  TMEM, sin / cos / bias / Hadamard fused in the epilogue warps;
//   BN_PREC_FP32 : CUDA-core fp32 (gemm_simt.cuh) — parity mode and on-device checker.
// Activation layout in the caller's workspace (row = point, row-major, element type T):
//   X3 [P, 64+F]  : cols 0..63 = encoding (60 real + 4 zero pad), cols 64.. = h_{skip-1}; the skip
//                   layer reads the whole row as its K = 64+F operand (no concat copy)
//   H_l, C_l [P,F]: h_l = sin(w0 z_l) and c_l = w0 cos(w0 z_l) (kept only when training)
//   FE [P,F (+64)], HD / CD [P, 256*blocks]: features (+ the encoded view direction when --input_viewdir: the colour
//                   head reads the whole row, like the skip layer reads X3) and the heads' hidden layer (+ cosine)
#include <vector>
#include <type_traits>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include "epilogues.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "epilogues_tc.cuh"

namespace bn {

constexpr int kEncPad = 64;
constexpr int kMaxBlocks = 9;     // rgb + beta + up to 7 BRDF heads
constexpr int kMaxOut = 16;       // scalar outputs of the heads' second layers
constexpr float kPiF = 3.14159265358979323846f;

enum { XF_SIGMOID = 0, XF_K = 1, XF_THETA_RPV = 2, XF_THETA_H = 3, XF_SOFTPLUS = 4 };
constexpr int kTOff = 32;         // the time embedding sits kTOff columns behind the features (after the direction encoding)

struct OutDesc { int block; long long w_off; long long b_off; int ch; int rep; int xform; };
struct HeadPlan {
  int n_out; OutDesc o[kMaxOut];
  int n_blocks;                 // blocks of the hidden layer that are evaluated
  int HH;                       // hidden width of one block (feat / 2)
  int ch_sigma, ch_nlr;         // packed channel of sigma / learned normal (-1 = off)
  long long wsig, bsig, wg, bg; // offsets of sigma_from_xyz.0 / grad_from_xyz
  long long b1_off[kMaxBlocks]; // bias offset of each block's first layer ({name}.0.bias)
};

}  // namespace bn

struct bn_mlp {
  bn_mlp_cfg cfg;
  int F, L, E, HH, skip;
  int DE, ldfe;                 // view-direction encoding width (0 = off) and the pitch of FE = [features | dir enc | pad | t | pad]
  int TE;                       // width of the time embedding read by the beta head (0 = no beta head): FE cols F + kTOff ..
  int num_sms;
  bool bf16;
  size_t es;
  void* Wp[16]; void* WTp[16]; int Kpad[16]; int Kreal[16];
  void* Wf; void* WfT;
  void* W1; void* W1T; float* b1cat;
  void* W2p;                    // [n_blocks*HH, 64] bf16: second-layer head weights as the B operand of the GHD GEMM
  void* W2pT; void* Wsig;       // [64, n_blocks*HH], [64, F] bf16: the same weights / w_sigma as B operands of the forward heads GEMMs
  void* WsigA;                  // [64, F] bf16, row 0 = w_sigma: density of a trunk-only call (bn_mlp_trunk_forward)
  long long* Wb_offs;           // device: (weight, bias) offsets of the trunk layers in the flat parameter buffer
  void* Wb;                     // [L * F, 64] bf16 bias blocks of the fused trunk kernels (mlp_chain.cuh): cols 0..59 = encoding
                                // weights of layer 0 / the skip layer (zero elsewhere), col 60 = bf16(b), col 61 = bf16(b - col 60)
  int n_blocks;
  int blk_lin0[bn::kMaxBlocks], blk_lin2[bn::kMaxBlocks], blk_head[bn::kMaxBlocks];
  bool synced;
  bool no_chain;
  int w2p_flags;                // flags of the training forward that last packed W2p for the backward (-1: stale)
  bool no_dchain;               // BN_NO_DGRAD_CHAIN=1: per-layer data-gradient GEMMs instead of the fused chain (A/B timing aid)
  // backward: weight gradients run on a side stream next to the data-gradient chain (forked from / joined into the caller's stream)
  cudaStream_t s2; cudaEvent_t ev_dz[16]; cudaEvent_t ev_w[16]; cudaEvent_t ev_h[8]; bool overlap;
  long long* chain_trace;
};

namespace bn {

// out_o[i - c0_o] += sum_p D[p][o] X[p][i]   for i in [c0_o, c1_o), plus bias_o += sum_p D[p][o].
struct SkinnyRow { float* dst; float* bias; int col; int c0, c1; };
struct SkinnyPlan { int n; SkinnyRow r[kMaxOut]; };

template <typename T>
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(SkinnyPlan sp, const T* __restrict__ D, int ldd,
                                                           const T* __restrict__ X, long long ldx, int ncols,
                                                           long long P, long long rows_per_block) {
  __shared__ float sD[32][kMaxOut];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const long long p0 = (long long)blockIdx.y * rows_per_block;
  const long long p1 = min(P, p0 + rows_per_block);
  float acc[kMaxOut], bacc[kMaxOut];
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) { acc[o] = 0.f; bacc[o] = 0.f; }
  for (long long pb = p0; pb < p1; pb += 32) {
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * kMaxOut; t += 256) {
      const int pp = t / kMaxOut, o = t % kMaxOut;
      sD[pp][o] = (pb + pp < p1 && o < sp.n) ? to_f<T>(D[(pb + pp) * ldd + sp.r[o].col]) : 0.f;
    }
    __syncthreads();
    const int cnt = (int)min(32LL, p1 - pb);
    for (int pp = 0; pp < cnt; ++pp) {
      const float x = i < ncols ? to_f<T>(X[(pb + pp) * ldx + i]) : 0.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) { acc[o] = fmaf(sD[pp][o], x, acc[o]); bacc[o] += sD[pp][o]; }
    }
  }
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    if (o < sp.n) {
      if (i >= sp.r[o].c0 && i < sp.r[o].c1) atomicAdd(sp.r[o].dst + (i - sp.r[o].c0), acc[o]);
      if (sp.r[o].bias && i == sp.r[o].c0) atomicAdd(sp.r[o].bias, bacc[o]);
    }
  }
}

// dst[c] += sum_p X[p][c]   (bias gradients).  8 columns per thread, 4 rows in flight per thread so
// that enough 16-byte loads are outstanding to stream at HBM speed.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, long long ldx, int ncols, long long P,
                                                     long long rows_per_block, float* __restrict__ dst) {
  __shared__ float red[4][64 * 8];
  const int tx = threadIdx.x % 64, ty = threadIdx.x / 64;      // 64 column groups x 4 row lanes
  const int c = (blockIdx.x * 64 + tx) * 8;
  const long long p0 = (long long)blockIdx.y * rows_per_block;
  const long long p1 = min(P, p0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < ncols) {
    long long p = p0 + ty;
    for (; p + 12 < p1; p += 16) {
      float v0[8], v1[8], v2[8], v3[8];
      load8<T>(X + p * ldx + c, v0); load8<T>(X + (p + 4) * ldx + c, v1);
      load8<T>(X + (p + 8) * ldx + c, v2); load8<T>(X + (p + 12) * ldx + c, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; p < p1; p += 4) {
      float v[8]; load8<T>(X + p * ldx + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  if (ty == 0 && c < ncols) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      atomicAdd(dst + c + j, (red[0][tx * 8 + j] + red[1][tx * 8 + j]) + (red[2][tx * 8 + j] + red[3][tx * 8 + j]));
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T> struct Ws {
  T* X3; T* H[16]; long long Hld[16]; T* C[16];
  T* FE; long long ldfe; T* HD; T* CD; T* GHD; T* G7D; T* GFE; T* GA; T* GB; T* DPRE;
  T* GZ[16];       // fused data-gradient chain (tcgen05 mode): dZ_l of every trunk layer l < L-1 (dZ_{L-1} = GA)
  long long ldx3, ldhd;
  // analytic-normal sweep (BN_MLP_NORMAL_AN)
  T* A[16];        // a_l = d sigma / d lin_l            (one per layer when training, ping-pong otherwise)
  T* U[16];        // u_l = a_l before the cosine mask; overwritten by ZB_l in the backward (training)
  T* EE; T* EE0;   // [P,64] d sigma / d enc: skip-layer part, total
  float* GRAW;     // [P,4] raw d sigma / d x (fp32)
  T* UBX;          // [P,64+F]: cols 0..63 = adjoint of EE, cols 64.. = ubar_{skip-1}
  T* UBA; T* UBB;  // ubar ping-pong
  T* SG;           // [P,64]: col 0 = sigmoid(s_p), other columns zero (A operand of the w_sigma second-order wgrad)
  float* SIGC;     // [P] density of every row, written by the fused trunk kernel (tcgen05 mode, full forward)
};

static inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

// the fused data-gradient chain (mlp_dgrad_chain.cuh) is specialised like the forward chain: bf16, 512-wide trunk
// the fused forward trunk kernel (mlp_chain.cuh): bf16, 512-wide trunk with a skip connection
static inline bool train_chain_ok(const bn_mlp* h) { return h->bf16 && h->F == 512 && h->skip >= 1 && h->L <= 16 && !h->no_chain; }
// ... which also leaves the density of every row behind (BN_CHAIN_NO_SIG=1, read per call: the separate sigma GEMMs instead —
// A/B knob; a forward and the calls that consume its workspace must run under the same setting)
static inline bool chain_sigma_ok(const bn_mlp* h) { return train_chain_ok(h) && getenv("BN_CHAIN_NO_SIG") == nullptr; }
static inline bool dgrad_chain_ok(const bn_mlp* h) { return h->bf16 && h->F == 512 && h->L >= 3 && !h->no_chain && !h->no_dchain; }

template <typename T>
static inline size_t carve(const bn_mlp* h, long long P, int flags, void* base, Ws<T>* w) {
  const bool train = flags & BN_MLP_TRAIN, sig_only = flags & BN_MLP_SIGMA_ONLY;
  const bool normals = (flags & BN_MLP_NORMAL_AN) && !sig_only;
  const int F = h->F, L = h->L;
  size_t off = 0;
  auto take = [&](long long elems) -> T* {
    T* p = base ? reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(base) + off) : nullptr;
    off += align_up((size_t)elems * sizeof(T));
    return p;
  };
  Ws<T> t{};
  t.ldx3 = kEncPad + F;
  t.ldhd = (long long)h->n_blocks * h->HH;
  t.X3 = take(P * t.ldx3);
  if (train) {
    for (int l = 0; l < L; ++l) {
      if (l == h->skip - 1) { t.H[l] = t.X3 ? t.X3 + kEncPad : nullptr; t.Hld[l] = t.ldx3; }
      else { t.H[l] = take(P * F); t.Hld[l] = F; }
      t.C[l] = take(P * F);
    }
  } else {
    T* ping = take(P * F); T* pong = take(P * F);
    for (int l = 0; l < L; ++l) {
      if (l == h->skip - 1) { t.H[l] = t.X3 ? t.X3 + kEncPad : nullptr; t.Hld[l] = t.ldx3; }
      else { t.H[l] = (l & 1) ? pong : ping; t.Hld[l] = F; }
      t.C[l] = normals ? take(P * F) : nullptr;
    }
  }
  if (!sig_only) {
    t.ldfe = h->ldfe;
    t.FE = take(P * t.ldfe);
    t.HD = take(P * t.ldhd);
    t.SIGC = reinterpret_cast<float*>(take(P * (long long)(sizeof(float) / sizeof(T))));
    if (train) {
      t.CD = take(P * t.ldhd); t.GHD = take(P * t.ldhd);
      t.G7D = take(P * F); t.GFE = take(P * F); t.GA = take(P * F); t.GB = take(P * F);
      t.DPRE = take(P * 64);
      if (dgrad_chain_ok(h)) {       // one buffer per layer: the chain produces them all in one launch, the wgrads read them later
        for (int l = 0; l < L - 1; ++l) t.GZ[l] = (l == L - 2) ? t.GB : (l == L - 3 ? t.G7D : take(P * F));
      }
    }
  }
  if (normals) {
    if (train) {
      for (int l = 0; l < L; ++l) { t.A[l] = take(P * F); t.U[l] = take(P * F); }
      t.UBX = take(P * t.ldx3); t.UBA = take(P * F); t.UBB = take(P * F); t.SG = take(P * 64);
    } else {
      T* ping = take(P * F); T* pong = take(P * F);
      for (int l = 0; l < L; ++l) { t.A[l] = (l & 1) ? pong : ping; t.U[l] = nullptr; }
    }
    t.EE = take(P * kEncPad); t.EE0 = take(P * kEncPad);
    t.GRAW = reinterpret_cast<float*>(take(P * 4 * (long long)(sizeof(float) / sizeof(T))));
  }
  if (w) *w = t;
  return off;
}

// GEMM dispatch: tcgen05 for bf16, CUDA cores for fp32
template <typename T, class Epi>
static int gemm_tn(const bn_mlp* h, const T* A, long long lda, const T* B, long long ldb, long long M, int N, int K,
                   const Epi& epi, cudaStream_t s, int k_real = -1) {
  prof_begin(0, 2.0 * (double)M * N * (k_real > 0 ? k_real : K), s);
  int rc;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (N >= 256) rc = tc::launch_tn<256>(A, lda, B, ldb, M, N, K, epi, h->num_sms, s);
    else if (N >= 128) rc = tc::launch_tn<128>(A, lda, B, ldb, M, N, K, epi, h->num_sms, s);
    else rc = tc::launch_tn<64>(A, lda, B, ldb, M, N, K, epi, h->num_sms, s);
  } else {
    rc = launch_tn_simt<T, Epi>(A, lda, B, ldb, (int)M, N, K, epi, s);
  }
  prof_end(s);
  return rc;
}
template <typename T, class Epi>
static int gemm_nt(const bn_mlp* h, const T* A, long long lda, const T* B, long long ldb, int Mo, int No, long long P,
                   const Epi& epi, cudaStream_t s, double flops = -1.0) {
  prof_begin(1, flops >= 0 ? flops : 2.0 * (double)P * Mo * No, s);
  int rc;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (No >= 256) rc = tc::launch_nt<256>(A, lda, B, ldb, Mo, No, P, epi, h->num_sms, s);
    else if (No >= 128) rc = tc::launch_nt<128>(A, lda, B, ldb, Mo, No, P, epi, h->num_sms, s);
    else rc = tc::launch_nt<64>(A, lda, B, ldb, Mo, No, P, epi, h->num_sms, s);
  } else {
    rc = launch_nt_simt<T, Epi>(A, lda, B, ldb, Mo, No, P, epi, s);
  }
  prof_end(s);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Fused layers: one call = one GEMM + epilogue.  bf16 -> tcgen05 mainloop with the TMA-staged
// functors of epilogues_tc.cuh; fp32 -> CUDA-core mainloop with the direct functors of epilogues.cuh.
constexpr bool is_bf16_v(const __nv_bfloat16*) { return true; }
constexpr bool is_bf16_v(const float*) { return false; }

// H = sin(w0 (A W^T + b))  [, C = w0 cos(...)]      A:[P,K] W:[N,K]
template <typename T>
static int layer_sin(const bn_mlp* h, const T* A, long long lda, const T* W, long long ldw, long long P, int N, int K,
                     const float* bias, float w0, T* H, long long ldh, T* C, long long ldc, cudaStream_t s, int k_real = -1) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (C) {
      tc::EpiSinT<true> e; e.bias = bias; e.w0 = w0;
      if (int rc = tc::stream_map(&e.out_map[0], H, P, N, ldh)) return rc;
      if (int rc = tc::stream_map(&e.out_map[1], C, P, N, ldc)) return rc;
      return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s, k_real);
    }
    tc::EpiSinT<false> e; e.bias = bias; e.w0 = w0;
    if (int rc = tc::stream_map(&e.out_map[0], H, P, N, ldh)) return rc;
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s, k_real);
  } else {
    EpiSin<T, false> e{bias, w0, H, ldh, C, ldc, (int)P, N};
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s, k_real);
  }
}

// out = A W^T + b
template <typename T>
static int layer_bias(const bn_mlp* h, const T* A, long long lda, const T* W, long long ldw, long long P, int N, int K,
                      const float* bias, T* out, long long ldo, cudaStream_t s) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    tc::EpiBiasT e; e.bias = bias;
    if (int rc = tc::stream_map(&e.out_map[0], out, P, N, ldo)) return rc;
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s);
  } else {
    EpiBias<T> e{bias, out, ldo, (int)P, N};
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s);
  }
}

// out = ((A BT^T) [+ addend]) [* mulc] [+ add2];  raw = (A BT^T) + addend.  bias_grad (nullable, tcgen05
// mode only, used by the unit test): column sums of out from the epilogue registers; the training path
// takes its bias gradients from layer_wgrad instead.
template <typename T> struct DgradArgs {
  const T* addend = nullptr; long long lda = 0;
  const T* mulc = nullptr; long long ldm = 0;
  const T* add2 = nullptr; long long ld2 = 0;
  T* raw = nullptr; long long ldr = 0;
  float* bias_grad = nullptr;
  // rank-<=4 addend sum_k rank_rows[row][k] * rank_col[k][col] (tcgen05 mode only)
  const T* rank_rows = nullptr; long long rank_ld = 0; const float* rank_col[4] = {nullptr, nullptr, nullptr, nullptr}; int n_rank = 0;
};
template <bool kAdd, bool kMul, bool kAdd2, bool kRaw, bool kRank = false>
static int dgrad_tc(const bn_mlp* h, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* BT, long long ldb, long long P,
                    int N, int K, const DgradArgs<__nv_bfloat16>& a, __nv_bfloat16* out, long long ldo, cudaStream_t s) {
  tc::EpiDgradT<kAdd, kMul, kAdd2, kRaw, kRank> e; e.colsum = a.bias_grad;
  e.rank_rows = a.rank_rows; e.rank_ld = a.rank_ld; e.n_rank = a.n_rank; e.M = (int)P;
  for (int k = 0; k < 4; ++k) e.rank_col[k] = a.rank_col[k];
  int i = 0;
  if (kAdd) { if (int rc = tc::stream_map(&e.in_map[i++], a.addend, P, N, a.lda)) return rc; }
  if (kMul) { if (int rc = tc::stream_map(&e.in_map[i++], a.mulc, P, N, a.ldm)) return rc; }
  if (kAdd2) { if (int rc = tc::stream_map(&e.in_map[i++], a.add2, P, N, a.ld2)) return rc; }
  if (int rc = tc::stream_map(&e.out_map[0], out, P, N, ldo)) return rc;
  if (kRaw) { if (int rc = tc::stream_map(&e.out_map[kRaw ? 1 : 0], a.raw, P, N, a.ldr)) return rc; }
  return gemm_tn<__nv_bfloat16>(h, A, lda, BT, ldb, P, N, K, e, s);
}

template <typename T>
static int colsum(const T* X, long long ldx, int ncols, long long P, float* dst, cudaStream_t s);

template <typename T>
static int layer_dgrad(const bn_mlp* h, const T* A, long long lda, const T* BT, long long ldb, long long P, int N, int K,
                       const DgradArgs<T>& a, T* out, long long ldo, cudaStream_t s) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    const int code = (a.addend ? 1 : 0) | (a.mulc ? 2 : 0) | (a.add2 ? 4 : 0) | (a.raw ? 8 : 0) | (a.n_rank > 0 ? 16 : 0);
    switch (code) {
      case 18: return dgrad_tc<false, true, false, false, true>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 22: return dgrad_tc<false, true, true, false, true>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 0: return dgrad_tc<false, false, false, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 1: return dgrad_tc<true, false, false, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 2: return dgrad_tc<false, true, false, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 3: return dgrad_tc<true, true, false, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 6: return dgrad_tc<false, true, true, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 7: return dgrad_tc<true, true, true, false>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      case 10: return dgrad_tc<false, true, false, true>(h, A, lda, BT, ldb, P, N, K, a, out, ldo, s);
      default: set_error("layer_dgrad: operand combination %d has no tcgen05 epilogue", code); return BN_ERR_ARG;
    }
  } else {
    EpiDgrad<T> e{a.addend, a.lda, a.mulc, a.ldm, out, ldo, (int)P, N};
    e.raw_out = a.raw; e.ldr = a.ldr; e.add2 = a.add2; e.ld2 = a.ld2;
    return gemm_tn<T>(h, A, lda, BT, ldb, P, N, K, e, s);
  }
}

// second-order sweep layer (normals.cu): ubar = (A W^T) ⊙ c ;  U <- ((A W^T) ⊙ U) (-w0^2 H)
template <typename T>
static int layer_second(const bn_mlp* h, const T* A, long long lda, const T* W, long long ldw, long long P, int N, int K,
                        const T* Cc, long long ldc, T* U, long long ldu, const T* H, long long ldh, T* ubar, long long ldo,
                        float neg_w0sq, cudaStream_t s, int k_real) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    tc::EpiSecondT e; e.neg_w0sq = neg_w0sq;
    if (int rc = tc::stream_map(&e.in_map[0], Cc, P, N, ldc)) return rc;
    if (int rc = tc::stream_map(&e.in_map[1], U, P, N, ldu)) return rc;
    if (int rc = tc::stream_map(&e.in_map[2], H, P, N, ldh)) return rc;
    if (int rc = tc::stream_map(&e.out_map[0], ubar, P, N, ldo)) return rc;
    if (int rc = tc::stream_map(&e.out_map[1], U, P, N, ldu)) return rc;
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s, k_real);
  } else {
    EpiSecond<T> e{Cc, ldc, U, ldu, H, ldh, ubar, ldo, neg_w0sq, (int)P, N};
    return gemm_tn<T>(h, A, lda, W, ldw, P, N, K, e, s, k_real);
  }
}

// dW [Mo, Kreal] += G[:, :Mo]^T In[:, :No]   (No = packed width of In with padding [pad_lo, pad_hi));
// bias_grad (nullable) [Mo] += column sums of G — fused into the tcgen05 mainloop (an N=16 MMA against a
// tile of ones), a separate pass over G in the fp32 mode.
template <typename T>
static int layer_wgrad(const bn_mlp* h, const T* G, long long ldg, const T* In, long long ldin, int Mo, int No, long long P,
                       float* dW, long long ldw, int pad_lo, int pad_hi, float* bias_grad, cudaStream_t s, double flops = -1.0) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (tc::wgrad_tma_ok(dW, ldw, No, pad_lo, pad_hi)) {
      if (bias_grad) {
        tc::EpiWgradT<true> e;
        if (int rc = tc::make_wgrad(&e, dW, ldw, Mo, No, pad_lo, pad_hi, bias_grad)) return rc;
        return gemm_nt<T>(h, G, ldg, In, ldin, Mo, No, P, e, s, flops);
      }
      tc::EpiWgradT<false> e;
      if (int rc = tc::make_wgrad(&e, dW, ldw, Mo, No, pad_lo, pad_hi, nullptr)) return rc;
      return gemm_nt<T>(h, G, ldg, In, ldin, Mo, No, P, e, s, flops);
    }
  }
  EpiWgrad e{dW, ldw, Mo, No, pad_lo, pad_hi};
  // the CUDA-core GEMM wants widths in multiples of 8: round up (the operand's pitch covers it, the epilogue drops cols >= No)
  if (int rc = gemm_nt<T>(h, G, ldg, In, ldin, Mo, (No + 7) / 8 * 8, P, e, s, flops)) return rc;
  if (bias_grad) return colsum<T>(G, ldg, Mo, P, bias_grad, s);
  return BN_OK;
}

template <typename T>
static int colsum(const T* X, long long ldx, int ncols, long long P, float* dst, cudaStream_t s) {
  const int bx = ceil_div(ncols, 64 * 8);
  int by = (int)max(1LL, min(ceil_div_ll(P, 64), (long long)(148 * 4 / bx)));
  const long long rows = ceil_div_ll(P, by);
  by = (int)ceil_div_ll(P, rows);
  colsum_kernel<T><<<dim3(bx, by), 256, 0, s>>>(X, ldx, ncols, P, rows, dst);
  return after_launch("colsum_kernel");
}

}  // namespace bn
