// Fused PE + SIREN trunk + sigma head: ONE persistent kernel runs all layers of the density pass
// (reference: SpSBRDFNeRF.forward(sigma_only=True), models/spsbrdfnerf.py:677-685 = Mapping
// nerf.py:53-70 + calc_features :636-646 + softplus(sigma_from_xyz) :682) for a block of 256 points per
// CTA pair.  Activations never leave the SM: they live in shared memory as the 128B-swizzled K-major A
// operand of the next layer, accumulators in tensor memory; only the weights stream (TMA, L2 resident,
// 16 KB per CTA per K block) and 4 bytes per point are written.
//
//   warp 0     weight producer: one [128 x 64] tile of W_l per (layer, column half, K block), 4-stage ring;
//              the leader CTA's barrier collects the bytes of both CTAs (tcgen05 pair protocol of gemm_tc.cuh)
//   warp 1     MMA issuer (leader CTA): tcgen05.mma.cta_group::2, M = 256 (128 rows per CTA), N = 256;
//              a layer = two column halves, each accumulating over the layer's K blocks into its own 256 TMEM
//              columns.  K block kb of the input is consumed as soon as the epilogue of the previous layer
//              has published it (act_ready[kb]), so the next layer's first half starts while the previous
//              layer's second half is still in the epilogue.
//   warp 2     TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld -> bias + sin (MUFU) -> bf16 -> swizzled st.shared into the K block of
//              the NEXT layer's input (in place: the write waits until every MMA of the current layer has
//              retired, i.e. for the second half's commit); last layer: dot with w_sigma, softplus.
//              Before layer 0 they evaluate x = o + d z and the positional encoding into K block 0.
// Shared memory: 9 K blocks x 16 KB activations ([PE | h], so the skip layer reads K = 576 without a concat)
// + 5 x 16 KB weight ring (4 in the training variant, which also stages the cosines).
//
// Biases ride on the tensor core: K block 0 (the encoding, 60 real columns) carries the constant 1 in its pad columns 60
// and 61, and every layer's weight sequence starts with a [512 x 64] "bias block" Wb_l whose columns 60 / 61 hold
// bf16(b) and bf16(b - bf16(b)) (the bias to ~2^-17 relative), columns 0..59 the encoding weights of the layers that read
// the encoding (layer 0, skip layer) and zeros otherwise.  Layers that do not read the encoding issue ONE extra
// 256 x 256 x 16 MMA per half (the K slice of columns 48..63) — 3 % more tensor work — and the epilogue no longer
// fetches, broadcasts (32 shuffles per unit) or adds biases: the accumulator IS the pre-activation.
#pragma once
#include "gemm_tc.cuh"
#include "epilogues_tc.cuh"

namespace bn {
namespace chain {

using namespace tc;

constexpr int kF = 512;                 // trunk width this kernel is specialised for
constexpr int kNKB = 1 + kF / 64;       // K blocks of the activation buffer: PE + 8 x 64 features
constexpr int kKBBytes = 128 * 128;     // one K block of one CTA: 128 rows x 64 bf16
constexpr int kMaxLayers = 16;
constexpr int kOneCol = 60;             // columns 60, 61 of K block 0 hold 1.0: the A operand of the bias MMA
// weight ring depth: the training variant gives one stage to the cosine staging boxes
// kCBox2 (experiment, BN_CHAIN_CBOX2=1): the training variant trades one weight stage for a second cosine box per warp
template <bool kTrain, bool kCBox2 = false> __host__ __device__ constexpr int w_stages() { return kTrain ? (kCBox2 ? 3 : 4) : 5; }

struct SigmaChainParams {
  CUtensorMap wmap[kMaxLayers];         // packed W_l [F, Kpad_l] bf16, boxes 64 (K) x 128 (rows)
  CUtensorMap bmap;                     // bias blocks Wb [L * F, 64] bf16 (see the header comment), same boxes
  const float* wsig; const float* bsig;
  const float* origins; const float* dirs; const float* z;
  float* out;
  long long P;
  int o_stride, d_stride, S, L, skip, n_freq;
  uint64_t pol_w, pol_s;                // L2 eviction policies: weight tiles (resident), activation stores (streaming)
  long long* trace;                     // diagnostics (bn_debug_chain_trace): clock64() stamps of pair 0's leader CTA, first block
};

// training forward: same chain, and every layer leaves h_l = sin(.) and c_l = w0 cos(.) in HBM for the
// backward pass (h_l straight out of the activation K blocks in boxes of 32 rows x 64 columns, c_l through
// one 32 x 32 staging box per epilogue warp); K block 0 (the encoding) goes to X3.
struct TrainChainParams {
  CUtensorMap wmap[kMaxLayers];
  CUtensorMap bmap;                     // bias blocks Wb [L * F, 64]
  CUtensorMap hmap[kMaxLayers];         // H_l [P, F] (layer skip-1 lives inside X3, pitch 64 + F)
  CUtensorMap cmap[kMaxLayers];         // C_l [P, F], boxes of 32 columns x 32 rows, 64-byte swizzle
  CUtensorMap x3map;                    // X3 [P, 64] encoding columns
  const float* origins; const float* dirs; const float* z;
  long long P;
  int o_stride, d_stride, S, L, skip, n_freq;
  int store_c;                          // 0: inference without analytic normals - no cosines are computed or stored
  int h_from;                           // h_l leaves the SM only for l >= h_from (0 when training, L-1 for inference)
  long long* trace;                     // diagnostics (bn_debug_chain_trace): clock64() stamps of pair 0's leader CTA, first block
  uint64_t pol_w, pol_s;                // L2 eviction policies: weight tiles (resident), activation stores (streaming)
  // density of every point from the fp32 sines of the last layer, exactly as the density pass computes it (nullptr: off)
  const float* wsig; const float* bsig; float* sig_out;
  float* sig_out2;                      // optional second copy (the caller's own density buffer)
};

template <bool kTrain> __host__ __device__ constexpr int chain_smem() {
  // activations + weight ring + (cosine boxes | sigma exchange) + barriers + alignment slack
  return kNKB * kKBBytes + w_stages<kTrain>() * kKBBytes + (kTrain ? 4 * 4096 + 512 : 1024) + 512 + 1024;
}
constexpr int sigma_chain_smem() { return chain_smem<false>(); }

// every layer starts with K block 0 against its bias block; layers that do not read the encoding use only its last K slice
__device__ __forceinline__ bool layer_reads_enc(int l, int skip) { return l == 0 || l == skip; }
__device__ __forceinline__ int layer_kb_last(int l) { return l == 0 ? 0 : kNKB - 1; }
// column of K block kb inside the packed weight matrix of layer l
__device__ __forceinline__ int layer_wcol(int l, int skip, int kb) { return (l == 0) ? 0 : (l == skip ? kb * 64 : (kb - 1) * 64); }

// ---- weight producer: one [128 x 64] tile of W_l per (block, layer, column half, K block) ----
template <int STAGES>
__device__ __forceinline__ void chain_producer(const CUtensorMap* wmap, const CUtensorMap* bmap, uint8_t* sW, uint64_t* wfull,
                                               uint64_t* wempty, int crank, int pair0, int npairs, int n_blocks, int L, int skip,
                                               uint64_t pol_w) {
  int stage = 0; uint32_t phase = 0;
  for (int blk = pair0; blk < n_blocks; blk += npairs)
    for (int l = 0; l < L; ++l)
      for (int n = 0; n < 2; ++n)
        for (int kb = 0; kb <= layer_kb_last(l); ++kb) {
          mbar_wait(&wempty[stage], phase ^ 1);
          if (crank == 0) mbar_expect_tx(&wfull[stage], 2 * kKBBytes);
          const uint32_t bar = mapa_u32(smem_u32(&wfull[stage]), 0);
          if (kb == 0) tma_load_2d_pair_hint(sW + stage * kKBBytes, bmap, bar, 0, l * kF + n * 256 + crank * 128, pol_w);
          else tma_load_2d_pair_hint(sW + stage * kKBBytes, &wmap[l], bar, layer_wcol(l, skip, kb), n * 256 + crank * 128, pol_w);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
}

// ---- MMA issuer (leader CTA): layer = two column halves x the layer's K blocks ----
// Runs on the WHOLE warp (uniform control flow; one elected lane issues) with descriptors = base + constant increments: the lean
// issue loop of gemm_tc.cuh — the one-lane version spent 195 cycles per 256 x 256 x 16 MMA on its own scalar instructions
// against the pipe's 128 (scripts/mma_rate.cu).
template <int STAGES>
__device__ __forceinline__ void chain_mma(uint8_t* sAct, uint8_t* sW, uint64_t* wfull, uint64_t* wempty, uint64_t* tfull,
                                          uint64_t* tempty, uint64_t* act_ready, uint64_t* kfree, uint32_t tmem_base,
                                          int pair0, int npairs, int n_blocks, int L, int skip, long long* trace = nullptr) {
  constexpr uint32_t idesc = make_idesc(256, 256, false);
  const uint64_t a_desc0 = make_desc(smem_u32(sAct), 16, 1024), b_desc0 = make_desc(smem_u32(sW), 16, 1024);
  constexpr int kb_inc = kKBBytes >> 4;                 // descriptor address units per K block / weight stage
  if ((threadIdx.x & 31) != 0) trace = nullptr;
  int stage = 0; uint32_t phase = 0;
  uint32_t te_ph[2] = {0, 0};
  uint32_t ar_ph = 0;                                   // bit kb: parity the next wait on act_ready[kb] expects
  for (int blk = pair0; blk < n_blocks; blk += npairs)
    for (int l = 0; l < L; ++l)
      for (int n = 0; n < 2; ++n) {
        mbar_wait(&tempty[n], te_ph[n] ^ 1); te_ph[n] ^= 1;      // the epilogue has read this half out
        fence_after_sync();
        const bool tr = trace != nullptr && blk == pair0 && pair0 == 0;
        if (tr) trace[(l * 2 + n) * 16 + 0] = clock64();      // [0] TMEM half free
        const int k_first0 = layer_reads_enc(l, skip) ? 0 : 3;   // K block 0: all of it, or only the slice with the constant 1
        long long t_act = 0, t_w = 0;                            // trace: cycles spent waiting for activations / weights
        for (int kb = 0; kb <= layer_kb_last(l); ++kb) {
          if (n == 0 && !(kb == 0 && l > 0)) {          // K block published once per layer (PE: once per block)
            const long long t0 = tr ? clock64() : 0;
            mbar_wait(&act_ready[kb], (ar_ph >> kb) & 1); ar_ph ^= 1u << kb;
            fence_after_sync();
            if (tr) t_act += clock64() - t0;
          }
          if (tr) trace[(l * 2 + n) * 16 + 1 + kb] = clock64(); // [1+kb] K block kb available to the issuer
          {
            const long long t0 = tr ? clock64() : 0;
            mbar_wait(&wfull[stage], phase);
            if (tr) t_w += clock64() - t0;
          }
          fence_after_sync();
          const uint64_t da0 = a_desc0 + (uint64_t)(kb * kb_inc), db0 = b_desc0 + (uint64_t)(stage * kb_inc);
          if (elect_one()) {
            if (kb == 0 && k_first0 == 3) {
              umma_bf16_pair(tmem_base + n * 256, da0 + 6, db0 + 6, idesc, 0u);       // slice 3 only: overwrites the accumulator
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_pair(tmem_base + n * 256, da0 + 2 * k, db0 + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_pair(&wempty[stage]);
            // K block kb (1..4) has now been read by both halves of this layer: the first half's epilogue may overwrite it
            // (one barrier per K block, so that its units publish their results one by one while K blocks 5..8 are still
            // being multiplied); a layer that reads no activation K block at all (layer 0) releases the four at once
            if (n == 1) {
              if (kb >= 1 && kb <= 4) umma_commit_pair(&kfree[kb - 1]);
              else if (kb == 0 && layer_kb_last(l) == 0) { for (int j = 0; j < 4; ++j) umma_commit_pair(&kfree[j]); }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tfull[n]);
        __syncwarp();
        if (tr) {
          trace[(l * 2 + n) * 16 + 10] = clock64();           // [10] all MMAs of the half issued
          trace[(l * 2 + n) * 16 + 14] = t_act;               // [14] cycles the issuer waited for activation K blocks
          trace[(l * 2 + n) * 16 + 15] = t_w;                 // [15] cycles the issuer waited for weight tiles
        }
      }
}

// x = o + d z (separately rounded, as the reference) and the positional encoding of one point -> one 128-byte
// row of K block 0
__device__ __forceinline__ void encode_row(const float* origins, int o_stride, const float* dirs, int d_stride, const float* z,
                                           int S, long long pc, int n_freq, uint8_t* kb0, uint32_t row_off, uint32_t swz) {
  const long long r = pc / S;
  const float zz = z[pc];
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(origins[r * o_stride + a], __fmul_rn(dirs[r * d_stride + a], zz));
  float e[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[kOneCol] = 1.0f; e[kOneCol + 1] = 1.0f;            // A operand of the bias MMAs (hi and lo part of every bias)
  if (n_freq == 0) { e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; }
  else {
#pragma unroll
    for (int k = 0; k < 10; ++k)
      if (k < n_freq) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int a = 0; a < 3; ++a) { float s, c; sincosf(f * x[a], &s, &c); e[k * 6 + a] = s; e[k * 6 + 3 + a] = c; }
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts128(kb0 + row_off + ((j << 4) ^ swz), bf_pack(e[8 * j], e[8 * j + 1]), bf_pack(e[8 * j + 2], e[8 * j + 3]),
           bf_pack(e[8 * j + 4], e[8 * j + 5]), bf_pack(e[8 * j + 6], e[8 * j + 7]));
}

__global__ void __launch_bounds__(kThreads, 1) sigma_chain_kernel(const __grid_constant__ SigmaChainParams prm) {
  constexpr int kWStages = w_stages<false>();
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAct = smem;                                        // [kNKB][128 rows][128 B], swizzled
  uint8_t* sW = sAct + kNKB * kKBBytes;                        // [kWStages][128 rows of W][128 B]
  float* sSig = reinterpret_cast<float*>(sW + kWStages * kKBBytes);   // [128] partial sigma of the hsel = 1 warps
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sSig + 256);
  uint64_t* wempty = wfull + kWStages;
  uint64_t* tfull = wempty + kWStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* act_ready = tempty + 2;                            // [kNKB]
  uint64_t* kfree = act_ready + kNKB;                          // [4]: K block 1 + j of the current layer is no longer read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kfree + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int crank = (int)cluster_ctarank();
  const int pair0 = (int)cluster_id_x(), npairs = (int)cluster_nctaid_x();
  const int n_blocks = (int)((prm.P + 255) / 256);
  const int L = prm.L, skip = prm.skip;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.wmap[l])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.bmap)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 16); }     // 8 epilogue warps x 2 CTAs
    for (int s = 0; s < kNKB; ++s) mbar_init(&act_ready[s], 16);
    for (int s = 0; s < 4; ++s) mbar_init(&kfree[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<true>(tmem_slot, 512);
  pdl_wait();                                                  // the weights may come from the optimizer kernel just before
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) chain_producer<kWStages>(prm.wmap, &prm.bmap, sW, wfull, wempty, crank, pair0, npairs, n_blocks, L, skip, prm.pol_w);
  } else if (warp == 1) {
    if (crank == 0)
      chain_mma<kWStages>(sAct, sW, wfull, wempty, tfull, tempty, act_ready, kfree, tmem_base, pair0, npairs, n_blocks, L, skip, prm.trace);
  } else if (warp >= 4) {
    // ===================== positional encoding + epilogues =====================
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    const int row = q * 32 + lane;                           // row of this CTA's 128-row slab
    const uint32_t row_off = row * 128, swz = (lane & 7) << 4;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t tf_ph[2] = {0, 0};
    uint32_t kf_ph = 0;
    auto arrive_leader = [&](uint64_t* bar) {               // one arrival per warp on the leader CTA's barrier
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    for (int blk = pair0; blk < n_blocks; blk += npairs) {
      const long long p = (long long)blk * 256 + crank * 128 + row;
      // ---- x = o + d z, encoding into K block 0 (hsel 0 warps; one row per thread) ----
      if (hsel == 0) {
        encode_row(prm.origins, prm.o_stride, prm.dirs, prm.d_stride, prm.z, prm.S, p < prm.P ? p : prm.P - 1, prm.n_freq,
                   sAct, row_off, swz);
        fence_async_smem();
      }
      arrive_leader(&act_ready[0]);
      float sig = 0.f;
      for (int l = 0; l < L; ++l) {
        const float w0 = l == 0 ? 30.0f : 1.0f;
        const bool last = l == L - 1;
        for (int n = 0; n < 2; ++n) {
          const bool tr = prm.trace != nullptr && blk == pair0 && pair0 == 0 && crank == 0 && warp == 4 && lane == 0;
          if (tr) prm.trace[(l * 2 + n) * 16 + 11] = clock64();   // [11] epilogue starts waiting for the half
          mbar_wait(&tfull[n], tf_ph[n]); tf_ph[n] ^= 1;
          fence_after_sync();
          if (tr) prm.trace[(l * 2 + n) * 16 + 12] = clock64();   // [12] half complete (tfull)
          // software-pipelined TMEM reads: unit u+1 is in flight while unit u goes through the MUFU
          uint32_t va[32], vb[32];
          const uint32_t tbase = tmem_base + t_lane + n * 256 + hsel * 32;
          tmem_ld32_issue(tbase, va);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            tmem_wait_ld();
            uint32_t (&v)[32] = (u & 1) ? vb : va;
            if (u < 3) tmem_ld32_issue(tbase + (u + 1) * 64, (u & 1) ? va : vb);
            else { fence_before_sync(); arrive_leader(&tempty[n]); }       // this warp's share of the half is in registers
            const int col0 = n * 256 + u * 64 + hsel * 32;
            if (!last) {                 // the accumulator is the pre-activation: the bias came in through the MMA
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j)
                pk[j] = bf_pack(__sinf(w0 * __uint_as_float(v[2 * j])), __sinf(w0 * __uint_as_float(v[2 * j + 1])));
              // in place: the unit becomes K block 1 + 4n + u of the next layer.  Second half: every MMA of this layer has
              // retired (tfull[1]).  First half: K block 1 + u still feeds the second half's MMAs until kfree[u]
              if (n == 0) mbar_wait(&kfree[u], kf_ph);
              uint8_t* kbp = sAct + (1 + 4 * n + u) * kKBBytes;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(kbp + row_off + (((hsel * 4 + j) << 4) ^ swz), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              fence_async_smem();
              arrive_leader(&act_ready[1 + 4 * n + u]);
            } else {
              const float4* wp = reinterpret_cast<const float4*>(prm.wsig + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w = __ldg(wp + j);
                sig = fmaf(__sinf(w0 * __uint_as_float(v[4 * j])), w.x, sig); sig = fmaf(__sinf(w0 * __uint_as_float(v[4 * j + 1])), w.y, sig);
                sig = fmaf(__sinf(w0 * __uint_as_float(v[4 * j + 2])), w.z, sig); sig = fmaf(__sinf(w0 * __uint_as_float(v[4 * j + 3])), w.w, sig);
              }
              if (n == 0) mbar_wait(&kfree[u], kf_ph);     // keep the barrier phases in step with the issuer
            }
          }
          if (n == 0) kf_ph ^= 1;
          if (tr) prm.trace[(l * 2 + n) * 16 + 13] = clock64();   // [13] the half's four units are through the MUFU
        }
      }
      // ---- sigma = softplus(w_sigma . h_{L-1} + b): the two warps of a quadrant hold half of the columns each ----
      if (hsel == 1) sSig[row] = sig;
      named_bar_sync(1 + q, 64);
      if (hsel == 0 && p < prm.P) {
        const float s = sig + sSig[row] + __ldg(prm.bsig);
        prm.out[p] = s > 20.f ? s : log1pf(expf(s));
      }
      named_bar_sync(1 + q, 64);                             // sSig is reused by the next block
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { fence_after_sync(); tmem_dealloc<true>(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------
// Training forward of the trunk: the same chain; every layer's h_l and c_l = w0 cos(.) are written to HBM.
template <bool kCBox2>
__global__ void __launch_bounds__(kThreads, 1) train_chain_kernel(const __grid_constant__ TrainChainParams prm) {
  constexpr int kWStages = w_stages<true, kCBox2>();
  constexpr int kCBytes = (kCBox2 ? 8 : 4) * 4096;            // 3 stages + 32 KB of boxes == 4 stages + 16 KB
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAct = smem;
  uint8_t* sW = sAct + kNKB * kKBBytes;
  uint8_t* sC = sW + kWStages * kKBBytes;                      // [8 warps][32 rows][64 B] cosine staging boxes
  float* sSig = reinterpret_cast<float*>(sC + kCBytes);        // [128] partial sigma of the hsel = 1 warps
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sSig + 128);
  uint64_t* wempty = wfull + kWStages;
  uint64_t* tfull = wempty + kWStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* act_ready = tempty + 2;
  uint64_t* kfree = act_ready + kNKB;                          // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kfree + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int crank = (int)cluster_ctarank();
  const int pair0 = (int)cluster_id_x(), npairs = (int)cluster_nctaid_x();
  const int n_blocks = (int)((prm.P + 255) / 256);
  const int L = prm.L, skip = prm.skip;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < L; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.wmap[l])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.bmap)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 16); }
    for (int s = 0; s < kNKB; ++s) mbar_init(&act_ready[s], 16);
    for (int s = 0; s < 4; ++s) mbar_init(&kfree[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<true>(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) chain_producer<kWStages>(prm.wmap, &prm.bmap, sW, wfull, wempty, crank, pair0, npairs, n_blocks, L, skip, prm.pol_w);
  } else if (warp == 1) {
    if (crank == 0)
      chain_mma<kWStages>(sAct, sW, wfull, wempty, tfull, tempty, act_ready, kfree, tmem_base, pair0, npairs, n_blocks, L, skip, prm.trace);
  } else if (warp >= 4) {
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t row_off = row * 128, swz = (lane & 7) << 4;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const bool leader = hsel == 0 && lane == 0;              // issues the quadrant's h_l stores (64-column boxes)
    // cosines leave through one box per WARP (32 rows x 32 columns, 64-byte swizzle): no cross-warp hand-shake
    uint8_t* cbox = sC + (q * 2 + hsel) * (kCBox2 ? 4096 : 2048);
    uint32_t cu = 0;                                         // cosine units stored so far by this warp (box parity)
    const uint32_t crow_off = lane * 64, cswz = ((lane >> 1) & 3) << 4;
    uint32_t tf_ph[2] = {0, 0};
    uint32_t kf_ph = 0;
    auto arrive_leader = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    for (int blk = pair0; blk < n_blocks; blk += npairs) {
      const int grow0 = blk * 256 + crank * 128 + q * 32;      // first global row of this quadrant
      const long long p = (long long)grow0 + lane;
      // ---- encoding -> K block 0 -> X3 ----
      if (hsel == 0) {
        if (lane == 0) bulk_wait_read0();                      // this warp's earlier stores have left shared memory
        __syncwarp();
        encode_row(prm.origins, prm.o_stride, prm.dirs, prm.d_stride, prm.z, prm.S, p < prm.P ? p : prm.P - 1, prm.n_freq,
                   sAct, row_off, swz);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d_hint(&prm.x3map, sAct + q * 4096, 0, grow0, prm.pol_s); bulk_commit(); }
      }
      arrive_leader(&act_ready[0]);
      float sig = 0.f;
      for (int l = 0; l < L; ++l) {
        const bool last = l == L - 1;
        const bool sig_on = last && prm.sig_out != nullptr;
        for (int n = 0; n < 2; ++n) {
          const bool tr = prm.trace != nullptr && blk == pair0 && pair0 == 0 && crank == 0 && warp == 4 && lane == 0;
          if (tr) prm.trace[(l * 2 + n) * 16 + 11] = clock64();   // [11] epilogue starts waiting for the half
          mbar_wait(&tfull[n], tf_ph[n]); tf_ph[n] ^= 1;
          fence_after_sync();
          if (tr) prm.trace[(l * 2 + n) * 16 + 12] = clock64();   // [12] half complete (tfull)
          uint32_t va[32], vb[32];
          const uint32_t tbase = tmem_base + t_lane + n * 256 + hsel * 32;
          const bool store_h = l >= prm.h_from;
          const bool l0 = l == 0;
          const bool row_ok = (long long)grow0 + lane < prm.P;
          tmem_ld32_issue(tbase, va);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            tmem_wait_ld();
            uint32_t (&v)[32] = (u & 1) ? vb : va;
            if (u < 3) tmem_ld32_issue(tbase + (u + 1) * 64, (u & 1) ? va : vb);
            else { fence_before_sync(); arrive_leader(&tempty[n]); }
            // the accumulator is the pre-activation (bias added by the tensor core): per element ONE range reduction feeds
            // both MUFU.SIN and MUFU.COS; w0 = 30 only exists in layer 0, every other layer skips both multiplies
            uint32_t pk[16], pc[16];
            if (l0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a0 = 30.0f * __uint_as_float(v[2 * j]), a1 = 30.0f * __uint_as_float(v[2 * j + 1]);
                pk[j] = bf_pack(__sinf(a0), __sinf(a1));
                if (prm.store_c) pc[j] = bf_pack(30.0f * __cosf(a0), 30.0f * __cosf(a1));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a0 = __uint_as_float(v[2 * j]), a1 = __uint_as_float(v[2 * j + 1]);
                pk[j] = bf_pack(__sinf(a0), __sinf(a1));
                if (prm.store_c) pc[j] = bf_pack(__cosf(a0), __cosf(a1));
              }
            }
            if (sig_on) {
              // last layer: density = w_sigma . h from the sines as they are stored (bf16), fp32 weights and accumulation;
              // unpacking the 16 words costs two ALU operations each, re-evaluating 32 sines would sit on the MUFU pipe
              // that bounds this epilogue
              const float2* wp = reinterpret_cast<const float2*>(prm.wsig + n * 256 + u * 64 + hsel * 32);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 ws = __ldg(wp + j);
                sig = fmaf(__uint_as_float(pk[j] << 16), ws.x, sig);
                sig = fmaf(__uint_as_float(pk[j] & 0xFFFF0000u), ws.y, sig);
              }
            }
            uint8_t* box = cbox + (kCBox2 ? (cu & 1) * 2048 : 0);
            if (prm.store_c) {
              // cosines: registers -> this warp's staging box -> TMA.  The box was read out by the TMA (with two boxes: the
              // store before the previous one); the leader's wait also covers the h_l boxes of the previous half
              if (lane == 0) { if (kCBox2 && !leader) bulk_wait_read1(); else bulk_wait_read0(); }
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(box + crow_off + ((j << 4) ^ cswz), pc[4 * j], pc[4 * j + 1], pc[4 * j + 2], pc[4 * j + 3]);
              ++cu;
            }
            // sines: in place into K block 1 + 4n + u of the next layer.  Second half: every MMA of this layer has retired
            // (tfull[1]); first half: K block 1 + u still feeds the second half's MMAs until kfree[u].  (The leader's h_l
            // stores of the previous layer out of this K block were read out long ago: it waits for them at every cosine box.)
            if (n == 0) mbar_wait(&kfree[u], kf_ph);
            uint8_t* kbp = sAct + (1 + 4 * n + u) * kKBBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(kbp + row_off + (((hsel * 4 + j) << 4) ^ swz), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_async_smem();                                  // ONE proxy fence covers both the K block and the cosine box
            if (!last) arrive_leader(&act_ready[1 + 4 * n + u]); else __syncwarp();
            if (prm.store_c && lane == 0) {
              tma_store_2d_hint(&prm.cmap[l], box, n * 256 + u * 64 + hsel * 32, grow0, prm.pol_s); bulk_commit();
            }
          }
          if (n == 0) kf_ph ^= 1;
          if (tr) prm.trace[(l * 2 + n) * 16 + 13] = clock64();   // [13] the half's units are through the MUFU / cosine stores
          if (store_h) {
            // h_l of this half: both warps of the quadrant have written their columns -> four 64-column boxes straight out
            // of the K blocks (read-only for everyone until the next layer's epilogue of the same half)
            // (last layer, second half: the same barrier hands the hsel = 1 warp's half of the sigma dot product over;
            // the next write of sSig[row] is a whole block of these barriers away)
            const bool sig_now = sig_on && n == 1;
            if (sig_now && hsel == 1) sSig[row] = sig;
            named_bar_sync(1 + q, 64);
            if (leader) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                tma_store_2d_hint(&prm.hmap[l], sAct + (1 + n * 4 + u) * kKBBytes + q * 4096, n * 256 + u * 64, grow0, prm.pol_s);
              bulk_commit();
            }
            if (sig_now && hsel == 0 && row_ok) {
              const float sv = sig + sSig[row] + __ldg(prm.bsig);
              const float sp = sv > 20.f ? sv : log1pf(expf(sv));
              prm.sig_out[p] = sp;
              if (prm.sig_out2) prm.sig_out2[p] = sp;
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait0();
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { fence_after_sync(); tmem_dealloc<true>(tmem_base, 512); }
}

}  // namespace chain
}  // namespace bn
