// Fused data-gradient chain of the SIREN trunk: ONE persistent kernel computes
//     dZ_{l-1} = (dZ_l W_l[:, h-part]) ⊙ c_{l-1}        for l = L-1 ... 1
// (autograd of calc_features, reference models/spsbrdfnerf.py:636-646: d/dz of sin(w0 z) is c = w0 cos(w0 z), kept by the
// training forward) for a block of 256 points per CTA pair — the mirror image of chain::train_chain_kernel.
//
// The per-layer GEMMs read dZ_l, read c_{l-1} and write dZ_{l-1}: 402 MB per layer at P = 131 072, at the HBM roofline.
// Here dZ stays in shared memory as the 128B-swizzled K-major A operand of the next (lower) layer; per layer only
// c_{l-1} comes in (TMA, 4 KB boxes per TMEM quadrant, prefetched several units ahead by a dedicated warp) and dZ_{l-1}
// goes out once (TMA straight out of the activation K blocks, for the weight-gradient GEMM of layer l-1): 268 MB per layer.
//
//   warp 0     weight producer: [128 x 64] tiles of W_l^T (rows = input features = output columns of the dgrad),
//              ring of kDWStages, pair protocol of gemm_tc.cuh (the leader's barrier collects both CTAs' bytes)
//   warp 1     MMA issuer (leader CTA): tcgen05.mma.cta_group::2, M = 256, N = 256 per column half, K blocks of 64
//   warp 2     TMEM allocator
//   warp 3     operand loader: dZ_{L-1} of the block (8 K blocks) at block start, then the c_{l-1} boxes of every unit
//   warps 4-11 epilogue: tcgen05.ld -> * c -> bf16 -> swizzled st.shared IN PLACE into the K block of the next layer,
//              published per K block; the first half's units wait for kfree[u] (their K block still feeds the second
//              half's MMAs), exactly as in the forward chain
#pragma once
#include "mlp_chain.cuh"

namespace bn {
namespace chain {

constexpr int kDKB = kF / 64;            // 8 activation K blocks (no encoding block: gradients w.r.t. the inputs are not needed)
// weight ring depth kDWStages and c boxes per TMEM quadrant kDCBoxes (4 KB each) are template parameters: (3, 3) = two units
// of c prefetch, (4, 2) = a deeper weight ring; both fill the 227 KB

struct DgradChainParams {
  CUtensorMap wmap[kMaxLayers];          // [l]: W_l^T restricted to the h-part rows: [F (in), F (out)] bf16, boxes 64 x 128
  CUtensorMap cmap[kMaxLayers];          // [l]: c_{l-1} [P, F], boxes 64 columns x 32 rows
  CUtensorMap gout[kMaxLayers];          // [l]: dZ_{l-1} [P, F], boxes 64 columns x 32 rows
  CUtensorMap gin;                       // dZ_{L-1} [P, F], boxes 64 columns x 128 rows
  uint64_t pol_w, pol_s;                 // L2 eviction policies: weight tiles (resident), streamed activations
  long long P;
  int L;
};

template <int kDWStages, int kDCBoxes> __host__ __device__ constexpr int dgrad_chain_smem() {
  return kDKB * kKBBytes + kDWStages * kKBBytes + 4 * kDCBoxes * 4096 + 1024 + 1024;
}

template <int kDWStages, int kDCBoxes>
__global__ void __launch_bounds__(kThreads, 1) dgrad_chain_kernel(const __grid_constant__ DgradChainParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAct = smem;                                        // [kDKB][128 rows][128 B], swizzled
  uint8_t* sW = sAct + kDKB * kKBBytes;                        // [kDWStages][128 rows of W^T][128 B]
  uint8_t* sC = sW + kDWStages * kKBBytes;                     // [4 quadrants][kDCBoxes][32 rows][128 B]
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sC + 4 * kDCBoxes * 4096);
  uint64_t* wempty = wfull + kDWStages;
  uint64_t* tfull = wempty + kDWStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* act_ready = tempty + 2;                            // [kDKB] leader CTA: K block written by all 16 epilogue warps
  uint64_t* gin_full = act_ready + kDKB;                       // [kDKB] leader CTA: dZ_{L-1} K block of both CTAs has landed
  uint64_t* kfree = gin_full + kDKB;                           // [4]
  uint64_t* cfull = kfree + 4;                                 // [4][kDCBoxes] local
  uint64_t* cempty = cfull + 4 * kDCBoxes;                     // [4][kDCBoxes] local, 2 arrivals (the quadrant's two warps)
  uint64_t* blk_free = cempty + 4 * kDCBoxes;                  // local: the block's last dZ stores have been read out (4 leaders)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(blk_free + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int crank = (int)cluster_ctarank();
  const int pair0 = (int)cluster_id_x(), npairs = (int)cluster_nctaid_x();
  const int n_blocks = (int)((prm.P + 255) / 256);
  const int L = prm.L;

  if (warp == 0 && lane == 0) {
    for (int l = 1; l < L; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.wmap[l])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.cmap[l])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.gout[l])) : "memory");
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&prm.gin)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDWStages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 16); }
    for (int s = 0; s < kDKB; ++s) { mbar_init(&act_ready[s], 16); mbar_init(&gin_full[s], 1); }
    for (int s = 0; s < 4; ++s) mbar_init(&kfree[s], 1);
    for (int s = 0; s < 4 * kDCBoxes; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], 2); }
    mbar_init(blk_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<true>(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int blk = pair0; blk < n_blocks; blk += npairs)
        for (int l = L - 1; l >= 1; --l)
          for (int n = 0; n < 2; ++n)
            for (int kb = 0; kb < kDKB; ++kb) {
              mbar_wait(&wempty[stage], phase ^ 1);
              if (crank == 0) mbar_expect_tx(&wfull[stage], 2 * kKBBytes);
              tma_load_2d_pair_hint(sW + stage * kKBBytes, &prm.wmap[l], mapa_u32(smem_u32(&wfull[stage]), 0), kb * 64, n * 256 + crank * 128, prm.pol_w);
              if (++stage == kDWStages) { stage = 0; phase ^= 1; }
            }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // lean issue loop (see chain_mma in mlp_chain.cuh): whole warp, one elected lane issues, descriptors = base + increments
    if (crank == 0) {
      constexpr uint32_t idesc = make_idesc(256, 256, false);
      const uint64_t a_desc0 = make_desc(smem_u32(sAct), 16, 1024), b_desc0 = make_desc(smem_u32(sW), 16, 1024);
      constexpr int kb_inc = kKBBytes >> 4;
      int stage = 0; uint32_t phase = 0;
      uint32_t te_ph[2] = {0, 0};
      uint32_t ar_ph = 0, gi_ph = 0;
      for (int blk = pair0; blk < n_blocks; blk += npairs) {
        for (int l = L - 1; l >= 1; --l)
          for (int n = 0; n < 2; ++n) {
            mbar_wait(&tempty[n], te_ph[n] ^ 1); te_ph[n] ^= 1;
            fence_after_sync();
            for (int kb = 0; kb < kDKB; ++kb) {
              if (n == 0) {                    // the layer's input K block: from HBM for the top layer, else from the epilogue above
                if (l == L - 1) mbar_wait(&gin_full[kb], gi_ph);
                else { mbar_wait(&act_ready[kb], (ar_ph >> kb) & 1); ar_ph ^= 1u << kb; }
                fence_after_sync();
              }
              mbar_wait(&wfull[stage], phase);
              fence_after_sync();
              const uint64_t da0 = a_desc0 + (uint64_t)(kb * kb_inc), db0 = b_desc0 + (uint64_t)(stage * kb_inc);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_pair(tmem_base + n * 256, da0 + 2 * k, db0 + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit_pair(&wempty[stage]);
                if (n == 1 && kb < 4) umma_commit_pair(&kfree[kb]);     // K block kb may be overwritten by the first half's epilogue
              }
              __syncwarp();
              if (++stage == kDWStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit_pair(&tfull[n]);
            __syncwarp();
          }
        gi_ph ^= 1;
      }
    }
  } else if (warp == 3) {
    // ===================== operand loader: dZ_{L-1} of the block, then c_{l-1} boxes unit by unit =====================
    if (lane == 0) {
      uint32_t bf_ph = 0;
      uint32_t cslot = 0, cphase = 0;                            // same ring position for the four quadrants
      for (int blk = pair0; blk < n_blocks; blk += npairs) {
        const int row0 = blk * 256 + crank * 128;
        mbar_wait(blk_free, bf_ph ^ 1); bf_ph ^= 1;              // the previous block's stores out of the K blocks are done
        for (int kb = 0; kb < kDKB; ++kb) {
          if (crank == 0) mbar_expect_tx(&gin_full[kb], 2 * kKBBytes);
          tma_load_2d_pair_hint(sAct + kb * kKBBytes, &prm.gin, mapa_u32(smem_u32(&gin_full[kb]), 0), kb * 64, row0, prm.pol_s);
        }
        for (int l = L - 1; l >= 1; --l)
          for (int n = 0; n < 2; ++n)
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint64_t* full = &cfull[q * kDCBoxes + cslot];
                mbar_wait(&cempty[q * kDCBoxes + cslot], cphase ^ 1);
                mbar_expect_tx(full, 4096);
                tma_load_2d_hint(sC + (q * kDCBoxes + cslot) * 4096, &prm.cmap[l], full, n * 256 + u * 64, row0 + q * 32, prm.pol_s);
              }
              if (++cslot == kDCBoxes) { cslot = 0; cphase ^= 1; }
            }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t row_off = row * 128, swz = (lane & 7) << 4;
    const uint32_t crow_off = lane * 128;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const bool leader = hsel == 0 && lane == 0;
    uint32_t tf_ph[2] = {0, 0};
    uint32_t kf_ph = 0;
    uint32_t cslot = 0, cphase = 0;
    auto arrive_leader = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    for (int blk = pair0; blk < n_blocks; blk += npairs) {
      const int grow0 = blk * 256 + crank * 128 + q * 32;
      for (int l = L - 1; l >= 1; --l) {
        const bool last = l == 1;
        for (int n = 0; n < 2; ++n) {
          mbar_wait(&tfull[n], tf_ph[n]); tf_ph[n] ^= 1;
          fence_after_sync();
          uint32_t va[32], vb[32];
          const uint32_t tbase = tmem_base + t_lane + n * 256 + hsel * 32;
          tmem_ld32_issue(tbase, va);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // c_{l-1} of the unit: this warp's 32 columns of the quadrant's box
            uint32_t cc[16];
            {
              const uint8_t* box = sC + (q * kDCBoxes + cslot) * 4096;
              mbar_wait(&cfull[q * kDCBoxes + cslot], cphase);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 t = lds128(box + crow_off + (((hsel * 4 + j) << 4) ^ swz));
                cc[4 * j] = t.x; cc[4 * j + 1] = t.y; cc[4 * j + 2] = t.z; cc[4 * j + 3] = t.w;
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&cempty[q * kDCBoxes + cslot]);
              if (++cslot == kDCBoxes) { cslot = 0; cphase ^= 1; }
            }
            tmem_wait_ld();
            uint32_t (&v)[32] = (u & 1) ? vb : va;
            if (u < 3) tmem_ld32_issue(tbase + (u + 1) * 64, (u & 1) ? va : vb);
            else { fence_before_sync(); arrive_leader(&tempty[n]); }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = bf_pack(__uint_as_float(v[2 * j]) * bf_lo(cc[j]), __uint_as_float(v[2 * j + 1]) * bf_hi(cc[j]));
            // in place: the unit is K block 4n + u of the next layer down.  Second half: all MMAs of this layer have retired;
            // first half: K block u still feeds the second half's MMAs until kfree[u].  The store of the previous layer out of
            // this K block has been read (the quadrant leader waits for its stores before the half's next ones).
            if (n == 0) mbar_wait(&kfree[u], kf_ph);
            uint8_t* kbp = sAct + (4 * n + u) * kKBBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(kbp + row_off + (((hsel * 4 + j) << 4) ^ swz), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_async_smem();
            if (!last) arrive_leader(&act_ready[4 * n + u]);
          }
          if (n == 0) kf_ph ^= 1;
          // dZ_{l-1} of this half leaves for the weight-gradient GEMM: four 64-column boxes per quadrant out of the K blocks
          // (the leader first makes sure its previous group has been read out: whoever passes the barrier below may then
          // overwrite those K blocks — nothing is pending in practice, that group was issued a whole half ago)
          if (leader) bulk_wait_read0();
          named_bar_sync(1 + q, 64);
          if (leader) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              tma_store_2d_hint(&prm.gout[l], sAct + (4 * n + u) * kKBBytes + q * 4096, n * 256 + u * 64, grow0, prm.pol_s);
            bulk_commit();
          }
        }
      }
      // the block's K blocks are reloaded with the next block's dZ_{L-1}: every store out of them must have been read
      if (leader) { bulk_wait_read0(); mbar_arrive(blk_free); }
    }
    if (leader) bulk_wait0();
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) { fence_after_sync(); tmem_dealloc<true>(tmem_base, 512); }
}

}  // namespace chain
}  // namespace bn
