// Microbenchmark: how many cycles does ONE tcgen05.mma take on this chip when nothing else is in the way?
// Every SM (pair) issues the MMA stream of a GEMM mainloop — four K = 16 instructions per 64-wide K block out of a 4-slot
// shared-memory ring, one tcgen05.commit per K block, a slot reused only after its commit has arrived — with NO loads, NO
// epilogue, operands = whatever shared memory holds (zeros).  The answer is the practical ceiling of the issue phase of
// gemm_tc_kernel / the fused trunk kernels, i.e. the number their traces (scripts/trace_chain.py, trace_wgrad.py) are read
// against; the data sheet's 8192 dense bf16 FLOP / clk / SM would be 128 cycles per 256 x 256 x 16 pair instruction.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I brdf_nerf_b200/csrc -I include scripts/mma_rate.cu -o scripts/_build/mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "gemm_tc.cuh"

using namespace bn::tc;

constexpr int kStages = 4;
constexpr int kStageBytes = 49152;      // A 16 KB + B up to 32 KB (single CTA, N = 256)

// mode bit 0: MN-major operands (weight-gradient layout) instead of K-major; pair: cta_group::2 (M = 256) or ::1 (M = 128)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
  return p != 0;
}

template <bool kPair>
__global__ void __launch_bounds__(384, 1) mma_rate_kernel(int n_kblocks, int N, int mn_major, int wait_slots, long long* out, int waiters, int fill, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);      // [kStages] "empty" + 1 final
  uint64_t* fullb = bars + kStages + 1;                                             // [kStages] "full" (variant bit 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fullb + kStages);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int crank = kPair ? (int)cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < kStages * kStageBytes / 4; i += blockDim.x) {
    // fill 1: two pseudo-random bf16 values in [-2, 2) per word (sign, exponent 126..127, random mantissa) instead of zeros
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t lo = (h & 0x80FFu) | 0x3F00u, hi = ((h >> 16) & 0x80FFu) | 0x3F00u;
    reinterpret_cast<uint32_t*>(smem)[i] = fill ? (lo | (hi << 16)) : 0u;
  }
  fence_async_smem();
  if (warp == 1 && lane == 0) {
    for (int s = 0; s <= kStages; ++s) mbar_init(&bars[s], 1);
    for (int s = 0; s < kStages; ++s) mbar_init(&fullb[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<kPair>(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if ((variant & 32) && warp == 1 && crank == 0) {
    // variant bit 5: LEAN issue loop.  The whole warp runs the loop (uniform control flow, descriptors are warp-uniform values
    // the compiler can keep in uniform registers), one elected lane issues; descriptors are a base + constant increments
    // instead of being rebuilt from the address for every instruction.
    const uint32_t idesc = make_idesc(kPair ? 256 : 128, N, mn_major != 0);
    const uint64_t a_base = mn_major ? make_desc(smem_u32(smem), 8192, 1024) : make_desc(smem_u32(smem), 16, 1024);
    const uint64_t b_base = mn_major ? make_desc(smem_u32(smem) + 16384, 8192, 1024) : make_desc(smem_u32(smem) + 16384, 16, 1024);
    const uint32_t k_inc = (mn_major ? 2048 : 32) >> 4, s_inc = kStageBytes >> 4;
    int stage = 0; uint32_t phase = 0;
    const long long t0 = clock64();
    for (int kb = 0; kb < n_kblocks; ++kb) {
      if (wait_slots && kb >= kStages) mbar_wait(&bars[stage], phase ^ 1);
      const uint64_t da0 = a_base + (uint64_t)(stage * s_inc), db0 = b_base + (uint64_t)(stage * s_inc);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (kPair) umma_bf16_pair(tmem_base, da0 + k * k_inc, db0 + k * k_inc, idesc, (kb | k) ? 1u : 0u);
          else umma_bf16(tmem_base, da0 + k * k_inc, db0 + k * k_inc, idesc, (kb | k) ? 1u : 0u);
        }
        if constexpr (kPair) umma_commit_pair(&bars[stage]); else umma_commit(&bars[stage]);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    const long long t1 = clock64();
    if (elect_one()) { if constexpr (kPair) umma_commit_pair(&bars[kStages]); else umma_commit(&bars[kStages]); }
    __syncwarp();
    mbar_wait(&bars[kStages], 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 1 && lane == 0 && crank == 0 && !(variant & 4)) {
    const uint32_t idesc = make_idesc(kPair ? 256 : 128, N, mn_major != 0);
    int stage = 0; uint32_t phase = 0;
    const long long t0 = clock64();
    for (int kb = 0; kb < n_kblocks; ++kb) {
      if (variant & 2) mbar_wait(&fullb[stage], phase);                          // a producer thread relays "slot free" -> "slot full"
      else if (wait_slots && kb >= kStages) mbar_wait(&bars[stage], phase ^ 1);  // the slot's previous MMAs have retired
      if (variant & 1) fence_after_sync();
      const uint32_t a_addr = smem_u32(smem + stage * kStageBytes), b_addr = a_addr + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = mn_major ? make_desc(a_addr + k * 2048, 8192, 1024) : make_desc(a_addr + k * 32, 16, 1024);
        const uint64_t db = mn_major ? make_desc(b_addr + k * 2048, 8192, 1024) : make_desc(b_addr + k * 32, 16, 1024);
        const uint32_t d = tmem_base + ((variant & 8) ? (k & 1) * 256 : 0);      // bit 3: ONE thread, two accumulators in turn
        const uint32_t acc = (variant & 8) ? ((kb | (k >> 1)) ? 1u : 0u) : ((kb | k) ? 1u : 0u);
        if constexpr (kPair) umma_bf16_pair(d, da, db, idesc, acc);
        else umma_bf16(d, da, db, idesc, acc);
      }
      if constexpr (kPair) umma_commit_pair(&bars[stage]); else umma_commit(&bars[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    const long long t1 = clock64();                       // every instruction issued
    if constexpr (kPair) umma_commit_pair(&bars[kStages]); else umma_commit(&bars[kStages]);
    mbar_wait(&bars[kStages], 0);                         // ... and retired
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if ((variant & 4) && (warp == 0 || warp == 3) && lane == 0 && crank == 0) {
    // variant bit 2: TWO issuing threads, each with half of the K blocks, its own two ring slots and its own accumulator
    // (columns 0..255 / 256..511): does the floor belong to the issuing thread or to the tensor pipe?
    const int w = warp == 0 ? 0 : 1;
    const uint32_t idesc = make_idesc(kPair ? 256 : 128, N, mn_major != 0);
    const long long t0 = clock64();
    for (int j = 0; j < n_kblocks / 2; ++j) {
      const int stage = w * 2 + (j & 1);
      if (j >= 2) mbar_wait(&bars[stage], ((j >> 1) & 1) ^ 1);
      const uint32_t a_addr = smem_u32(smem + stage * kStageBytes), b_addr = a_addr + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = mn_major ? make_desc(a_addr + k * 2048, 8192, 1024) : make_desc(a_addr + k * 32, 16, 1024);
        const uint64_t db = mn_major ? make_desc(b_addr + k * 2048, 8192, 1024) : make_desc(b_addr + k * 32, 16, 1024);
        const uint32_t d = tmem_base + ((variant & 16) ? 0 : w * 256);           // bit 4: both threads into the SAME accumulator
        if constexpr (kPair) umma_bf16_pair(d, da, db, idesc, (j | k) ? 1u : 0u);
        else umma_bf16(d, da, db, idesc, (j | k) ? 1u : 0u);
      }
      if constexpr (kPair) umma_commit_pair(&bars[stage]); else umma_commit(&bars[stage]);
    }
    const long long t1 = clock64();
    if constexpr (kPair) umma_commit_pair(&fullb[w]); else umma_commit(&fullb[w]);
    mbar_wait(&fullb[w], 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 0 && lane == 0 && (variant & 2)) {
    // the real kernels' producer: waits for the slot to be free, then (leader only) completes the slot's "full" phase
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < n_kblocks; ++kb) {
      mbar_wait(&bars[stage], phase ^ 1);
      if (crank == 0) mbar_expect_tx(&fullb[stage], 0);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4 && waiters && crank == 0) {
    // what the epilogue warps of the real kernels do during the mainloop: wait for the accumulator on an mbarrier
    // waiters 1: every lane polls (mbar_wait as the kernels have it); 2: one lane polls; 3: one lane polls with nanosleep back-off
    if (waiters == 1) mbar_wait(&bars[kStages], 0);
    else if (lane == 0) {
      while (!mbar_try_wait(&bars[kStages], 0)) { if (waiters == 3) __nanosleep(200); }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 2) { fence_after_sync(); tmem_dealloc<kPair>(tmem_base, 512); }
}

template <bool kPair>
static void run(const char* name, int n_sm, int N, int mn_major, int wait_slots, int waiters = 0, int fill = 0, int variant = 0) {
  const int n_kblocks = 512;
  long long* out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
  const int smem = kStages * kStageBytes + 1024 + 256;
  auto kern = mma_rate_kernel<kPair>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kPair ? n_sm / 2 * 2 : n_sm); cfg.blockDim = dim3(waiters ? 384 : 128); cfg.dynamicSmemBytes = smem;   // warps 0..3 always exist
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f; long long h[2] = {0, 0};
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, n_kblocks, N, mn_major, wait_slots, out, waiters, fill, variant);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  }
  const double n_mma = 4.0 * n_kblocks;
  const double flop = 2.0 * (kPair ? 256 : 128) * N * 16 * n_mma * (kPair ? cfg.gridDim.x / 2 : cfg.gridDim.x);
  printf("%-44s M=%d N=%3d %s %s: %6.1f cycles / MMA issued, %6.1f retired | kernel %7.1f us = %7.1f TFLOP/s on %d SMs\n", name,
         kPair ? 256 : 128, N, mn_major ? "MN-major" : "K-major ", wait_slots ? "ring " : "free ", h[0] / n_mma, h[1] / n_mma, best * 1e3,
         flop / (best * 1e-3) / 1e12, (int)cfg.gridDim.x);
  cudaFree(out);
}

int main() {
  int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  for (int mn = 0; mn < 2; ++mn) {
    run<true>("pair (cta_group::2), all SMs", n_sm, 256, mn, 1);
    run<true>("pair (cta_group::2), all SMs, no slot waits", n_sm, 256, mn, 0);
    run<true>("pair (cta_group::2), ONE pair alone", 2, 256, mn, 1);
    run<true>("pair (cta_group::2), all SMs", n_sm, 128, mn, 1);
    run<false>("single CTA (cta_group::1), all SMs", n_sm, 256, mn, 1);
    run<false>("single CTA (cta_group::1), ONE SM alone", 1, 256, mn, 1);
    run<false>("single CTA (cta_group::1), all SMs", n_sm, 128, mn, 1);
  }
  run<true>("pair, RANDOM operands", n_sm, 256, 0, 1, 0, 1);
  run<true>("pair, RANDOM operands", n_sm, 256, 1, 1, 0, 1);
  run<true>("pair, RANDOM operands, ONE pair alone", 2, 256, 1, 1, 0, 1);
  run<false>("single CTA, RANDOM operands", n_sm, 256, 1, 1, 0, 1);
  run<true>("pair, RANDOM operands, N = 128", n_sm, 128, 1, 1, 0, 1);
  run<true>("pair, random, TWO issuer threads / accumulators", n_sm, 256, 1, 1, 0, 1, 4);
  run<true>("pair, random, TWO issuers, N = 128", n_sm, 128, 1, 1, 0, 1, 4);
  run<true>("pair, random, LEAN loop (elect, uniform descriptors)", n_sm, 256, 1, 1, 0, 1, 32);
  run<true>("pair, random, LEAN loop, K-major", n_sm, 256, 0, 1, 0, 1, 32);
  run<true>("pair, random, LEAN loop, N = 128", n_sm, 128, 1, 1, 0, 1, 32);
  run<false>("single CTA, random, LEAN loop", n_sm, 256, 1, 1, 0, 1, 32);
  run<true>("pair, random, ONE issuer, two accumulators in turn", n_sm, 256, 1, 1, 0, 1, 8);
  run<true>("pair, random, ONE issuer, two accumulators, K-major", n_sm, 256, 0, 1, 0, 1, 8);
  run<true>("pair, random, ONE issuer, two accumulators, N = 128", n_sm, 128, 1, 1, 0, 1, 8);
  run<true>("pair, random, TWO issuers, SAME accumulator", n_sm, 256, 1, 1, 0, 1, 20);
  run<false>("single CTA, random, TWO issuers", n_sm, 256, 1, 1, 0, 1, 4);
  run<true>("pair, random, fence::after_thread_sync per K block", n_sm, 256, 1, 1, 0, 1, 1);
  run<true>("pair, random, producer relay (empty -> full)", n_sm, 256, 1, 1, 0, 1, 2);
  run<true>("pair, random, relay + fence", n_sm, 256, 1, 1, 0, 1, 3);
  run<true>("pair, random, relay + fence + polling warps", n_sm, 256, 1, 1, 1, 1, 3);
  run<true>("pair + 8 warps polling an mbarrier (all lanes)", n_sm, 256, 1, 1, 1);
  run<true>("pair + 8 warps polling (one lane each)", n_sm, 256, 1, 1, 2);
  run<true>("pair + 8 warps polling (one lane, nanosleep)", n_sm, 256, 1, 1, 3);
  run<true>("pair + 8 warps polling an mbarrier (all lanes)", n_sm, 256, 0, 1, 1);
  run<true>("pair + 8 warps polling (one lane, nanosleep)", n_sm, 256, 0, 1, 3);
  return 0;
}
