// tcgen05 / TMEM / TMA GEMM for sm_100a: bf16 operands, fp32 accumulation in tensor memory,
// fused epilogue functors (epilogues.cuh, epilogues_tc.cuh).  Persistent, warp-specialised:
//   warp 0    TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1    MMA issuer     (one elected lane issues tcgen05.mma, tcgen05.commit frees ring slots)
//   warp 2    TMEM allocator (2 accumulator stages x BN columns)
//   warp 4-11 epilogue       (tcgen05.ld 32x32b -> registers -> fused math -> smem boxes -> TMA),
//                             overlapped with the next tile's mainloop through the second TMEM stage
// kPair: two CTAs of a cluster (one TPC) run ONE 256 x BN tile with tcgen05.mma.cta_group::2: each CTA
// stages its own 128 rows of A and only HALF of B, the leader CTA issues the MMAs for both, every CTA
// drains its own 128 x BN accumulator.  Halving the B footprint is what pays for deep operand rings and
// double-buffered epilogue boxes inside 227 KB.
//
//   kNT == false : C[m][n] = sum_k A[m][k] B[n][k]   A:[M,K] B:[N,K], both K-major   (fwd, dgrad)
//   kNT == true  : C[m][n] = sum_p A[p][m] B[p][n]   A:[P,M] B:[P,N], both MN-major  (wgrad,
//                  reduction over points split across CTAs, fp32 atomics in the epilogue)
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace bn {
namespace tc {

constexpr int kBM = 128;          // rows of the output tile (UMMA M, cta_group::1)
constexpr int kBK = 64;           // bf16 elements per smem row = 128 bytes = one swizzle atom row
constexpr int kThreads = 384;      // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue (two per TMEM quadrant)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must end in a trapped launch, never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("bn::tc mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
// Programmatic dependent launch (opt-in, BN_PDL=1): a tcgen05 kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may become resident and run its prologue (barrier
// init, TMEM allocation, tensor-map prefetch) while the previous kernel of the stream drains; pdl_wait() blocks until
// that kernel has completed and its writes are visible, and must precede the first access to global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative smem offset of every CTA in `mask`, and
// each destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// ---- smem -> global through the TMA (epilogue staging) ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// element-wise fp32 add into global memory (L2 atomics on whole lines instead of scattered red.global)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// L2 eviction-priority hints of the TMA (createpolicy encodings): the packed weights (2-3 MB, re-read by every block of every
// CTA) must survive in the 126 MB L2 next to GBs of activations that stream through it once — without the hints the
// activation stores of the fused trunk evicted the weight tiles and the MMA issuer waited ~4.5 k cycles per layer for them
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // streaming: activations written / read once per kernel
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;    // resident: weights
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
// BN_NO_L2_HINTS=1 (A/B timing aid, read per launch): every transfer with the default policy
inline uint64_t pol_weights() { return getenv("BN_NO_L2_HINTS") ? kEvictNormal : kEvictLast; }
inline uint64_t pol_stream() { return getenv("BN_NO_L2_HINTS") ? kEvictNormal : kEvictFirst; }
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA) that reads them next
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void sts128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool kPair> __device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  if constexpr (kPair) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <bool kPair> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared::cluster address of `local` (an address inside this CTA's shared window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// default semantics (release at CTA scope): the only state handed over is the TMEM stage, ordered by
// tcgen05.fence::before_thread_sync; an explicit .release.cluster costs a MEMBAR.ALL + ERRBAR per arrive
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair load: lands in THIS CTA's smem, signals the mbarrier at shared::cluster address `bar_cluster`
// (the leader CTA's "full" barrier collects the bytes of both CTAs)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// multicast pair load (clusters of two CTA pairs): the box lands at the same CTA-relative offset of every CTA in `mask`, and
// each destination's bytes are counted on the barrier at `bar_leader`'s offset in the destination's PAIR LEADER (`bar_leader` =
// shared::cluster address of the issuing CTA's own pair leader's barrier, i.e. its own address with the peer bit cleared, which is
// what CUTLASS's SM100_TMA_2SM_LOAD_MULTICAST passes)
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_leader), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
// one lane of a converged warp (the warp stays converged around the elected block: descriptors and addresses are warp-uniform
// values, which lets the compiler keep them in uniform registers instead of voting them over on every tcgen05 instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit of cta_group::2 MMAs: arrives on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// ... on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair_mask(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// commit that arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue the load, do other work, then tmem_wait_ld() before touching r[]
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptors (SWIZZLE_128B, version 1 = Blackwell).  Units of 16 bytes.
//   K-major  : 8-row groups 1024 B apart (SBO); LBO unused (=1)
//   MN-major : 64-element (128 B) column chunks `lbo_bytes` apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, M = 128 (one CTA) or 256 (CTA pair), N=BN, majorness per operand
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// L2 prefetch of one box (no shared memory, no barrier): the DRAM access happens `distance` K blocks before the load
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
// NT only: how many K blocks ahead of its loads the producer prefetches the operand boxes into L2.  The operand ring
// (5 stages of 32 KB) holds about 2 000 cycles of work, less than the latency of a loaded HBM; the prefetch moves the
// DRAM latency out of the ring, which then only has to cover an L2 hit.  MEASURED on B200 (profiles/r02m_ab_wgrad.txt): the
// opposite — 752 us of weight gradients per step without, 774 / 793 / 849 / 951 us at distance 4 / 8 / 16 / 32.  The kernel is
// not bound by its loads at all: with every load after the first ring pass, the bias MMAs and the epilogue stores removed it
// takes the same time (profiles/r02o_ab_wgrad.txt), and the trace (scripts/trace_wgrad.py, profiles/r02p_trace_wgrad.txt) shows
// the issuer thread inside the MMA issue for 897 cycles per K block = 224 cycles per 256x256x16 pair MMA (128 nominal, ~180 at
// the cuBLAS rate this pool measures) while it waits ~100 cycles per K block for operands.  Off; BN_NT_PREFETCH=<k> enables.
inline int nt_prefetch_distance() {
  const char* e = getenv("BN_NT_PREFETCH");       // read per launch: A/B runs inside one process
  return e ? atoi(e) : 0;
}

struct Work {
  int m_tiles, n_tiles, splits;   // 128-row tiles, BN-column tiles; splits only for kNT
  int kb_total, kb_per_split;     // k-blocks (of kBK) in the reduction dimension
  int n_cols;                     // real number of output columns (a multiple of the epilogue unit)
  int reverse;                    // TN only: walk the row tiles from the last to the first (see tile_reverse())
  uint64_t pol_a, pol_b;          // TN only: L2 eviction policies of the A (activation) and B (weight) tiles
  int pf;                         // NT only: L2 prefetch distance in K blocks (0 = off)
  long long* trace;               // NT pair only: [0] kernel entry, [1] setup done, [2] first MMA issued, [3] last MMA issued, [4] cycles the
                                  // issuer waited for operands, [5] K blocks, [6] accumulator complete (epilogue), [7] epilogue done,
                                  // [8] cycles the producer waited for free slots
};
// diagnostics (bn_debug_chain_trace with BN_NT_TRACE set): clock64() stamps of CTA pair 0 of the 512 x 512 weight-gradient launches
inline long long*& nt_trace() { static long long* p = nullptr; return p; }


// Tile order of the NEXT tn launches on this thread (experiment knob).  Idea: in a chain of GEMMs in which each one consumes
// the [P, F] tensor its predecessor has just written, let the consumer start with the rows the producer wrote last, which
// should still be in the 126 MB L2 (a [131072, 512] bf16 tensor is 134 MB).  Measured on B200: slower (see mlp.cu), off by default.
inline bool& tile_reverse() { static thread_local bool r = false; return r; }

// Epilogue kinds.  An epilogue functor `Epi` declares `static constexpr int kMode`:
//   EPI_DIRECT      : apply<32>(row, col0, acc[32]) writes global memory itself (debug / skinny paths)
//   EPI_TMA_BF16    : the tile leaves (and its element-wise operands enter) through shared memory and
//                     the TMA in units of 32 rows x 64 bf16 columns = one 128B-swizzled 4 KB box per TMEM
//                     quadrant; kIn input streams, kOut output streams.  A thread owns one row of the unit,
//                     so its 16-byte chunks land conflict-free in the swizzled box and every global access
//                     is a full-line TMA transaction instead of 32 scattered sectors.  Boxes are double
//                     buffered (operands prefetched two units ahead, stores drained one unit behind)
//                     whenever kIn + kOut <= 3.
//   EPI_TMA_RED_F32 : units of 32 rows x 32 fp32 columns leave through cp.reduce.async.bulk (add).
// TMA functors provide
//   compute(row, col0, acc[32], in[kIn][16], out[kOut][16])   (EPI_TMA_BF16: the two warps of a quadrant
//        each own 32 of the unit's 64 columns: acc = columns [col0, col0+32), in/out packed bf16x2)
//   compute(row, col0, acc[32], out[32])                                       (EPI_TMA_RED_F32)
//   load(i, slot, bar, col0, row0) / store(o, slot, col0, row0)                issue the TMA transfers
enum { EPI_DIRECT = 0, EPI_TMA_BF16 = 1, EPI_TMA_RED_F32 = 2 };

constexpr int kSlotBytes = 4096;   // 32 rows x 128 bytes
// NT functors with `kBias`: the kernel also accumulates sum_p A[p][m] (the bias gradient that belongs to
// the weight gradient A^T B) with one extra N=16 MMA per K step against a tile of ones, on the CTAs of the
// first column tile; epi.bias (nullable, fp32 [M]) receives it through red.global.
template <class Epi, class = void> struct epi_has_bias { static constexpr bool value = false; };
template <class Epi> struct epi_has_bias<Epi, decltype((void)Epi::kBias)> { static constexpr bool value = Epi::kBias; };
constexpr int kOnesBytes = 8192;   // [64 K rows x 64 MN] bf16 ones, the B operand of the bias MMA
template <class Epi> __host__ __device__ constexpr int epi_nbuf() {
  // weight gradients drain once per CTA after the whole reduction: nothing to overlap, the smem is worth
  // more as an extra operand stage
  return Epi::kMode == EPI_TMA_RED_F32 ? 1 : (Epi::kIn + Epi::kOut <= 3 ? 2 : 1);
}
template <class Epi> __host__ __device__ constexpr int epi_slots() {        // 4 KB boxes per TMEM quadrant (= per pair of epilogue warps)
  return Epi::kMode == EPI_DIRECT ? 0 : (Epi::kMode == EPI_TMA_RED_F32 ? 2 * epi_nbuf<Epi>() : (Epi::kIn + Epi::kOut) * epi_nbuf<Epi>());
}
template <class Epi> __host__ __device__ constexpr int epi_smem() {
  return 4 * epi_slots<Epi>() * kSlotBytes + (epi_has_bias<Epi>::value ? kOnesBytes : 0);
}

constexpr int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA
template <int BN, bool kPair> __host__ __device__ constexpr int stage_bytes() { return kBM * kBK * 2 + (kPair ? BN / 2 : BN) * kBK * 2; }
template <int BN, bool kPair, class Epi> __host__ __device__ constexpr int pick_stages() {
  constexpr int avail = kMaxSmem - 1024 - 512 - epi_smem<Epi>();
  return avail / stage_bytes<BN, kPair>() > 8 ? 8 : avail / stage_bytes<BN, kPair>();
}
template <int BN, bool kPair, class Epi> __host__ __device__ constexpr int smem_bytes() {
  return pick_stages<BN, kPair, Epi>() * stage_bytes<BN, kPair>() + epi_smem<Epi>() + 512 + 1024;
}

// kMC (NT pair kernels only): clusters of FOUR CTAs = two pairs that work on the same column tile and K range of two adjacent
// 256-row groups.  They need the same B tile: each CTA loads ONE of the two 64-column boxes of its half and multicasts it to its
// counterpart in the other pair, so the B operand crosses the L2 -> SM fabric once per cluster instead of once per pair (the
// weight-gradient GEMM moves 537 MB per launch through L2 and is bound by exactly that once the issue loop is lean).
template <int BN, int STAGES, bool kNT, bool kPair, class Epi, bool kMC = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Work wk,
               const __grid_constant__ Epi epi) {
  static_assert(!kMC || (kNT && kPair && BN == 256), "multicast variant: weight-gradient pair kernel with 256-column tiles");
  constexpr int CL = kMC ? 4 : (kPair ? 2 : 1);
  constexpr int BNL = kPair ? BN / 2 : BN;                     // B rows (output columns) staged by this CTA
  constexpr uint32_t A_BYTES = kBM * kBK * 2;
  constexpr uint32_t B_BYTES = BNL * kBK * 2;
  constexpr bool kBias = epi_has_bias<Epi>::value;
  // split-P weight gradients run one item per CTA: a single accumulator stage (+ 32 columns for the bias MMA)
  constexpr int ACC = kNT ? 1 : 2;
  constexpr int TMEM_NEED = ACC * BN + (kBias ? 32 : 0);
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static_assert(TMEM_NEED <= 512, "accumulators exceed tensor memory");
  constexpr int NSLOT = epi_slots<Epi>();
  constexpr int NBUF = epi_nbuf<Epi>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sEpi = sB + STAGES * B_BYTES;                       // 4 quadrants x NSLOT boxes of 4 KB
  uint8_t* sOnes = sEpi + 4 * NSLOT * kSlotBytes;              // kBias only
  uint64_t* full = reinterpret_cast<uint64_t*>(sOnes + (kBias ? kOnesBytes : 0));
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* ibar = tempty + 2;                                 // [quadrant][buffer]: operand boxes landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ibar + 8);

  pdl_launch_dependents();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  [[maybe_unused]] const bool tr = kNT && kPair && wk.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) wk.trace[0] = clock64();
  const int m_groups = (wk.m_tiles + CL - 1) / CL;                 // row tiles are handed out CL at a time
  const int n_items = m_groups * wk.n_tiles * wk.splits;
  const int t_count = m_groups * wk.n_tiles;
  // item -> (row group, column tile): every role (producer, MMA issuer, epilogue warps) maps through this one function
  auto item_tile = [&](int it) {
    const int t = it % t_count;
    if (!kNT && wk.reverse) return (m_groups - 1 - t / wk.n_tiles) * wk.n_tiles + t % wk.n_tiles;
    return t;
  };
  const int crank = kPair ? (int)cluster_ctarank() : 0;          // rank in the cluster (0..CL-1) = 128-row tile of the row group
  [[maybe_unused]] const int prank = crank & 1;                   // rank in the CTA pair
  [[maybe_unused]] const int lead = crank & ~1;                   // cluster rank of this pair's leader
  [[maybe_unused]] const uint16_t pair_mask = (uint16_t)(3u << lead);
  const int item0 = kPair ? (int)cluster_id_x() : (int)blockIdx.x;
  const int item_step = kPair ? (int)cluster_nctaid_x() : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    // empty: one commit per pair that reads what lands in this CTA's slot (kMC: both pairs write into each other's slots)
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMC ? 2 : 1); }
    // tempty: one arrival per epilogue warp; in a pair the leader's barrier collects both CTAs' warps
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kPair ? 16 : 8); }
    for (int s = 0; s < 8; ++s) mbar_init(&ibar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<kPair>(tmem_slot, TMEM_COLS);
  if constexpr (kBias) {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;   // bf16 1.0 x2
    fence_async_smem();                    // read by the tensor core through the async proxy
  }
  fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();           // the peer's barriers are initialised before anyone signals them
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                              // the previous kernel's results are visible from here on
  if (tr && threadIdx.x == 0) wk.trace[1] = clock64();

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = item0; it < n_items; it += item_step) {
        const int split = it / (m_groups * wk.n_tiles);
        const int t = item_tile(it);
        const int m_blk = (t / wk.n_tiles) * CL + crank, n_blk = t % wk.n_tiles;
        const int ncol0 = n_blk * BN + prank * BNL;              // first B row / output column staged by this CTA
        const int kb0 = split * wk.kb_per_split;
        const int kb1 = min(wk.kb_total, kb0 + wk.kb_per_split);
        [[maybe_unused]] auto prefetch_kb = [&](int kp) {
#pragma unroll
          for (int c = 0; c < kBM / 64; ++c) tma_prefetch_2d(&tmA, m_blk * kBM + c * 64, kp * kBK);
#pragma unroll
          for (int c = 0; c < BNL / 64; ++c) tma_prefetch_2d(&tmB, ncol0 + c * 64, kp * kBK);
        };
        if constexpr (kNT) {
          if (wk.pf > 0)
            for (int kp = kb0 + STAGES; kp < min(kb1, kb0 + wk.pf); ++kp) prefetch_kb(kp);
        }
        long long t_slot = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          if constexpr (kNT) { if (wk.pf > 0 && kb + wk.pf < kb1) prefetch_kb(kb + wk.pf); }
          const long long ts0 = tr ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);          // the MMAs that read this slot have retired
          if (tr) t_slot += clock64() - ts0;
          uint8_t* a = sA + stage * A_BYTES;
          uint8_t* b = sB + stage * B_BYTES;
          if constexpr (kPair) {
            // the leader's barrier expects the bytes of both CTAs; the peer only issues its loads
            if (prank == 0) mbar_expect_tx(&full[stage], 2 * (A_BYTES + B_BYTES));
            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), lead);
            if constexpr (kMC) {
#pragma unroll
              for (int c = 0; c < kBM / 64; ++c) tma_load_2d_pair(a + c * 8192, &tmA, bar, m_blk * kBM + c * 64, kb * kBK);
              // this CTA's half of B is two 64-column boxes: pair 0 loads the first, pair 1 the second, each for both pairs
              const int c = crank >> 1;
              tma_load_2d_pair_mc(b + c * 8192, &tmB, bar, ncol0 + c * 64, kb * kBK, (uint16_t)(5u << prank));
            } else if (!kNT) {               // A: activations streaming through once; B: weights, re-read by every row tile
              tma_load_2d_pair_hint(a, &tmA, bar, kb * kBK, m_blk * kBM, wk.pol_a);
              tma_load_2d_pair_hint(b, &tmB, bar, kb * kBK, ncol0, wk.pol_b);
            } else {
#pragma unroll
              for (int c = 0; c < kBM / 64; ++c) tma_load_2d_pair(a + c * 8192, &tmA, bar, m_blk * kBM + c * 64, kb * kBK);
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d_pair(b + c * 8192, &tmB, bar, ncol0 + c * 64, kb * kBK);
            }
          } else {
            mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
            if (!kNT) {
              tma_load_2d_hint(a, &tmA, &full[stage], kb * kBK, m_blk * kBM, wk.pol_a);
              tma_load_2d_hint(b, &tmB, &full[stage], kb * kBK, ncol0, wk.pol_b);
            } else {
              // boxes of 64 (contiguous MN) x 64 (reduction rows); one box per 64 output rows/cols
#pragma unroll
              for (int c = 0; c < kBM / 64; ++c) tma_load_2d(a + c * 8192, &tmA, &full[stage], m_blk * kBM + c * 64, kb * kBK);
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d(b + c * 8192, &tmB, &full[stage], ncol0 + c * 64, kb * kBK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tr) wk.trace[8] = t_slot;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of a pair only) =====================
    // LEAN issue loop (scripts/mma_rate.cu, profiles/r02q_mma_rate.txt): the loop runs on the WHOLE warp with uniform control flow
    // and one elected lane issues; shared-memory descriptors are a base plus constant increments.  The first version ran on one
    // lane and rebuilt both descriptors from the address for every instruction: ~40 scalar instructions (shifts, masks, R2UR, a
    // vote loop around every tcgen05 instruction) between two MMAs = 186-224 cycles per 256x256x16 MMA against the pipe's 128.
    if (prank == 0) {
      constexpr uint32_t idesc = make_idesc(kPair ? 256 : 128, BN, kNT);
      constexpr uint32_t k_inc = (kNT ? 2048u : 32u) >> 4;             // descriptor address units (16 bytes) per K slice of 16
      const uint64_t a_desc0 = kNT ? make_desc(smem_u32(sA), 8192, 1024) : make_desc(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = kNT ? make_desc(smem_u32(sB), 8192, 1024) : make_desc(smem_u32(sB), 16, 1024);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      [[maybe_unused]] constexpr uint32_t idesc_bias = make_idesc(kPair ? 256 : 128, 16, kNT);
      [[maybe_unused]] const uint64_t ones_desc0 = make_desc(smem_u32(sOnes), 8192, 1024);
      const bool tr1 = tr && lane == 0;
      for (int it = item0; it < n_items; it += item_step) {
        const int split = it / (m_groups * wk.n_tiles);
        const int kb0 = split * wk.kb_per_split;
        const int kb1 = min(wk.kb_total, kb0 + wk.kb_per_split);
        // the bias MMAs are dealt round-robin over the column tiles of a row block (K block kb goes to
        // column tile kb % n_tiles), so that no CTA carries all of the extra A-operand reads
        [[maybe_unused]] const int n_blk = item_tile(it) % wk.n_tiles;
        [[maybe_unused]] bool bias_on = false, bias_started = false;
        if constexpr (kBias) bias_on = epi.bias != nullptr;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        [[maybe_unused]] int rot = kb0 % wk.n_tiles;               // kb % n_tiles without a division per K block
        long long t_wait = 0;
        if (tr1) wk.trace[2] = clock64();
        for (int kb = kb0; kb < kb1; ++kb) {
          [[maybe_unused]] const bool bias_now = bias_on && rot == n_blk;
          if (++rot == wk.n_tiles) rot = 0;
          const long long tw0 = tr1 ? clock64() : 0;
          mbar_wait(&full[stage], phase);
          if (tr1) t_wait += clock64() - tw0;
          fence_after_sync();
          const uint64_t da0 = a_desc0 + (uint64_t)(stage * (int)(A_BYTES >> 4));
          const uint64_t db0 = b_desc0 + (uint64_t)(stage * (int)(B_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t da = da0 + k * k_inc, db = db0 + k * k_inc;
              if constexpr (kPair) umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              if constexpr (kBias) {
                if (bias_now) {     // D'[m][0..15] += sum_k A[k][m] * 1
                  const uint64_t dones = ones_desc0 + k * k_inc;
                  if constexpr (kPair) umma_bf16_pair(tmem_base + ACC * BN, da, dones, idesc_bias, (bias_started || k > 0) ? 1u : 0u);
                  else umma_bf16(tmem_base + ACC * BN, da, dones, idesc_bias, (bias_started || k > 0) ? 1u : 0u);
                }
              }
            }
            if constexpr (kMC) umma_commit_pair_mask(&empty[stage], (uint16_t)0xF);       // both pairs' producers write this slot
            else if constexpr (kPair) umma_commit_pair(&empty[stage]);
            else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if constexpr (kBias) { if (bias_now) bias_started = true; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) { if constexpr (kPair) umma_commit_pair_mask(&tfull[acc], pair_mask); else umma_commit(&tfull[acc]); }
        __syncwarp();
        if (tr1) { wk.trace[3] = clock64(); wk.trace[4] = t_wait; wk.trace[5] = kb1 - kb0; }
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // warps w and w+4 share TMEM lane quadrant q = w % 4 (a warp may only touch lanes 32(w%4)..+31) and
    // split the columns, so that two warps per scheduler hide each other's TMEM / MUFU / smem latency
    const int q = warp & 3;
    const int hsel = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    // this warp's share of the accumulator is in registers: hand the TMEM stage back to the MMA issuer
    auto release_acc = [&](int a) {
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[a]), lead));
        else mbar_arrive(&tempty[a]);
      }
    };
    if constexpr (Epi::kMode == EPI_DIRECT) {
      for (int it = item0; it < n_items; it += item_step) {
        const int t = item_tile(it);
        const int m_blk = (t / wk.n_tiles) * CL + crank, n_blk = t % wk.n_tiles;
        mbar_wait(&tfull[acc], acc_phase);
        fence_after_sync();
        const int row = m_blk * kBM + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = hsel; c < BN / 32; c += 2) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          epi.template apply<32>(row, n_blk * BN + c * 32, v);
        }
        release_acc(acc);
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    } else if constexpr (Epi::kMode == EPI_TMA_BF16) {
      constexpr int KI = Epi::kIn, KO = Epi::kOut;
      uint8_t* slots = sEpi + q * (NSLOT * kSlotBytes);        // [in stream][buf] then [out stream][buf]
      uint64_t* ib = &ibar[q * 2];
      const uint32_t row_off = lane * 128, swz = (lane & 7) << 4;
      const bool leader = hsel == 0 && lane == 0;          // issues the quadrant's TMA traffic
      auto tile_of = [&](int it, int& m_blk, int& n_blk, int& n_units) {
        const int t = item_tile(it);
        m_blk = (t / wk.n_tiles) * CL + crank; n_blk = t % wk.n_tiles;
        const int left = wk.n_cols - n_blk * BN;
        n_units = left >= BN ? BN / 64 : (left + 63) / 64;
      };
      // operand prefetch cursor (leader lane only): runs NBUF units ahead of the unit being computed
      int pf_it = item0, pf_u = 0;
      auto prefetch = [&](int buf) {
        if constexpr (KI > 0) {
          if (pf_it >= n_items) return;
          int mb, nb, nu; tile_of(pf_it, mb, nb, nu);
          mbar_expect_tx(&ib[buf], KI * kSlotBytes);
#pragma unroll
          for (int i = 0; i < KI; ++i)
            epi.load(i, slots + (i * NBUF + buf) * kSlotBytes, &ib[buf], nb * BN + pf_u * 64, mb * kBM + q * 32);
          if (++pf_u >= nu) { pf_u = 0; pf_it += item_step; }
        }
      };
      if (leader) { for (int b = 0; b < NBUF; ++b) prefetch(b); }
      uint32_t g = 0;                                      // units processed so far by this quadrant
      for (int it = item0; it < n_items; it += item_step) {
        int m_blk, n_blk, n_units; tile_of(it, m_blk, n_blk, n_units);
        mbar_wait(&tfull[acc], acc_phase);
        fence_after_sync();
        const int row0 = m_blk * kBM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + hsel * 32;
#pragma unroll 1
        for (int u = 0; u < n_units; ++u, ++g) {
          const int col0 = n_blk * BN + u * 64;
          const int buf = (NBUF == 2) ? (int)(g & 1) : 0;
          [[maybe_unused]] uint32_t in[KI > 0 ? KI : 1][16];
          if constexpr (KI > 0) {
            mbar_wait(&ib[buf], (g / NBUF) & 1);
#pragma unroll
            for (int i = 0; i < KI; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 v = lds128(slots + (i * NBUF + buf) * kSlotBytes + row_off + (((hsel * 4 + j) << 4) ^ swz));
                in[i][4 * j] = v.x; in[i][4 * j + 1] = v.y; in[i][4 * j + 2] = v.z; in[i][4 * j + 3] = v.w;
              }
          }
          float v[32];
          tmem_ld32(taddr + u * 64, v);
          if (u == n_units - 1) release_acc(acc);
          uint32_t out[KO][16];
          epi.compute(row0 + lane, col0 + hsel * 32, v, in, out);
          // the boxes this unit writes were last used NBUF units ago: their TMA stores have read them out
          if (leader) { if (NBUF == 2) bulk_wait_read1(); else bulk_wait_read0(); }
          named_bar_sync(1 + q, 64);                 // ... and both warps hold this unit's operands in registers
          if (leader) prefetch(buf);                 // refill the operand boxes just consumed
#pragma unroll
          for (int o = 0; o < KO; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(slots + ((KI + o) * NBUF + buf) * kSlotBytes + row_off + (((hsel * 4 + j) << 4) ^ swz),
                     out[o][4 * j], out[o][4 * j + 1], out[o][4 * j + 2], out[o][4 * j + 3]);
          fence_async_smem();
          named_bar_sync(1 + q, 64);                 // the unit's boxes are complete
          if (leader) {
#pragma unroll
            for (int o = 0; o < KO; ++o) epi.store(o, slots + ((KI + o) * NBUF + buf) * kSlotBytes, col0, row0);
            bulk_commit();
          }
        }
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (leader) bulk_wait0();                   // all boxes written before the CTA may exit
    } else {
      // fp32 reduce-add units of 32 columns: the two warps of a quadrant alternate units, two boxes each
      uint8_t* slot = sEpi + (q * NSLOT + hsel * NBUF) * kSlotBytes;
      const uint32_t row_off = lane * 128, swz = (lane & 7) << 4;
      uint32_t g = 0;
      for (int it = item0; it < n_items; it += item_step) {
        const int t = item_tile(it);
        const int m_blk = (t / wk.n_tiles) * CL + crank, n_blk = t % wk.n_tiles;
        const int left = wk.n_cols - n_blk * BN;
        const int n_units = left >= BN ? BN / 32 : (left + 31) / 32;
        mbar_wait(&tfull[acc], acc_phase);
        fence_after_sync();
        if (tr && warp == 4 && lane == 0) wk.trace[6] = clock64();
        const int row0 = m_blk * kBM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
        if constexpr (kBias) {
          // row sums of A^T over this CTA's share of the K blocks (none if its range holds no block of its own)
          const int split = it / (m_groups * wk.n_tiles);
          const int kb0 = split * wk.kb_per_split, kb1 = min(wk.kb_total, kb0 + wk.kb_per_split);
          const int first = kb0 + ((n_blk - kb0 % wk.n_tiles) + wk.n_tiles) % wk.n_tiles;
          if (epi.bias != nullptr && first < kb1 && hsel == 1) {
            float b[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ACC * BN, b);
            if (row0 + lane < epi.M) atomicAdd(epi.bias + row0 + lane, b[0]);
          }
        }
        int last = n_units - 1; if ((last & 1) != hsel) --last;      // this warp's last unit (may be < 0)
        if (last < 0) release_acc(acc);
#pragma unroll 1
        for (int u = hsel; u < n_units; u += 2, ++g) {
          const int col0 = n_blk * BN + u * 32;
          uint8_t* sl = slot + (NBUF == 2 ? (g & 1) : 0) * kSlotBytes;
          float v[32];
          tmem_ld32(taddr + u * 32, v);
          if (u == last) release_acc(acc);
          uint32_t out[32];
          epi.compute(row0 + lane, col0, v, out);
          if (lane == 0) { if (NBUF == 2) bulk_wait_read1(); else bulk_wait_read0(); }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(sl + row_off + ((j << 4) ^ swz), out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { epi.store(0, sl, col0, row0); bulk_commit(); }
        }
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait0();
      if (tr && warp == 4 && lane == 0) wk.trace[7] = clock64();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();           // no CTA leaves while the peer can still signal / read its smem
  if (warp == 2) { fence_after_sync(); tmem_dealloc<kPair>(tmem_base, TMEM_COLS); }
}

// ---- host side: tensor maps ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

// 2-D bf16 tensor [rows, cols] with row pitch `ld` elements; box = box_cols x box_rows, 128B swizzle
int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                  int box_cols, int box_rows);
// same with a 64-byte swizzle (boxes of 32 bf16 columns)
int make_map_bf16_sw(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                     int box_cols, int box_rows, int swizzle_bytes);
// 2-D fp32 tensor, same conventions (box_cols * 4 bytes must be 128)
int make_map_f32(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                 int box_cols, int box_rows);
// true when (base, ld) of an element type of `es` bytes can be described by a tensor map
inline bool tma_ok(const void* base, long long ld, int es) {
  return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && ((ld * es) & 15) == 0;
}

// BN_PDL=1 launches the tcgen05 kernels with programmatic stream serialisation.  Off by default: measured on B200
// (r01e) the training step is power-capped, and overlapping prologues with the previous kernel's tail changed
// nothing (2.51-2.54 ms with, 2.48-2.53 ms without); the device side (pdl_wait) is a no-op for a normal launch.
inline bool pdl_enabled() { static const bool on = getenv("BN_PDL") != nullptr; return on; }

// persistent launch: one CTA per SM (rounded down to whole pairs), cluster dims (CL,1,1)
template <int CL, class Kern, class Epi>
int launch_kernel(Kern kern, int smem, int num_sms, int items, const CUtensorMap& ma, const CUtensorMap& mb, const Work& wk,
                  const Epi& epi, cudaStream_t s, const char* name) {
  BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  const int clusters = min(items, num_sms / CL);
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  BN_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, wk, epi));
  return after_launch(name);
}

// A:[M,K] (ld lda), B:[N,K]: C = A B^T through `epi`
template <int BN, class Epi>
int launch_tn(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb,
              long long M, int N, int K, const Epi& epi, int num_sms, cudaStream_t s) {
  if (K % kBK) { set_error("tc::launch_tn: K=%d not a multiple of 64", K); return BN_ERR_ARG; }
  if (Epi::kMode != EPI_DIRECT && N % 64) { set_error("tc::launch_tn: N=%d not a multiple of 64", N); return BN_ERR_ARG; }
  CUtensorMap ma, mb;
  Work wk{(int)ceil_div_ll(M, kBM), ceil_div(N, BN), 1, K / kBK, K / kBK, N, tile_reverse() ? 1 : 0, pol_stream(), pol_weights(), 0, nullptr};
  if (int rc = make_map_bf16(&ma, A, M, K, lda, kBK, kBM)) return rc;
  if constexpr (BN >= 128) {
    if (wk.m_tiles >= 4) {               // CTA pairs: every CTA stages BN/2 rows of the B tile
      constexpr int STAGES = pick_stages<BN, true, Epi>();
      static_assert(STAGES >= 2, "epilogue staging leaves no room for the operand ring");
      if (int rc = make_map_bf16(&mb, B, N, K, ldb, kBK, BN / 2)) return rc;
      return launch_kernel<2>(gemm_tc_kernel<BN, STAGES, false, true, Epi>, smem_bytes<BN, true, Epi>(), num_sms,
                              ceil_div(wk.m_tiles, 2) * wk.n_tiles, ma, mb, wk, epi, s, "gemm_tc_kernel<tn,pair>");
    }
  }
  constexpr int STAGES = pick_stages<BN, false, Epi>();
  static_assert(STAGES >= 2, "epilogue staging leaves no room for the operand ring");
  if (int rc = make_map_bf16(&mb, B, N, K, ldb, kBK, BN)) return rc;
  return launch_kernel<1>(gemm_tc_kernel<BN, STAGES, false, false, Epi>, smem_bytes<BN, false, Epi>(), num_sms,
                          wk.m_tiles * wk.n_tiles, ma, mb, wk, epi, s, "gemm_tc_kernel<tn>");
}

// A:[P,Mo] (ld lda), B:[P,No]: C[Mo,No] = A^T B through `epi` (atomic accumulate), split over P
template <int BN, class Epi>
int launch_nt(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* B, long long ldb,
              int Mo, int No, long long P, const Epi& epi, int num_sms, cudaStream_t s) {
  if (Mo % 64) { set_error("tc::launch_nt: Mo=%d not a multiple of 64", Mo); return BN_ERR_ARG; }
  CUtensorMap ma, mb;
  if (int rc = make_map_bf16(&ma, A, P, Mo, lda, 64, kBK)) return rc;
  if (int rc = make_map_bf16(&mb, B, P, No, ldb, 64, kBK)) return rc;
  Work wk;
  wk.reverse = 0; wk.pol_a = wk.pol_b = kEvictNormal;
  wk.pf = nt_prefetch_distance();
  wk.trace = (Mo == 512 && No == 512) ? nt_trace() : nullptr;
  wk.m_tiles = ceil_div(Mo, kBM); wk.n_tiles = ceil_div(No, BN);
  wk.kb_total = (int)ceil_div_ll(P, kBK);
  wk.n_cols = No;
  const int tiles = wk.m_tiles * wk.n_tiles;
  const bool pair = BN >= 128 && wk.m_tiles % 2 == 0;
  // one work item per CTA (pair): the split count is rounded DOWN so that tiles x splits never
  // exceeds the resident grid — a second, nearly empty round would double the kernel's duration
  const int slots_avail = pair ? (num_sms / 2) / (tiles / 2) : num_sms / tiles;
  int splits = max(1, min(slots_avail, ceil_div(wk.kb_total, 8)));
  wk.kb_per_split = ceil_div(wk.kb_total, splits);
  wk.splits = ceil_div(wk.kb_total, wk.kb_per_split);
  if constexpr (BN == 256) {
    // clusters of two pairs with the B tile multicast between them (kMC): BN_NT_MC=1, read per call (A/B knob)
    if (pair && wk.m_tiles % 4 == 0 && getenv("BN_NT_MC") != nullptr) {
      auto kern = gemm_tc_kernel<BN, pick_stages<BN, true, Epi>(), true, true, Epi, true>;
      constexpr int smem = smem_bytes<BN, true, Epi>();
      static int max_clusters = -1;                  // resident 4-CTA clusters of this kernel (GPC granularity: fewer than SMs / 4)
      if (max_clusters < 0) {
        BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(num_sms / 4 * 4); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nc = 0;
        BN_CUDA(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));
        max_clusters = nc;
      }
      const int cluster_tiles = (wk.m_tiles / 4) * wk.n_tiles;
      if (max_clusters >= cluster_tiles) {
        splits = max(1, min(max_clusters / cluster_tiles, ceil_div(wk.kb_total, 8)));
        wk.kb_per_split = ceil_div(wk.kb_total, splits);
        wk.splits = ceil_div(wk.kb_total, wk.kb_per_split);
        return launch_kernel<4>(kern, smem, 4 * max_clusters, cluster_tiles * wk.splits, ma, mb, wk, epi, s, "gemm_tc_kernel<nt,pair,mc>");
      }
    }
  }
  if constexpr (BN >= 128) {
    if (pair)
      return launch_kernel<2>(gemm_tc_kernel<BN, pick_stages<BN, true, Epi>(), true, true, Epi>, smem_bytes<BN, true, Epi>(), num_sms,
                              (wk.m_tiles / 2) * wk.n_tiles * wk.splits, ma, mb, wk, epi, s, "gemm_tc_kernel<nt,pair>");
  }
  return launch_kernel<1>(gemm_tc_kernel<BN, pick_stages<BN, false, Epi>(), true, false, Epi>, smem_bytes<BN, false, Epi>(), num_sms,
                          tiles * wk.splits, ma, mb, wk, epi, s, "gemm_tc_kernel<nt>");
}

}  // namespace tc
}  // namespace bn
