// K-C (part 1): volume compositing — alpha / transmittance / weights and the weighted accumulation
// of every per-sample channel, forward and backward.
//
// Replaces (reference, paths relative to /root/reference):
//   cal_weight                         models/spsbrdfnerf.py:50-69
//   calc_depth_std                     train_utils.py:35-39         (std output)
//   the Σ_s w·x accumulations          models/spsbrdfnerf.py:198-199,242,270-275,292,314-317,326-338,352
//   autograd of all of the above       (cumprod_backward, exp, relu, mul, sum)
//
// Layout: one warp per ray. Sample i of the ray is owned by lane (i mod 32), so every global access
// of a (N,S) tensor is a coalesced 128-byte row segment and the per-sample channel block
// packed[N,S,C] (the MLP's packed output, reference channel order spsbrdfnerf.py:694-757) is read
// as C contiguous floats per lane (128-bit vectors when C % 4 == 0).  The transmittance is an
// exclusive multiplicative scan: 5 shuffles per 32 samples with a running carry.
// HBM traffic (algorithmic, fp32): fwd reads z + packed (+noise when noise_std != 0), writes
// alpha, T, w; bwd re-reads z, packed, alpha, T, w and writes d_packed.
#include "common.cuh"
#include "composite_ray.cuh"

namespace bn {

constexpr int kRaysPerBlock = 4;
constexpr int kMaxSamples = 512;

// Narrow rows (Lambertian: C = 4, density pass: C = 1) at the path's standard ray length S = 128: the whole ray is loaded
// up front — four 32-sample rows x (z + C channels) per lane, 2.6 KB in flight per warp — and then composited out of
// registers.  The generic kernel below keeps two rows in flight (3.95 -> 4.4 TB/s at 65 536 rays); this one exists because
// the Lambertian forward is the one on the headline configuration and was the furthest from the HBM roofline.
template <int C>
__global__ void __launch_bounds__(kRaysPerBlock * kWarp) composite_fwd128_kernel(CompositeFwd a) {
  static_assert(C <= 4, "register budget: 4 rows x C channels per lane");
  const int lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kRaysPerBlock + threadIdx.x / kWarp;
  if (r >= a.N) return;
  constexpr int S = 128, R = S / kWarp;
  const long long base = (long long)r * S;
  constexpr int kSig = (C == 1) ? 0 : 3;
  float z[R], nz[R], ir[R], x[R][C];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = k * kWarp + lane;
    z[k] = __ldg(a.z + base + i);
    load_row<C>(a.packed + packed_row(a.sort_idx, a.N, S, a.S1, r, i) * C, x[k]);
    nz[k] = a.noise ? __ldg(a.noise + base + i) : 0.f;
    ir[k] = 0.f;
    if constexpr (C > 1) { if (a.irr) ir[k] = __ldg(a.irr + base + i); }
  }
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  float acc_i[4] = {0.f, 0.f, 0.f, 0.f};
  float depth = 0.f, wsum = 0.f, carry = 1.0f;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = k * kWarp + lane;
    float znext = __shfl_down_sync(kFull, z[k], 1);
    const float zfirst_next = __shfl_sync(kFull, z[k + 1 < R ? k + 1 : k], 0);
    if (lane == kWarp - 1) znext = zfirst_next;
    float sg = x[k][kSig];
    if (a.noise) sg += nz[k] * a.noise_std;
    const float delta = (i + 1 < S) ? (znext - z[k]) : 1e10f;
    const float al = 1.0f - expf(-delta * fmaxf(sg, 0.f));       // accurate expf: alpha feeds the guided sampler
    const float f = 1.0f - al + 1e-10f;
    float incl = warp_scan_mul(f, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    const float w = al * T;
    carry *= __shfl_sync(kFull, incl, kWarp - 1);
    if (a.alpha) a.alpha[base + i] = al;
    if (a.trans) a.trans[base + i] = T;
    a.weights[base + i] = w;
    depth += w * z[k]; wsum += w;
    if constexpr (C > 1) {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += w * x[k][c];
      if (a.irr) {
        const float wi = w * ir[k];
        acc_i[0] += wi * x[k][0]; acc_i[1] += wi * x[k][1]; acc_i[2] += wi * x[k][2]; acc_i[3] += wi;
      }
    }
  }
  depth = warp_sum(depth); wsum = warp_sum(wsum);
  if (a.std) {
    float s2 = 0.f;                                              // Σ w (z-d)² out of registers, same order as the generic kernel
    __syncwarp();
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

// generic ray length: composite_ray (composite_ray.cuh), one warp per ray
template <int C>
__global__ void __launch_bounds__(kRaysPerBlock * kWarp) composite_fwd_kernel(CompositeFwd a) {
  const int lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kRaysPerBlock + threadIdx.x / kWarp;
  if (r >= a.N) return;
  composite_ray<C>(a, r, lane);
}

struct CompositeBwd {
  const float* z; const float* packed; const float* noise; const float* irr;
  float noise_std;
  const float *alpha, *trans, *weights;    // saved forward results
  const float *g_acc;                      // (N,C) grad wrt acc (entry sigma_ch ignored)
  const float *g_acc_irr;                  // (N,4) or null
  const float *g_depth, *g_wsum;           // (N) nullable
  const float *g_weights;                  // (N,S) nullable: explicit grad wrt weights (losses)
  const float *g_packed_direct;            // (N,S,C) nullable: explicit grad wrt per-sample channels
  float* g_packed;                         // (N,S,C) out
  int N, S, sigma_ch;
  const long long* sort_idx; int S1;       // optional, as in CompositeFwd: packed AND g_packed are in the MLP's row order
};

template <int C>
__global__ void __launch_bounds__(kRaysPerBlock * kWarp) composite_bwd_kernel(CompositeBwd a) {
  const int lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kRaysPerBlock + threadIdx.x / kWarp;
  if (r >= a.N) return;
  const int S = a.S;
  const long long base = (long long)r * S;
  constexpr int kSig = 3;
  float ga[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ga[c] = (c == kSig) ? 0.f : a.g_acc[(long long)r * C + c];
  float gi[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.irr && a.g_acc_irr) { for (int c = 0; c < 4; ++c) gi[c] = a.g_acc_irr[r * 4 + c]; }
  const float gd = a.g_depth ? a.g_depth[r] : 0.f;
  const float gw0 = a.g_wsum ? a.g_wsum[r] : 0.f;
  // walk the ray back to front: suffix sum of g_j w_j over j > i
  float carry = 0.f;
  const int rows = ceil_div(S, kWarp);
  for (int rr = rows - 1; rr >= 0; --rr) {
    const int i = rr * kWarp + lane;
    const bool ok = i < S;
    float x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = 0.f;
    float zi = 0.f, al = 0.f, T = 0.f, w = 0.f, znext = 0.f, irr = 0.f;
    const long long prow = ok ? packed_row(a.sort_idx, a.N, S, a.S1, r, i) : 0;
    if (ok) {
      load_row<C>(a.packed + prow * C, x);
      zi = a.z[base + i]; al = a.alpha[base + i]; T = a.trans[base + i]; w = a.weights[base + i];
      znext = (i + 1 < S) ? a.z[base + i + 1] : 0.f;
      if (a.irr) irr = a.irr[base + i];
    }
    float g = gd * zi + gw0;
#pragma unroll
    for (int c = 0; c < C; ++c) g += ga[c] * x[c];
    if (a.irr) g += irr * (gi[0] * x[0] + gi[1] * x[1] + gi[2] * x[2] + gi[3]);
    if (a.g_weights && ok) g += a.g_weights[base + i];
    float gw = ok ? g * w : 0.f;
    // inclusive suffix sum within the row (reverse scan), then make it exclusive and add the carry
    float rev = gw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float t = __shfl_down_sync(kFull, rev, o); if (lane + o < 32) rev += t; }
    const float row_total = __shfl_sync(kFull, rev, 0);
    const float suffix = rev - gw + carry;          // Σ_{j>i} g_j w_j
    carry += row_total;
    if (ok) {
      const float f = 1.0f - al + 1e-10f;
      const float d_alpha = g * T - suffix / f;
      float sg = x[kSig];
      if (a.noise) sg += a.noise[base + i] * a.noise_std;
      const float delta = (i + 1 < S) ? (znext - zi) : 1e10f;
      const float d_sigma = sg > 0.f ? d_alpha * delta * (1.0f - al) : 0.f;
      float out[C];
#pragma unroll
      for (int c = 0; c < C; ++c) out[c] = w * ga[c];
      if (a.irr) { float wi = w * irr; out[0] += wi * gi[0]; out[1] += wi * gi[1]; out[2] += wi * gi[2]; }
      out[kSig] = d_sigma;
      if (a.g_packed_direct) {
        float e[C];
        load_row<C>(a.g_packed_direct + (base + i) * C, e);
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] += e[c];
      }
      store_row<C>(a.g_packed + prow * C, out);
    }
  }
}

template <int C> int launch_fwd(const CompositeFwd& a, cudaStream_t s) {
  if constexpr (C <= 4) {
    if (a.S == 128) {        // standard ray length of the full pass (64 stratified + 64 guided): whole ray in registers
      composite_fwd128_kernel<C><<<ceil_div(a.N, kRaysPerBlock), kRaysPerBlock * kWarp, 0, s>>>(a);
      return after_launch("composite_fwd128_kernel");
    }
  }
  composite_fwd_kernel<C><<<ceil_div(a.N, kRaysPerBlock), kRaysPerBlock * kWarp, 0, s>>>(a);
  return after_launch("composite_fwd_kernel");
}
template <int C> int launch_bwd(const CompositeBwd& a, cudaStream_t s) {
  composite_bwd_kernel<C><<<ceil_div(a.N, kRaysPerBlock), kRaysPerBlock * kWarp, 0, s>>>(a);
  return after_launch("composite_bwd_kernel");
}

#define BN_DISPATCH_C(C, FN, ...)                                                          \
  switch (C) {                                                                             \
    case 4: return FN<4>(__VA_ARGS__);   case 5: return FN<5>(__VA_ARGS__);               \
    case 7: return FN<7>(__VA_ARGS__);   case 8: return FN<8>(__VA_ARGS__);               \
    case 10: return FN<10>(__VA_ARGS__); case 11: return FN<11>(__VA_ARGS__);             \
    case 13: return FN<13>(__VA_ARGS__); case 14: return FN<14>(__VA_ARGS__);             \
    case 16: return FN<16>(__VA_ARGS__); case 17: return FN<17>(__VA_ARGS__);             \
    case 19: return FN<19>(__VA_ARGS__); case 20: return FN<20>(__VA_ARGS__);             \
    case 22: return FN<22>(__VA_ARGS__);                                                   \
    default: set_error("composite: unsupported channel count %d", C); return BN_ERR_ARG;   \
  }

static int dispatch_fwd(int C, const CompositeFwd& a, cudaStream_t s) { BN_DISPATCH_C(C, launch_fwd, a, s) }
static int dispatch_bwd(int C, const CompositeBwd& a, cudaStream_t s) { BN_DISPATCH_C(C, launch_bwd, a, s) }

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default"))) int bn_composite_sigma(const float* z, const float* sigma, const float* noise, float noise_std,
                                  float* alpha, float* trans, float* weights, float* depth, float* std_out,
                                  int n_rays, int n_samples, cudaStream_t stream) {
  BN_CHECK_ARG(z && sigma && weights && depth, "null pointer");
  BN_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples, "n_samples out of range");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  CompositeFwd a{};
  a.z = z; a.packed = sigma; a.noise = (noise && noise_std != 0.f) ? noise : nullptr; a.noise_std = noise_std;
  a.alpha = alpha; a.trans = trans; a.weights = weights; a.depth = depth; a.std = std_out;
  a.N = n_rays; a.S = n_samples; a.sigma_ch = 0;
  return launch_fwd<1>(a, stream);
}

extern "C" __attribute__((visibility("default"))) int bn_composite_forward(const float* z, const float* packed, int n_channels, int sigma_channel,
                                    const float* noise, float noise_std, const float* irr,
                                    float* alpha, float* trans, float* weights,
                                    float* depth, float* wsum, float* acc, float* acc_irr,
                                    int n_rays, int n_samples, const int64_t* sort_idx, int n_stratified, cudaStream_t stream) {
  BN_CHECK_ARG(z && packed && alpha && trans && weights && depth && acc, "null pointer");
  BN_CHECK_ARG(!sort_idx || (n_stratified >= 0 && n_stratified <= n_samples), "n_stratified out of range");
  BN_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples, "n_samples out of range");
  BN_CHECK_ARG(sigma_channel == 3, "the packed row keeps the density in channel 3 (reference layout)");
  BN_CHECK_ARG(!irr || acc_irr, "irr given without acc_irr");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  CompositeFwd a{};
  a.z = z; a.packed = packed; a.noise = (noise && noise_std != 0.f) ? noise : nullptr; a.noise_std = noise_std;
  a.irr = irr; a.alpha = alpha; a.trans = trans; a.weights = weights; a.depth = depth; a.wsum = wsum;
  a.acc = acc; a.acc_irr = acc_irr; a.N = n_rays; a.S = n_samples; a.sigma_ch = sigma_channel;
  a.sort_idx = (const long long*)sort_idx; a.S1 = n_stratified;
  return dispatch_fwd(n_channels, a, stream);
}

extern "C" __attribute__((visibility("default"))) int bn_composite_backward(const float* z, const float* packed, int n_channels, int sigma_channel,
                                     const float* noise, float noise_std, const float* irr,
                                     const float* alpha, const float* trans, const float* weights,
                                     const float* g_acc, const float* g_acc_irr, const float* g_depth,
                                     const float* g_wsum, const float* g_weights, const float* g_packed_direct,
                                     float* g_packed, int n_rays, int n_samples, const int64_t* sort_idx, int n_stratified,
                                     cudaStream_t stream) {
  BN_CHECK_ARG(z && packed && alpha && trans && weights && g_acc && g_packed, "null pointer");
  BN_CHECK_ARG(!sort_idx || (n_stratified >= 0 && n_stratified <= n_samples), "n_stratified out of range");
  BN_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples, "n_samples out of range");
  BN_CHECK_ARG(sigma_channel == 3, "the packed row keeps the density in channel 3 (reference layout)");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  CompositeBwd a{};
  a.z = z; a.packed = packed; a.noise = (noise && noise_std != 0.f) ? noise : nullptr; a.noise_std = noise_std;
  a.irr = irr; a.alpha = alpha; a.trans = trans; a.weights = weights;
  a.g_acc = g_acc; a.g_acc_irr = g_acc_irr; a.g_depth = g_depth; a.g_wsum = g_wsum;
  a.g_weights = g_weights; a.g_packed_direct = g_packed_direct; a.g_packed = g_packed;
  a.N = n_rays; a.S = n_samples; a.sigma_ch = sigma_channel;
  a.sort_idx = (const long long*)sort_idx; a.S1 = n_stratified;
  return dispatch_bwd(n_channels, a, stream);
}
