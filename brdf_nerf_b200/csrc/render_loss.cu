// K-C of the Lambertian training step in ONE kernel, both directions: per ray (one warp, the whole 128-sample ray in
// registers) volume compositing, the Lambertian colour, the colour + depth-supervision loss, and the backward of all three
// down to the gradient of every packed MLP row.  Nothing per-sample (alpha, transmittance, weights) is written and read back,
// and the sort index is applied on load / on store (the MLP's rows stay in generation order, see bn_permute_samples).
//
// Replaces, for --model spsbrdf-nerf before the BRDF stage (no normals, no per-sample BRDF, irradiance 1, noise_std 0):
//   cal_weight + the weighted sums            models/spsbrdfnerf.py:50-69,196-199     composite_fwd128_kernel<4>
//   rgb = albedo_accu (+ rgb_padding)         models/spsbrdfnerf.py:281-283,459       shade_rays_kernel<false>
//   SNerfLoss + DepthLoss(subset)             metrics.py:39-61,82-161                 loss_kernel
//   autograd of the three                                                            shade_rays_kernel<true>, composite_bwd_kernel<4>
// and computes exactly their arithmetic (same expressions, same order): tests/test_gpu_composite.py compares the gradient rows
// and per-ray outputs with the five-launch path bit for bit; the scalar loss is the same sum of per-ray terms added with fp32
// atomics.  The five separate exports stay the general path (BRDF stages, render_rays' per-sample outputs, S != 128).
#include "common.cuh"
#include "composite_ray.cuh"

namespace bn {

constexpr float kRgbPad = 0.001f;        // rgb_padding, spsbrdfnerf.py:459 (kPad of shade.cu)

struct RenderLossArgs {
  const float* z;                 // (N,128) merged, ascending
  const float* rows;              // [N*128, 4] packed MLP rows (albedo rgb, sigma) in generation order
  const long long* sort_idx;      // (N,128) or null (rows already in depth order)
  int S1;
  const float* target_rgb;        // (N,3)
  const long long* valid_depth;   // (N) or null: no depth term
  const float* target_depth; const float* target_weight; int td_stride; const float* target_std;
  float lambda_rgb, k_ds; int use_all_depth;
  float* loss;                    // (1), zeroed before the launch
  float* g_rows;                  // [N*128, 4]
  float* rgb; float* depth;       // (N,3), (N): per-ray results (nullable)
  int N;
};

__global__ void __launch_bounds__(4 * kWarp) lambertian_render_loss_kernel(const __grid_constant__ RenderLossArgs a) {
  __shared__ float part[4];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int r = blockIdx.x * 4 + wid;
  constexpr int S = 128, R = S / kWarp, C = 4, kSig = 3;
  float contrib = 0.f;
  if (r < a.N) {
    const long long base = (long long)r * S;
    // ---- forward compositing (composite_fwd128_kernel<4>) ----
    float z[R], x[R][C], al[R], T[R], w[R];
    long long prow[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int i = k * kWarp + lane;
      z[k] = __ldg(a.z + base + i);
      prow[k] = packed_row(a.sort_idx, a.N, S, a.S1, r, i);
      load_row<C>(a.rows + prow[k] * C, x[k]);
    }
    float acc[3] = {0.f, 0.f, 0.f};
    float depth = 0.f, wsum = 0.f, carry = 1.0f;
    float delta[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int i = k * kWarp + lane;
      float znext = __shfl_down_sync(kFull, z[k], 1);
      const float zfirst_next = __shfl_sync(kFull, z[k + 1 < R ? k + 1 : k], 0);
      if (lane == kWarp - 1) znext = zfirst_next;
      const float sg = x[k][kSig];
      delta[k] = (i + 1 < S) ? (znext - z[k]) : 1e10f;
      al[k] = 1.0f - expf(-delta[k] * fmaxf(sg, 0.f));
      const float f = 1.0f - al[k] + 1e-10f;
      float incl = warp_scan_mul(f, lane);
      float excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 1.0f;
      T[k] = carry * excl;
      w[k] = al[k] * T[k];
      carry *= __shfl_sync(kFull, incl, kWarp - 1);
      depth += w[k] * z[k]; wsum += w[k];
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += w[k] * x[k][c];
    }
    depth = warp_sum(depth); wsum = warp_sum(wsum);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = warp_sum(acc[c]);
    // ---- Lambertian colour (shade_rays_kernel, BN_BRDF_NONE, irradiance 1) ----
    float raw[3], rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      raw[c] = 1.0f * ((1.0f + 2.0f * kRgbPad) * acc[c] - kRgbPad * wsum);
      rgb[c] = fminf(fmaxf(raw[c], 0.f), 1.f);
    }
    if (lane == 0) {
      if (a.rgb) { a.rgb[r * 3] = rgb[0]; a.rgb[r * 3 + 1] = rgb[1]; a.rgb[r * 3 + 2] = rgb[2]; }
      if (a.depth) a.depth[r] = depth;
    }
    // ---- loss and its gradients w.r.t. rgb / depth (loss_kernel) ----
    float g_rgb[3];
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float diff = rgb[c] - a.target_rgb[r * 3 + c];
      const float d2 = diff * diff;
      sq += (lane == c) ? d2 : 0.f;                  // the reference kernel sums lanes 0..2 through a butterfly: same order
      g_rgb[c] = diff * (2.0f * a.lambda_rgb / (3.0f * a.N));
    }
    sq = warp_sum(sq);
    contrib = a.lambda_rgb * sq / (3.0f * a.N);
    float gd = 0.f;
    if (a.valid_depth != nullptr) {
      float s2 = 0.f;
#pragma unroll
      for (int k = 0; k < R; ++k) { const float dz = z[k] - depth; s2 += dz * dz * w[k]; }
      s2 = warp_sum(s2);
      const float pred_std = sqrtf(s2);
      const float td = a.target_depth[(long long)r * a.td_stride];
      const float tw = a.target_weight ? a.target_weight[(long long)r * a.td_stride] : 1.0f;
      const float ts = a.target_std[r];
      const float dd = depth - td;
      bool sel = a.valid_depth[r] > 0;
      if (!a.use_all_depth) sel = sel && ((fabsf(dd) - ts) > 0.f || ts < pred_std);
      const float m = sel ? tw : 0.f;
      contrib += a.k_ds * m * dd * dd;
      gd = 2.0f * a.k_ds * m * dd;
    }
    // ---- backward of the colour (shade_rays_kernel<true>) ----
    float ga[3], gw0 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float up = (raw[c] >= 0.f && raw[c] <= 1.f) ? g_rgb[c] : 0.f;
      ga[c] = 0.f + 1.0f * (1.0f + 2.0f * kRgbPad) * up;
      gw0 += -1.0f * kRgbPad * up;
    }
    // ---- backward of the compositing (composite_bwd_kernel<4>): the ray back to front, suffix sum of g_j w_j over j > i ----
    float bcarry = 0.f;
#pragma unroll
    for (int k = R - 1; k >= 0; --k) {
      float g = gd * z[k] + gw0;
#pragma unroll
      for (int c = 0; c < C; ++c) g += ((c == kSig) ? 0.f : ga[c < 3 ? c : 0]) * x[k][c];
      const float gw = g * w[k];
      float rev = gw;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { float t = __shfl_down_sync(kFull, rev, o); if (lane + o < 32) rev += t; }
      const float row_total = __shfl_sync(kFull, rev, 0);
      const float suffix = rev - gw + bcarry;
      bcarry += row_total;
      const float f = 1.0f - al[k] + 1e-10f;
      const float d_alpha = g * T[k] - suffix / f;
      const float sg = x[k][kSig];
      const float d_sigma = sg > 0.f ? d_alpha * delta[k] * (1.0f - al[k]) : 0.f;
      float out[C];
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = w[k] * ga[c];
      out[kSig] = d_sigma;
      store_row<C>(a.g_rows + prow[k] * C, out);
    }
  }
  if (lane == 0) part[wid] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(a.loss, (part[0] + part[1]) + (part[2] + part[3]));
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default")))
int bn_lambertian_render_loss(const float* z, const float* rows, const int64_t* sort_idx, int n_stratified,
                              const float* target_rgb, const int64_t* valid_depth, const float* target_depth,
                              const float* target_weight, int td_stride, const float* target_std, float lambda_rgb,
                              float lambda_ds, int use_all_depth, float* loss, float* g_rows, float* rgb, float* depth,
                              int n_rays, int n_samples, cudaStream_t stream) {
  BN_CHECK_ARG(z && rows && target_rgb && loss && g_rows, "null pointer");
  BN_CHECK_ARG(n_rays > 0, "empty batch");
  BN_CHECK_ARG(n_samples == 128, "the fused Lambertian step handles rays of 128 samples (64 stratified + 64 guided); use "
                                 "bn_composite_forward / bn_shade_rays_* / bn_loss_color_depth otherwise");
  BN_CHECK_ARG(!sort_idx || (n_stratified >= 0 && n_stratified <= n_samples), "n_stratified out of range");
  BN_CHECK_ARG(valid_depth == nullptr || (target_depth && target_std && td_stride >= 1),
               "the depth term needs target_depth and target_std");
  RenderLossArgs a{};
  a.z = z; a.rows = rows; a.sort_idx = (const long long*)sort_idx; a.S1 = n_stratified; a.target_rgb = target_rgb;
  a.valid_depth = (const long long*)valid_depth; a.target_depth = target_depth; a.target_weight = target_weight;
  a.td_stride = td_stride; a.target_std = target_std; a.lambda_rgb = lambda_rgb; a.k_ds = lambda_ds / 3.0f / (float)n_rays;
  a.use_all_depth = use_all_depth; a.loss = loss; a.g_rows = g_rows; a.rgb = rgb; a.depth = depth; a.N = n_rays;
  BN_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  lambertian_render_loss_kernel<<<ceil_div(n_rays, 4), 4 * kWarp, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
