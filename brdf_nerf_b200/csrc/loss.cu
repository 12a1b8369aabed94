// Training losses fused with their gradients: one launch replaces the ~15 element-wise / reduce
// launches (and, in the reference, three host round trips) of the colour + depth-supervision losses.
//
// Replaces (reference, paths relative to /root/reference):
//   SNerfLoss (lambda_sc == 0)                              metrics.py:39-61    L_c = lambda_rgb * mean((rgb - target)^2)
//   DepthLoss(subset=True, GNLL=False) / ComputeSubsetDepthLoss   metrics.py:82-161
//       rays with valid_depth > 0 are selected when usealldepth, or when |d - d*| > std* or calc_depth_std(z, d, w) > std*
//       (is_not_in_expected_distribution, :98-101);  L_d = (lambda_ds / 3) * mean_sel( n_sel / N * w* (d - d*)^2 )
//       = (lambda_ds / 3 / N) * sum_sel w* (d - d*)^2 : the n_sel normalisers cancel, so the selection is a 0/1 weight
//       and no index list (np.where on the host in the reference) is ever built.
//   calc_depth_std                                          train_utils.py:35-39
// Gradients: d L / d rgb = 2 lambda_rgb (rgb - target) / (3 N);  d L / d depth = 2 (lambda_ds / 3 / N) m (d - d*),
// m = selected * w*.  (The selection mask is piecewise constant: no gradient flows through pred_std.)
// One warp per ray; the scalar loss is reduced per block and added with one atomic.
#include "common.cuh"

namespace bn {

struct LossArgs {
  const float* rgb; const float* target_rgb;      // (N,3)
  const float* depth;                              // (N)
  const float* z; const float* weights;            // (N,S)
  const int64_t* valid_depth;                      // (N) or null: no depth term
  const float* target_depth; const float* target_weight; int td_stride;   // [r * td_stride]; target_weight null = 1
  const float* target_std;                         // (N)
  float lambda_rgb, k_ds;                          // k_ds = lambda_ds / 3 / N
  int use_all_depth;
  float* loss;                                     // (1) accumulated: zero before the launch
  float* g_rgb; float* g_depth;                    // (N,3), (N) (g_depth nullable when no depth term)
  int N, S;
};

__global__ void __launch_bounds__(128) loss_kernel(LossArgs a) {
  __shared__ float part[4];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int r = blockIdx.x * 4 + wid;
  float contrib = 0.f;
  if (r < a.N) {
    float sq = 0.f;
    if (lane < 3) {
      const float diff = a.rgb[r * 3 + lane] - a.target_rgb[r * 3 + lane];
      sq = diff * diff;
      a.g_rgb[r * 3 + lane] = diff * (2.0f * a.lambda_rgb / (3.0f * a.N));
    }
    sq = warp_sum(sq);
    contrib = a.lambda_rgb * sq / (3.0f * a.N);
    if (a.valid_depth != nullptr) {
      const float d = a.depth[r];
      float s2 = 0.f;
      for (int i = lane; i < a.S; i += kWarp) {
        const float dz = a.z[(long long)r * a.S + i] - d;
        s2 += dz * dz * a.weights[(long long)r * a.S + i];
      }
      s2 = warp_sum(s2);
      const float pred_std = sqrtf(s2);
      const float td = a.target_depth[(long long)r * a.td_stride];
      const float tw = a.target_weight ? a.target_weight[(long long)r * a.td_stride] : 1.0f;
      const float ts = a.target_std[r];
      const float dd = d - td;
      bool sel = a.valid_depth[r] > 0;
      if (!a.use_all_depth) sel = sel && ((fabsf(dd) - ts) > 0.f || ts < pred_std);
      const float m = sel ? tw : 0.f;
      contrib += a.k_ds * m * dd * dd;
      if (lane == 0) a.g_depth[r] = 2.0f * a.k_ds * m * dd;
    }
  }
  if (lane == 0) part[wid] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(a.loss, (part[0] + part[1]) + (part[2] + part[3]));
}

// Device-side NaN counter ("Errors"): the reference's train_utils.check_nan (train_utils.py:61-78) counts NaNs
// with torch.isnan(x).sum() and a host round trip per tensor (three per render_rays call, rendering.py:121-123).  Here the
// count is accumulated into a caller buffer with no synchronisation; the caller reads it when (and if) it wants to report.
__global__ void __launch_bounds__(256) count_nan_kernel(const float* __restrict__ x, long long n, int* __restrict__ counter) {
  int c = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ld_stream4(x + 4 * i);
    c += (v.x != v.x) + (v.y != v.y) + (v.z != v.z) + (v.w != v.w);
  }
  for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += x[i] != x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
  if ((threadIdx.x & 31) == 0 && c != 0) atomicAdd(counter, c);
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default")))
int bn_loss_color_depth(const float* rgb, const float* target_rgb, const float* depth, const float* z, const float* weights,
                        const int64_t* valid_depth, const float* target_depth, const float* target_weight, int td_stride,
                        const float* target_std, float lambda_rgb, float lambda_ds, int use_all_depth,
                        float* loss, float* g_rgb, float* g_depth, int n_rays, int n_samples, cudaStream_t stream) {
  BN_CHECK_ARG(rgb && target_rgb && loss && g_rgb, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  BN_CHECK_ARG(valid_depth == nullptr || (depth && z && weights && target_depth && target_std && g_depth),
               "the depth term needs depth, z, weights, target_depth, target_std and g_depth");
  LossArgs a{rgb, target_rgb, depth, z, weights, valid_depth, target_depth, target_weight, td_stride, target_std,
             lambda_rgb, lambda_ds / 3.0f / (float)n_rays, use_all_depth, loss, g_rgb, g_depth, n_rays, n_samples};
  BN_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  loss_kernel<<<ceil_div(n_rays, 4), 128, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Regularisers of the BRDF stage, fused with their gradients.  Replaces (reference):
//   NormalRegLoss      metrics.py:179-216   L_n = lambda * sum_{r,s} w_rs min(0, n_rs . v_r)^2,  v_r = -d_r (main.py:269-285;
//                      the reference flattens rays x samples before its `.sum(dim=-1)`, so the value is a SUM, not a ray mean)
//   HardSurfaceLoss    metrics.py:263-290   L_h = lambda / N * sum_r sum_s (z_rs - depth_r)^2 w_rs   (calc_depth_std_2,
//                      train_utils.py:38-39)
// Gradients written for bn_composite_backward: g_weights (N,S), the normal channels of g_packed (N,S,pitch; every other
// channel zero) and d L_h / d depth ADDED to g_depth.  One warp per ray.
namespace bn {

struct RegArgs {
  const float* weights; const float* z; const float* depth;    // (N,S), (N,S), (N)
  const float* packed; int pitch;                               // (N,S,pitch)
  int ch[2]; float lambda_nr[2];                                // normal channels (analytic, learned), -1 = off
  const float* rays;                                            // (N,11): d = rays[r*11 + 3..5]
  float k_hs;                                                   // lambda_hs / N (0 = off)
  float* loss; float* g_weights; float* g_packed; float* g_depth; float* bad_count;
  int N, S;
};

__global__ void __launch_bounds__(128) reg_loss_kernel(RegArgs a) {
  __shared__ float part[4];
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int r = blockIdx.x * 4 + wid;
  float contrib = 0.f;
  if (r < a.N) {
    const float vx = -a.rays[r * 11 + 3], vy = -a.rays[r * 11 + 4], vz = -a.rays[r * 11 + 5];
    const float d = a.depth ? a.depth[r] : 0.f;
    float gd = 0.f, bad[2] = {0.f, 0.f};
    for (int i = lane; i < a.S; i += kWarp) {
      const long long p = (long long)r * a.S + i;
      const float w = a.weights[p];
      float gw = 0.f;
      float* grow = a.g_packed ? a.g_packed + p * a.pitch : nullptr;
      if (grow) for (int c = 0; c < a.pitch; ++c) grow[c] = 0.f;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (a.ch[t] < 0) continue;
        const float* n = a.packed + p * a.pitch + a.ch[t];
        const float ndv = n[0] * vx + n[1] * vy + n[2] * vz;
        if (ndv < 0.f) bad[t] += 1.f;
        const float m = fminf(ndv, 0.f);
        contrib += a.lambda_nr[t] * w * m * m;
        gw += a.lambda_nr[t] * m * m;
        const float k = 2.0f * a.lambda_nr[t] * w * m;
        grow[a.ch[t]] = k * vx; grow[a.ch[t] + 1] = k * vy; grow[a.ch[t] + 2] = k * vz;
      }
      if (a.k_hs != 0.f) {
        const float dz = a.z[p] - d;
        contrib += a.k_hs * dz * dz * w;
        gw += a.k_hs * dz * dz;
        gd -= 2.0f * a.k_hs * dz * w;
      }
      a.g_weights[p] = gw;
    }
    contrib = warp_sum(contrib);
    if (a.k_hs != 0.f) { gd = warp_sum(gd); if (lane == 0) a.g_depth[r] += gd; }
    if (a.bad_count) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float b = warp_sum(bad[t]);
        if (lane == 0 && a.ch[t] >= 0 && b > 0.f) atomicAdd(a.bad_count + t, b);
      }
    }
  }
  if (lane == 0) part[wid] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(a.loss, (part[0] + part[1]) + (part[2] + part[3]));
}

}  // namespace bn

extern "C" __attribute__((visibility("default")))
int bn_loss_regularizers(const float* weights, const float* z, const float* depth, const float* packed, int pitch,
                         int normal_an_ch, float lambda_nr_an, int normal_lr_ch, float lambda_nr_lr, const float* rays,
                         float lambda_hs, float* loss, float* g_weights, float* g_packed, float* g_depth, float* bad_count,
                         int n_rays, int n_samples, cudaStream_t stream) {
  BN_CHECK_ARG(weights && loss && g_weights, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  const bool nr = (normal_an_ch >= 0 && lambda_nr_an != 0.f) || (normal_lr_ch >= 0 && lambda_nr_lr != 0.f);
  BN_CHECK_ARG(!nr || (packed && rays && g_packed && pitch >= 7), "the normal term needs packed, rays and g_packed");
  BN_CHECK_ARG(lambda_hs == 0.f || (z && depth && g_depth), "the hard-surface term needs z, depth and g_depth");
  RegArgs a{};
  a.weights = weights; a.z = z; a.depth = depth; a.packed = packed; a.pitch = pitch;
  a.ch[0] = (normal_an_ch >= 0 && lambda_nr_an != 0.f) ? normal_an_ch : -1; a.lambda_nr[0] = lambda_nr_an;
  a.ch[1] = (normal_lr_ch >= 0 && lambda_nr_lr != 0.f) ? normal_lr_ch : -1; a.lambda_nr[1] = lambda_nr_lr;
  BN_CHECK_ARG((a.ch[0] < 0 || a.ch[0] + 3 <= pitch) && (a.ch[1] < 0 || a.ch[1] + 3 <= pitch), "normal channel out of range");
  a.rays = rays; a.k_hs = lambda_hs / (float)n_rays;
  a.loss = loss; a.g_weights = g_weights; a.g_packed = nr ? g_packed : nullptr; a.g_depth = g_depth; a.bad_count = bad_count;
  a.N = n_rays; a.S = n_samples;
  reg_loss_kernel<<<ceil_div(n_rays, 4), 128, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
int bn_count_nan(const float* x, long long n, int* counter, cudaStream_t stream) {
  BN_CHECK_ARG(x && counter, "null pointer");
  BN_CHECK_ARG(n > 0, "n must be > 0");
  const long long blocks = ceil_div_ll(ceil_div_ll(n, 4), 256);
  count_nan_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, stream>>>(x, n, counter);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
