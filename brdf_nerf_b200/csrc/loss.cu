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
