// Between the two network passes of render_rays (rendering.py:262-273 of the reference) everything is per-ray work on 64 + 64
// numbers: compositing of the stratified densities (cal_weight, spsbrdfnerf.py:50-69 -> weights, depth), the depth-guided samples
// (compute_samples_around_depth ... sample_pdf, rendering.py:13-147) and the merge with its sort index (rendering.py:263-273).
// As three launches (composite_fwd_kernel<1>, guided_kernel, merge_kernel) they were 34 us of dependent, latency-bound kernels in a
// 2.3 ms step; this kernel runs the SAME three device functions (composite_ray.cuh, sampler_ray.cuh) back to back in the ray's
// warp.  Results are bit-identical to the three separate exports (tests/test_gpu_sampler.py): the sampler arithmetic is explicit
// round-to-nearest intrinsics, the compositing code is compiled with the same flags as composite.cu.
#include "common.cuh"
#include "composite_ray.cuh"
#include "sampler_ray.cuh"

namespace bn {

struct CoarseToFineArgs {
  CompositeFwd c;        // stratified densities -> alpha / trans (nullable) / weights / depth
  GuidedArgs g;          // -> z2 (+ sampling std)
  float* z_out; long long* idx_out; float* unsort_out;      // merged depths, sort index, unsorted concatenation
};

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) coarse_to_fine_kernel(const __grid_constant__ CoarseToFineArgs a) {
  // guided: 4 arrays of kMaxBins floats per warp; merge: keys + indices of 2 kMaxBins entries per warp — the same 4 KB, reused
  __shared__ float s_buf[kWarpsPerBlock][4 * kMaxBins];
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= a.c.N) return;
  composite_ray<1>(a.c, r, lane);
  __syncwarp();                                   // weights[r][*], depth[r] written by this warp are read below
  float* b = s_buf[wib];
  guided_ray(a.g, r, lane, b, b + kMaxBins, b + 2 * kMaxBins, b + 3 * kMaxBins);
  __syncwarp();                                   // z2[r][*]
  merge_ray(a.g.z1, a.g.z2, a.z_out, a.idx_out, a.unsort_out, a.g.S1, a.g.G, r, lane, b, reinterpret_cast<int*>(b + 2 * kMaxBins));
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default")))
int bn_coarse_to_fine(const float* z1, const float* sigma1, const float* noise1, float noise_std,
                      const float* t_vals, const float* gauss_w, const float* u_pred, const float* near0, const float* far0,
                      float d_range, const int64_t* valid_depth, const float* gt_depth, int gt_depth_stride, const float* gt_std,
                      const float* u_gt, float* weights1, float* depth1, float* std1, float* z2, float* z_out, int64_t* idx_out,
                      float* unsort_out, int n_rays, int n_samples, int n_guided, cudaStream_t stream) {
  BN_CHECK_ARG(z1 && sigma1 && t_vals && gauss_w && u_pred && near0 && far0 && weights1 && depth1 && z2 && z_out, "null pointer");
  BN_CHECK_ARG(n_samples >= 2 && n_samples <= kMaxBins && n_guided >= 2 && n_guided <= kMaxBins,
               "n_samples / guided_samples must be in [2, 256]");
  BN_CHECK_ARG(!valid_depth || (gt_depth && gt_std && u_gt && gt_depth_stride >= 1),
               "valid_depth given without gt_depth / gt_std / u_gt");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  CoarseToFineArgs a{};
  a.c.z = z1; a.c.packed = sigma1; a.c.noise = (noise1 && noise_std != 0.0f) ? noise1 : nullptr; a.c.irr = nullptr; a.c.noise_std = noise_std;
  a.c.alpha = nullptr; a.c.trans = nullptr; a.c.weights = weights1; a.c.depth = depth1; a.c.wsum = nullptr; a.c.std = nullptr;
  a.c.acc = nullptr; a.c.acc_irr = nullptr; a.c.N = n_rays; a.c.S = n_samples; a.c.sigma_ch = 0;
  a.g = GuidedArgs{z1, depth1, weights1, t_vals, gauss_w, u_pred, near0, far0, (const long long*)valid_depth, gt_depth,
                   gt_depth_stride, gt_std, u_gt, z2, std1, d_range, n_rays, n_samples, n_guided};
  a.z_out = z_out; a.idx_out = (long long*)idx_out; a.unsort_out = unsort_out;
  coarse_to_fine_kernel<<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * kWarp, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
