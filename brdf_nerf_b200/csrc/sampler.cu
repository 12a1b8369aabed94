// K-A: stratified + depth-guided sample generator and the two-list merge.
//
// Replaces (reference, paths relative to /root/reference):
//   get_z_vals                    rendering.py:149-166     -> stratified_kernel
//   calc_depth_std                train_utils.py:35-39     -> guided_kernel step 1
//   compute_samples_around_depth  rendering.py:116-130     -> guided_kernel step 2
//   sample_3sigma_asym            rendering.py:76-91       -> guided_kernel step 3
//   sample_3sigma / sample_pdf    rendering.py:54-74,13-52 -> guided_kernel steps 4-10
//   GenerateGuidedSamples         rendering.py:132-147     -> per-ray GT override (mask, no compaction)
//   sort / cat / sort             rendering.py:263-273     -> merge_kernel
//
// Arithmetic contract (bit-exact against oracle/sampler_np.py and the reference's torch-CPU
// results): fp32 round-to-nearest after every operation — this TU is compiled with -fmad=false and
// uses explicit __f*_rn intrinsics; the row sum follows ATen's CPU order (8-lane partials, 4-way
// ILP, lanes in order), the CDF is accumulated sequentially in fp64 and rounded per prefix, the
// bin search is an upper bound (searchsorted right=True).  The two constant tables (t_vals and
// the Gaussian window) are inputs produced by torch on the host.
#include "common.cuh"
#include "sampler_ray.cuh"

namespace bn {

__global__ void stratified_kernel(const float* __restrict__ near_p, const float* __restrict__ far_p,
                                  int stride, const float* __restrict__ t, const float* __restrict__ u,
                                  float* __restrict__ z, long long total, int S) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int r = (int)(idx / S), i = (int)(idx % S);
  float nr = near_p[(long long)r * stride], fr = far_p[(long long)r * stride];
  auto zf = [&](int j) {
    float tj = t[j];
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, tj)), __fmul_rn(fr, tj));
  };
  float zi = zf(i);
  float lower = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(zf(i - 1), zi));
  float upper = (i == S - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, zf(i + 1)));
  z[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u[idx]));
}

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) guided_kernel(GuidedArgs a) {
  __shared__ float s_a[kWarpsPerBlock][kMaxBins];   // q / weights / pdf
  __shared__ float s_e[kWarpsPerBlock][kMaxBins];   // bin edges
  __shared__ float s_c[kWarpsPerBlock][kMaxBins];   // cdf
  __shared__ float s_s[kWarpsPerBlock][kMaxBins];   // samples (padded to a power of two)
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= a.N) return;
  guided_ray(a, r, lane, s_a[wib], s_e[wib], s_c[wib], s_s[wib]);
}

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp)
merge_kernel(const float* z1, const float* z2, float* z_out, long long* idx_out, float* unsort_out, int N, int S1, int G) {
  __shared__ float s_k[kWarpsPerBlock][2 * kMaxBins];
  __shared__ int s_i[kWarpsPerBlock][2 * kMaxBins];
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= N) return;
  merge_ray(z1, z2, z_out, idx_out, unsort_out, S1, G, r, lane, s_k[wib], s_i[wib]);
}

// plain per-row ascending sort (gsam_only path: rendering.py:263,266-269)
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp)
rowsort_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int S) {
  __shared__ float s_k[kWarpsPerBlock][2 * kMaxBins];
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= N) return;
  float* K = s_k[wib];
  int m = 1; while (m < S) m <<= 1;
  for (int i = lane; i < m; i += kWarp) K[i] = i < S ? in[(long long)r * S + i] : __int_as_float(0x7f800000);
  __syncwarp();
  bitonic_sort_warp(K, m, lane);
  for (int i = lane; i < S; i += kWarp) out[(long long)r * S + i] = K[i];
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default"))) int bn_sample_stratified(const float* near_p, const float* far_p, int stride,
                                    const float* t_vals, const float* u, float* z_out,
                                    int n_rays, int n_samples, cudaStream_t stream) {
  BN_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && stride >= 1, "bad sizes");
  if (n_rays == 0) return BN_OK;
  BN_CHECK_ARG(near_p && far_p && t_vals && u && z_out, "null pointer");
  long long total = (long long)n_rays * n_samples;
  int threads = 256;
  stratified_kernel<<<(unsigned)ceil_div_ll(total, threads), threads, 0, stream>>>(
      near_p, far_p, stride, t_vals, u, z_out, total, n_samples);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_sample_guided(const float* z1, const float* depth, const float* weights,
                                const float* t_vals, const float* gauss_w, const float* u_pred,
                                const float* near0, const float* far0, float d_range,
                                const int64_t* valid_depth, const float* gt_depth, int gt_depth_stride,
                                const float* gt_std, const float* u_gt,
                                float* z2_out, float* std_out,
                                int n_rays, int n_samples, int n_guided, cudaStream_t stream) {
  BN_CHECK_ARG(z1 && depth && weights && t_vals && gauss_w && u_pred && near0 && far0 && z2_out, "null pointer");
  BN_CHECK_ARG(n_samples >= 2 && n_samples <= kMaxBins && n_guided >= 2 && n_guided <= kMaxBins,
               "n_samples / guided_samples must be in [2, 256]");
  BN_CHECK_ARG(!valid_depth || (gt_depth && gt_std && u_gt && gt_depth_stride >= 1),
               "valid_depth given without gt_depth / gt_std / u_gt");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  GuidedArgs a{z1, depth, weights, t_vals, gauss_w, u_pred, near0, far0,
               (const long long*)valid_depth, gt_depth, gt_depth_stride, gt_std, u_gt,
               z2_out, std_out, d_range, n_rays, n_samples, n_guided};
  guided_kernel<<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * kWarp, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_merge_samples(const float* z1, const float* z2, float* z_out, int64_t* idx_out,
                                float* unsort_out, int n_rays, int n_samples, int n_guided,
                                cudaStream_t stream) {
  BN_CHECK_ARG(z1 && z2 && z_out, "null pointer");
  BN_CHECK_ARG(n_samples >= 1 && n_guided >= 1 && n_samples + n_guided <= 2 * kMaxBins, "bad sizes");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  merge_kernel<<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * kWarp, 0, stream>>>(
      z1, z2, z_out, (long long*)idx_out, unsort_out, n_rays, n_samples, n_guided);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

// Applies sort_idx to per-point rows.  The MLP evaluates the points of a ray in generation order,
// stratified block first ([N][S1] rows) then guided block ([N][G] rows), so that the trunk activations of the
// stratified points are computed ONCE and shared by the density pass and the full pass (the reference evaluates
// them twice, rendering.py:225 and :274); compositing needs them in depth order (rendering.py:271-273).
//   scatter == 0: sorted[r][s][:] = blocks[row(r, idx[r][s])][:]       (forward: packed MLP rows -> depth order)
//   scatter == 1: blocks[row(r, idx[r][s])][:] = sorted[r][s][:]       (backward: gradients -> MLP row order)
// row(r, j) = j < S1 ? r*S1 + j : N*S1 + r*G + (j - S1).  idx[r] is a permutation, so scatter is a bijection.
__global__ void permute_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst,
                                    int N, int S1, int G, int pitch, int scatter) {
  const int S = S1 + G;
  const long long tot = (long long)N * S * pitch;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long pt = e / pitch; const int c = (int)(e % pitch);
    const long long r = pt / S;
    const int j = (int)idx[pt];
    const long long row = j < S1 ? r * S1 + j : (long long)N * S1 + r * G + (j - S1);
    if (scatter) dst[row * pitch + c] = src[e]; else dst[e] = src[row * pitch + c];
  }
}

extern "C" __attribute__((visibility("default")))
int bn_permute_samples(const float* src, const int64_t* sort_idx, float* dst, int n_rays, int n_samples, int n_guided,
                       int pitch, int scatter, cudaStream_t stream) {
  BN_CHECK_ARG(src && sort_idx && dst && src != dst, "null or aliased pointer");
  BN_CHECK_ARG(n_samples >= 1 && n_guided >= 1 && pitch >= 1, "bad sizes");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  const long long tot = (long long)n_rays * (n_samples + n_guided) * pitch;
  const int grid = (int)min((tot + 255) / 256, (long long)148 * 16);
  permute_rows_kernel<<<grid, 256, 0, stream>>>(src, (const long long*)sort_idx, dst, n_rays, n_samples, n_guided, pitch, scatter);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_sort_rows(const float* in, float* out, int n_rays, int n, cudaStream_t stream) {
  BN_CHECK_ARG(in && out && n >= 1 && n <= 2 * kMaxBins, "bad arguments");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  rowsort_kernel<<<ceil_div(n_rays, kWarpsPerBlock), kWarpsPerBlock * kWarp, 0, stream>>>(in, out, n_rays, n);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
