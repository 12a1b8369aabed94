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

namespace bn {

constexpr int kMaxBins = 256;       // n_samples / guided_samples upper bound for the guided kernel
constexpr int kWarpsPerBlock = 4;

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

// Sum of x[0..n) in ATen's CPU order; x in shared memory, result replicated on all lanes.
__device__ float aten_row_sum_warp(const float* x, int n, int lane) {
  const int V = 8, ILP = 4;
  const int nvec = n / V, groups = nvec / ILP;
  const int k = lane >> 3, j = lane & 7;
  float p = 0.0f;
  for (int i = 0; i < groups; ++i) p = __fadd_rn(p, x[(i * ILP + k) * V + j]);
  if (k == 0)
    for (int v = groups * ILP; v < nvec; ++v) p = __fadd_rn(p, x[v * V + j]);
  float p1 = __shfl_sync(kFull, p, j + 8), p2 = __shfl_sync(kFull, p, j + 16), p3 = __shfl_sync(kFull, p, j + 24);
  if (k == 0) p = __fadd_rn(__fadd_rn(__fadd_rn(p, p1), p2), p3);
  float acc = 0.0f;
  for (int s = nvec * V; s < n; ++s) acc = __fadd_rn(acc, x[s]);
  for (int jj = 0; jj < V; ++jj) acc = __fadd_rn(acc, __shfl_sync(kFull, p, jj));
  return acc;
}

// ascending bitonic sort of keys in shared memory, m = power of two, one warp
__device__ void bitonic_sort_warp(float* key, int m, int lane) {
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += kWarp) {
        int l = i ^ j;
        if (l > i) {
          float a = key[i], b = key[l];
          bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[l] = a; }
        }
      }
      __syncwarp();
    }
  }
}

struct GuidedArgs {
  const float* z1; const float* depth; const float* weights;
  const float* t_vals; const float* gauss_w; const float* u_pred;
  const float* near0; const float* far0;
  const long long* valid_depth; const float* gt_depth; int gt_depth_stride; const float* gt_std;
  const float* u_gt;
  float* z2; float* std_out;
  float d_range; int N, S1, G;
};

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) guided_kernel(GuidedArgs a) {
  __shared__ float s_a[kWarpsPerBlock][kMaxBins];   // q / weights / pdf
  __shared__ float s_e[kWarpsPerBlock][kMaxBins];   // bin edges
  __shared__ float s_c[kWarpsPerBlock][kMaxBins];   // cdf
  __shared__ float s_s[kWarpsPerBlock][kMaxBins];   // samples (padded to a power of two)
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= a.N) return;
  float* A = s_a[wib]; float* E = s_e[wib]; float* C = s_c[wib]; float* Sm = s_s[wib];
  const int S1 = a.S1, n = a.G;
  const float eps = 1e-5f;
  const float k = a.d_range;

  // 1. sampling std of pass 1 (always evaluated: it is also an output)
  const float d_pred = a.depth[r];
  for (int i = lane; i < S1; i += kWarp) {
    float dz = __fsub_rn(a.z1[(long long)r * S1 + i], d_pred);
    A[i] = __fmul_rn(__fmul_rn(dz, dz), a.weights[(long long)r * S1 + i]);
  }
  __syncwarp();
  const float std_pred = __fsqrt_rn(aten_row_sum_warp(A, S1, lane));
  if (a.std_out && lane == 0) a.std_out[r] = std_pred;
  __syncwarp();

  // 2. window centre / half width: predicted depth, or the ground-truth depth for supervised rays
  float c = d_pred, sd = std_pred;
  const float* u = a.u_pred + (long long)r * n;
  if (a.valid_depth && a.valid_depth[r] > 0) {
    c = a.gt_depth[(long long)r * a.gt_depth_stride];
    sd = a.gt_std[r];
    u = a.u_gt + (long long)r * n;
  }
  const float near0 = *a.near0, far0 = *a.far0;
  float lo = __fsub_rn(c, __fmul_rn(k, sd)), hi = __fadd_rn(c, __fmul_rn(k, sd));
  // 3. clamp to the chunk's [near, far], then make the window symmetric about the centre
  lo = fminf(fmaxf(lo, near0), far0);
  hi = fminf(fmaxf(hi, near0), far0);
  const float half = fminf(fabsf(__fsub_rn(hi, c)), fabsf(__fsub_rn(lo, c)));
  lo = __fsub_rn(c, half);
  hi = __fadd_rn(c, half);
  // 4-5. bin edges
  const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(n - 1));
  for (int j = lane; j < n; j += kWarp) {
    float tj = a.t_vals[j];
    E[j] = __fadd_rn(__fmul_rn(lo, __fsub_rn(1.0f, tj)), __fmul_rn(hi, tj));
  }
  __syncwarp();
  // 6. Gaussian bin weights (+eps)
  const float inv_den = __fadd_rn(step, eps);
  for (int j = lane; j < n - 1; j += kWarp) {
    float factor = __fdiv_rn(__fsub_rn(E[j + 1], E[j]), inv_den);
    A[j] = __fadd_rn(__fmul_rn(factor, a.gauss_w[j]), eps);
  }
  __syncwarp();
  // 7. pdf
  const float tot = aten_row_sum_warp(A, n - 1, lane);
  __syncwarp();
  for (int j = lane; j < n - 1; j += kWarp) A[j] = __fdiv_rn(A[j], tot);
  __syncwarp();
  // 8. cdf: sequential fp64 accumulation, each prefix rounded to fp32
  if (lane == 0) {
    double acc = 0.0;
    C[0] = 0.0f;
    for (int j = 0; j < n - 1; ++j) { acc += (double)A[j]; C[j + 1] = (float)acc; }
  }
  __syncwarp();
  // 9. inverse-CDF lookup
  int m = 1; while (m < n) m <<= 1;
  for (int s = lane; s < m; s += kWarp) {
    float v = __int_as_float(0x7f800000);
    if (s < n) {
      float us = u[s];
      int lo_i = 0, hi_i = n;                    // ATen's cus_upper_bound, step for step (CDF may be non-monotone)
      while (lo_i < hi_i) { int mid = (lo_i + hi_i) >> 1; if (!(C[mid] > us)) lo_i = mid + 1; else hi_i = mid; }
      int below = max(lo_i - 1, 0), above = min(lo_i, n - 1);
      float c0 = C[below], c1 = C[above], b0 = E[below], b1 = E[above];
      float den = __fsub_rn(c1, c0);
      if (den < eps) den = 1.0f;
      v = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(us, c0), den), __fsub_rn(b1, b0)));
    }
    Sm[s] = v;
  }
  __syncwarp();
  // 10. ascending sort
  bitonic_sort_warp(Sm, m, lane);
  for (int s = lane; s < n; s += kWarp) a.z2[(long long)r * n + s] = Sm[s];
}

// Sort the concatenation [z1 | z2] per ray carrying the source index (stable on ties).
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp)
merge_kernel(const float* __restrict__ z1, const float* __restrict__ z2, float* __restrict__ z_out,
             long long* __restrict__ idx_out, float* __restrict__ unsort_out, int N, int S1, int G) {
  __shared__ float s_k[kWarpsPerBlock][2 * kMaxBins];
  __shared__ int s_i[kWarpsPerBlock][2 * kMaxBins];
  const int wib = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r = blockIdx.x * kWarpsPerBlock + wib;
  if (r >= N) return;
  float* K = s_k[wib]; int* I = s_i[wib];
  const int S = S1 + G;
  int m = 1; while (m < S) m <<= 1;
  for (int i = lane; i < m; i += kWarp) {
    float v = __int_as_float(0x7f800000);
    if (i < S1) v = z1[(long long)r * S1 + i];
    else if (i < S) v = z2[(long long)r * G + (i - S1)];
    K[i] = v; I[i] = i;
    if (i < S && unsort_out) unsort_out[(long long)r * S + i] = v;
  }
  __syncwarp();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += kWarp) {
        int l = i ^ j;
        if (l > i) {
          float a = K[i], b = K[l]; int ia = I[i], ib = I[l];
          bool gt = (a > b) || (a == b && ia > ib);
          bool up = (i & k) == 0;
          if (gt == up) { K[i] = b; K[l] = a; I[i] = ib; I[l] = ia; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < S; i += kWarp) {
    z_out[(long long)r * S + i] = K[i];
    if (idx_out) idx_out[(long long)r * S + i] = (long long)I[i];
  }
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
