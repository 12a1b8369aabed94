// Tile inference -> DSM: the product the reference's users evaluate, built from the rendered depth of a
// whole tile without leaving the GPU.
//
// Replaces (reference, paths relative to /root/reference):
//   get_latlonalt_from_nerf_prediction   datasets/satellite_rgb_dep.py:601-634
//       xyz = ((double)o + (double)d * (double)depth) * range + center      float64, every operation rounded separately;
//       cs == 'utm' (the default, opt.py:252): that is (east, north, alt); cs == 'ecef': ecef_to_latlon_custom
//       (sat_utils.py:127-146) then the UTM projection of sat_utils.py:148-162 (geodesy.cuh)
//   get_dsm_from_nerf_prediction         datasets/satellite_rgb_dep.py:636-697
//       cloud bounds -> raster grid (:666-671, four scalars on the host), then plyflatten(cloud, xoff, yoff, resolution,
//       xsize, ysize, radius=1, sigma=inf) (:680; plyflatten==0.2.0, requirements.txt:11): every point adds its height
//       with weight w to the (2 radius + 1)^2 cells around its own cell, cell value = weighted mean, empty cells = NaN.
//   calc_normal_from_pts3d               sat_utils.py:16-50 (via calc_normal_from_depth_v2, satellite_rgb_dep.py:578-585)
//       four-neighbour cross-product normals of the float32 point image.
//
// Kernels (all HBM / L2-atomic bound, no tensor-core shape anywhere):
//   dsm_points_kernel      thread per ray: 24 B out (float64 x, y, alt), bounds by warp shuffle + one atomicMax per warp
//                          on order-preserving integer keys (max of x, y, -x, -y: one zero-initialised key array).
//   dsm_scatter_kernel     thread per point.  sigma == inf (the reference's call): the weight is 1 for every neighbour, so
//                          the (2r+1)^2 scatter factorises: ONE (sum, count) atomic pair into the point's own cell of an
//                          apron-extended grid, and the neighbourhood sum becomes a dense box filter in the finalize pass
//                          (2 atomics per point instead of 18).  Finite sigma: the weight depends on the point's position
//                          inside its cell, so every neighbour gets its own atomic pair.
//   dsm_finalize_kernel    thread per raster cell: box-sums the apron grid (sigma == inf) or reads its own cell, writes
//                          mean (float32) / NaN and the cell's weight sum.
// Sums are float64 atomics, counts float32 (exact integers for sigma == inf); the reference keeps a float32 running mean
// whose value depends on the point order, so rasters agree to ~1e-4 m, the count image exactly.
// Compiled with -fmad=false: the float64 point cloud is bit-exact against the reference's torch/numpy result.
#include "common.cuh"
#include "geodesy.cuh"

#include <math.h>

namespace bn {

__device__ __forceinline__ unsigned long long dkey(double v) {        // order-preserving map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(kFull, v, o);
    v = t > v ? t : v;
  }
  return v;
}

constexpr int kPtsBlock = 256;
constexpr int kMaxRayStride = 16;

// A block handles 256 consecutive rays: the ray records are staged through shared memory with fully coalesced loads
// (the 44-byte record straddles sectors when a thread reads its own row), the float64 / float32 points leave through
// shared memory the same way (a thread's 24-byte point would otherwise be a stride-3 store).
__global__ void __launch_bounds__(kPtsBlock) dsm_points_kernel(const float* __restrict__ rays, int ray_stride,
                                                               const float* __restrict__ depth, long long n, double range,
                                                               double cx, double cy, double cz, double* __restrict__ cloud,
                                                               float* __restrict__ pts_f32, unsigned long long* __restrict__ keys,
                                                               int vec_ok, int ecef, const __grid_constant__ UtmParams utm) {
  __shared__ __align__(16) float s_ray[kPtsBlock * kMaxRayStride];
  __shared__ __align__(16) double s_out[kPtsBlock * 3];
  const long long base = (long long)blockIdx.x * kPtsBlock;
  const int cnt = (int)((n - base) < kPtsBlock ? (n - base) : kPtsBlock);
  const int tid = threadIdx.x;
  const bool full = cnt == kPtsBlock && vec_ok;                       // full blocks move 16-byte vectors (block bases are 16 B aligned)
  if (full) {
    const float4* src = reinterpret_cast<const float4*>(rays + base * ray_stride);
    for (int t = tid; t < kPtsBlock * ray_stride / 4; t += kPtsBlock) reinterpret_cast<float4*>(s_ray)[t] = src[t];
  } else {
    for (int t = tid; t < cnt * ray_stride; t += kPtsBlock) s_ray[t] = rays[base * ray_stride + t];
  }
  __syncthreads();
  unsigned long long k0 = 0, k1 = 0, k2 = 0, k3 = 0;                   // 0 is below the key of every double
  if (tid < cnt) {
    const float* ray = s_ray + tid * ray_stride;                        // odd stride (11): conflict free
    const double dep = (double)depth[base + tid];
    double x = ((double)ray[0] + (double)ray[3] * dep) * range + cx;
    double y = ((double)ray[1] + (double)ray[4] * dep) * range + cy;
    double z = ((double)ray[2] + (double)ray[5] * dep) * range + cz;
    if (ecef) {                                                         // cs == 'ecef' (satellite_rgb_dep.py:629-631)
      double lat, lon, alt;
      from_ecef(x, y, z, lat, lon, alt);
      to_utm(utm, lat, lon, x, y);
      z = alt;
    }
    s_out[tid * 3 + 0] = x; s_out[tid * 3 + 1] = y; s_out[tid * 3 + 2] = z;
    if (isfinite(x) && isfinite(y)) { k0 = dkey(x); k1 = dkey(y); k2 = dkey(-x); k3 = dkey(-y); }
  }
  __syncthreads();
  if (full) {
    double2* dst = reinterpret_cast<double2*>(cloud + base * 3);
    for (int t = tid; t < kPtsBlock * 3 / 2; t += kPtsBlock) dst[t] = reinterpret_cast<const double2*>(s_out)[t];
    if (pts_f32 && tid < kPtsBlock * 3 / 4)
      reinterpret_cast<float4*>(pts_f32 + base * 3)[tid] = make_float4((float)s_out[4 * tid], (float)s_out[4 * tid + 1],
                                                                        (float)s_out[4 * tid + 2], (float)s_out[4 * tid + 3]);
  } else {
    for (int t = tid; t < cnt * 3; t += kPtsBlock) {
      const double v = s_out[t];
      cloud[base * 3 + t] = v;
      if (pts_f32) pts_f32[base * 3 + t] = (float)v;
    }
  }
  if (keys) {
    // block maximum through shared memory, then ONE conditional atomic per key and block: same-address atomics serialise in
    // L2 (measured: a per-warp atomicMax made this kernel 375 us for 4.2 M rays), and after the first few blocks almost no
    // block still raises a bound, which a plain (possibly stale, hence conservative) read detects
    __shared__ unsigned long long s_key[kPtsBlock / 32][4];
    k0 = warp_max_u64(k0); k1 = warp_max_u64(k1); k2 = warp_max_u64(k2); k3 = warp_max_u64(k3);
    if ((tid & 31) == 0) { s_key[tid >> 5][0] = k0; s_key[tid >> 5][1] = k1; s_key[tid >> 5][2] = k2; s_key[tid >> 5][3] = k3; }
    __syncthreads();
    if (tid < 4) {
      unsigned long long m = 0;
#pragma unroll
      for (int wi = 0; wi < kPtsBlock / 32; ++wi) m = s_key[wi][tid] > m ? s_key[wi][tid] : m;
      if (m != 0 && m > *(volatile unsigned long long*)(keys + tid)) atomicMax(keys + tid, m);
    }
  }
}

// bounds (4 doubles): xmin, xmax, ymin, ymax
__global__ void dsm_bounds_kernel(const unsigned long long* __restrict__ keys, double* __restrict__ bounds) {
  if (threadIdx.x == 0) {
    bounds[1] = dkey_inv(keys[0]); bounds[3] = dkey_inv(keys[1]);
    bounds[0] = -dkey_inv(keys[2]); bounds[2] = -dkey_inv(keys[3]);
  }
}

struct DsmGrid {
  double xoff, yoff, resolution;
  int xsize, ysize, radius;
  float sigma;              // +inf: box mode
  int box;                  // 1: sigma == inf
  int gw, gh;               // accumulator grid: (xsize + 2 radius) x (ysize + 2 radius) in box mode, else xsize x ysize
};

__global__ void __launch_bounds__(256) dsm_scatter_kernel(const double* __restrict__ cloud, int cloud_stride, int value_col,
                                                          long long n, DsmGrid g, double* __restrict__ acc_sum,
                                                          float* __restrict__ acc_cnt) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double xx = cloud[p * cloud_stride], yy = cloud[p * cloud_stride + 1];
  const double v = (double)(float)cloud[p * cloud_stride + value_col];   // the rasteriser reads heights as float32
  if (!(isfinite(xx) && isfinite(yy))) return;
  const double fi = floor((xx - g.xoff) / g.resolution), fj = floor((-yy - (-g.yoff)) / g.resolution);
  if (!(fabs(fi) < 1e9 && fabs(fj) < 1e9)) return;
  const int i = (int)fi, j = (int)fj;
  if (g.box) {
    const int gi = i + g.radius, gj = j + g.radius;                     // own cell inside the apron-extended grid
    if (gi < 0 || gj < 0 || gi >= g.gw || gj >= g.gh) return;
    const long long k = (long long)g.gw * gj + gi;
    atomicAdd(acc_sum + k, v);
    atomicAdd(acc_cnt + k, 1.0f);
    return;
  }
  for (int k1 = -g.radius; k1 <= g.radius; ++k1)
    for (int k2 = -g.radius; k2 <= g.radius; ++k2) {
      const int ii = i + k1, jj = j + k2;
      if (ii < 0 || jj < 0 || ii >= g.xsize || jj >= g.ysize) continue;
      const float dist_x = (float)(xx - (g.xoff + g.resolution * (0.5 + ii)));
      const float dist_y = (float)(yy - (g.yoff - g.resolution * (0.5 + jj)));
      const float dist = hypotf(dist_x, dist_y);
      const float w = expf(-dist * dist / (2.0f * g.sigma * g.sigma));
      const long long k = (long long)g.xsize * jj + ii;
      atomicAdd(acc_sum + k, v * (double)w);
      atomicAdd(acc_cnt + k, w);
    }
}

__global__ void __launch_bounds__(256) dsm_finalize_kernel(DsmGrid g, const double* __restrict__ acc_sum,
                                                           const float* __restrict__ acc_cnt, float* __restrict__ raster,
                                                           float* __restrict__ cnt_out) {
  const int ii = blockIdx.x * 32 + (threadIdx.x & 31), jj = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ii >= g.xsize || jj >= g.ysize) return;
  double s = 0.0;
  float c = 0.f;
  if (g.box) {
    const int w = 2 * g.radius + 1;
    for (int dj = 0; dj < w; ++dj)
      for (int di = 0; di < w; ++di) {
        const long long k = (long long)g.gw * (jj + dj) + (ii + di);
        s += acc_sum[k];
        c += acc_cnt[k];
      }
  } else {
    const long long k = (long long)g.xsize * jj + ii;
    s = acc_sum[k];
    c = acc_cnt[k];
  }
  const long long o = (long long)g.xsize * jj + ii;
  raster[o] = c == 0.f ? __int_as_float(0x7fc00000) : (float)(s / (double)c);
  if (cnt_out) cnt_out[o] = c;
}

// l2_normalize (train_utils.py:28-33): x / sqrt(max(sum x^2, eps)), eps = float32 machine epsilon
__device__ __forceinline__ float3 l2n(float3 a) {
  const float n = fmaxf(a.x * a.x + a.y * a.y + a.z * a.z, 1.1920928955078125e-07f);
  // MUFU rsqrt (<= 2 ulp) instead of IEEE sqrt + division: the stencil was issue-bound on those sequences (ncu: 82 % SM
  // busy at 28 % of the HBM roofline); the result stays within 1e-6 of the reference's x / sqrt(n)
  const float inv = rsqrtf(n);
  return make_float3(a.x * inv, a.y * inv, a.z * inv);
}
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }

__global__ void __launch_bounds__(256) dsm_normals_kernel(const float* __restrict__ pts, int h, int w, float* __restrict__ normals) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  float3 out = make_float3(0.f, 0.f, 0.f);
  if (x > 0 && y > 0 && x < w - 1 && y < h - 1) {
    const long long c = ((long long)y * w + x) * 3;
    const float3 p0 = ld3(pts + c);
    const float3 south = l2n(sub3(ld3(pts + c + 3ll * w), p0)), north = l2n(sub3(ld3(pts + c - 3ll * w), p0));
    const float3 east = l2n(sub3(ld3(pts + c + 3), p0)), west = l2n(sub3(ld3(pts + c - 3), p0));
    const float3 n1 = l2n(cross3(east, north)), n2 = l2n(cross3(west, south));
    const float3 n3 = l2n(cross3(north, west)), n4 = l2n(cross3(south, east));
    out = l2n(make_float3((((n1.x + n2.x) + n3.x) + n4.x) / 4.f, (((n1.y + n2.y) + n3.y) + n4.y) / 4.f,
                          (((n1.z + n2.z) + n3.z) + n4.z) / 4.f));
  }
  float* o = normals + ((long long)y * w + x) * 3;
  o[0] = out.x; o[1] = out.y; o[2] = out.z;
}

static int make_grid(DsmGrid& g, double xoff, double yoff, double resolution, int xsize, int ysize, int radius, float sigma) {
  g.xoff = xoff; g.yoff = yoff; g.resolution = resolution; g.xsize = xsize; g.ysize = ysize; g.radius = radius;
  g.sigma = sigma; g.box = isinf(sigma) ? 1 : 0;
  g.gw = g.box ? xsize + 2 * radius : xsize;
  g.gh = g.box ? ysize + 2 * radius : ysize;
  return 0;
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default")))
int bn_dsm_points(const float* rays, int ray_stride, const float* depth, long long n_rays, double scene_range,
                  double center_x, double center_y, double center_z, int cs, int utm_zone, double* cloud, float* points_f32,
                  double* bounds, unsigned long long* bounds_scratch, cudaStream_t stream) {
  BN_CHECK_ARG(cs == 0 || cs == 1, "cs must be 0 (ecef) or 1 (utm)");
  BN_CHECK_ARG(cs == 1 || (utm_zone >= 1 && utm_zone <= 60), "utm_zone must be 1..60 when cs is ecef");
  BN_CHECK_ARG(rays && depth && cloud, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && ray_stride >= 6 && ray_stride <= kMaxRayStride, "n_rays must be > 0 and 6 <= ray_stride <= 16");
  BN_CHECK_ARG((bounds == nullptr) == (bounds_scratch == nullptr), "bounds and bounds_scratch go together");
  if (bounds) BN_CUDA(cudaMemsetAsync(bounds_scratch, 0, 4 * sizeof(unsigned long long), stream));
  const long long blocks = ceil_div_ll(n_rays, kPtsBlock);
  BN_CHECK_ARG(blocks < (1ll << 31), "too many rays for one launch");
  const int vec_ok = ((uintptr_t)rays % 16 == 0) && ((uintptr_t)cloud % 16 == 0) && ((uintptr_t)points_f32 % 16 == 0);
  dsm_points_kernel<<<(unsigned)blocks, kPtsBlock, 0, stream>>>(rays, ray_stride, depth, n_rays, scene_range, center_x, center_y,
                                                          center_z, cloud, points_f32, bounds_scratch, vec_ok, cs == 0 ? 1 : 0,
                                                          make_utm_params(cs == 0 ? utm_zone : 1));
  BN_LAUNCH_CHECK();
  if (bounds) {
    dsm_bounds_kernel<<<1, 32, 0, stream>>>(bounds_scratch, bounds);
    BN_LAUNCH_CHECK();
  }
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
size_t bn_dsm_workspace_bytes(int xsize, int ysize, int radius, float sigma) {
  if (xsize <= 0 || ysize <= 0 || radius < 0) return 0;
  DsmGrid g;
  make_grid(g, 0, 0, 1, xsize, ysize, radius, sigma);
  return (size_t)g.gw * g.gh * (sizeof(double) + sizeof(float));
}

static int ws_check(const char* who, const DsmGrid& g, const void* workspace, size_t workspace_bytes) {
  const size_t need = (size_t)g.gw * g.gh * (sizeof(double) + sizeof(float));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("%s: workspace missing or too small (%zu < %zu bytes)", who, workspace_bytes, need);
    return BN_ERR_STATE;
  }
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
int bn_dsm_accumulate(const double* cloud, int cloud_stride, int value_col, long long n_points, double xoff, double yoff,
                      double resolution, int xsize, int ysize, int radius, float sigma, int zero_first,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(cloud != nullptr, "null pointer");
  BN_CHECK_ARG(n_points > 0 && cloud_stride >= 3 && value_col >= 2 && value_col < cloud_stride, "bad cloud layout");
  BN_CHECK_ARG(xsize > 0 && ysize > 0 && radius >= 0 && radius <= 8 && resolution > 0 && sigma > 0, "bad raster grid");
  DsmGrid g;
  make_grid(g, xoff, yoff, resolution, xsize, ysize, radius, sigma);
  if (int rc = ws_check("bn_dsm_accumulate", g, workspace, workspace_bytes)) return rc;
  const size_t cells = (size_t)g.gw * g.gh;
  double* acc_sum = (double*)workspace;
  float* acc_cnt = (float*)(acc_sum + cells);
  if (zero_first) BN_CUDA(cudaMemsetAsync(workspace, 0, cells * (sizeof(double) + sizeof(float)), stream));
  const long long blocks = ceil_div_ll(n_points, 256);
  BN_CHECK_ARG(blocks < (1ll << 31), "too many points for one launch");
  dsm_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(cloud, cloud_stride, value_col, n_points, g, acc_sum, acc_cnt);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
int bn_dsm_finalize(int xsize, int ysize, int radius, float sigma, const void* workspace, size_t workspace_bytes,
                    float* raster, float* count, cudaStream_t stream) {
  BN_CHECK_ARG(raster != nullptr, "null pointer");
  BN_CHECK_ARG(xsize > 0 && ysize > 0 && radius >= 0 && radius <= 8 && sigma > 0, "bad raster grid");
  DsmGrid g;
  make_grid(g, 0.0, 0.0, 1.0, xsize, ysize, radius, sigma);
  if (int rc = ws_check("bn_dsm_finalize", g, workspace, workspace_bytes)) return rc;
  const double* acc_sum = (const double*)workspace;
  const float* acc_cnt = (const float*)(acc_sum + (size_t)g.gw * g.gh);
  dsm_finalize_kernel<<<dim3(ceil_div(xsize, 32), ceil_div(ysize, 8)), 256, 0, stream>>>(g, acc_sum, acc_cnt, raster, count);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
int bn_dsm_rasterize(const double* cloud, int cloud_stride, int value_col, long long n_points, double xoff, double yoff,
                     double resolution, int xsize, int ysize, int radius, float sigma, float* raster, float* count,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(raster != nullptr, "null pointer");
  if (int rc = bn_dsm_accumulate(cloud, cloud_stride, value_col, n_points, xoff, yoff, resolution, xsize, ysize, radius,
                                 sigma, 1, workspace, workspace_bytes, stream)) return rc;
  return bn_dsm_finalize(xsize, ysize, radius, sigma, workspace, workspace_bytes, raster, count, stream);
}

extern "C" __attribute__((visibility("default")))
int bn_dsm_normals_from_points(const float* points, int height, int width, float* normals, cudaStream_t stream) {
  BN_CHECK_ARG(points && normals, "null pointer");
  BN_CHECK_ARG(height > 0 && width > 0, "empty image");
  dsm_normals_kernel<<<dim3(ceil_div(width, 32), ceil_div(height, 8)), 256, 0, stream>>>(points, height, width, normals);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
