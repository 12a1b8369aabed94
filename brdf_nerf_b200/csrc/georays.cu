// Ray feed geometry: pixel grid + RPC camera model -> the (N, 11) float32 ray records that render_rays
// consumes, built on the GPU (the reference localises every pixel of every image in numpy on the host, twice, and ships
// per-ray dicts through DataLoader workers).
//
// Replaces (reference, paths relative to /root/reference):
//   get_rays            datasets/satellite_rgb_dep.py:23-78      near = localisation at max_alt, far = at min_alt,
//                                                                 o = near, d = (far - near)/|far - near|, bounds [0, |far - near|]
//   rpc.localization    third-party `rpcm` (requirements.txt:2): iterative inversion of the RPC00B projection
//                       (rpc_model.py localization_iterative / apply_rfm / apply_poly), restated in oracle/georays_np.py
//   latlon_to_ecef_custom   sat_utils.py:110-125  (cs == 'ecef')
//   utm_from_latlon     sat_utils.py:148-162 -> pyproj "+proj=utm +zone=N" (cs == 'utm', the default): Krueger series
//                       transverse Mercator to order n^6 (Karney 2011), GRS80, no false northing (see oracle/georays_np.py)
//   normalize_rays      datasets/satellite_rgb_dep.py:550-559    float32: (o - center) / range, near / range, far / range
//   get_sun_dirs        datasets/satellite_rgb_dep.py:561-576    appended as columns 8..10 (hstack at :390)
// One thread per pixel: two localisations (5-7 iterations of six rational-polynomial evaluations each, float64), the map
// projection, and a coalesced store of the record through shared memory.  FP64-pipe bound (~12 k flops per ray), not HBM.
// The reference iterates until EVERY pixel of a call has converged, so a pixel's result depends on the slowest pixel's
// iteration count (an early-converged point moves by ~1e-7 m in the extra iterations: enough to flip float32 direction
// bits).  Two launches reproduce that without a host round trip: rpc_iterations_kernel finds the per-call counts (block
// maximum + conditional atomicMax), rays_from_rpc_kernel then iterates every pixel exactly that often.
// Compiled with -fmad=false so that the float64 arithmetic matches the oracle's operation for operation (libm aside).
#include "common.cuh"
#include "geodesy.cuh"

#include <math.h>

namespace bn {

struct RpcDev {
  double row_offset, col_offset, lat_offset, lon_offset, alt_offset;
  double row_scale, col_scale, lat_scale, lon_scale, alt_scale;
  double row_num[20], row_den[20], col_num[20], col_den[20];
};

struct RaysArgs {
  RpcDev rpc;
  const double* cols; const double* rows;     // device (N) or both null: pixel p -> (p % width, p / width)
  long long n; int width;
  double min_alt, max_alt;
  int cs;                                      // 0 ecef, 1 utm
  UtmParams utm;
  int normalize; float cx, cy, cz, range;
  int with_sun; float sx, sy, sz;
  float* out; int out_stride;
  int* fail_count;
  int* iters;                                  // device: [0] iterations of the max_alt call, [1] of the min_alt call
};

// rpcm apply_poly: x = lat, y = lon, z = alt (normalised); same grouping of the additions as the reference package
__device__ __forceinline__ double rpc_poly(const double* __restrict__ p, double x, double y, double z) {
  double out = 0.0;
  out += p[0];
  out += p[1] * y + p[2] * x + p[3] * z;
  out += p[4] * y * x + p[5] * y * z + p[6] * x * z;
  out += p[7] * y * y + p[8] * x * x + p[9] * z * z;
  out += p[10] * x * y * z;
  out += p[11] * y * y * y;
  out += p[12] * y * x * x + p[13] * y * z * z + p[14] * y * y * x;
  out += p[15] * x * x * x;
  out += p[16] * x * z * z + p[17] * y * y * z + p[18] * x * x * z;
  out += p[19] * z * z * z;
  return out;
}
__device__ __forceinline__ double rpc_rfm(const double* num, const double* den, double x, double y, double z) {
  return rpc_poly(num, x, y, z) / rpc_poly(den, x, y, z);
}

// rpcm localization_iterative for one point.  fixed_iters < 0: iterate until THIS point is within 1e-18 (returns its
// iteration count in n_out, false when 100 iterations were not enough); fixed_iters >= 0: exactly that many iterations,
// which is what the reference does to every point of a call (its loop runs until the slowest point has converged).
__device__ bool rpc_localize(const RpcDev& r, double col, double row, double alt, int fixed_iters, double& lon_deg,
                             double& lat_deg, int& n_out) {
  const double ncol = (col - r.col_offset) / r.col_scale, nrow = (row - r.row_offset) / r.row_scale;
  const double nalt = (alt - r.alt_offset) / r.alt_scale;
  double lon = -1.0, lat = -1.0, eps = 2.0;
  double x0 = rpc_rfm(r.col_num, r.col_den, lat, lon, nalt), y0 = rpc_rfm(r.row_num, r.row_den, lat, lon, nalt);
  double x1 = rpc_rfm(r.col_num, r.col_den, lat, lon + eps, nalt), y1 = rpc_rfm(r.row_num, r.row_den, lat, lon + eps, nalt);
  double x2 = rpc_rfm(r.col_num, r.col_den, lat + eps, lon, nalt), y2 = rpc_rfm(r.row_num, r.row_den, lat + eps, lon, nalt);
  int n = 0;
  bool ok = true;
  while (true) {
    if (fixed_iters >= 0) { if (n >= fixed_iters) break; }
    else {
      if ((x0 - ncol) * (x0 - ncol) + (y0 - nrow) * (y0 - nrow) < 1e-18) break;
      if (n > 100) { ok = false; break; }
    }
    const double e1x = x1 - x0, e1y = y1 - y0, e2x = x2 - x0, e2y = y2 - y0, ux = ncol - x0, uy = nrow - y0;
    const double a1 = (ux * e1x + uy * e1y) / (e1x * e1x + e1y * e1y);
    const double a2 = (ux * e2x + uy * e2y) / (e2x * e2x + e2y * e2y);
    lon = lon + a1 * eps;
    lat = lat + a2 * eps;
    eps = .1;
    x0 = rpc_rfm(r.col_num, r.col_den, lat, lon, nalt); y0 = rpc_rfm(r.row_num, r.row_den, lat, lon, nalt);
    x1 = rpc_rfm(r.col_num, r.col_den, lat, lon + eps, nalt); y1 = rpc_rfm(r.row_num, r.row_den, lat, lon + eps, nalt);
    x2 = rpc_rfm(r.col_num, r.col_den, lat + eps, lon, nalt); y2 = rpc_rfm(r.row_num, r.row_den, lat + eps, lon, nalt);
    ++n;
  }
  n_out = n;
  lon_deg = lon * r.lon_scale + r.lon_offset;
  lat_deg = lat * r.lat_scale + r.lat_offset;
  return ok;
}

constexpr int kRaysBlock = 128;

__device__ __forceinline__ void pixel_of(const RaysArgs& a, long long p, double& col, double& row) {
  col = a.cols ? a.cols[p] : (double)(p % a.width);
  row = a.rows ? a.rows[p] : (double)(p / a.width);
}

// Pass 1: how many iterations does the slowest pixel need, per altitude (the reference's `while not np.all(...)`)?
__global__ void __launch_bounds__(kRaysBlock) rpc_iterations_kernel(const __grid_constant__ RaysArgs a) {
  __shared__ int s_n[kRaysBlock / 32][2];
  const long long p = (long long)blockIdx.x * kRaysBlock + threadIdx.x;
  int n_max = 0, n_min = 0;
  if (p < a.n) {
    double col, row, lon, lat;
    pixel_of(a, p, col, row);
    bool ok = rpc_localize(a.rpc, col, row, a.max_alt, -1, lon, lat, n_max);
    ok = rpc_localize(a.rpc, col, row, a.min_alt, -1, lon, lat, n_min) && ok;
    if (!ok && a.fail_count) atomicAdd(a.fail_count, 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_max = max(n_max, __shfl_xor_sync(kFull, n_max, o));
    n_min = max(n_min, __shfl_xor_sync(kFull, n_min, o));
  }
  if ((threadIdx.x & 31) == 0) { s_n[threadIdx.x >> 5][0] = n_max; s_n[threadIdx.x >> 5][1] = n_min; }
  __syncthreads();
  if (threadIdx.x < 2) {
    int m = 0;
#pragma unroll
    for (int wi = 0; wi < kRaysBlock / 32; ++wi) m = max(m, s_n[wi][threadIdx.x]);
    if (m > *(volatile int*)(a.iters + threadIdx.x)) atomicMax(a.iters + threadIdx.x, m);   // conditional: almost never taken
  }
}

// Pass 2: the rays, every pixel iterated exactly as often as the reference iterates the whole call.
__global__ void __launch_bounds__(kRaysBlock) rays_from_rpc_kernel(const __grid_constant__ RaysArgs a) {
  __shared__ float s_out[kRaysBlock * 11];
  const long long base = (long long)blockIdx.x * kRaysBlock;
  const int cnt = (int)((a.n - base) < kRaysBlock ? (a.n - base) : kRaysBlock);
  const int tid = threadIdx.x;
  const int it_max = min(a.iters[0], 101), it_min = min(a.iters[1], 101);
  if (tid < cnt) {
    const long long p = base + tid;
    double col, row, lon, lat, nx, ny, nz, fx, fy, fz;
    int n_unused;
    pixel_of(a, p, col, row);
    rpc_localize(a.rpc, col, row, a.max_alt, it_max, lon, lat, n_unused);  // nearest to the camera: maximum altitude
    if (a.cs == 0) to_ecef(lat, lon, a.max_alt, nx, ny, nz); else { to_utm(a.utm, lat, lon, nx, ny); nz = a.max_alt; }
    rpc_localize(a.rpc, col, row, a.min_alt, it_min, lon, lat, n_unused);
    if (a.cs == 0) to_ecef(lat, lon, a.min_alt, fx, fy, fz); else { to_utm(a.utm, lat, lon, fx, fy); fz = a.min_alt; }
    const double dx = fx - nx, dy = fy - ny, dz = fz - nz;
    const double len = sqrt((dx * dx + dy * dy) + dz * dz);
    float o0 = (float)nx, o1 = (float)ny, o2 = (float)nz, near = 0.f, far = (float)len;
    if (a.normalize) {                                                     // float32, one rounding per operation
      o0 = (o0 - a.cx) / a.range; o1 = (o1 - a.cy) / a.range; o2 = (o2 - a.cz) / a.range;
      near = near / a.range; far = far / a.range;
    }
    float* s = s_out + tid * a.out_stride;
    s[0] = o0; s[1] = o1; s[2] = o2;
    s[3] = (float)(dx / len); s[4] = (float)(dy / len); s[5] = (float)(dz / len);
    s[6] = near; s[7] = far;
    if (a.with_sun) { s[8] = a.sx; s[9] = a.sy; s[10] = a.sz; }
  }
  __syncthreads();
  for (int t = tid; t < cnt * a.out_stride; t += kRaysBlock) a.out[base * a.out_stride + t] = s_out[t];
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default")))
int bn_rays_from_rpc(const bn_rpc* rpc, const double* cols, const double* rows, long long n_rays, int width,
                     double min_alt, double max_alt, int cs, int utm_zone, int normalize, float center_x, float center_y,
                     float center_z, float scene_range, const float* sun_dir, float* rays_out, int out_stride,
                     int* fail_count, int* iterations, cudaStream_t stream) {
  BN_CHECK_ARG(rpc && rays_out && iterations, "null pointer");
  BN_CHECK_ARG((cols == nullptr) == (rows == nullptr), "cols and rows go together");
  BN_CHECK_ARG(n_rays > 0 && (cols != nullptr || width > 0), "n_rays must be > 0 (and width > 0 for a pixel grid)");
  BN_CHECK_ARG(cs == 0 || cs == 1, "cs must be 0 (ecef) or 1 (utm)");
  BN_CHECK_ARG(cs == 0 || (utm_zone >= 1 && utm_zone <= 60), "utm_zone must be 1..60");
  BN_CHECK_ARG(out_stride == (sun_dir ? 11 : 8), "out_stride must be 8, or 11 with a sun direction");
  BN_CHECK_ARG(!normalize || scene_range > 0.f, "scene_range must be > 0");
  static_assert(sizeof(bn_rpc) == sizeof(RpcDev), "bn_rpc layout");
  RaysArgs a;
  memcpy(&a.rpc, rpc, sizeof(RpcDev));
  a.cols = cols; a.rows = rows; a.n = n_rays; a.width = width; a.min_alt = min_alt; a.max_alt = max_alt; a.cs = cs;
  a.utm = make_utm_params(cs == 1 ? utm_zone : 1);
  a.normalize = normalize; a.cx = center_x; a.cy = center_y; a.cz = center_z; a.range = scene_range;
  a.with_sun = sun_dir != nullptr;
  a.sx = sun_dir ? sun_dir[0] : 0.f; a.sy = sun_dir ? sun_dir[1] : 0.f; a.sz = sun_dir ? sun_dir[2] : 0.f;
  a.out = rays_out; a.out_stride = out_stride; a.fail_count = fail_count; a.iters = iterations;
  const long long blocks = ceil_div_ll(n_rays, kRaysBlock);
  BN_CHECK_ARG(blocks < (1ll << 31), "too many rays for one launch");
  BN_CUDA(cudaMemsetAsync(iterations, 0, 2 * sizeof(int), stream));
  rpc_iterations_kernel<<<(unsigned)blocks, kRaysBlock, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  rays_from_rpc_kernel<<<(unsigned)blocks, kRaysBlock, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
