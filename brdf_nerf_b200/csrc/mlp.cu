// K-B: positional encoding + SIREN MLP — weight packing, encoding, heads, forward / backward
// orchestration and the C entry points.  Shared declarations live in mlp_internal.cuh.
#include "mlp_internal.cuh"
#include "mlp_chain.cuh"
#include "mlp_dgrad_chain.cuh"

namespace bn {

// ------------------------------------------------------------------------------------------------
// x = o + d z (separately rounded mul and add, as the reference) and the positional encoding
// [sin(2^k x), cos(2^k x)]_k written into cols 0..63 of X3 (zero padded).  One thread per point.
template <typename T>
__global__ void encode_kernel(const float* __restrict__ origins, int o_stride, const float* __restrict__ dirs,
                              int d_stride, const float* __restrict__ z, int S, long long P, int n_freq,
                              T* __restrict__ X3, long long ld) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long r = p / S;
  const float zz = z[p];
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = __fadd_rn(origins[r * o_stride + a], __fmul_rn(dirs[r * d_stride + a], zz));
  float e[kEncPad];
#pragma unroll
  for (int i = 0; i < kEncPad; ++i) e[i] = 0.f;
  if (n_freq == 0) { e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; }
  else {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      if (k < n_freq) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int a = 0; a < 3; ++a) { float s, c; sincosf(f * x[a], &s, &c); e[k * 6 + a] = s; e[k * 6 + 3 + a] = c; }
      }
    }
  }
  T* dst = X3 + p * ld;
#pragma unroll
  for (int i = 0; i < kEncPad; i += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = e[i + j];
    Pack<T, 8>::store(dst + i, t);
  }
}

// Mapping(ray direction) (nerf.py:53-70, mapping_sizes[1] frequencies; raw direction without --mapping) into the 64
// columns behind the features: the colour head's first layer reads [features | dir enc | 0] as one K = F + 64 operand
template <typename T>
__global__ void dir_encode_kernel(const float* __restrict__ dirs, int d_stride, int S, long long P, int n_freq,
                                  T* __restrict__ dst, long long ld, int ncols) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long r = p / S;
  float d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) d[a] = dirs[r * d_stride + a];
  float e[kEncPad];
#pragma unroll
  for (int i = 0; i < kEncPad; ++i) e[i] = 0.f;
  if (n_freq == 0) { e[0] = d[0]; e[1] = d[1]; e[2] = d[2]; }
  else {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      if (k < n_freq) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int a = 0; a < 3; ++a) { float sn, cs; sincosf(f * d[a], &sn, &cs); e[k * 6 + a] = sn; e[k * 6 + 3 + a] = cs; }
      }
    }
  }
  T* o = dst + p * ld;
#pragma unroll
  for (int i = 0; i < kEncPad; i += 8) {
    if (i >= ncols) continue;               // the columns behind belong to the time embedding (bn_mlp_write_t)
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = e[i + j];
    Pack<T, 8>::store(o + i, t);
  }
}

// time embedding of the rays into FE columns [F + kTOff, F + 64) of their points' rows (zeros behind the TE real columns);
// without a view direction the columns [F, F + kTOff) are zeroed here as well (nothing else writes them)
template <typename T>
__global__ void t_embed_kernel(const float* __restrict__ t_rows, int te, int S, long long P, bool zero_front,
                               T* __restrict__ dst, long long ld) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long r = p / S;
  T* o = dst + p * ld;
#pragma unroll
  for (int i = 0; i < kEncPad; i += 8) {
    if (i < kTOff && !zero_front) continue;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = i + j - kTOff; t[j] = (c >= 0 && c < te) ? t_rows[r * te + c] : 0.f; }
    Pack<T, 8>::store(o + i, t);
  }
}

template <typename T> __device__ __forceinline__ void load8g(const T* p, float (&v)[8]) { load8<T>(p, v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ------------------------------------------------------------------------------------------------
// second layers of all heads + sigma + learned normal.  A warp evaluates kQP points at a time (lanes
// split K, weights fetched once per kQP points, kQP independent load streams in flight) and walks
// the points with a grid-stride loop.
constexpr int kQP = 4;

template <typename T>
__global__ void __launch_bounds__(256) heads_fwd_kernel(HeadPlan hp, const float* __restrict__ params,
                                                        const T* __restrict__ Hlast, long long ldh, int F,
                                                        const T* __restrict__ HD, long long ldd,
                                                        float* __restrict__ out, int pitch, long long P, bool sigma_only,
                                                        bool sigma_done) {
  const int lane = threadIdx.x % 32;
  const long long warp = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const long long nwarps = (long long)gridDim.x * (blockDim.x / 32);
  for (long long p0 = warp * kQP; p0 < P; p0 += nwarps * kQP) {
    long long pq[kQP];
#pragma unroll
    for (int q = 0; q < kQP; ++q) pq[q] = min(p0 + q, P - 1);
    // sigma (+ learned normal) from the trunk's last activation
    float s_acc[kQP], g_acc[kQP][3];
#pragma unroll
    for (int q = 0; q < kQP; ++q) { s_acc[q] = 0.f; g_acc[q][0] = g_acc[q][1] = g_acc[q][2] = 0.f; }
    // sigma_done: the fused trunk kernel already wrote softplus(w_sigma . h + b) from its fp32 activations
    for (int i = lane * 8; i < ((sigma_done && hp.ch_nlr < 0) ? 0 : F); i += 256) {
      float h[kQP][8];
#pragma unroll
      for (int q = 0; q < kQP; ++q) load8g<T>(Hlast + pq[q] * ldh + i, h[q]);
      float w[8];
      load8<float>(params + hp.wsig + i, w);
#pragma unroll
      for (int q = 0; q < kQP; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_acc[q] = fmaf(h[q][j], w[j], s_acc[q]);
      if (hp.ch_nlr >= 0) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          load8<float>(params + hp.wg + (long long)o * F + i, w);
#pragma unroll
          for (int q = 0; q < kQP; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) g_acc[q][o] = fmaf(h[q][j], w[j], g_acc[q][o]);
        }
      }
    }
    const float bsig = __ldg(params + hp.bsig);
#pragma unroll
    for (int q = 0; q < kQP; ++q) {
      const float sigma = softplusf_(warp_sum(s_acc[q]) + bsig);
      if (lane == 0 && p0 + q < P && !sigma_done) {
        if (sigma_only) out[p0 + q] = sigma; else out[(p0 + q) * pitch + hp.ch_sigma] = sigma;
      }
    }
    if (sigma_only) continue;
    if (hp.ch_nlr >= 0) {
#pragma unroll
      for (int q = 0; q < kQP; ++q) {
        float g[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) g[o] = warp_sum(g_acc[q][o]) + __ldg(params + hp.bg + o);
        const float inv = 1.0f / sqrtf(fmaxf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2], 1.1920929e-07f));
        if (lane == 0 && p0 + q < P) {
          float* row = out + (p0 + q) * pitch + hp.ch_nlr;
          row[0] = -g[0] * inv; row[1] = -g[1] * inv; row[2] = -g[2] * inv;
        }
      }
    }
    for (int o = 0; o < hp.n_out; ++o) {
      const OutDesc d = hp.o[o];
      float acc[kQP];
#pragma unroll
      for (int q = 0; q < kQP; ++q) acc[q] = 0.f;
      for (int i = lane * 8; i < hp.HH; i += 256) {
        float w[8];
        load8<float>(params + d.w_off + i, w);
#pragma unroll
        for (int q = 0; q < kQP; ++q) {
          float h[8]; load8g<T>(HD + pq[q] * ldd + (long long)d.block * hp.HH + i, h);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q] = fmaf(h[j], w[j], acc[q]);
        }
      }
      const float bo = __ldg(params + d.b_off);
#pragma unroll
      for (int q = 0; q < kQP; ++q) {
        const float pre = warp_sum(acc[q]) + bo;
        const float s = sigmoidf_(pre);
        float v = s;
        if (d.xform == XF_SOFTPLUS) v = softplusf_(pre);
        else if (d.xform == XF_K) v = (s - 0.5f) * 2.0f + 1.0f;
        else if (d.xform == XF_THETA_RPV) v = (s - 0.5f) * 2.0f;
        else if (d.xform == XF_THETA_H) v = s * (float)(M_PI * 30.0 / 180.0);
        if (lane < d.rep && p0 + q < P) out[(p0 + q) * pitch + d.ch + lane] = v;
      }
    }
  }
}

// backward of heads_fwd_kernel.  Writes, per point:
//   GHD [P, n_blocks*HH] = (sum_o dpre_o W2_o) ⊙ CD      (dgrad operand of the heads' first layer)
//   G7D [P, F]           = dpre_sigma w_sigma + sum_o dpre_nlr,o Wg_o   (direct grads into h_{L-1})
//   DPRE [P, 64]         cols 0..15 head pre-activation grads, 16 sigma, 17..19 learned normal
//                        (A operand of the skinny second-layer weight gradients)
// and accumulates the small bias gradients (second layers, sigma, grad_from_xyz) and the bias
// gradient of the heads' first layer (column sums of GHD) through shared memory.
// kLight (tcgen05 mode): only DPRE and the small bias gradients are produced here; GHD comes out of a
// K = 64 GEMM (DPRE x W2p^T, masked by CD, column sums = first-layer bias gradients) and G7D is never
// materialised (rank-4 addend of the feature-layer dgrad epilogue).
template <typename T, bool kLight>
__global__ void __launch_bounds__(256) heads_bwd_kernel(HeadPlan hp, const float* __restrict__ params,
                                                        const float* __restrict__ out, const float* __restrict__ g_out, int pitch,
                                                        const T* __restrict__ Hlast, long long ldh, int F,
                                                        const T* __restrict__ CD, long long ldd,
                                                        T* __restrict__ GHD, T* __restrict__ G7D, T* __restrict__ DPRE,
                                                        float* __restrict__ g_params, long long P) {
  extern __shared__ float s_red[];              // [n_blocks*HH] column sums of GHD, then [24] small biases
  const int HKa = hp.n_blocks * hp.HH;
  for (int i = threadIdx.x; i < HKa + 24; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x % 32;
  const long long warp = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const long long nwarps = (long long)gridDim.x * (blockDim.x / 32);
  float bsum[20];                               // per-warp sums of dpre (uniform across lanes)
#pragma unroll
  for (int o = 0; o < 20; ++o) bsum[o] = 0.f;
  for (long long p0 = warp * kQP; p0 < P; p0 += nwarps * kQP) {
    long long pq[kQP]; bool ok[kQP];
#pragma unroll
    for (int q = 0; q < kQP; ++q) { ok[q] = p0 + q < P; pq[q] = min(p0 + q, P - 1); }
    float dpre[kQP][kMaxOut], dsig[kQP], dv[kQP][3];
#pragma unroll
    for (int q = 0; q < kQP; ++q) {
      const float* row = out + pq[q] * pitch;
      const float* grow = g_out + pq[q] * pitch;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) {
        dpre[q][o] = 0.f;
        if (o < hp.n_out && ok[q]) {
          const OutDesc& d = hp.o[o];
          float g = 0.f;
          for (int c = 0; c < d.rep; ++c) g += grow[d.ch + c];
          float sg, scale;
          const float v = row[d.ch];
          if (d.xform == XF_K) { sg = (v - 1.0f) * 0.5f + 0.5f; scale = 2.0f; }
          else if (d.xform == XF_THETA_RPV) { sg = v * 0.5f + 0.5f; scale = 2.0f; }
          else if (d.xform == XF_THETA_H) { const float k = (float)(M_PI * 30.0 / 180.0); sg = v / k; scale = k; }
          else { sg = v; scale = 1.0f; }
          dpre[q][o] = d.xform == XF_SOFTPLUS ? g * (1.0f - expf(-v)) : g * scale * sg * (1.0f - sg);
        }
      }
      // softplus'(x) = 1 - exp(-softplus(x))
      dsig[q] = ok[q] ? grow[hp.ch_sigma] * (1.0f - expf(-row[hp.ch_sigma])) : 0.f;
      dv[q][0] = dv[q][1] = dv[q][2] = 0.f;
    }
    if (hp.ch_nlr >= 0) {
      // n = -v/|v| ; recompute v = Wg h + bg
      float g_acc[kQP][3];
#pragma unroll
      for (int q = 0; q < kQP; ++q) g_acc[q][0] = g_acc[q][1] = g_acc[q][2] = 0.f;
      for (int i = lane * 8; i < F; i += 256) {
        float h[kQP][8];
#pragma unroll
        for (int q = 0; q < kQP; ++q) load8g<T>(Hlast + pq[q] * ldh + i, h[q]);
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          float w[8]; load8<float>(params + hp.wg + (long long)o * F + i, w);
#pragma unroll
          for (int q = 0; q < kQP; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) g_acc[q][o] = fmaf(h[q][j], w[j], g_acc[q][o]);
        }
      }
#pragma unroll
      for (int q = 0; q < kQP; ++q) {
        float v[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) v[o] = warp_sum(g_acc[q][o]) + __ldg(params + hp.bg + o);
        const float* grow = g_out + pq[q] * pitch;
        const float gn[3] = {grow[hp.ch_nlr], grow[hp.ch_nlr + 1], grow[hp.ch_nlr + 2]};
        const float sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        if (sq > 1.1920929e-07f) {
          const float inv = 1.0f / sqrtf(sq);
          const float u[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
          const float gu = gn[0] * u[0] + gn[1] * u[1] + gn[2] * u[2];
#pragma unroll
          for (int o = 0; o < 3; ++o) dv[q][o] = ok[q] ? -(gn[o] - gu * u[o]) * inv : 0.f;
        } else {
          const float inv = 1.0f / sqrtf(1.1920929e-07f);
#pragma unroll
          for (int o = 0; o < 3; ++o) dv[q][o] = ok[q] ? -gn[o] * inv : 0.f;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kQP; ++q) {
#pragma unroll
      for (int o = 0; o < 16; ++o) bsum[o] += dpre[q][o];
      bsum[16] += dsig[q]; bsum[17] += dv[q][0]; bsum[18] += dv[q][1]; bsum[19] += dv[q][2];
      if (ok[q] && lane < 8) {
        // row of 64: lane l writes columns 8l..8l+7
        float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (lane == 0) {
#pragma unroll
          for (int o = 0; o < 8; ++o) t[o] = dpre[q][o];
        } else if (lane == 1) {
#pragma unroll
          for (int o = 0; o < 8; ++o) t[o] = dpre[q][8 + o];
        } else if (lane == 2) { t[0] = dsig[q]; t[1] = dv[q][0]; t[2] = dv[q][1]; t[3] = dv[q][2]; }
        Pack<T, 8>::store(DPRE + pq[q] * 64 + lane * 8, t);
      }
    }
    if constexpr (!kLight) {
    for (int i = lane * 8; i < F; i += 256) {
      float ws[8]; load8<float>(params + hp.wsig + i, ws);
      float g[kQP][8];
#pragma unroll
      for (int q = 0; q < kQP; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) g[q][j] = dsig[q] * ws[j];
      if (hp.ch_nlr >= 0) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          float w[8]; load8<float>(params + hp.wg + (long long)o * F + i, w);
#pragma unroll
          for (int q = 0; q < kQP; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) g[q][j] = fmaf(dv[q][o], w[j], g[q][j]);
        }
      }
#pragma unroll
      for (int q = 0; q < kQP; ++q) if (ok[q]) Pack<T, 8>::store(G7D + pq[q] * F + i, g[q]);
    }
    for (int b = 0; b < hp.n_blocks; ++b) {
      for (int i = lane * 8; i < hp.HH; i += 256) {
        float a[kQP][8];
#pragma unroll
        for (int q = 0; q < kQP; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) a[q][j] = 0.f;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          if (o < hp.n_out && hp.o[o].block == b) {
            float w[8]; load8<float>(params + hp.o[o].w_off + i, w);
#pragma unroll
            for (int q = 0; q < kQP; ++q)
#pragma unroll
              for (int j = 0; j < 8; ++j) a[q][j] = fmaf(dpre[q][o], w[j], a[q][j]);
          }
        }
        float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < kQP; ++q) {
          float c[8]; load8g<T>(CD + pq[q] * ldd + (long long)b * hp.HH + i, c);
#pragma unroll
          for (int j = 0; j < 8; ++j) { a[q][j] *= c[j]; cs[j] += ok[q] ? a[q][j] : 0.f; }
          if (ok[q]) Pack<T, 8>::store(GHD + pq[q] * ldd + (long long)b * hp.HH + i, a[q]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&s_red[b * hp.HH + i + j], cs[j]);
      }
    }
    }  // !kLight
  }
  if (lane == 0) {
#pragma unroll
    for (int o = 0; o < 20; ++o) atomicAdd(&s_red[HKa + o], bsum[o]);
  }
  __syncthreads();
  if constexpr (!kLight)
    for (int i = threadIdx.x; i < HKa; i += blockDim.x) atomicAdd(g_params + hp.b1_off[i / hp.HH] + (i % hp.HH), s_red[i]);
  if (threadIdx.x < hp.n_out) atomicAdd(g_params + hp.o[threadIdx.x].b_off, s_red[HKa + threadIdx.x]);
  if (threadIdx.x == 16) atomicAdd(g_params + hp.bsig, s_red[HKa + 16]);
  if (threadIdx.x >= 17 && threadIdx.x < 20 && hp.ch_nlr >= 0) atomicAdd(g_params + hp.bg + (threadIdx.x - 17), s_red[HKa + threadIdx.x]);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float wd, float bc1, float bc2_sqrt, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * gscale;
  float pi = p[i];
  if (wd != 0.f) gi += wd * pi;
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

// Graph-replayable Adam: the step counter and the learning rate live in device memory (state = {lr, step, block counter, -}),
// so a captured training step advances them itself.  ONE launch: every block derives the bias corrections of update
// t = step + 1, applies it to its 4096 elements, and the block that finishes last stores the new step — after every block
// has read the old one.  The corrections 1 - b^t are -expm1f(t ln b) with ln b split into two floats: no cancellation (the
// host path's 1.0f - (float)pow(b, t) carries up to 3e-5 of relative rounding at t = 1, this form 1e-7) and no fp64 on the
// device (two double-precision pow per block cost 11 us per step on B200: measured, profiles/r02r_launches_step_summary.txt).
constexpr int kAdamThreads = 1024, kAdamPerThread = 4;
__global__ void __launch_bounds__(kAdamThreads) adam_state_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                  float* __restrict__ v, long long n, float* __restrict__ state, float b1,
                                                                  float b2, float eps, float wd, float gscale, float ln_b1_hi,
                                                                  float ln_b1_lo, float ln_b2_hi, float ln_b2_lo) {
  __shared__ float s_bc[3];
  if (threadIdx.x == 0) {
    const float t = state[1] + 1.0f;                 // exact up to 2^24 updates
    s_bc[0] = t;
    s_bc[1] = -expm1f(fmaf(t, ln_b1_hi, t * ln_b1_lo));
    s_bc[2] = sqrtf(-expm1f(fmaf(t, ln_b2_hi, t * ln_b2_lo)));
  }
  __syncthreads();
  const float lr = state[0], bc1 = s_bc[1], bc2_sqrt = s_bc[2];
  const long long i0 = ((long long)blockIdx.x * kAdamThreads + threadIdx.x) * kAdamPerThread;
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    if (wd != 0.f) gi += wd * pi;
    mi = b1 * mi + (1.0f - b1) * gi;
    vi = b2 * vi + (1.0f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - (lr / bc1) * (mi / denom);
  };
  if (i0 + kAdamPerThread <= n) {
    float4 P4 = *reinterpret_cast<const float4*>(p + i0), M4 = *reinterpret_cast<const float4*>(m + i0);
    float4 V4 = *reinterpret_cast<const float4*>(v + i0);
    const float4 G4 = *reinterpret_cast<const float4*>(g + i0);
    update(P4.x, G4.x, M4.x, V4.x); update(P4.y, G4.y, M4.y, V4.y); update(P4.z, G4.z, M4.z, V4.z); update(P4.w, G4.w, M4.w, V4.w);
    *reinterpret_cast<float4*>(p + i0) = P4; *reinterpret_cast<float4*>(m + i0) = M4; *reinterpret_cast<float4*>(v + i0) = V4;
  } else {
    for (long long i = i0; i < n; ++i) update(p[i], g[i], m[i], v[i]);
  }
  // the last block to finish publishes the new step count (and leaves the counter at zero for the next launch)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(state + 2);
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) { *counter = 0u; state[1] = s_bc[0]; }
  }
}

// tcgen05 mode, no learned normal: one thread per point turns the gradient of the packed output row into
// the pre-activation gradients DPRE[p][0..15] (heads' second layers), DPRE[p][16] (sigma) — everything
// else in the backward of the heads is a GEMM on DPRE.  Rows leave through shared memory so that the
// [P,64] bf16 matrix is written with full 16-byte coalesced stores; the scalar bias gradients are
// reduced per block.
__global__ void __launch_bounds__(256) heads_dpre_kernel(HeadPlan hp, const float* __restrict__ out,
                                                         const float* __restrict__ g_out, int pitch,
                                                         __nv_bfloat16* __restrict__ DPRE, float* __restrict__ g_params,
                                                         long long P) {
  __shared__ __align__(16) __nv_bfloat16 rows[256][64 + 8];     // +8: rows 144 B apart -> conflict-free 16 B accesses
  __shared__ float bred[kMaxOut + 1];
  const int tid = threadIdx.x;
  if (tid <= kMaxOut) bred[tid] = 0.f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + tid;
  float d[kMaxOut + 1];
#pragma unroll
  for (int o = 0; o <= kMaxOut; ++o) d[o] = 0.f;
  if (p < P) {
    const float* row = out + p * pitch;
    const float* grow = g_out + p * pitch;
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
      if (o < hp.n_out) {
        const OutDesc& q = hp.o[o];
        float g = 0.f;
        for (int c = 0; c < q.rep; ++c) g += grow[q.ch + c];
        const float v = row[q.ch];
        float sg, scale;
        if (q.xform == XF_K) { sg = (v - 1.0f) * 0.5f + 0.5f; scale = 2.0f; }
        else if (q.xform == XF_THETA_RPV) { sg = v * 0.5f + 0.5f; scale = 2.0f; }
        else if (q.xform == XF_THETA_H) { const float k = (float)(M_PI * 30.0 / 180.0); sg = v / k; scale = k; }
        else { sg = v; scale = 1.0f; }
        d[o] = q.xform == XF_SOFTPLUS ? g * (1.0f - expf(-v)) : g * scale * sg * (1.0f - sg);
      }
    }
    d[kMaxOut] = grow[hp.ch_sigma] * (1.0f - expf(-row[hp.ch_sigma]));      // softplus'(x) = 1 - exp(-softplus(x))
  }
  uint32_t* r32 = reinterpret_cast<uint32_t*>(&rows[tid][0]);
#pragma unroll
  for (int j = 0; j < 8; ++j) r32[j] = tc::bf_pack(d[2 * j], d[2 * j + 1]);
  r32[8] = tc::bf_pack(d[kMaxOut], 0.f);
#pragma unroll
  for (int j = 9; j < 32; ++j) r32[j] = 0u;
  // block sums of every column that is a bias gradient
#pragma unroll
  for (int o = 0; o <= kMaxOut; ++o) {
    if (o < hp.n_out || o == kMaxOut) {
      const float sum = warp_sum(d[o]);
      if ((tid & 31) == 0) atomicAdd(&bred[o], sum);
    }
  }
  __syncthreads();
  // 256 rows x 8 chunks of 16 B, consecutive threads write consecutive chunks
  const long long p0 = (long long)blockIdx.x * 256;
  for (int c = tid; c < 256 * 8; c += 256) {
    const int r = c >> 3, j = c & 7;
    if (p0 + r < P)
      *reinterpret_cast<uint4*>(DPRE + (p0 + r) * 64 + j * 8) = *reinterpret_cast<const uint4*>(&rows[r][j * 8]);
  }
  if (tid < hp.n_out) atomicAdd(g_params + hp.o[tid].b_off, bred[tid]);
  if (tid == kMaxOut) atomicAdd(g_params + hp.bsig, bred[kMaxOut]);
}

// W2p[b*HH + i][o] = W2_o[i] when output o hangs off block b, else 0  (o < 64; [n_blocks*HH, 64] bf16)
__global__ void pack_w2_kernel(HeadPlan hp, const float* __restrict__ params, __nv_bfloat16* __restrict__ W2p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = hp.n_blocks * hp.HH;
  if (idx >= rows * 64) return;
  const int r = idx / 64, o = idx % 64;
  float v = 0.f;
  if (o < hp.n_out && hp.o[o].block == r / hp.HH) v = params[hp.o[o].w_off + (r % hp.HH)];
  W2p[idx] = __float2bfloat16_rn(v);
}

// forward operands of the heads' second layers as tcgen05 B matrices (N = 64 padded outputs, K-major):
//   W2pT [64, ldk] : row o = W2_o laid over the columns of its hidden block, zeros elsewhere
//   Wsig [64, F]   : row 0 = w_sigma, rows 1..3 = grad_from_xyz (learned normal) when evaluated
//   W2p  [hk, 64]  : the backward's copy (pack_w2_kernel), written here as well when the forward is a training forward
__global__ void pack_heads_fwd_kernel(HeadPlan hp, const float* __restrict__ params, __nv_bfloat16* __restrict__ W2pT,
                                      int ldk, __nv_bfloat16* __restrict__ Wsig, int F, __nv_bfloat16* __restrict__ W2p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int hk = hp.n_blocks * hp.HH;
  if (W2p != nullptr && idx < hk * 64) {
    const int r = idx / 64, o = idx % 64;
    float v = 0.f;
    if (o < hp.n_out && hp.o[o].block == r / hp.HH) v = params[hp.o[o].w_off + (r % hp.HH)];
    W2p[idx] = __float2bfloat16_rn(v);
  }
  if (idx < 64 * hk) {
    const int o = idx / hk, cix = idx % hk;
    float v = 0.f;
    if (o < hp.n_out && hp.o[o].block == cix / hp.HH) v = params[hp.o[o].w_off + (cix % hp.HH)];
    W2pT[(long long)o * ldk + cix] = __float2bfloat16_rn(v);
  }
  if (idx < 64 * F) {
    const int r = idx / F, i = idx % F;
    float v = 0.f;
    if (r == 0) v = params[hp.wsig + i];
    else if (r <= 3 && hp.ch_nlr >= 0) v = params[hp.wg + (long long)(r - 1) * F + i];
    Wsig[idx] = __float2bfloat16_rn(v);
  }
}

// direct epilogue of the heads GEMM: columns 0..n_out-1 of the accumulator row are the second-layer
// pre-activations; bias, sigmoid and the output transforms are applied and the packed row is written
struct EpiHeadsOut {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;
  HeadPlan hp; const float* params; float* out; int pitch; int M;
  const float* sig;               // density of every row, already computed by the fused trunk kernel (nullptr: not available)
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (col0 != 0 || row >= M) return;
    float* orow = out + (long long)row * pitch;
    if (sig) orow[hp.ch_sigma] = __ldg(sig + row);
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) {
      if (o < hp.n_out) {
        const OutDesc& d = hp.o[o];
        const float pre = acc[o] + __ldg(params + d.b_off);
        const float sg = 1.0f / (1.0f + expf(-pre));
        float v = sg;
        if (d.xform == XF_SOFTPLUS) v = pre > 20.f ? pre : log1pf(expf(pre));
        else if (d.xform == XF_K) v = (sg - 0.5f) * 2.0f + 1.0f;
        else if (d.xform == XF_THETA_RPV) v = (sg - 0.5f) * 2.0f;
        else if (d.xform == XF_THETA_H) v = sg * (float)(M_PI * 30.0 / 180.0);
        for (int c = 0; c < d.rep; ++c) orow[d.ch + c] = v;
      }
    }
  }
};

// direct epilogue of the sigma GEMM: column 0 = w_sigma . h, columns 1..3 = grad_from_xyz . h (learned normal)
struct EpiSigmaOut {
  static constexpr int kMode = 0, kIn = 0, kOut = 0;
  const float* bsig; const float* bg; float* out; int pitch, ch_sigma, ch_nlr, M;
  template <int n> __device__ __forceinline__ void apply(int row, int col0, const float (&acc)[n]) const {
    if (col0 != 0 || row >= M) return;
    float* orow = out + (long long)row * pitch;
    if (ch_sigma >= 0) {
      const float sv = acc[0] + __ldg(bsig);
      orow[ch_sigma] = sv > 20.f ? sv : log1pf(expf(sv));
    }
    if (ch_nlr >= 0) {
      const float g0 = acc[1] + __ldg(bg), g1 = acc[2] + __ldg(bg + 1), g2 = acc[3] + __ldg(bg + 2);
      const float inv = 1.0f / sqrtf(fmaxf(g0 * g0 + g1 * g1 + g2 * g2, 1.1920929e-07f));
      orow[ch_nlr] = -g0 * inv; orow[ch_nlr + 1] = -g1 * inv; orow[ch_nlr + 2] = -g2 * inv;
    }
  }
};

static int build_plan(const bn_mlp* h, int flags, HeadPlan* hp, int* n_channels) {
  HeadPlan p{};
  const bn_mlp_cfg& c = h->cfg;
  p.HH = h->HH;
  p.wsig = c.w_off[BN_LIN_SIGMA]; p.bsig = c.b_off[BN_LIN_SIGMA];
  p.wg = c.w_off[BN_LIN_GRAD]; p.bg = c.b_off[BN_LIN_GRAD];
  int ch = 0;
  // albedo: rgb_from_xyzdir.2 rows 0..2 on block 0
  for (int o = 0; o < 3; ++o)
    p.o[p.n_out++] = OutDesc{0, c.w_off[BN_LIN_RGB2] + (long long)o * h->HH, c.b_off[BN_LIN_RGB2] + o, ch++, 1, XF_SIGMOID};
  p.ch_sigma = ch++;
  const bool want_beta = (flags & BN_MLP_BETA) != 0;
  if (want_beta && c.head_dim[BN_HEAD_BETA] <= 0) { set_error("BN_MLP_BETA requested but the model has no beta_from_xyz head"); return BN_ERR_ARG; }
  const int ch_beta = want_beta ? ch++ : -1;    // spsbrdfnerf.py:156-158: the uncertainty channel precedes the normals
  if (flags & BN_MLP_NORMAL_AN) ch += 3;        // written by the analytic-normal sweep
  p.ch_nlr = -1;
  if (flags & BN_MLP_NORMAL_LR) {
    if (!c.normal_lr) { set_error("learned normal requested but the model has no grad_from_xyz head"); return BN_ERR_ARG; }
    p.ch_nlr = ch; ch += 3;
  }
  int last_block = 0;
  auto add_head = [&](int head, int xform) {
    int blk = -1;
    for (int b = 0; b < h->n_blocks; ++b) if (h->blk_head[b] == head) blk = b;
    if (blk < 0) return;
    const int dim = c.head_dim[head];
    const int lin2 = h->blk_lin2[blk];
    if (head == BN_HEAD_BETA) {
      p.o[p.n_out++] = OutDesc{blk, c.w_off[lin2], c.b_off[lin2], ch_beta, 1, xform};
    } else if (head == BN_HEAD_ROUGH || head == BN_HEAD_THETA) {
      p.o[p.n_out++] = OutDesc{blk, c.w_off[lin2], c.b_off[lin2], ch, 1, xform}; ch += 1;
    } else if (dim == 1) {
      p.o[p.n_out++] = OutDesc{blk, c.w_off[lin2], c.b_off[lin2], ch, 3, xform}; ch += 3;
    } else {
      for (int o = 0; o < 3; ++o)
        p.o[p.n_out++] = OutDesc{blk, c.w_off[lin2] + (long long)o * h->HH, c.b_off[lin2] + o, ch + o, 1, xform};
      ch += 3;
    }
    last_block = max(last_block, blk);
  };
  if (want_beta) add_head(BN_HEAD_BETA, XF_SOFTPLUS);
  if (flags & BN_MLP_ROUGH) add_head(BN_HEAD_ROUGH, XF_SIGMOID);
  else if (flags & BN_MLP_RPV) { add_head(BN_HEAD_K, XF_K); add_head(BN_HEAD_THETA_RPV, XF_THETA_RPV); add_head(BN_HEAD_RHOC, XF_SIGMOID); }
  else if (flags & BN_MLP_HAPKE) {
    add_head(BN_HEAD_B, XF_SIGMOID); add_head(BN_HEAD_C, XF_SIGMOID);
    if (flags & BN_MLP_HAPKE_THETA) add_head(BN_HEAD_THETA, XF_THETA_H);
  }
  p.n_blocks = last_block + 1;
  for (int b = 0; b < h->n_blocks; ++b) p.b1_off[b] = c.b_off[h->blk_lin0[b]];
  if (hp) *hp = p;
  if (n_channels) *n_channels = ch;
  return BN_OK;
}

// all packed copies of one optimizer step in ONE launch: blockIdx.y = job (a Linear layer or a bias copy)
struct PackJob { long long w_off; int N, Kreal, E, Kpad; void* Wp; long long ldp; void* WTp; long long ldt; int row0; int Epad; };
constexpr int kMaxPackJobs = 40;
struct PackJobs { int n; PackJob j[kMaxPackJobs]; };

// bias blocks of the fused trunk kernels (see mlp_chain.cuh): (layer row, column) elements of Wb [L * F, 64], grid-stride over
// the blocks of one pack job
__device__ __forceinline__ void pack_bias_blocks(const float* __restrict__ params, __nv_bfloat16* __restrict__ Wb, int L, int F, int E,
                                                 int skip, const long long* __restrict__ offs) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < L * F * 64; idx += gridDim.x * blockDim.x) {
    const int col = idx & 63, row = (idx >> 6) % F, l = idx / (64 * F);
    const long long w_off = offs[2 * l], b_off = offs[2 * l + 1];
    float v = 0.f;
    if (col < E && (l == 0 || l == skip)) {
      const int kreal = l == 0 ? E : E + F;                    // the encoding is the FIRST E input columns of both layers
      v = params[w_off + (long long)row * kreal + col];
    } else if (col == chain::kOneCol) {
      v = params[b_off + row];
    } else if (col == chain::kOneCol + 1) {
      const float b = params[b_off + row];
      v = b - __bfloat162float(__float2bfloat16(b));
    }
    Wb[idx] = __float2bfloat16(v);
  }
}

// One 64 x 64 tile (n x padded k) per block iteration: the fp32 rows are read coalesced along k, the row-major copy Wp is
// written coalesced along k, and the transposed copy WTp leaves through shared memory coalesced along n (a thread-per-
// element version wrote WTp with a 2-byte stride-ldt pattern and took 29 us per step for 2.7 M weights).
constexpr int kPackTile = 64;
template <typename T>
__global__ void __launch_bounds__(256) pack_all_kernel(const __grid_constant__ PackJobs jobs, const float* __restrict__ params) {
  __shared__ float tile[kPackTile][kPackTile + 1];
  const PackJob& q = jobs.j[blockIdx.y];
  if (q.Kreal == -2) {                            // bias blocks of the fused trunk kernels (N = L, ldp = F, Kpad = skip, WTp = offsets)
    if constexpr (std::is_same<T, __nv_bfloat16>::value)
      pack_bias_blocks(params, reinterpret_cast<__nv_bfloat16*>(q.Wp), q.N, (int)q.ldp, q.E, q.Kpad, reinterpret_cast<const long long*>(q.WTp));
    return;
  }
  if (q.Kreal < 0) {                              // bias copy job: Wp is a float* destination, N values
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < q.N; n += gridDim.x * blockDim.x)
      reinterpret_cast<float*>(q.Wp)[q.row0 + n] = params[q.w_off + n];
    return;
  }
  const int tiles_k = (q.Kpad + kPackTile - 1) / kPackTile, tiles_n = (q.N + kPackTile - 1) / kPackTile;
  const int tx = threadIdx.x % kPackTile, ty = threadIdx.x / kPackTile;       // 64 x 4
  for (int t = blockIdx.x; t < tiles_k * tiles_n; t += gridDim.x) {
    const int n0 = (t / tiles_k) * kPackTile, k0 = (t % tiles_k) * kPackTile;
    const int kp = k0 + tx;
    int k = -1;                                   // source column of padded column kp, -1 = padding
    if (kp < q.Kpad) {
      if (q.E >= 0) { if (kp < q.E) k = kp; else if (kp >= q.Epad) k = kp - (q.Epad - q.E); }     // E real columns, padding up to Epad
      else k = kp;
      if (k >= q.Kreal) k = -1;
    }
#pragma unroll 4
    for (int r = ty; r < kPackTile; r += 4) {
      const int n = n0 + r;
      const float v = (k >= 0 && n < q.N) ? params[q.w_off + (long long)n * q.Kreal + k] : 0.f;
      tile[r][tx] = v;
      if (q.Wp && n < q.N && kp < q.Kpad) reinterpret_cast<T*>(q.Wp)[(long long)(q.row0 + n) * q.ldp + kp] = from_f<T>(v);
    }
    __syncthreads();
    if (q.WTp) {
      const int n = n0 + tx;
#pragma unroll 4
      for (int r = ty; r < kPackTile; r += 4) {
        const int kk = k0 + r;
        if (n < q.N && kk < q.Kpad) reinterpret_cast<T*>(q.WTp)[(long long)kk * q.ldt + q.row0 + n] = from_f<T>(tile[tx][r]);
      }
    }
    __syncthreads();
  }
}

template <typename T>
static int sync_weights_t(bn_mlp* h, const float* params, cudaStream_t s) {
  const bn_mlp_cfg& c = h->cfg;
  PackJobs jobs{};
  auto pack = [&](int lin, int N, int Kreal, int E, int Kpad, void* Wp, long long ldp, void* WTp, long long ldt, int row0) {
    jobs.j[jobs.n++] = PackJob{c.w_off[lin], N, Kreal, E, Kpad, Wp, ldp, WTp, ldt, row0, kEncPad};
  };
  for (int l = 0; l < h->L; ++l) {
    const bool enc_in = (l == 0 || l == h->skip);
    pack(BN_LIN_TRUNK0 + l, h->F, h->Kreal[l], enc_in ? h->E : -1, h->Kpad[l], h->Wp[l], h->Kpad[l], h->WTp[l], h->F, 0);
  }
  pack(BN_LIN_FEATS, h->F, h->F, -1, h->F, h->Wf, h->F, h->WfT, h->F, 0);
  const long long HK = (long long)h->n_blocks * h->HH;
  for (int b = 0; b < h->n_blocks; ++b) {
    if (h->blk_head[b] == BN_HEAD_BETA) {       // [features | t]: the time embedding sits kTOff columns behind the features
      pack(h->blk_lin0[b], h->HH, h->F + h->TE, h->F, h->ldfe, h->W1, h->ldfe, h->W1T, HK, b * h->HH);
      jobs.j[jobs.n - 1].Epad = h->F + kTOff;
    } else {
      pack(h->blk_lin0[b], h->HH, h->F + (b == 0 ? h->DE : 0), -1, h->ldfe, h->W1, h->ldfe, h->W1T, HK, b * h->HH);
    }
    jobs.j[jobs.n++] = PackJob{c.b_off[h->blk_lin0[b]], h->HH, -1, -1, 1, h->b1cat, 0, nullptr, 0, b * h->HH, kEncPad};
  }
  if (h->bf16) pack(BN_LIN_SIGMA, 1, h->F, -1, h->F, h->WsigA, h->F, nullptr, 0, 0);   // row 0 of the [64, F] density operand
  if (h->bf16 && h->Wb)
    jobs.j[jobs.n++] = PackJob{0, h->L, -2, h->E, h->skip, h->Wb, h->F, (void*)h->Wb_offs, 0, 0, kEncPad};
  if (jobs.n > kMaxPackJobs) { set_error("too many pack jobs"); return BN_ERR_STATE; }
  pack_all_kernel<T><<<dim3(64, jobs.n), 256, 0, s>>>(jobs, params);
  if (int rc = after_launch("pack_all_kernel")) return rc;
  h->synced = true;
  h->w2p_flags = -1;                      // the heads' second-layer operands are re-packed by the next heads forward
  return BN_OK;
}

static unsigned heads_grid(const bn_mlp* h, long long P) {
  return (unsigned)max(1LL, min(ceil_div_ll(P, 8 * kQP), (long long)h->num_sms * 4));
}

// density pass as ONE fused kernel (mlp_chain.cuh); tcgen05 mode, 512-wide trunk
static int sigma_chain(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs, int d_stride,
                       const float* z, int N, int S, float* out, cudaStream_t s) {
  const bn_mlp_cfg& c = h->cfg;
  chain::SigmaChainParams prm;
  for (int l = 0; l < h->L; ++l) {
    if (int rc = tc::make_map_bf16(&prm.wmap[l], h->Wp[l], h->F, h->Kpad[l], h->Kpad[l], 64, 128)) return rc;
  }
  if (int rc = tc::make_map_bf16(&prm.bmap, h->Wb, (long long)h->L * h->F, 64, 64, 64, 128)) return rc;
  prm.wsig = params + c.w_off[BN_LIN_SIGMA]; prm.bsig = params + c.b_off[BN_LIN_SIGMA];
  prm.origins = origins; prm.dirs = dirs; prm.z = z; prm.out = out;
  prm.P = (long long)N * S; prm.o_stride = o_stride; prm.d_stride = d_stride; prm.S = S;
  prm.L = h->L; prm.skip = h->skip; prm.n_freq = c.n_freq_xyz;
  prm.trace = h->chain_trace;
  prm.pol_w = tc::pol_weights(); prm.pol_s = tc::pol_stream();
  const int n_blocks = (int)ceil_div_ll(prm.P, 256);
  constexpr int smem = chain::sigma_chain_smem();
  BN_CUDA(cudaFuncSetAttribute(chain::sigma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * min(n_blocks, h->num_sms / 2));
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc::pdl_enabled() ? 2 : 1;
  prof_begin(2, 2.0 * (double)prm.P * ((double)h->E * h->F + (double)(h->L - 2) * h->F * h->F + (double)(h->F + h->E) * h->F + h->F), s);
  BN_CUDA(cudaLaunchKernelEx(&cfg, chain::sigma_chain_kernel, prm));
  const int rc = after_launch("sigma_chain_kernel");
  prof_end(s);
  return rc;
}

// trunk of a training (or analytic-normal) forward as ONE fused kernel: X3, H_l, C_l of every layer are
// written for the backward pass, the layer inputs themselves never leave the SM
static int train_chain(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs, int d_stride,
                       const float* z, int N, int S, const Ws<__nv_bfloat16>& w, bool keep_c, bool train, float* sigma_out,
                       cudaStream_t s) {
  const bn_mlp_cfg& c = h->cfg;
  const long long P = (long long)N * S;
  chain::TrainChainParams prm;
  for (int l = 0; l < h->L; ++l) {
    if (int rc = tc::make_map_bf16(&prm.wmap[l], h->Wp[l], h->F, h->Kpad[l], h->Kpad[l], 64, 128)) return rc;
    if (int rc = tc::stream_map(&prm.hmap[l], w.H[l], P, h->F, w.Hld[l])) return rc;
    if (keep_c) { if (int rc = tc::make_map_bf16_sw(&prm.cmap[l], w.C[l], P, h->F, h->F, 32, 32, 64)) return rc; }
    else memset(&prm.cmap[l], 0, sizeof(prm.cmap[l]));
  }
  if (int rc = tc::make_map_bf16(&prm.bmap, h->Wb, (long long)h->L * h->F, 64, 64, 64, 128)) return rc;
  if (int rc = tc::stream_map(&prm.x3map, w.X3, P, kEncPad, w.ldx3)) return rc;
  prm.origins = origins; prm.dirs = dirs; prm.z = z;
  prm.P = P; prm.o_stride = o_stride; prm.d_stride = d_stride; prm.S = S;
  prm.L = h->L; prm.skip = h->skip; prm.n_freq = c.n_freq_xyz;
  prm.store_c = keep_c ? 1 : 0; prm.h_from = train ? 0 : h->L - 1;
  prm.trace = h->chain_trace;
  prm.wsig = params + c.w_off[BN_LIN_SIGMA]; prm.bsig = params + c.b_off[BN_LIN_SIGMA]; prm.sig_out = chain_sigma_ok(h) ? w.SIGC : nullptr;
  prm.sig_out2 = prm.sig_out ? sigma_out : nullptr;
  prm.pol_w = tc::pol_weights(); prm.pol_s = tc::pol_stream();
  const int n_blocks = (int)ceil_div_ll(P, 256);
  constexpr int smem = chain::chain_smem<true>();
  // experiment knob: one weight stage traded for a second cosine staging box per epilogue warp
  static const bool cbox2 = getenv("BN_CHAIN_CBOX2") != nullptr;
  auto chain_kernel = cbox2 ? chain::train_chain_kernel<true> : chain::train_chain_kernel<false>;
  BN_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * min(n_blocks, h->num_sms / 2));
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = tc::pdl_enabled() ? 2 : 1;
  prof_begin(2, 2.0 * (double)P * ((double)h->E * h->F + (double)(h->L - 2) * h->F * h->F + (double)(h->F + h->E) * h->F), s);
  BN_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel, prm));
  const int rc = after_launch("train_chain_kernel");
  prof_end(s);
  return rc;
}

// rows [row0, row0 + n) of the per-point forward buffers (the workspace is carved for the whole call)
template <typename T>
static void offset_rows(const bn_mlp* h, Ws<T>* w, long long row0) {
  if (row0 == 0) return;
  w->X3 += row0 * w->ldx3;
  if (w->FE) w->FE += row0 * w->ldfe;
  for (int l = 0; l < h->L; ++l) {
    w->H[l] += row0 * w->Hld[l];
    if (w->C[l]) w->C[l] += row0 * h->F;
  }
  if (w->SIGC) w->SIGC += row0;
}

// PE + trunk of N*S points whose activations land in `w` (already offset to the first row of this call)
template <typename T>
static int trunk_t(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs, int d_stride,
                   const float* z, int N, int S, bool keep_c, bool train, const Ws<T>& w, cudaStream_t s,
                   float* sigma_out = nullptr) {
  const long long P = (long long)N * S;
  const bn_mlp_cfg& c = h->cfg;
  const int F = h->F, L = h->L;
  if (h->DE > 0 && w.FE) {      // --input_viewdir: the encoded ray direction of these rows, next to their (later) features
    dir_encode_kernel<T><<<(unsigned)ceil_div_ll(P, 128), 128, 0, s>>>(dirs, d_stride, S, P, c.n_freq_dir, w.FE + F, w.ldfe,
                                                                       h->TE > 0 ? kTOff : kEncPad);
    BN_LAUNCH_CHECK();
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (train_chain_ok(h))
      return train_chain(h, params, origins, o_stride, dirs, d_stride, z, N, S, w, keep_c, train, sigma_out, s);
  }
  encode_kernel<T><<<(unsigned)ceil_div_ll(P, 128), 128, 0, s>>>(origins, o_stride, dirs, d_stride, z, S, P,
                                                               c.n_freq_xyz, w.X3, w.ldx3);
  BN_LAUNCH_CHECK();
  for (int l = 0; l < L; ++l) {
    const T* A; long long lda;
    if (l == 0 || l == h->skip) { A = w.X3; lda = w.ldx3; } else { A = w.H[l - 1]; lda = w.Hld[l - 1]; }
    // first layer: sin(30 lin), |30 lin| <= 30: the MUFU path is exact to ~2e-6 there, far below
    // the bf16 resolution of the stored activation; the fp32 mode keeps the accurate sincosf
    if (int rc = layer_sin<T>(h, A, lda, (const T*)h->Wp[l], h->Kpad[l], P, F, h->Kpad[l], params + c.b_off[l],
                              l == 0 ? 30.0f : 1.0f, w.H[l], w.Hld[l], keep_c ? w.C[l] : nullptr, F, s, h->Kreal[l])) return rc;
  }
  return BN_OK;
}

// sigma = softplus(w_sigma . h_{L-1} + b) of P rows, written with `pitch` floats per row (density of a trunk-only call)
template <typename T>
static int sigma_rows_t(bn_mlp* h, const float* params, const T* Hl, long long ldl, long long P, float* out, int pitch,
                        cudaStream_t s) {
  const bn_mlp_cfg& c = h->cfg;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    EpiSigmaOut es{params + c.b_off[BN_LIN_SIGMA], nullptr, out, pitch, 0, -1, (int)P};
    return gemm_tn<T>(h, Hl, ldl, (const T*)h->WsigA, h->F, P, 64, h->F, es, s, h->F / 64);
  } else {
    HeadPlan hp; int nch;
    if (int rc = build_plan(h, 0, &hp, &nch)) return rc;
    heads_fwd_kernel<T><<<heads_grid(h, P), 256, 0, s>>>(hp, params, Hl, ldl, h->F, nullptr, 0, out, pitch, P, true, false);
    BN_LAUNCH_CHECK();
    return BN_OK;
  }
}

// feature layer + hidden layer of the heads + their second layers / sigma / learned normal for P rows
template <typename T>
static int heads_t(bn_mlp* h, const float* params, long long P, int flags, float* out, int pitch, const Ws<T>& w, cudaStream_t s) {
  const bool train = flags & BN_MLP_TRAIN;
  const bn_mlp_cfg& c = h->cfg;
  const int F = h->F, L = h->L;
  HeadPlan hp; int nch;
  if (int rc = build_plan(h, flags, &hp, &nch)) return rc;
  if (pitch < nch) { set_error("bn_mlp_forward: out_pitch %d < %d channels", pitch, nch); return BN_ERR_ARG; }
  const T* Hl = w.H[L - 1]; const long long ldl = w.Hld[L - 1];
  if (int rc = layer_bias<T>(h, Hl, ldl, (const T*)h->Wf, F, P, F, F, params + c.b_off[BN_LIN_FEATS], w.FE, w.ldfe, s)) return rc;
  {
    const int HKa = hp.n_blocks * h->HH;
    if (int rc = layer_sin<T>(h, w.FE, w.ldfe, (const T*)h->W1, h->ldfe, P, HKa, h->ldfe, h->b1cat, 1.0f, w.HD, w.ldhd,
                              train ? w.CD : nullptr, w.ldhd, s)) return rc;
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // second layers of the heads, sigma and the learned normal as two skinny tcgen05 GEMMs (N = 64) whose
    // epilogues write the packed fp32 rows: every activation row is read exactly once, by the TMA
    const int HKa = hp.n_blocks * h->HH;
    const int ldk = h->n_blocks * h->HH;
    pack_heads_fwd_kernel<<<ceil_div(64 * max(HKa, F), 256), 256, 0, s>>>(hp, params, (__nv_bfloat16*)h->W2pT, ldk,
                                                                          (__nv_bfloat16*)h->Wsig, F,
                                                                          train ? (__nv_bfloat16*)h->W2p : nullptr);
    BN_LAUNCH_CHECK();
    h->w2p_flags = train ? flags : -1;       // host-side note for the backward of the same flags: W2p is current
    // the fused trunk kernel already left the density of every row in w.SIGC (fp32 dot product of the last layer's sines):
    // it goes out with the heads' rows, and the sigma GEMM only remains for the learned normal
    const float* sig = chain_sigma_ok(h) ? w.SIGC : nullptr;
    EpiHeadsOut eh{hp, params, out, pitch, (int)P, sig};
    if (int rc = gemm_tn<T>(h, w.HD, w.ldhd, (const T*)h->W2pT, ldk, P, 64, HKa, eh, s, hp.n_out * h->HH / 64)) return rc;
    if (sig && hp.ch_nlr < 0) return BN_OK;
    EpiSigmaOut es{params + hp.bsig, params + hp.bg, out, pitch, sig ? -1 : hp.ch_sigma, hp.ch_nlr, (int)P};
    return gemm_tn<T>(h, Hl, ldl, (const T*)h->Wsig, F, P, 64, F, es, s, (hp.ch_nlr >= 0 ? 4 : 1) * F / 64);
  } else {
    heads_fwd_kernel<T><<<heads_grid(h, P), 256, 0, s>>>(hp, params, Hl, ldl, F, w.HD, w.ldhd, out, pitch, P, false, false);
    BN_LAUNCH_CHECK();
    return BN_OK;
  }
}

static inline bool keep_cos(int flags) {
  return (flags & BN_MLP_TRAIN) || ((flags & BN_MLP_NORMAL_AN) && !(flags & BN_MLP_SIGMA_ONLY));
}

template <typename T>
static int forward_t(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs,
                     int d_stride, const float* z, int N, int S, int flags, float* out, int pitch, void* wsp,
                     cudaStream_t s) {
  const long long P = (long long)N * S;
  const bool sig_only = flags & BN_MLP_SIGMA_ONLY;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (sig_only && h->F == chain::kF && h->skip >= 1 && h->L <= chain::kMaxLayers && !h->no_chain)
      return sigma_chain(h, params, origins, o_stride, dirs, d_stride, z, N, S, out, s);
  }
  Ws<T> w; carve<T>(h, P, flags, wsp, &w);
  if (int rc = trunk_t<T>(h, params, origins, o_stride, dirs, d_stride, z, N, S, keep_cos(flags), (flags & BN_MLP_TRAIN) != 0, w, s)) return rc;
  if (sig_only) {
    HeadPlan hp; int nch;
    if (int rc = build_plan(h, flags, &hp, &nch)) return rc;
    heads_fwd_kernel<T><<<heads_grid(h, P), 256, 0, s>>>(hp, params, w.H[h->L - 1], w.Hld[h->L - 1], h->F, nullptr, 0, out, 1, P, true, false);
    BN_LAUNCH_CHECK();
    return BN_OK;
  }
  return heads_t<T>(h, params, P, flags, out, pitch, w, s);
}

// trunk of the rows [row0, row0 + N*S) of a workspace carved for `total` points (+ their density)
template <typename T>
static int trunk_rows_t(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs, int d_stride,
                        const float* z, int N, int S, int flags, long long total, long long row0, float* sigma_out,
                        void* wsp, cudaStream_t s) {
  Ws<T> w; carve<T>(h, total, flags, wsp, &w);
  offset_rows<T>(h, &w, row0);
  if (int rc = trunk_t<T>(h, params, origins, o_stride, dirs, d_stride, z, N, S, keep_cos(flags), (flags & BN_MLP_TRAIN) != 0, w, s,
                          sigma_out)) return rc;
  if (sigma_out) {
    if (chain_sigma_ok(h) && w.SIGC) return BN_OK;       // the fused trunk kernel wrote it on the way
    return sigma_rows_t<T>(h, params, w.H[h->L - 1], w.Hld[h->L - 1], (long long)N * S, sigma_out, 1, s);
  }
  return BN_OK;
}

// data gradients of the whole trunk as ONE fused kernel (mlp_dgrad_chain.cuh): dZ_{L-1} = w.GA in, dZ_l -> w.GZ[l], l < L-1
static int dgrad_chain(bn_mlp* h, const Ws<__nv_bfloat16>& w, long long P, cudaStream_t s) {
  chain::DgradChainParams prm;
  memset(&prm, 0, sizeof(prm));
  const int F = h->F;
  for (int l = 1; l < h->L; ++l) {
    const __nv_bfloat16* BT = (const __nv_bfloat16*)h->WTp[l] + (l == h->skip ? (long long)kEncPad * F : 0);
    if (int rc = tc::make_map_bf16(&prm.wmap[l], BT, F, F, F, 64, 128)) return rc;
    if (int rc = tc::stream_map(&prm.cmap[l], w.C[l - 1], P, F, F)) return rc;
    if (int rc = tc::stream_map(&prm.gout[l], w.GZ[l - 1], P, F, F)) return rc;
  }
  if (int rc = tc::make_map_bf16(&prm.gin, w.GA, P, F, F, 64, 128)) return rc;
  prm.P = P; prm.L = h->L;
  prm.pol_w = tc::pol_weights(); prm.pol_s = tc::pol_stream();
  const int n_blocks = (int)ceil_div_ll(P, 256);
  static_assert(chain::dgrad_chain_smem<3, 3>() == chain::dgrad_chain_smem<4, 2>(), "both variants fill the same shared memory");
  constexpr int smem = chain::dgrad_chain_smem<3, 3>();
  // 4 weight stages + 2 c boxes per quadrant measured faster than 3 + 3 (459-498 vs 538-557 us per launch at P = 131 072,
  // profiles/r02h_ab.txt); BN_DCHAIN_W3C3=1 selects the other split for A/B runs
  const bool w3c3 = getenv("BN_DCHAIN_W3C3") != nullptr;
  auto kern = w3c3 ? chain::dgrad_chain_kernel<3, 3> : chain::dgrad_chain_kernel<4, 2>;
  BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * min(n_blocks, h->num_sms / 2));
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  prof_begin(3, 2.0 * (double)P * (double)(h->L - 1) * F * F, s);
  BN_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  const int rc = after_launch("dgrad_chain_kernel");
  prof_end(s);
  return rc;
}

template <typename T>
static int backward_t(bn_mlp* h, const float* params, const float* out, const float* g_out, int pitch, int N, int S,
                      int flags, float* g, void* wsp, cudaStream_t s) {
  const long long P = (long long)N * S;
  const bn_mlp_cfg& c = h->cfg;
  const int F = h->F, L = h->L;
  const bool normals = flags & BN_MLP_NORMAL_AN;   // ZB_l (second-order terms) were left in w.U[l]
  Ws<T> w; carve<T>(h, P, flags, wsp, &w);
  HeadPlan hp; int nch;
  if (int rc = build_plan(h, flags, &hp, &nch)) return rc;
  const T* Hl = w.H[L - 1]; const long long ldl = w.Hld[L - 1];
  const int HKa = hp.n_blocks * h->HH;
  constexpr bool kTC = std::is_same<T, __nv_bfloat16>::value;   // bias grads fused into the tcgen05 dgrad epilogues
  // every weight gradient only shares its inputs with the data-gradient chain: in tcgen05 mode they run on the handle's
  // side stream (forked from `s` by an event each time their input is ready, joined back before returning)
  // per-launch profiling (bench.py roofline leg) times each kernel alone: no concurrent weight-gradient stream then
  const bool overlap = kTC && h->s2 != nullptr && h->overlap && !prof_enabled();
  int n_fork = 0;
  auto fork = [&]() -> cudaStream_t {             // the side stream, ordered after everything enqueued on s so far
    if (!overlap) return s;
    cudaEventRecord(h->ev_h[n_fork], s);
    cudaStreamWaitEvent(h->s2, h->ev_h[n_fork], 0);
    ++n_fork;
    return h->s2;
  };
  if (kTC && hp.ch_nlr < 0) {
    if constexpr (kTC) {
      heads_dpre_kernel<<<(unsigned)ceil_div_ll(P, 256), 256, 0, s>>>(hp, out, g_out, pitch, w.DPRE, g, P);
      BN_LAUNCH_CHECK();
    }
  } else {
    const size_t smem = (size_t)(HKa + 24) * sizeof(float);
    heads_bwd_kernel<T, kTC><<<heads_grid(h, P), 256, smem, s>>>(hp, params, out, g_out, pitch, Hl, ldl, F, w.CD, w.ldhd,
                                                                w.GHD, w.G7D, w.DPRE, g, P);
    BN_LAUNCH_CHECK();
  }
  if constexpr (kTC) {
    // GHD = (DPRE W2p^T) ⊙ CD, first-layer bias gradients = its column sums
    if (h->w2p_flags != flags) {           // not left behind by the forward of this step (other flags, or weights re-packed since)
      pack_w2_kernel<<<ceil_div(HKa * 64, 256), 256, 0, s>>>(hp, params, (__nv_bfloat16*)h->W2p);
      BN_LAUNCH_CHECK();
      h->w2p_flags = flags;
    }
    for (int b = 0; b < hp.n_blocks; ++b) {
      DgradArgs<T> a; a.mulc = w.CD + (long long)b * h->HH; a.ldm = w.ldhd;
      if (int rc = layer_dgrad<T>(h, w.DPRE, 64, (const T*)h->W2p + (long long)b * h->HH * 64, 64, P, h->HH, 64, a,
                                  w.GHD + (long long)b * h->HH, w.ldhd, s)) return rc;
    }
  }
  // second-layer weight gradients of the heads / sigma / learned-normal heads: dW2[o] = DPRE[:,o]^T X
  if constexpr (kTC) {
    EpiSkinny e1{}; e1.n_rows = hp.n_out;
    for (int o = 0; o < hp.n_out; ++o) e1.r[o] = EpiSkinnyRow{g + hp.o[o].w_off, hp.o[o].block * h->HH, (hp.o[o].block + 1) * h->HH};
    cudaStream_t sw = fork();
    if (int rc = gemm_nt<T>(h, w.DPRE, 64, w.HD, w.ldhd, 64, HKa, P, e1, sw, 2.0 * P * hp.n_out * h->HH)) return rc;
    EpiSkinny e2{}; e2.n_rows = 20;
    e2.r[16] = EpiSkinnyRow{g + hp.wsig, 0, F};
    if (hp.ch_nlr >= 0) for (int o = 0; o < 3; ++o) e2.r[17 + o] = EpiSkinnyRow{g + hp.wg + (long long)o * F, 0, F};
    if (int rc = gemm_nt<T>(h, w.DPRE, 64, Hl, ldl, 64, F, P, e2, sw, 2.0 * P * (hp.ch_nlr >= 0 ? 4 : 1) * F)) return rc;
  } else {
    SkinnyPlan sp{};
    for (int o = 0; o < hp.n_out; ++o)
      sp.r[sp.n++] = SkinnyRow{g + hp.o[o].w_off, nullptr, o, hp.o[o].block * h->HH, (hp.o[o].block + 1) * h->HH};
    const int bx = ceil_div(HKa, 256);
    int by = (int)max(1LL, min(ceil_div_ll(P, 128), (long long)(148 * 4 / bx)));
    const long long rows = ceil_div_ll(ceil_div_ll(P, by), 32) * 32;
    by = (int)ceil_div_ll(P, rows);
    skinny_wgrad_kernel<T><<<dim3(bx, by), 256, 0, s>>>(sp, w.DPRE, 64, w.HD, w.ldhd, HKa, P, rows);
    BN_LAUNCH_CHECK();
    SkinnyPlan s2{};
    s2.r[s2.n++] = SkinnyRow{g + hp.wsig, nullptr, 16, 0, F};
    if (hp.ch_nlr >= 0)
      for (int o = 0; o < 3; ++o) s2.r[s2.n++] = SkinnyRow{g + hp.wg + (long long)o * F, nullptr, 17 + o, 0, F};
    const int bx2 = ceil_div(F, 256);
    int by2 = (int)max(1LL, min(ceil_div_ll(P, 128), (long long)(148 * 4 / bx2)));
    const long long rows2 = ceil_div_ll(ceil_div_ll(P, by2), 32) * 32;
    by2 = (int)ceil_div_ll(P, rows2);
    skinny_wgrad_kernel<T><<<dim3(bx2, by2), 256, 0, s>>>(s2, w.DPRE, 64, Hl, ldl, F, P, rows2);
    BN_LAUNCH_CHECK();
  }
  // heads' first layer: wgrad (+ bias gradient; the fp32 mode got it from heads_bwd_kernel) per block, dgrad into the features
  cudaStream_t sw1 = fork();
  for (int b = 0; b < hp.n_blocks; ++b) {
    const int lin = h->blk_lin0[b];
    // the colour head (block 0) also reads the encoded view direction: In = [features | dir enc | pad], dW is [HH, F + DE]
    if (h->blk_head[b] == BN_HEAD_BETA) {
      // dW is [HH, F + TE]: packed columns [F, F + kTOff) are padding, [F + kTOff, F + kTOff + TE) the time embedding
      if (int rc = layer_wgrad<T>(h, w.GHD + (long long)b * h->HH, w.ldhd, w.FE, w.ldfe, h->HH, F + kTOff + h->TE, P, g + c.w_off[lin],
                                  F + h->TE, F, F + kTOff, kTC ? g + hp.b1_off[b] : nullptr, sw1)) return rc;
      continue;
    }
    const int kin = F + (b == 0 ? h->DE : 0), no = (b == 0 && h->DE) ? h->ldfe : F;
    if (int rc = layer_wgrad<T>(h, w.GHD + (long long)b * h->HH, w.ldhd, w.FE, w.ldfe, h->HH, no, P, g + c.w_off[lin], kin, kin, no,
                                kTC ? g + hp.b1_off[b] : nullptr, sw1)) return rc;
  }
  // Every data-gradient GEMM from here down consumes the [P, *] tensor its predecessor has just written.  Alternating the row-
  // tile direction (tc::tile_reverse()), so that each one starts with the rows that should still be in L2, was measured on
  // B200 and LOST: 2.50 vs 2.43 ms per step, the dgrad GEMMs 0.91-0.98 vs 0.87 ms (profiles/r01f_zigzag_ab.txt) — descending
  // row order costs the HBM streams more than the L2 hits return.  Kept as an opt-in knob (BN_TILE_ZIGZAG=1).
  static const bool zigzag = kTC && getenv("BN_TILE_ZIGZAG") != nullptr;
  {
    DgradArgs<T> a;
    tc::tile_reverse() = zigzag;                            // GHD was written first row tile first
    const int rc = layer_dgrad<T>(h, w.GHD, w.ldhd, (const T*)h->W1T, (long long)h->n_blocks * h->HH, P, F, HKa, a, w.GFE, F, s);
    tc::tile_reverse() = false;
    if (rc) return rc;
  }
  // feature layer
  {
    if (int rc = layer_wgrad<T>(h, w.GFE, F, Hl, ldl, F, F, P, g + c.w_off[BN_LIN_FEATS], F, F, F, g + c.b_off[BN_LIN_FEATS], fork())) return rc;
    DgradArgs<T> a; a.mulc = w.C[L - 1]; a.ldm = F;
    if constexpr (kTC) {          // direct grads into h_{L-1}: dsigma w_sigma + sum_k dv_k Wg_k, DPRE cols 16..19
      a.rank_rows = w.DPRE + 16; a.rank_ld = 64; a.n_rank = hp.ch_nlr >= 0 ? 4 : 1;
      a.rank_col[0] = params + hp.wsig;
      for (int k = 0; k < 3; ++k) a.rank_col[1 + k] = hp.ch_nlr >= 0 ? params + hp.wg + (long long)k * F : nullptr;
    } else { a.addend = w.G7D; a.lda = F; }
    if (normals) { a.add2 = w.U[L - 1]; a.ld2 = F; }
    if (int rc = layer_dgrad<T>(h, w.GFE, F, (const T*)h->WfT, F, P, F, F, a, w.GA, F, s)) return rc;
  }
  // trunk, last layer first.  buf[ci] = dZ_l.  The weight gradient of layer l and the data gradient that produces dZ_{l-1}
  // only share their input, so (tcgen05 mode) the wgrads run on the handle's side stream: their epilogue-only tails and
  // the ramp of the next dgrad overlap instead of leaving the SMs idle between two persistent kernels.  dZ rotates through
  // three buffers; a buffer is rewritten only after the wgrad that read it has finished (ev_w), and the side stream
  // is joined back into `s` before returning (the whole pattern is stream-capturable).
  if constexpr (kTC) {
    if (dgrad_chain_ok(h) && !normals) {
      // every trunk data gradient in one launch; the weight gradients follow, each reading the dZ_l the chain left in HBM
      // and the layer input the forward left there (they share nothing with each other: side stream next to the chain's tail)
      if (int rc = dgrad_chain(h, w, P, s)) return rc;
      cudaStream_t sw = fork();
      for (int l = L - 1; l >= 0; --l) {
        const bool enc_in = (l == 0 || l == h->skip);
        const T* In; long long ldin;
        if (enc_in) { In = w.X3; ldin = w.ldx3; } else { In = w.H[l - 1]; ldin = w.Hld[l - 1]; }
        const T* dZ = (l == L - 1) ? w.GA : w.GZ[l];
        if (int rc = layer_wgrad<T>(h, dZ, F, In, ldin, F, h->Kpad[l], P, g + c.w_off[l], h->Kreal[l],
                                    enc_in ? h->E : h->Kpad[l], enc_in ? kEncPad : h->Kpad[l], g + c.b_off[l], sw, 2.0 * P * F * h->Kreal[l])) return rc;
      }
      if (overlap) {
        BN_CUDA(cudaEventRecord(h->ev_h[7], h->s2));
        BN_CUDA(cudaStreamWaitEvent(s, h->ev_h[7], 0));
      }
      return BN_OK;
    }
  }
  T* buf[3] = {w.GA, w.GB, kTC ? w.G7D : nullptr};       // G7D is unused in tcgen05 mode (rank-4 epilogue addend instead)
  int ci = 0;
  int reader[3] = {-1, -1, -1};                           // layer whose wgrad (on s2) last read buf[i]
  for (int l = L - 1; l >= 0; --l) {
    const T* cur = buf[ci];
    const bool enc_in = (l == 0 || l == h->skip);
    const T* In; long long ldin;
    if (enc_in) { In = w.X3; ldin = w.ldx3; } else { In = w.H[l - 1]; ldin = w.Hld[l - 1]; }
    cudaStream_t sw = s;
    if (overlap) {
      BN_CUDA(cudaEventRecord(h->ev_dz[l], s));
      BN_CUDA(cudaStreamWaitEvent(h->s2, h->ev_dz[l], 0));
      sw = h->s2;
    }
    if (int rc = layer_wgrad<T>(h, cur, F, In, ldin, F, h->Kpad[l], P, g + c.w_off[l], h->Kreal[l],
                                enc_in ? h->E : h->Kpad[l], enc_in ? kEncPad : h->Kpad[l], g + c.b_off[l], sw, 2.0 * P * F * h->Kreal[l])) return rc;
    if (overlap) { BN_CUDA(cudaEventRecord(h->ev_w[l], h->s2)); reader[ci] = l; }
    if (l > 0) {
      const int ni = overlap ? (ci + 1) % 3 : (ci ^ 1);
      if (overlap && reader[ni] >= 0) { BN_CUDA(cudaStreamWaitEvent(s, h->ev_w[reader[ni]], 0)); reader[ni] = -1; }
      const T* BT = (const T*)h->WTp[l] + (l == h->skip ? (long long)kEncPad * F : 0);
      DgradArgs<T> a; a.mulc = w.C[l - 1]; a.ldm = F;
      if (normals) { a.add2 = w.U[l - 1]; a.ld2 = F; }
      tc::tile_reverse() = zigzag && ((L - 1 - l) % 2 == 0);   // the feature-layer dgrad wrote dZ_{L-1} first tile first
      const int rc = layer_dgrad<T>(h, cur, F, BT, F, P, F, F, a, buf[ni], F, s);
      tc::tile_reverse() = false;
      if (rc) return rc;
      ci = ni;
    }
  }
  if (overlap) {                                          // join: the side stream is in order, its last event covers all of it
    BN_CUDA(cudaEventRecord(h->ev_h[7], h->s2));
    BN_CUDA(cudaStreamWaitEvent(s, h->ev_h[7], 0));
  }
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default"))) int bn_mlp_create(const bn_mlp_cfg* cfg, bn_mlp** out) {
  BN_CHECK_ARG(cfg && out, "null pointer");
  BN_CHECK_ARG(cfg->feat >= 128 && cfg->feat % 128 == 0 && cfg->feat <= 1024, "feat must be a multiple of 128 in [128, 1024]");
  BN_CHECK_ARG(cfg->layers >= 2 && cfg->layers <= 16, "layers must be in [2, 16]");
  BN_CHECK_ARG(cfg->skip_layer == -1 || (cfg->skip_layer >= 1 && cfg->skip_layer < cfg->layers), "skip_layer out of range");
  BN_CHECK_ARG(cfg->n_freq_xyz >= 0 && cfg->n_freq_xyz <= 10, "n_freq_xyz must be in [0, 10]");
  BN_CHECK_ARG(cfg->n_freq_dir >= 0 && cfg->n_freq_dir <= 10, "n_freq_dir must be in [0, 10]");
  BN_CHECK_ARG(cfg->precision == BN_PREC_FP32 || cfg->precision == BN_PREC_BF16, "unknown precision");
  BN_CHECK_ARG(cfg->head_dim[BN_HEAD_BETA] <= 0 || (cfg->t_dims >= 1 && cfg->t_dims <= kEncPad - kTOff && cfg->t_dims % 4 == 0),
               "beta head: t_dims must be a multiple of 4 in [4, 32]");
  BN_CHECK_ARG(!cfg->viewdir || cfg->n_freq_dir * 6 <= kTOff || cfg->head_dim[BN_HEAD_BETA] <= 0, "direction encoding overlaps the time embedding");
  int dev = 0;
  BN_CUDA(cudaGetDevice(&dev));
  if (int rc = bn_device_check(dev)) return rc;
  bn_mlp* h = new bn_mlp();
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->F = cfg->feat; h->L = cfg->layers; h->HH = cfg->feat / 2; h->skip = cfg->skip_layer;
  h->E = cfg->n_freq_xyz == 0 ? 3 : 6 * cfg->n_freq_xyz;
  h->DE = cfg->viewdir ? (cfg->n_freq_dir == 0 ? 3 : 6 * cfg->n_freq_dir) : 0;
  h->TE = cfg->head_dim[BN_HEAD_BETA] > 0 ? cfg->t_dims : 0;
  h->ldfe = h->F + ((h->DE || h->TE) ? kEncPad : 0);
  h->bf16 = cfg->precision == BN_PREC_BF16;
  h->es = h->bf16 ? 2 : 4;
  cudaDeviceProp prop;
  BN_CUDA(cudaGetDeviceProperties(&prop, dev));
  h->num_sms = prop.multiProcessorCount;
  h->no_chain = getenv("BN_NO_CHAIN") != nullptr;      // debugging aid: per-layer GEMMs instead of the fused trunk kernels
  h->no_dchain = getenv("BN_NO_DGRAD_CHAIN") != nullptr;
  // blocks of the heads' hidden layer: rgb first, then every BRDF head that exists
  h->n_blocks = 0;
  h->blk_lin0[0] = BN_LIN_RGB0; h->blk_lin2[0] = BN_LIN_RGB2; h->blk_head[0] = -1; h->n_blocks = 1;
  // the beta head is evaluated in every full forward (not only with apply_brdf): its block comes right after the colour block,
  // so that a Lambertian-stage call of a model with BRDF heads stops after two blocks
  for (int i = 0; i < BN_NUM_HEADS; ++i) {
    const int hd = i == 0 ? BN_HEAD_BETA : i - 1;
    if (cfg->head_dim[hd] > 0) {
      h->blk_lin0[h->n_blocks] = BN_LIN_HEAD0 + 2 * hd;
      h->blk_lin2[h->n_blocks] = BN_LIN_HEAD0 + 2 * hd + 1;
      h->blk_head[h->n_blocks] = hd;
      ++h->n_blocks;
    }
  }
  for (int l = 0; l < h->L; ++l) {
    const bool enc_in = (l == 0 || l == h->skip);
    h->Kreal[l] = l == 0 ? h->E : (l == h->skip ? h->E + h->F : h->F);
    h->Kpad[l] = l == 0 ? kEncPad : (l == h->skip ? kEncPad + h->F : h->F);
    (void)enc_in;
    BN_CUDA(cudaMalloc(&h->Wp[l], (size_t)h->F * h->Kpad[l] * h->es));
    BN_CUDA(cudaMalloc(&h->WTp[l], (size_t)h->F * h->Kpad[l] * h->es));
  }
  const size_t HK = (size_t)h->n_blocks * h->HH;
  BN_CUDA(cudaMalloc(&h->Wf, (size_t)h->F * h->F * h->es));
  BN_CUDA(cudaMalloc(&h->WfT, (size_t)h->F * h->F * h->es));
  BN_CUDA(cudaMalloc(&h->W1, HK * h->ldfe * h->es));
  BN_CUDA(cudaMalloc(&h->W1T, HK * h->ldfe * h->es));
  BN_CUDA(cudaMalloc(&h->b1cat, HK * sizeof(float)));
  BN_CUDA(cudaMalloc(&h->W2p, HK * 64 * sizeof(__nv_bfloat16)));
  BN_CUDA(cudaMalloc(&h->W2pT, HK * 64 * sizeof(__nv_bfloat16)));
  BN_CUDA(cudaMalloc(&h->Wsig, (size_t)64 * h->F * sizeof(__nv_bfloat16)));
  BN_CUDA(cudaStreamCreateWithFlags(&h->s2, cudaStreamNonBlocking));
  for (int l = 0; l < 16; ++l) {
    BN_CUDA(cudaEventCreateWithFlags(&h->ev_dz[l], cudaEventDisableTiming));
    BN_CUDA(cudaEventCreateWithFlags(&h->ev_w[l], cudaEventDisableTiming));
  }
  for (int i = 0; i < 8; ++i) BN_CUDA(cudaEventCreateWithFlags(&h->ev_h[i], cudaEventDisableTiming));
  h->overlap = getenv("BN_NO_OVERLAP") == nullptr;       // A/B timing aid: serial backward on the caller's stream
  BN_CUDA(cudaMalloc(&h->WsigA, (size_t)64 * h->F * sizeof(__nv_bfloat16)));
  BN_CUDA(cudaMemset(h->WsigA, 0, (size_t)64 * h->F * sizeof(__nv_bfloat16)));
  if (h->bf16) {
    BN_CUDA(cudaMalloc(&h->Wb, (size_t)h->L * h->F * 64 * sizeof(__nv_bfloat16)));
    long long offs[32];
    for (int l = 0; l < h->L; ++l) { offs[2 * l] = cfg->w_off[BN_LIN_TRUNK0 + l]; offs[2 * l + 1] = cfg->b_off[BN_LIN_TRUNK0 + l]; }
    BN_CUDA(cudaMalloc(&h->Wb_offs, sizeof(offs)));
    BN_CUDA(cudaMemcpy(h->Wb_offs, offs, sizeof(long long) * 2 * h->L, cudaMemcpyHostToDevice));
  }
  *out = h;
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) void bn_mlp_destroy(bn_mlp* h) {
  if (!h) return;
  for (int l = 0; l < h->L; ++l) { cudaFree(h->Wp[l]); cudaFree(h->WTp[l]); }
  if (h->s2) cudaStreamDestroy(h->s2);
  for (int l = 0; l < 16; ++l) { if (h->ev_dz[l]) cudaEventDestroy(h->ev_dz[l]); if (h->ev_w[l]) cudaEventDestroy(h->ev_w[l]); }
  for (int i = 0; i < 8; ++i) if (h->ev_h[i]) cudaEventDestroy(h->ev_h[i]);
  cudaFree(h->Wf); cudaFree(h->WfT); cudaFree(h->W1); cudaFree(h->W1T); cudaFree(h->b1cat); cudaFree(h->W2p); cudaFree(h->W2pT); cudaFree(h->Wsig); cudaFree(h->WsigA);
  cudaFree(h->Wb); cudaFree(h->Wb_offs);
  delete h;
}

extern "C" __attribute__((visibility("default"))) int bn_mlp_sync_weights(bn_mlp* h, const float* params, cudaStream_t stream) {
  BN_CHECK_ARG(h && params, "null pointer");
  return h->bf16 ? sync_weights_t<__nv_bfloat16>(h, params, stream) : sync_weights_t<float>(h, params, stream);
}

extern "C" __attribute__((visibility("default"))) int bn_mlp_out_channels(const bn_mlp* h, int flags) {
  if (!h) return BN_ERR_ARG;
  if (flags & BN_MLP_SIGMA_ONLY) return 1;
  int n = 0;
  if (build_plan(h, flags, nullptr, &n)) return BN_ERR_ARG;
  return n;
}

extern "C" __attribute__((visibility("default"))) size_t bn_mlp_workspace_bytes(const bn_mlp* h, int64_t n_points, int flags) {
  if (!h || n_points <= 0) return 0;
  return h->bf16 ? carve<__nv_bfloat16>(h, n_points, flags, nullptr, nullptr) : carve<float>(h, n_points, flags, nullptr, nullptr);
}

extern "C" __attribute__((visibility("default"))) int bn_mlp_forward(bn_mlp* h, const float* params, const float* origins, int o_stride,
                              const float* dirs, int d_stride, const float* z, int n_rays, int n_samples,
                              int flags, float* out, int out_pitch, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
  BN_CHECK_ARG(h && params && origins && dirs && z && out && workspace, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  if (!h->synced) { set_error("bn_mlp_forward: call bn_mlp_sync_weights first"); return BN_ERR_STATE; }
  if (workspace_bytes < bn_mlp_workspace_bytes(h, (int64_t)n_rays * n_samples, flags)) {
    set_error("bn_mlp_forward: workspace too small"); return BN_ERR_STATE;
  }
  BN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  return h->bf16 ? forward_t<__nv_bfloat16>(h, params, origins, o_stride, dirs, d_stride, z, n_rays, n_samples, flags, out, out_pitch, workspace, stream)
                 : forward_t<float>(h, params, origins, o_stride, dirs, d_stride, z, n_rays, n_samples, flags, out, out_pitch, workspace, stream);
}

extern "C" __attribute__((visibility("default")))
int bn_mlp_trunk_forward(bn_mlp* h, const float* params, const float* origins, int o_stride, const float* dirs, int d_stride,
                         const float* z, int n_rays, int n_samples, int flags, int64_t total_points, int64_t row0,
                         float* sigma_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(h && params && origins && dirs && z && workspace, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  BN_CHECK_ARG(!(flags & BN_MLP_SIGMA_ONLY), "bn_mlp_trunk_forward keeps activations: use bn_mlp_forward for a density-only pass");
  BN_CHECK_ARG(row0 >= 0 && row0 + (int64_t)n_rays * n_samples <= total_points, "rows out of range");
  BN_CHECK_ARG(row0 % 128 == 0, "row0 must be a multiple of 128 (TMA boxes / MMA tiles start on 128-row boundaries)");
  if (!h->synced) { set_error("bn_mlp_trunk_forward: call bn_mlp_sync_weights first"); return BN_ERR_STATE; }
  if (workspace_bytes < bn_mlp_workspace_bytes(h, total_points, flags)) {
    set_error("bn_mlp_trunk_forward: workspace too small"); return BN_ERR_STATE;
  }
  BN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  return h->bf16 ? trunk_rows_t<__nv_bfloat16>(h, params, origins, o_stride, dirs, d_stride, z, n_rays, n_samples, flags, total_points, row0, sigma_out, workspace, stream)
                 : trunk_rows_t<float>(h, params, origins, o_stride, dirs, d_stride, z, n_rays, n_samples, flags, total_points, row0, sigma_out, workspace, stream);
}

extern "C" __attribute__((visibility("default")))
int bn_mlp_write_t(bn_mlp* h, const float* t_rows, int n_rays, int n_samples, int flags, int64_t total_points, int64_t row0,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(h && t_rows && workspace, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0 && row0 >= 0 && row0 + (int64_t)n_rays * n_samples <= total_points, "row range out of bounds");
  BN_CHECK_ARG(h->TE > 0 && (flags & BN_MLP_BETA) && !(flags & BN_MLP_SIGMA_ONLY), "bn_mlp_write_t needs a model with a beta head and BN_MLP_BETA");
  if (workspace_bytes < bn_mlp_workspace_bytes(h, total_points, flags)) { set_error("bn_mlp_write_t: workspace too small"); return BN_ERR_STATE; }
  const long long P = (long long)n_rays * n_samples;
  if (h->bf16) {
    Ws<__nv_bfloat16> w; carve<__nv_bfloat16>(h, total_points, flags, workspace, &w);
    t_embed_kernel<__nv_bfloat16><<<(unsigned)ceil_div_ll(P, 128), 128, 0, stream>>>(t_rows, h->TE, n_samples, P, h->DE == 0,
                                                                                      w.FE + row0 * w.ldfe + h->F, w.ldfe);
  } else {
    Ws<float> w; carve<float>(h, total_points, flags, workspace, &w);
    t_embed_kernel<float><<<(unsigned)ceil_div_ll(P, 128), 128, 0, stream>>>(t_rows, h->TE, n_samples, P, h->DE == 0,
                                                                             w.FE + row0 * w.ldfe + h->F, w.ldfe);
  }
  return after_launch("t_embed_kernel");
}

extern "C" __attribute__((visibility("default")))
int bn_mlp_heads_forward(bn_mlp* h, const float* params, int64_t total_points, int flags, float* out, int out_pitch,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(h && params && out && workspace, "null pointer");
  BN_CHECK_ARG(total_points > 0 && !(flags & BN_MLP_SIGMA_ONLY), "bad arguments");
  if (!h->synced) { set_error("bn_mlp_heads_forward: call bn_mlp_sync_weights first"); return BN_ERR_STATE; }
  if (workspace_bytes < bn_mlp_workspace_bytes(h, total_points, flags)) {
    set_error("bn_mlp_heads_forward: workspace too small"); return BN_ERR_STATE;
  }
  if (h->bf16) {
    Ws<__nv_bfloat16> w; carve<__nv_bfloat16>(h, total_points, flags, workspace, &w);
    return heads_t<__nv_bfloat16>(h, params, total_points, flags, out, out_pitch, w, stream);
  }
  Ws<float> w; carve<float>(h, total_points, flags, workspace, &w);
  return heads_t<float>(h, params, total_points, flags, out, out_pitch, w, stream);
}

extern "C" __attribute__((visibility("default"))) int bn_mlp_backward(bn_mlp* h, const float* params, const float* out, const float* g_out, int out_pitch,
                               int n_rays, int n_samples, int flags, float* g_params,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  BN_CHECK_ARG(h && params && out && g_out && g_params && workspace, "null pointer");
  BN_CHECK_ARG(n_rays > 0 && n_samples > 0, "empty batch");
  BN_CHECK_ARG((flags & BN_MLP_TRAIN) && !(flags & BN_MLP_SIGMA_ONLY), "backward needs a BN_MLP_TRAIN full forward");
  if (workspace_bytes < bn_mlp_workspace_bytes(h, (int64_t)n_rays * n_samples, flags)) {
    set_error("bn_mlp_backward: workspace too small"); return BN_ERR_STATE;
  }
  return h->bf16 ? backward_t<__nv_bfloat16>(h, params, out, g_out, out_pitch, n_rays, n_samples, flags, g_params, workspace, stream)
                 : backward_t<float>(h, params, out, g_out, out_pitch, n_rays, n_samples, flags, g_params, workspace, stream);
}

extern "C" __attribute__((visibility("default"))) int bn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                            float grad_scale, cudaStream_t stream) {
  BN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "bad arguments");
  const float bc1 = 1.0f - (float)pow((double)beta1, (double)step);
  const float bc2 = 1.0f - (float)pow((double)beta2, (double)step);
  adam_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default")))
int bn_adam_step_graph(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state,
                       float beta1, float beta2, float eps, float weight_decay, float grad_scale, cudaStream_t stream) {
  BN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && n > 0, "bad arguments");
  BN_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "buffers must be 16-byte aligned");
  const double l1 = log((double)beta1), l2 = log((double)beta2);
  const float l1h = (float)l1, l2h = (float)l2;
  adam_state_kernel<<<(unsigned)ceil_div_ll(n, kAdamThreads * kAdamPerThread), kAdamThreads, 0, stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, weight_decay, grad_scale, l1h, (float)(l1 - (double)l1h), l2h,
      (float)(l2 - (double)l2h));
  BN_LAUNCH_CHECK();
  return BN_OK;
}

// Unit-test hook: run one GEMM of the MLP engine in isolation.
//   kind 0 (TN): out[M,N] = A[M,K] B[N,K]^T            (fp32 store)
//   kind 1 (NT): out[Mo=M, No=N] += A[P=K, M]^T B[P=K, N]  (fp32 atomics; zero `out` first)
// precision BN_PREC_BF16 -> tcgen05 path on bf16 operands, BN_PREC_FP32 -> CUDA-core path on fp32.
extern "C" __attribute__((visibility("default")))
int bn_debug_gemm(int kind, int precision, const void* A, long long lda, const void* B, long long ldb,
                  float* out, long long ldo, long long M, int N, long long K, cudaStream_t stream) {
  BN_CHECK_ARG(A && B && out, "null pointer");
  int dev = 0; BN_CUDA(cudaGetDevice(&dev));
  if (int rc = bn_device_check(dev)) return rc;
  static int sms = 0;
  if (!sms) BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bn_mlp h{}; h.num_sms = sms;
  if (kind == 0) {
    EpiStoreF32 epi{out, ldo, (int)M, N};
    if (precision == BN_PREC_BF16) return gemm_tn<__nv_bfloat16>(&h, (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, M, N, (int)K, epi, stream);
    return gemm_tn<float>(&h, (const float*)A, lda, (const float*)B, ldb, M, N, (int)K, epi, stream);
  }
  if (kind == 2) {        // mainloop only (bf16): accumulators are read back and dropped
    EpiNull epi{out};
    return gemm_tn<__nv_bfloat16>(&h, (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, M, N, (int)K, epi, stream);
  }
  EpiWgrad epi{out, ldo, (int)M, N, N, N};
  if (precision == BN_PREC_BF16) return gemm_nt<__nv_bfloat16>(&h, (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, (int)M, N, K, epi, stream);
  return gemm_nt<float>(&h, (const float*)A, lda, (const float*)B, ldb, (int)M, N, K, epi, stream);
}

extern "C" __attribute__((visibility("default")))
int bn_debug_gemm_epi(int kind, const void* A, long long lda, const void* B, long long ldb, void* out, long long ldo,
                      const void* add, const void* mul, float* colsum, int pad_lo, int pad_hi,
                      long long M, int N, long long K, cudaStream_t stream) {
  BN_CHECK_ARG(A && B && out, "null pointer");
  int dev = 0; BN_CUDA(cudaGetDevice(&dev));
  if (int rc = bn_device_check(dev)) return rc;
  static int sms = 0;
  if (!sms) BN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  bn_mlp h{}; h.num_sms = sms;
  typedef __nv_bfloat16 T;
  if (kind == 0) {
    DgradArgs<T> a; a.addend = (const T*)add; a.lda = ldo; a.mulc = (const T*)mul; a.ldm = ldo; a.bias_grad = colsum;
    return layer_dgrad<T>(&h, (const T*)A, lda, (const T*)B, ldb, M, N, (int)K, a, (T*)out, ldo, stream);
  }
  BN_CHECK_ARG(tc::wgrad_tma_ok((const float*)out, ldo, N, pad_lo, pad_hi), "output not addressable by the TMA");
  return layer_wgrad<T>(&h, (const T*)A, lda, (const T*)B, ldb, (int)M, N, K, (float*)out, ldo, pad_lo, pad_hi, colsum, stream);
}

// Diagnostics: device buffer of >= 16 * 2 * layers int64 that the fused density pass fills with clock64() stamps
// of its first block (scripts/trace_chain.py); NULL switches the stamps off.
extern "C" __attribute__((visibility("default"))) int bn_debug_chain_trace(bn_mlp* h, long long* device_buf) {
  BN_CHECK_ARG(h != nullptr, "null handle");
  h->chain_trace = device_buf;
  // BN_NT_TRACE: the weight-gradient GEMM's stamps follow the chain's 512 words (scripts/trace_wgrad.py sizes the buffer)
  tc::nt_trace() = (device_buf && getenv("BN_NT_TRACE")) ? device_buf + 512 : nullptr;
  return BN_OK;
}

// Unit-test hook: offset / pitch of X3, H_l, C_l inside a workspace carved for (n_points, flags)
extern "C" __attribute__((visibility("default")))
int bn_debug_ws_tensor(const bn_mlp* h, int64_t n_points, int flags, int which, int layer, int64_t* offset_bytes,
                       int64_t* pitch_elems) {
  BN_CHECK_ARG(h && offset_bytes && pitch_elems, "null pointer");
  BN_CHECK_ARG(which >= 0 && which <= 2 && layer >= 0 && layer < h->L, "which / layer out of range");
  uint8_t* const base = reinterpret_cast<uint8_t*>(uintptr_t(4096));      // carve() only does pointer arithmetic on it
  const uint8_t* p = nullptr; long long ld = 0;
  if (h->bf16) {
    Ws<__nv_bfloat16> w; carve<__nv_bfloat16>(h, n_points, flags, base, &w);
    if (which == 0) { p = (const uint8_t*)w.X3; ld = w.ldx3; } else if (which == 1) { p = (const uint8_t*)w.H[layer]; ld = w.Hld[layer]; }
    else { p = (const uint8_t*)w.C[layer]; ld = h->F; }
  } else {
    Ws<float> w; carve<float>(h, n_points, flags, base, &w);
    if (which == 0) { p = (const uint8_t*)w.X3; ld = w.ldx3; } else if (which == 1) { p = (const uint8_t*)w.H[layer]; ld = w.Hld[layer]; }
    else { p = (const uint8_t*)w.C[layer]; ld = h->F; }
  }
  BN_CHECK_ARG(p != nullptr, "this tensor is not kept for these flags");
  *offset_bytes = (int64_t)(p - base); *pitch_elems = ld;
  return BN_OK;
}
