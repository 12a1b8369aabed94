// K-C (part 3): per-ray shading epilogue (irradiance, Lambertian / RPV / Hapke / microfacet rgb) and
// the per-sample BRDF evaluation of the MultiBRDF variant, forward and backward.
//
// Replaces (reference, paths relative to /root/reference):
//   albedo_accu clamp, normal_s, nr_vw / nr_sun / hpk_scl   models/spsbrdfnerf.py:198-199,241-255
//   irradiance (ones | |z·sun| | sun visibility)            models/spsbrdfnerf.py:259-268
//   Lambertian rgb with rgb_padding                         models/spsbrdfnerf.py:270-275
//   per-ray / per-sample BRDF dispatch and final rgb        models/spsbrdfnerf.py:284-357
// The backward kernels evaluate the same templated BRDF code over forward-mode dual numbers
// (brdf.cuh) and contract the Jacobian row with the incoming d loss / d rgb.
#include "brdf.cuh"

namespace bn {

constexpr float kPad = 0.001f;          // rgb_padding, spsbrdfnerf.py:459
constexpr float kF32Eps = 1.1920929e-07f;

struct ParamIdx { int p0, p1, p2; bool h0, h1, h2; };

// channel offsets of the (up to three) BRDF parameters of colour channel c inside a packed row
__device__ __forceinline__ ParamIdx param_index(const bn_shade_cfg& cfg, int c) {
  ParamIdx q{-1, -1, -1, false, false, false};
  int off = cfg.param_ch;
  if (off < 0) return q;
  if (cfg.brdf_type == BN_BRDF_MICROFACET) { q.p0 = off; q.h0 = true; }
  else if (cfg.brdf_type == BN_BRDF_RPV) {
    if (cfg.funcM) { q.p0 = off + c; q.h0 = true; off += 3; }
    if (cfg.funcF) { q.p1 = off + c; q.h1 = true; off += 3; }
    if (cfg.funcH == 1) { q.p2 = off + c; q.h2 = true; off += 3; }
  } else if (cfg.brdf_type == BN_BRDF_HAPKE) {
    if (cfg.hapke_b) { q.p0 = off + c; q.h0 = true; off += 3; }
    if (cfg.hapke_c) { q.p1 = off + c; q.h1 = true; off += 3; }
    if (cfg.hapke_theta) { q.p2 = off; q.h2 = true; }
  }
  return q;
}

struct ChanAux { float a[8]; };

// one colour channel of the active BRDF; n is used as given (unit for per-ray, raw for per-sample)
template <class T>
__device__ __forceinline__ T brdf_channel(const bn_shade_cfg& cfg, const float (&l)[3], const float (&v)[3],
                                          const T (&n)[3], const T& w, const ParamIdx& q, const T& p0,
                                          const T& p1, const T& p2, ChanAux* aux) {
  if (cfg.brdf_type == BN_BRDF_MICROFACET) {
    MicroAux m;
    T glossy = microfacet_glossy(l, v, n, p0, cfg.fresnel_f0, aux ? &m : nullptr);
    if (aux) { aux->a[0] = m.glossy; aux->a[1] = m.f; aux->a[2] = m.g; aux->a[3] = m.d; aux->a[4] = m.ldn;
               aux->a[5] = m.vdn; aux->a[6] = m.nh; aux->a[7] = 0.f; }
    return w + glossy;
  }
  Angles<T> a = calc_angles(l, v, n);
  if (cfg.brdf_type == BN_BRDF_RPV) {
    const bool has_rc = cfg.funcH != 0;
    float M1, G, H;
    T out = rpv_channel(a, w, q.h0, p0, q.h1, p1, has_rc, cfg.funcH == 2 ? w : p2, &M1, &G, &H);
    if (aux) { aux->a[0] = M1; aux->a[1] = G; aux->a[2] = H; aux->a[3] = val(a.ci); aux->a[4] = val(a.cv); }
    return out;
  }
  HapkeAux h;
  T out = hapke_channel(a, w, q.h0, p0, q.h1, p1, q.h2, p2, cfg.hpk_scl, cfg.shell_hapke, &h);
  if (aux) { aux->a[0] = h.P; aux->a[1] = h.Hi; aux->a[2] = h.Hv; aux->a[3] = h.ci; aux->a[4] = h.cv; aux->a[5] = h.shad; }
  return out;
}

__device__ __forceinline__ float irradiance_of_ray(const bn_shade_cfg& cfg, const float* ray, const float* irr_last, int r) {
  if (cfg.irr_mode == BN_IRR_COS) return fabsf(ray[10]);     // up-vector · sun_d (spsbrdfnerf.py:260-264)
  if (cfg.irr_mode == BN_IRR_SUNVIS) return irr_last ? irr_last[r] : 1.0f;
  return 1.0f;
}

struct ShadeRays {
  bn_shade_cfg cfg;
  const float* rays;      // (N,11)
  const float* acc;       // (N,C)
  const float* wsum;      // (N)
  const float* acc_irr;   // (N,4) or null
  const float* irr_last;  // (N) or null: irradiance of the last sample (per-ray BRDF with sun visibility)
  // forward outputs
  float *rgb, *albedo_accu, *normal_s, *nr_vw, *nr_sun, *hpk_scl, *brdf, *aux;
  // backward
  const float* g_rgb; float *g_acc, *g_wsum, *g_acc_irr;
  int N;
};

template <bool kBackward>
__global__ void shade_rays_kernel(ShadeRays a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.N) return;
  const bn_shade_cfg& cfg = a.cfg;
  const int C = cfg.n_channels;
  const float* ray = a.rays + (long long)r * 11;
  const float* acc = a.acc + (long long)r * C;
  const float l[3] = {ray[8], ray[9], ray[10]};
  const float v[3] = {-ray[3], -ray[4], -ray[5]};
  const float ws = a.wsum[r];
  const float irr = irradiance_of_ray(cfg, ray, a.irr_last, r);
  float* ga = kBackward ? a.g_acc + (long long)r * C : nullptr;
  float gws = 0.f;
  if (kBackward) for (int c = 0; c < C; ++c) ga[c] = 0.f;

  // ---- Lambertian colour (always evaluated; overwritten below when a BRDF is active)
  float rgb[3];
  const bool per_sample_irr = cfg.irr_mode == BN_IRR_SUNVIS && a.acc_irr != nullptr && !(cfg.brdf_type && !cfg.multi_brdf);
  for (int c = 0; c < 3; ++c) {
    float raw = per_sample_irr ? (1.0f + 2.0f * kPad) * a.acc_irr[r * 4 + c] - kPad * a.acc_irr[r * 4 + 3]
                               : irr * ((1.0f + 2.0f * kPad) * acc[c] - kPad * ws);
    rgb[c] = fminf(fmaxf(raw, 0.f), 1.f);
    if (kBackward && !cfg.brdf_type) {
      const float up = (raw >= 0.f && raw <= 1.f) ? a.g_rgb[r * 3 + c] : 0.f;
      if (per_sample_irr) {
        a.g_acc_irr[r * 4 + c] = (1.0f + 2.0f * kPad) * up;
        if (c == 0) a.g_acc_irr[r * 4 + 3] = 0.f;
        a.g_acc_irr[r * 4 + 3] += -kPad * up;
      } else {
        ga[c] += irr * (1.0f + 2.0f * kPad) * up;
        gws += -irr * kPad * up;
      }
    }
  }
  if (!kBackward && a.albedo_accu)
    for (int c = 0; c < 3; ++c) a.albedo_accu[r * 3 + c] = fminf(fmaxf(acc[c], 0.f), 1.f);

  // ---- accumulated normal and its dot products
  const int nc = cfg.normal_ch;
  float nsv[3] = {0.f, 0.f, 1.f};
  if (nc >= 0) {
    float sq = acc[nc] * acc[nc] + acc[nc + 1] * acc[nc + 1] + acc[nc + 2] * acc[nc + 2];
    float inv = 1.0f / sqrtf(fmaxf(sq, kF32Eps));
    for (int j = 0; j < 3; ++j) nsv[j] = acc[nc + j] * inv;
    if (!kBackward) {
      float nv = nsv[0] * v[0] + nsv[1] * v[1] + nsv[2] * v[2];
      float nl = nsv[0] * l[0] + nsv[1] * l[1] + nsv[2] * l[2];
      if (a.normal_s) for (int j = 0; j < 3; ++j) a.normal_s[r * 3 + j] = nsv[j];
      if (a.nr_vw) a.nr_vw[r] = nv;
      if (a.nr_sun) a.nr_sun[r] = nl;
      if (a.hpk_scl) a.hpk_scl[r] = 1.0f / (cfg.hpk_scl * (nv + nl));
    }
  }

  // ---- BRDF colour
  if (cfg.brdf_type) {
    if (cfg.multi_brdf) {
      // per-sample BRDF already accumulated by the compositing kernel into channels brdf_ch..+2
      for (int c = 0; c < 3; ++c) {
        float raw = irr * ((1.0f + 2.0f * kPad) * acc[cfg.brdf_ch + c] - kPad * ws);
        if (per_sample_irr) raw = 0.f;   // unsupported combination guarded on the host
        rgb[c] = fminf(fmaxf(raw, 0.f), 1.f);
        if (kBackward) {
          const float up = (raw >= 0.f && raw <= 1.f) ? a.g_rgb[r * 3 + c] : 0.f;
          ga[cfg.brdf_ch + c] += irr * (1.0f + 2.0f * kPad) * up;
          gws += -irr * kPad * up;
        }
      }
    } else {
      for (int c = 0; c < 3; ++c) {
        const ParamIdx q = param_index(cfg, c);
        const float wv = (1.0f + 2.0f * kPad) * acc[c] - kPad * ws;     // albedo_s (padded)
        if (!kBackward) {
          ChanAux aux;
          const float zero = 0.f;
          float b = brdf_channel<float>(cfg, l, v, nsv, wv, q, q.h0 ? acc[q.p0] : zero, q.h1 ? acc[q.p1] : zero,
                                        q.h2 ? acc[q.p2] : zero, &aux);
          rgb[c] = fminf(fmaxf(irr * b, 0.f), 1.f);
          if (a.brdf) a.brdf[r * 3 + c] = b;
          if (a.aux) for (int t = 0; t < 8; ++t) a.aux[((long long)r * 3 + c) * 8 + t] = aux.a[t];
        } else {
          typedef Dual<7> D;
          D an[3] = {D::var(acc[nc], 0), D::var(acc[nc + 1], 1), D::var(acc[nc + 2], 2)};
          D sq = an[0] * an[0] + an[1] * an[1] + an[2] * an[2];
          if (sq.v < kF32Eps) sq = D(kF32Eps);
          D nrm = sqrt_(sq);
          D n[3] = {an[0] / nrm, an[1] / nrm, an[2] / nrm};
          D w = D::var(wv, 3);
          D p0 = q.h0 ? D::var(acc[q.p0], 4) : D(0.f);
          D p1 = q.h1 ? D::var(acc[q.p1], 5) : D(0.f);
          D p2 = q.h2 ? D::var(acc[q.p2], 6) : D(0.f);
          D b = brdf_channel<D>(cfg, l, v, n, w, q, p0, p1, p2, nullptr);
          const float raw = irr * b.v;
          const float up = (raw >= 0.f && raw <= 1.f) ? a.g_rgb[r * 3 + c] * irr : 0.f;
          for (int j = 0; j < 3; ++j) ga[nc + j] += up * b.d[j];
          ga[c] += (1.0f + 2.0f * kPad) * up * b.d[3];
          gws += -kPad * up * b.d[3];
          if (q.h0) ga[q.p0] += up * b.d[4];
          if (q.h1) ga[q.p1] += up * b.d[5];
          if (q.h2) ga[q.p2] += up * b.d[6];
        }
      }
    }
  }
  if (!kBackward) { for (int c = 0; c < 3; ++c) a.rgb[r * 3 + c] = rgb[c]; }
  else { a.g_wsum[r] = gws; }
}

struct BrdfPoints {
  bn_shade_cfg cfg;
  const float* rays;     // (N,11)
  float* packed;         // (N,S,C): reads normal/albedo/params, writes brdf_ch..+2 (forward)
  float* g_packed;       // (N,S,C): backward in/out
  float* aux;            // (N,S,3,8) optional
  int N, S;
};

template <bool kBackward>
__global__ void brdf_points_kernel(BrdfPoints a) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)a.N * a.S) return;
  const bn_shade_cfg& cfg = a.cfg;
  const int C = cfg.n_channels;
  const int r = (int)(p / a.S);
  const float* ray = a.rays + (long long)r * 11;
  const float l[3] = {ray[8], ray[9], ray[10]};
  const float v[3] = {-ray[3], -ray[4], -ray[5]};
  float* x = a.packed + p * C;
  const int nc = cfg.normal_ch;
  for (int c = 0; c < 3; ++c) {
    const ParamIdx q = param_index(cfg, c);
    if (!kBackward) {
      const float n[3] = {x[nc], x[nc + 1], x[nc + 2]};
      ChanAux aux;
      const float zero = 0.f;
      float b = brdf_channel<float>(cfg, l, v, n, x[c], q, q.h0 ? x[q.p0] : zero, q.h1 ? x[q.p1] : zero,
                                    q.h2 ? x[q.p2] : zero, a.aux ? &aux : nullptr);
      x[cfg.brdf_ch + c] = b;
      if (a.aux) for (int t = 0; t < 8; ++t) a.aux[(p * 3 + c) * 8 + t] = aux.a[t];
    } else {
      typedef Dual<7> D;
      float* g = a.g_packed + p * C;
      const float up = g[cfg.brdf_ch + c];
      if (up == 0.f) continue;
      D n[3] = {D::var(x[nc], 0), D::var(x[nc + 1], 1), D::var(x[nc + 2], 2)};
      D w = D::var(x[c], 3);
      D p0 = q.h0 ? D::var(x[q.p0], 4) : D(0.f);
      D p1 = q.h1 ? D::var(x[q.p1], 5) : D(0.f);
      D p2 = q.h2 ? D::var(x[q.p2], 6) : D(0.f);
      D b = brdf_channel<D>(cfg, l, v, n, w, q, p0, p1, p2, nullptr);
      for (int j = 0; j < 3; ++j) g[nc + j] += up * b.d[j];
      g[c] += up * b.d[3];
      if (q.h0) g[q.p0] += up * b.d[4];
      if (q.h1) g[q.p1] += up * b.d[5];
      if (q.h2) g[q.p2] += up * b.d[6];
    }
  }
}

static int check_cfg(const bn_shade_cfg* cfg) {
  if (!cfg) { set_error("null shade cfg"); return BN_ERR_ARG; }
  if (cfg->n_channels < 4 || cfg->n_channels > 32) { set_error("shade cfg: n_channels out of range"); return BN_ERR_ARG; }
  if (cfg->brdf_type != BN_BRDF_NONE) {
    if (cfg->normal_ch < 4) { set_error("shade cfg: a BRDF needs a normal channel"); return BN_ERR_ARG; }
    if (cfg->multi_brdf && cfg->brdf_ch < 4) { set_error("shade cfg: multi_brdf needs brdf_ch"); return BN_ERR_ARG; }
    if (cfg->multi_brdf && cfg->irr_mode == BN_IRR_SUNVIS) {
      set_error("shade cfg: per-sample sun visibility with MultiBRDF is a reference defect path (SURVEY App. C.3)");
      return BN_ERR_ARG;
    }
  }
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" __attribute__((visibility("default"))) int bn_shade_rays_forward(const bn_shade_cfg* cfg, const float* rays, const float* acc,
                                     const float* wsum, const float* acc_irr, const float* irr_last,
                                     float* rgb, float* albedo_accu, float* normal_s, float* nr_vw,
                                     float* nr_sun, float* hpk_scl, float* brdf, float* aux,
                                     int n_rays, cudaStream_t stream) {
  if (int rc = check_cfg(cfg)) return rc;
  BN_CHECK_ARG(rays && acc && wsum && rgb, "null pointer");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  ShadeRays a{};
  a.cfg = *cfg; a.rays = rays; a.acc = acc; a.wsum = wsum; a.acc_irr = acc_irr; a.irr_last = irr_last;
  a.rgb = rgb; a.albedo_accu = albedo_accu; a.normal_s = normal_s; a.nr_vw = nr_vw; a.nr_sun = nr_sun;
  a.hpk_scl = hpk_scl; a.brdf = brdf; a.aux = aux; a.N = n_rays;
  shade_rays_kernel<false><<<ceil_div(n_rays, 128), 128, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_shade_rays_backward(const bn_shade_cfg* cfg, const float* rays, const float* acc,
                                      const float* wsum, const float* acc_irr, const float* irr_last,
                                      const float* g_rgb, float* g_acc, float* g_wsum, float* g_acc_irr,
                                      int n_rays, cudaStream_t stream) {
  if (int rc = check_cfg(cfg)) return rc;
  BN_CHECK_ARG(rays && acc && wsum && g_rgb && g_acc && g_wsum, "null pointer");
  BN_CHECK_ARG(!(acc_irr && cfg->irr_mode == BN_IRR_SUNVIS) || g_acc_irr, "acc_irr given without g_acc_irr");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  ShadeRays a{};
  a.cfg = *cfg; a.rays = rays; a.acc = acc; a.wsum = wsum; a.acc_irr = acc_irr; a.irr_last = irr_last;
  a.g_rgb = g_rgb; a.g_acc = g_acc; a.g_wsum = g_wsum; a.g_acc_irr = g_acc_irr; a.N = n_rays;
  shade_rays_kernel<true><<<ceil_div(n_rays, 64), 64, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_brdf_points_forward(const bn_shade_cfg* cfg, const float* rays, float* packed, float* aux,
                                      int n_rays, int n_samples, cudaStream_t stream) {
  if (int rc = check_cfg(cfg)) return rc;
  BN_CHECK_ARG(rays && packed, "null pointer");
  BN_CHECK_ARG(cfg->multi_brdf && cfg->brdf_type != BN_BRDF_NONE, "per-sample BRDF requires multi_brdf and a BRDF type");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  BrdfPoints a{*cfg, rays, packed, nullptr, aux, n_rays, n_samples};
  long long P = (long long)n_rays * n_samples;
  brdf_points_kernel<false><<<(unsigned)ceil_div_ll(P, 128), 128, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}

extern "C" __attribute__((visibility("default"))) int bn_brdf_points_backward(const bn_shade_cfg* cfg, const float* rays, const float* packed,
                                       float* g_packed, int n_rays, int n_samples, cudaStream_t stream) {
  if (int rc = check_cfg(cfg)) return rc;
  BN_CHECK_ARG(rays && packed && g_packed, "null pointer");
  BN_CHECK_ARG(cfg->multi_brdf && cfg->brdf_type != BN_BRDF_NONE, "per-sample BRDF requires multi_brdf and a BRDF type");
  if (n_rays <= 0) return n_rays == 0 ? BN_OK : BN_ERR_ARG;
  BrdfPoints a{*cfg, rays, const_cast<float*>(packed), g_packed, nullptr, n_rays, n_samples};
  long long P = (long long)n_rays * n_samples;
  brdf_points_kernel<true><<<(unsigned)ceil_div_ll(P, 64), 64, 0, stream>>>(a);
  BN_LAUNCH_CHECK();
  return BN_OK;
}
