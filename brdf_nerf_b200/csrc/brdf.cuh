// K-C (part 2): RPV / Hapke / GGX-microfacet BRDF evaluation, written once over a scalar type T.
// T = float gives the forward value; T = Dual<NV> (forward-mode dual number) gives the exact
// Jacobian row w.r.t. NV inputs, which the backward kernels contract with the upstream gradient.
//
// Replaces (reference, paths relative to /root/reference):
//   calc_angles, Henyey_Greenstein       BRDF/basic_func.py:5-44
//   func_M1 / func_G / func_H / calc_rpv BRDF/RPV.py:6-63
//   E1 E2 f chi eta mu0_eff mu_eff S PF HF hapkeHG_6var   BRDF/Hapke.py:6-200
//   Microfacet.forward / _get_d / _get_g / _get_f           BRDF/microfacet.py:20-118
// NaN policy: the reference's check_nan(val, val_rep) replacements (train_utils.py:61-78) are kept
// in-kernel (`nan_to`): a replaced value carries a zero derivative, exactly like torch.where.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace bn {

template <int NV>
struct Dual {
  float v;
  float d[NV];
  __device__ Dual() {}
  __device__ Dual(float c) : v(c) {
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = 0.f;
  }
  __device__ static Dual var(float c, int idx) { Dual r(c); r.d[idx] = 1.f; return r; }
};

// ---- scalar overloads so the templated code reads the same for float and Dual ----
__device__ __forceinline__ float val(float a) { return a; }
template <int N> __device__ __forceinline__ float val(const Dual<N>& a) { return a.v; }

#define BN_DUAL_UNARY(name, fv, dfdv)                                              \
  template <int N> __device__ __forceinline__ Dual<N> name(const Dual<N>& a) {    \
    Dual<N> r; const float x = a.v; const float y = fv; r.v = y; const float g = dfdv; \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i];             \
    return r; }

template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const float inv = 1.0f / b.v; r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, float b) { Dual<N> r = a; r.v += b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(float b, const Dual<N>& a) { return a + b; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, float b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(float b, const Dual<N>& a) { return (-a) + b; }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, float b) {
  Dual<N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(float b, const Dual<N>& a) { return a * b; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, float b) { return a * (1.0f / b); }
template <int N> __device__ __forceinline__ Dual<N> operator/(float b, const Dual<N>& a) { return Dual<N>(b) / a; }

BN_DUAL_UNARY(sqrt_, sqrtf(x), 0.5f / y)
BN_DUAL_UNARY(sin_, sinf(x), cosf(x))
BN_DUAL_UNARY(cos_, cosf(x), -sinf(x))
BN_DUAL_UNARY(tan_, tanf(x), 1.0f + y * y)
BN_DUAL_UNARY(acos_, acosf(x), -1.0f / sqrtf(1.0f - x * x))
BN_DUAL_UNARY(exp_, expf(x), y)
BN_DUAL_UNARY(log_, logf(x), 1.0f / x)
BN_DUAL_UNARY(abs_, fabsf(x), (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f))
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ float sin_(float x) { return sinf(x); }
__device__ __forceinline__ float cos_(float x) { return cosf(x); }
__device__ __forceinline__ float tan_(float x) { return tanf(x); }
__device__ __forceinline__ float acos_(float x) { return acosf(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ float log_(float x) { return logf(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }

// x^p with a constant exponent
__device__ __forceinline__ float powc(float x, float p) { return powf(x, p); }
template <int N> __device__ __forceinline__ Dual<N> powc(const Dual<N>& a, float p) {
  Dual<N> r; r.v = powf(a.v, p); const float g = p * powf(a.v, p - 1.0f);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = g * a.d[i];
  return r; }
// x^y, both variable (torch.pow backward: y x^(y-1) dx + x^y ln x dy)
__device__ __forceinline__ float powv(float x, float y) { return powf(x, y); }
template <int N> __device__ __forceinline__ Dual<N> powv(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = powf(a.v, b.v);
  const float ga = b.v * powf(a.v, b.v - 1.0f);
  const float gb = (a.v == 0.f && b.v >= 0.f) ? 0.f : r.v * logf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = ga * a.d[i] + gb * b.d[i];
  return r; }

// torch.clamp: gradient passes where lo <= x <= hi
__device__ __forceinline__ float clamp_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
template <int N> __device__ __forceinline__ Dual<N> clamp_(const Dual<N>& a, float lo, float hi) {
  if (a.v < lo) return Dual<N>(lo);
  if (a.v > hi) return Dual<N>(hi);
  return a; }
__device__ __forceinline__ float nan_to(float x, float rep) { return isnan(x) ? rep : x; }
template <int N> __device__ __forceinline__ Dual<N> nan_to(const Dual<N>& a, float rep) { return isnan(a.v) ? Dual<N>(rep) : a; }
template <int N> __device__ __forceinline__ Dual<N> nan_to(const Dual<N>& a, const Dual<N>& rep) { return isnan(a.v) ? rep : a; }
// constant stop-gradient (RPV.py:54 detaches G)
__device__ __forceinline__ float detach(float x) { return x; }
template <int N> __device__ __forceinline__ float detach(const Dual<N>& a) { return a.v; }

template <class T> struct Angles { T ci, cv, cg, sza, vza, g, phi; };

// basic_func.py:5-31 — l, v are per-ray constants, n carries the derivative
template <class T>
__device__ __forceinline__ Angles<T> calc_angles(const float (&l)[3], const float (&v)[3], const T (&n)[3]) {
  Angles<T> a;
  a.ci = clamp_(n[0] * l[0] + n[1] * l[1] + n[2] * l[2], 1e-5f, 1.0f);
  a.cv = clamp_(n[0] * v[0] + n[1] * v[1] + n[2] * v[2], 1e-5f, 1.0f);
  const float cg = fminf(fmaxf(v[0] * l[0] + v[1] * l[1] + v[2] * l[2], -1.0f), 1.0f);
  a.cg = T(cg);
  a.sza = acos_(a.ci); a.vza = acos_(a.cv); a.g = T(acosf(cg));
  T si = sin_(a.sza), sv = sin_(a.vza);
  a.phi = acos_(clamp_((a.cg - a.ci * a.cv) / si / sv, -1.0f, 1.0f));
  return a;
}

// basic_func.py:33-44
template <class T> __device__ __forceinline__ T henyey_greenstein(const T& x, const T& th) {
  T t2 = th * th;
  T y = (1.0f - t2) / (powc(1.0f + 2.0f * th * x + t2, 1.5f) + 1e-6f);
  return nan_to(y, 0.0f);
}

// RPV.py:39-63 for one colour channel. has_* select the active factors.
template <class T>
__device__ __forceinline__ T rpv_channel(const Angles<T>& a, const T& w, bool has_k, const T& k,
                                         bool has_th, const T& th, bool has_rc, const T& rc,
                                         float* M1o = nullptr, float* Go = nullptr, float* Ho = nullptr) {
  T M1 = T(1.0f), F = T(1.0f), H = T(1.0f);
  float Gv = 1.0f;
  if (has_k) M1 = nan_to(powv(a.ci * a.cv * (a.ci + a.cv) + 1e-5f, k - 1.0f), 0.0f);
  if (has_th) F = henyey_greenstein(a.cg, th);
  if (has_rc) {
    const float ti = tanf(val(a.sza)), tv = tanf(val(a.vza)), cp = cosf(val(a.phi));
    Gv = nan_to(sqrtf(ti * ti + tv * tv - 2.0f * ti * tv * cp + 1e-5f), 0.0f);   // detached (RPV.py:54)
    H = nan_to(1.0f + (1.0f - rc) / (1.0f + Gv + 1e-5f), 0.0f);
  }
  if (M1o) *M1o = val(M1);
  if (Go) *Go = Gv;
  if (Ho) *Ho = val(H);
  return w * M1 * F * H;
}

// ---------------------------------------------------------------- Hapke
constexpr float kPi = 3.14159265358979323846f;
template <class T> __device__ __forceinline__ T hk_E1(const T& x, const T& th) {
  return nan_to(exp_(-(2.0f / kPi) * (1.0f / tan_(th + 1e-5f)) * (1.0f / tan_(x + 1e-5f))), 0.0f); }
template <class T> __device__ __forceinline__ T hk_E2(const T& x, const T& th) {
  T a = 1.0f / tan_(th + 1e-5f), b = 1.0f / tan_(x + 1e-5f);
  return nan_to(exp_(-(1.0f / kPi) * (a * a) * (b * b)), 0.0f); }
template <class T> __device__ __forceinline__ T hk_f(const T& phi) {
  return nan_to(exp_(-2.0f * tan_((phi + 1e-5f) / 2.0f)), 0.0f); }
template <class T> __device__ __forceinline__ T hk_chi(const T& x) {
  T t = tan_(x + 1e-5f);
  return nan_to(1.0f / sqrt_(1.0f + kPi * (t * t)), 0.0f); }
template <class T> __device__ __forceinline__ T hk_eta(const T& x, const T& th) {
  return nan_to(hk_chi(th) * (cos_(x) + sin_(x) * tan_(th + 1e-5f) * (hk_E2(x, th) / (2.0f - hk_E1(x, th)))), 0.0f); }

// Hapke.py:32-48
template <class T> __device__ __forceinline__ T hk_mu0_eff(const T& i, const T& e, const T& phi, const T& th) {
  T sp = sin_(phi / 2.0f); T y;
  if (val(i) <= val(e)) {
    y = cos_(phi) * hk_E2(e, th) + sp * sp * hk_E2(i, th);
    y = y / (2.0f - hk_E1(e, th) - phi / kPi * hk_E1(i, th));
  } else {
    y = hk_E2(i, th) - sp * sp * hk_E2(e, th);
    y = y / (2.0f - hk_E1(i, th) - phi / kPi * hk_E1(e, th));
  }
  y = hk_chi(th) * (cos_(i) + sin_(i) * tan_(th) * y);
  return nan_to(y, cos_(i));
}
// Hapke.py:50-66
template <class T> __device__ __forceinline__ T hk_mu_eff(const T& i, const T& e, const T& phi, const T& th) {
  T sp = sin_(phi / 2.0f); T y;
  if (val(i) <= val(e)) {
    y = hk_E2(e, th) - sp * sp * hk_E2(i, th);
    y = y / (2.0f - hk_E1(e, th) - (phi / kPi) * hk_E1(i, th));
  } else {
    y = cos_(phi) * hk_E2(i, th) + sp * sp * hk_E2(e, th);
    y = y / (2.0f - hk_E1(i, th) - (phi / kPi) * hk_E1(e, th));
  }
  y = hk_chi(th) * (cos_(e) + sin_(e) * tan_(th) * y);
  return nan_to(y, cos_(e));
}
// Hapke.py:68-91
template <class T> __device__ __forceinline__ T hk_shadow(const T& i, const T& e, const T& phi, const T& th) {
  T ci = cos_(i), cv = cos_(e);
  T mue = hk_mu_eff(i, e, phi, th);
  T etai = hk_eta(i, th), etae = hk_eta(e, th), chit = hk_chi(th), ff = hk_f(phi);
  T temp = (mue / etae) * (ci / etai) * chit;
  T y = (val(i) <= val(e)) ? temp / (1.0f - ff + ff * chit * (ci / etai))
                           : temp / (1.0f - ff + ff * chit * (cv / etae));
  return nan_to(y, 0.0f);
}
// Hapke.py:93-115
template <class T> __device__ __forceinline__ T hk_phase2(const T& x, const T& b, const T& c) {
  T b2 = b * b, bx = b * x;
  T y = c * (1.0f - b2) / (powc(1.0f - 2.0f * bx + b2, 1.5f) + 1e-6f);
  y = y + (1.0f - c) * (1.0f - b2) / (powc(1.0f + 2.0f * bx + b2, 1.5f) + 1e-6f);
  return nan_to(y, 0.0f);
}
// Hapke.py:117-131
template <class T> __device__ __forceinline__ T hk_H(const T& x, const T& w) {
  T gamma = sqrt_(1.0f - w);
  T r0 = (1.0f - gamma) / (1.0f + gamma);
  T lg = log_(abs_((1.0f + x) / x));
  T y = 1.0f / (1.0f - w * x * (r0 + (1.0f - 2.0f * r0 * x) / 2.0f * lg));
  return nan_to(y, 1.0f);
}

struct HapkeAux { float P, Hi, Hv, ci, cv, shad; };
// Hapke.py:139-200 for one colour channel (B == 1: B0/h are None on the path, spsbrdfnerf.py:322)
template <class T>
__device__ __forceinline__ T hapke_channel(const Angles<T>& a, const T& w, bool has_b, const T& b, bool has_c,
                                           const T& c, bool has_th, const T& th, float hpk_scl, int shell,
                                           HapkeAux* aux = nullptr) {
  T P = T(1.0f);
  if (has_b) P = has_c ? hk_phase2(a.cg, b, c) : henyey_greenstein(a.cg, b);
  T ci = a.ci, cv = a.cv, shad = T(1.0f);
  if (has_th) {
    ci = hk_mu0_eff(a.sza, a.vza, a.phi, th);
    cv = hk_mu_eff(a.sza, a.vza, a.phi, th);
    shad = hk_shadow(a.sza, a.vza, a.phi, th);
  }
  T Hi = hk_H(ci, w), Hv = hk_H(cv, w);
  T out;
  if (!has_b) {
    if (shell == 1) out = w / hpk_scl;
    else if (shell == 2) out = w / ((ci + cv) * hpk_scl + 1e-6f);
    else out = w * (Hi * Hv) / ((ci + cv) * hpk_scl + 1e-6f);
  } else {
    T geo = ci / (ci + cv) / cos_(a.sza);
    out = w / hpk_scl * geo * (P + Hi * Hv - 1.0f) * shad;
  }
  if (aux) { aux->P = val(P); aux->Hi = val(Hi); aux->Hv = val(Hv); aux->ci = val(ci); aux->cv = val(cv); aux->shad = val(shad); }
  return out;
}

// ---------------------------------------------------------------- GGX microfacet
struct MicroAux { float glossy, f, g, d, ldn, vdn, h[3], nh; };
__device__ __forceinline__ float nan_to_num_(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return x; }
template <int N> __device__ __forceinline__ Dual<N> nan_to_num_(const Dual<N>& a) {
  if (isnan(a.v) || isinf(a.v)) return Dual<N>(nan_to_num_(a.v));
  return a; }

// microfacet.py:20-69; returns the glossy term (identical for the three colour channels);
// brdf_c = albedo_c + glossy.  l, v are re-normalised with eps 1e-6 as in safe_l2_normalize.
template <class T>
__device__ __forceinline__ T microfacet_glossy(const float (&l_in)[3], const float (&v_in)[3], const T (&n_in)[3],
                                               const T& rough, float f0, MicroAux* aux = nullptr) {
  float l[3], v[3], h[3];
  float ln = fmaxf(sqrtf(l_in[0] * l_in[0] + l_in[1] * l_in[1] + l_in[2] * l_in[2]), 1e-6f);
  float vn = fmaxf(sqrtf(v_in[0] * v_in[0] + v_in[1] * v_in[1] + v_in[2] * v_in[2]), 1e-6f);
#pragma unroll
  for (int i = 0; i < 3; ++i) { l[i] = l_in[i] / ln; v[i] = v_in[i] / vn; h[i] = l[i] + v[i]; }
  float hn = fmaxf(sqrtf(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), 1e-6f);
#pragma unroll
  for (int i = 0; i < 3; ++i) h[i] /= hn;
  T nn = sqrt_(n_in[0] * n_in[0] + n_in[1] * n_in[1] + n_in[2] * n_in[2]);
  if (val(nn) < 1e-6f) nn = T(1e-6f);
  T n[3] = {n_in[0] / nn, n_in[1] / nn, n_in[2] / nn};
  T alpha = rough * rough;
  T a2 = alpha * alpha;
  // _get_d
  T cm = n[0] * h[0] + n[1] * h[1] + n[2] * h[2];
  const float chi_d = val(cm) > 0.f ? 1.f : 0.f;
  T cm2 = cm * cm;
  T tan2 = nan_to_num_((1.0f - cm2) / cm2);
  T den = kPi * (cm2 * cm2) * ((a2 + tan2) * (a2 + tan2));
  T d = nan_to_num_(a2 * chi_d / den);
  T ldn = abs_(n[0] * l[0] + n[1] * l[1] + n[2] * l[2]);
  if (val(ldn) < 0.001f) ldn = T(0.001f);
  T vdn = abs_(n[0] * v[0] + n[1] * v[1] + n[2] * v[2]);
  if (val(vdn) < 0.001f) vdn = T(0.001f);
  T glossy = nan_to_num_(0.04f * d / (4.0f * ldn * vdn));
  if (aux) {
    // Fresnel and geometry terms are outputs only (unused in brdf, microfacet.py:57-69)
    float ldh = l[0] * h[0] + l[1] * h[1] + l[2] * h[2];
    float om = 1.0f - ldh;
    aux->f = f0 + (1.0f - f0) * om * om * om * om * om;
    float cvn = val(n[0]) * v[0] + val(n[1]) * v[1] + val(n[2]) * v[2];
    float hv = h[0] * v[0] + h[1] * v[1] + h[2] * v[2];
    float chi_g = nan_to_num_(hv / cvn) > 0.f ? 1.f : 0.f;
    float cvn2 = fminf(fmaxf(cvn * cvn, 0.f), 1.f);
    float t2 = nan_to_num_(fmaxf(nan_to_num_((1.0f - cvn2) / cvn2), 0.f));
    aux->g = nan_to_num_(chi_g * 2.0f / (1.0f + sqrtf(1.0f + val(a2) * t2)));
    aux->glossy = val(glossy); aux->d = val(d); aux->ldn = val(ldn); aux->vdn = val(vdn);
    aux->h[0] = h[0]; aux->h[1] = h[1]; aux->h[2] = h[2]; aux->nh = val(cm);
  }
  return glossy;
}

}  // namespace bn
