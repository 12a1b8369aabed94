"""ORACLE (test infrastructure only — never imported by the product path).

torch (CPU, fp32 or fp64) restatement of the parametric BRDFs evaluated on the hot path.

Reference functions restated (paths relative to /root/reference):
  * calc_angles, Henyey_Greenstein         BRDF/basic_func.py:5-44
  * func_M1, func_G, func_H, RPV.calc_rpv  BRDF/RPV.py:6-63
  * E1,E2,f,chi,eta,mu0_eff,mu_eff,S,PF,HF BRDF/Hapke.py:6-131
  * Hapke.hapkeHG_6var                     BRDF/Hapke.py:139-200
  * Microfacet.forward/_get_d/_get_g/_get_f BRDF/microfacet.py:20-118

NaN policy: the reference replaces NaNs by a per-function fallback (train_utils.check_nan);
`_nan_to` below is that replacement without the printing / host sync.
"""
from __future__ import annotations

import math

import torch

PI = math.pi


def _nan_to(y, rep):
    return torch.where(torch.isnan(y), rep, y)


def _dot(a, b):
    return (a * b).sum(-1)


def angles(light, view, normal, eps=1e-5):
    """basic_func.py:5-31. light/view/normal: (M,3). Returns dict of (M,) tensors."""
    ci = _dot(light, normal).clamp(eps, 1.0)
    cv = _dot(view, normal).clamp(eps, 1.0)
    cg = _dot(view, light).clamp(-1.0, 1.0)
    sza, vza, g = torch.acos(ci), torch.acos(cv), torch.acos(cg)
    si, sv = torch.sin(sza), torch.sin(vza)
    cphi = ((cg - ci * cv) / si / sv).clamp(-1.0, 1.0)
    return dict(ci=ci, cv=cv, cg=cg, sza=sza, vza=vza, g=g, si=si, sv=sv, phi=torch.acos(cphi))


def henyey_greenstein(x, th, eps=1e-6):
    """basic_func.py:33-44. x: (M,1) cos(phase); th: (M,3)."""
    t2 = th * th
    y = (1 - t2) / (torch.pow(1 + 2 * th * x + t2, 1.5) + eps)
    return _nan_to(y, torch.zeros_like(y))


def rpv(light, view, normal, w, k=None, theta=None, rhoc=None):
    """RPV.py:39-63. Returns (brdf, M1, G, H, ci, cv)."""
    a = angles(light, view, normal)
    ci, cv, cg = a["ci"].unsqueeze(-1), a["cv"].unsqueeze(-1), a["cg"].unsqueeze(-1)
    if k is not None:
        base = ci * cv * (ci + cv) + 1e-5                       # RPV.py:6-18
        M1 = torch.pow(base, k - 1)
        M1 = _nan_to(M1, torch.zeros_like(M1))
    else:
        M1 = torch.ones_like(ci)
    Fh = henyey_greenstein(cg, theta) if theta is not None else torch.ones_like(cg)
    if rhoc is not None:
        ti, tv, cp = torch.tan(a["sza"]), torch.tan(a["vza"]), torch.cos(a["phi"])   # RPV.py:20-30
        G = torch.sqrt(ti ** 2 + tv ** 2 - 2 * ti * tv * cp + 1e-5)
        G = _nan_to(G, torch.zeros_like(G)).unsqueeze(-1)
        H = 1 + (1 - rhoc) / (1 + G.detach() + 1e-5)            # RPV.py:32-35, G detached at :54
        H = _nan_to(H, torch.zeros_like(H))
    else:
        G = torch.ones_like(ci)
        H = torch.ones_like(ci)
    return w * M1 * Fh * H, M1, G, H, a["ci"], a["cv"]


# ---------------------------------------------------------------- Hapke
def _E1(x, th, eps=1e-5):
    y = torch.exp(-(2.0 / PI) * 1.0 / torch.tan(th + eps) * 1.0 / torch.tan(x + eps))
    return _nan_to(y, torch.zeros_like(y))


def _E2(x, th, eps=1e-5):
    y = torch.exp(-(1.0 / PI) * (1.0 / torch.tan(th + eps)) ** 2 * (1.0 / torch.tan(x + eps)) ** 2)
    return _nan_to(y, torch.zeros_like(y))


def _f(phi, eps=1e-5):
    y = torch.exp(-2.0 * torch.tan((phi + eps) / 2))
    return _nan_to(y, torch.zeros_like(y))


def _chi(x, eps=1e-5):
    y = 1.0 / torch.sqrt(1.0 + PI * torch.tan(x + eps) ** 2)
    return _nan_to(y, torch.zeros_like(y))


def _eta(x, th, eps=1e-5):
    y = _chi(th) * (torch.cos(x) + torch.sin(x) * torch.tan(th + eps) * (_E2(x, th) / (2 - _E1(x, th))))
    return _nan_to(y, torch.zeros_like(y))


def _branch(i, e, fn_le, fn_gt):
    """Evaluate fn_le on rows with i<=e and fn_gt on the others (Hapke.py index-assign pattern)."""
    y = torch.zeros_like(e)
    m1 = i <= e
    m2 = ~m1
    if m1.any():
        y = y.masked_scatter(m1, fn_le(m1))
    if m2.any():
        y = y.masked_scatter(m2, fn_gt(m2))
    return y


def _mu0_eff(i, e, phi, th):
    """Hapke.py:32-48."""
    def le(m):
        ii, ee, pp, tt = i[m], e[m], phi[m], th[m]
        y = torch.cos(pp) * _E2(ee, tt) + torch.sin(pp / 2) ** 2 * _E2(ii, tt)
        y = y / (2 - _E1(ee, tt) - pp / PI * _E1(ii, tt))
        return _chi(tt) * (torch.cos(ii) + torch.sin(ii) * torch.tan(tt) * y)

    def gt(m):
        ii, ee, pp, tt = i[m], e[m], phi[m], th[m]
        y = _E2(ii, tt) - torch.sin(pp / 2) ** 2 * _E2(ee, tt)
        y = y / (2 - _E1(ii, tt) - pp / PI * _E1(ee, tt))
        return _chi(tt) * (torch.cos(ii) + torch.sin(ii) * torch.tan(tt) * y)
    return _nan_to(_branch(i, e, le, gt), torch.cos(i))


def _mu_eff(i, e, phi, th):
    """Hapke.py:50-66."""
    def le(m):
        ii, ee, pp, tt = i[m], e[m], phi[m], th[m]
        y = _E2(ee, tt) - torch.sin(pp / 2) ** 2 * _E2(ii, tt)
        y = y / (2 - _E1(ee, tt) - (pp / PI) * _E1(ii, tt))
        return _chi(tt) * (torch.cos(ee) + torch.sin(ee) * torch.tan(tt) * y)

    def gt(m):
        ii, ee, pp, tt = i[m], e[m], phi[m], th[m]
        y = torch.cos(pp) * _E2(ii, tt) + torch.sin(pp / 2) ** 2 * _E2(ee, tt)
        y = y / (2 - _E1(ii, tt) - (pp / PI) * _E1(ee, tt))
        return _chi(tt) * (torch.cos(ee) + torch.sin(ee) * torch.tan(tt) * y)
    return _nan_to(_branch(i, e, le, gt), torch.cos(e))


def _shadow(i, e, phi, th):
    """Hapke.py:68-91."""
    ci, cv = torch.cos(i), torch.cos(e)
    mue = _mu_eff(i, e, phi, th)
    etai, etae, chit, ff = _eta(i, th), _eta(e, th), _chi(th), _f(phi)
    temp = (mue / etae) * (ci / etai) * chit
    y = _branch(i, e,
                lambda m: temp[m] / (1 - ff[m] + ff[m] * chit[m] * (ci[m] / etai[m])),
                lambda m: temp[m] / (1 - ff[m] + ff[m] * chit[m] * (cv[m] / etae[m])))
    return _nan_to(y, torch.zeros_like(y))


def _phase_double_hg(x, b, c):
    """Hapke.py:93-115."""
    b2, bx = b * b, b * x
    y = c * (1 - b2) / (torch.pow(1 - 2 * bx + b2, 1.5) + 1e-6)
    y = y + (1 - c) * (1 - b2) / (torch.pow(1 + 2 * bx + b2, 1.5) + 1e-6)
    return _nan_to(y, torch.zeros_like(y))


def _chandrasekhar(x, w):
    """Hapke.py:117-131. x: (M,1), w: (M,3)."""
    gamma = torch.sqrt(1 - w)
    r0 = (1 - gamma) / (1 + gamma)
    lg = torch.log(torch.abs((1 + x) / x))
    y = torch.pow(1 - w * x * (r0 + (1 - 2 * r0 * x) / 2 * lg), -1)
    return _nan_to(y, torch.ones_like(y))


def hapke(light, view, normal, w, b=None, c=None, theta=None, hpk_scl=4.0, shell_hapke=0):
    """Hapke.py:139-200 (B0/h are always None on the path, spsbrdfnerf.py:322 => B == 1).
    Returns (brdf, P, B, Hi, Hv, ShadFunc, ci, cv)."""
    a = angles(light, view, normal)
    ci, cv, cg = a["ci"], a["cv"], a["cg"]
    if b is None:
        P = torch.ones_like(cg).unsqueeze(-1).repeat(1, 3)
    elif c is None:
        P = henyey_greenstein(cg.unsqueeze(-1), b)
    else:
        P = _phase_double_hg(cg.unsqueeze(-1), b, c)
    B = torch.ones_like(a["g"]).unsqueeze(-1)
    if theta is not None:
        ci = _mu0_eff(a["sza"], a["vza"], a["phi"], theta)
        cv = _mu_eff(a["sza"], a["vza"], a["phi"], theta)
        shad = _shadow(a["sza"], a["vza"], a["phi"], theta).unsqueeze(-1)
    else:
        shad = torch.ones_like(a["sza"]).unsqueeze(-1)
    Hi = _chandrasekhar(ci.unsqueeze(-1), w)
    Hv = _chandrasekhar(cv.unsqueeze(-1), w)
    if b is None:
        if shell_hapke == 1:
            brdf = w / hpk_scl
        elif shell_hapke == 2:
            brdf = w / ((ci + cv) * hpk_scl + 1e-6).unsqueeze(-1)
        elif shell_hapke == 3:
            brdf = w * (Hi * Hv) / ((ci + cv) * hpk_scl + 1e-6).unsqueeze(-1)
        else:
            raise ValueError("Hapke without b requires shell_hapke in {1,2,3}")
    else:
        geo = (ci / (ci + cv) / torch.cos(a["sza"])).unsqueeze(-1)
        brdf = w / hpk_scl * geo * (P * B + Hi * Hv - 1) * shad
    return brdf, P, B, Hi, Hv, shad, ci, cv


# ---------------------------------------------------------------- microfacet (GGX)
def _safe_unit(x, eps=1e-6):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


def microfacet(light, view, normal, albedo, rough, f0=0.04):
    """microfacet.py:20-118 with one light (L == 1, squeezed).  Returns the reference's 9-tuple
    (glossy, brdf, f, g, d, l_dot_n, v_dot_n, halfvec, n_h) with the light axis kept as size 1
    where the reference keeps it."""
    l, v, n = _safe_unit(light), _safe_unit(view), _safe_unit(normal)
    h = _safe_unit(l + v)
    f = f0 + (1 - f0) * (1 - _dot(l, h)) ** 5                               # _get_f
    alpha = rough ** 2                                                        # (M,1)
    a2 = alpha.squeeze(-1) ** 2
    # _get_d
    cm = _dot(h, n)
    chi_d = (cm > 0).to(cm.dtype)
    cm2 = cm * cm
    tan2 = torch.nan_to_num((1 - cm2) / cm2)
    d = torch.nan_to_num(a2 * chi_d / (PI * cm2 * cm2 * (a2 + tan2) ** 2))
    # _get_g (view side only; lvis is False on the path, spsbrdfnerf.py:284)
    cvn = _dot(n, v)
    chi_g = (torch.nan_to_num(_dot(h, v) / cvn) > 0).to(cm.dtype)
    cvn2 = (cvn * cvn).clamp(0.0, 1.0)
    t2 = torch.nan_to_num(torch.nan_to_num((1 - cvn2) / cvn2).clamp(min=0.0))
    g = torch.nan_to_num(chi_g * 2 / (1 + torch.sqrt(1 + a2 * t2)))
    ldn = _dot(l, n).abs().clamp_min(0.001)
    vdn = _dot(v, n).abs().clamp_min(0.001)
    glossy = torch.nan_to_num(0.04 * d / (4 * ldn * vdn))
    brdf = albedo + glossy.unsqueeze(-1)
    u = lambda t: t.unsqueeze(-1)
    return u(glossy), brdf.unsqueeze(1), u(f), u(g), u(d), u(ldn), vdn, h.unsqueeze(1), u(cm)
