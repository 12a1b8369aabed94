"""ORACLE (test infrastructure only — never imported by the product path).

torch-CPU (fp32 or fp64) restatement of the reference's ray-rendering hot path with every random
draw injected explicitly.  It is the checker for the CUDA path and the CPU baseline of bench.py;
autograd through it is the gradient oracle.

Reference functions restated (paths relative to /root/reference):
  * Mapping.forward                 models/nerf.py:53-70
  * SpSBRDFNeRF.calc_features       models/spsbrdfnerf.py:636-646  (layers built :513-524)
  * SpSBRDFNeRF.forward             models/spsbrdfnerf.py:662-757
  * SpSBRDFNeRF.calc_normals        models/spsbrdfnerf.py:648-660
  * cal_weight                      models/spsbrdfnerf.py:50-69
  * inference                       models/spsbrdfnerf.py:71-416
  * render_rays (spsbrdf-nerf arm)  rendering.py:168-291
  * l2_normalize                    train_utils.py:28-33
The sample generators live in oracle/sampler_np.py (bit-exact numpy restatement); here they are
called through `sampler_np` so the whole pipeline is one executable specification.

Pinning: tests/test_oracle_vs_reference.py compares every result key against the live reference
(run with the same draws injected into torch.rand / rand_like / randn); tests/golden/*.npz hold
outputs of the live reference produced by oracle/make_golden.py.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as Fnn

from . import brdf_torch as brdf
from . import sampler_np

F32_EPS = float(torch.finfo(torch.float32).eps)
RGB_PAD = 0.001          # spsbrdfnerf.py:459


@dataclasses.dataclass
class Draws:
    """Random draws of one render_rays call in the reference's order."""
    u_strat: torch.Tensor                    # (N,S1) rand_like, rendering.py:163
    noise1: torch.Tensor                     # (N,S1) randn, spsbrdfnerf.py:58 (pass 1)
    u_pred: torch.Tensor                     # (N,G)  rand, rendering.py:36
    noise2: torch.Tensor                     # (N,S)  randn, spsbrdfnerf.py:58 (pass 2)
    u_gt: Optional[torch.Tensor] = None      # (N,G)  rows of valid rays consumed, rendering.py:144
    u_sun: Optional[torch.Tensor] = None     # (N,S') rendering.py:253
    noise_sun: Optional[torch.Tensor] = None

    @staticmethod
    def make(n, s1, g, s2, seed, with_gt=False, with_sun=False, s_sun=None):
        gen = torch.Generator().manual_seed(seed)
        r = lambda *sh: torch.rand(*sh, generator=gen)
        rn = lambda *sh: torch.randn(*sh, generator=gen)
        d = Draws(u_strat=r(n, s1), noise1=rn(n, s1), u_pred=r(n, g), noise2=rn(n, s2))
        if with_gt:
            d.u_gt = r(n, g)
        if with_sun:
            d.u_sun = r(n, s_sun or s1)
            d.noise_sun = rn(n, s_sun or s1)
        return d


def l2_normalize(x):
    sq = (x * x).sum(-1, keepdim=True)
    return x / torch.sqrt(torch.clamp_min(sq, F32_EPS))


def fourier(x, n_freq):
    """[sin(2^0 x), cos(2^0 x), ..., cos(2^(n-1) x)], no identity term."""
    if n_freq == 0:
        return x
    out = []
    for k in range(n_freq):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


class OracleModel:
    """Functional view over a reference-compatible state_dict."""

    def __init__(self, state: Dict[str, torch.Tensor], args, dtype=torch.float32, requires_grad=False,
                 mapping_sizes=(10, 4), skips=(4,)):
        self.args = args
        self.p = {k: v.detach().clone().to(dtype).requires_grad_(requires_grad) for k, v in state.items()}
        self.dtype = dtype
        self.layers = int(args.fc_layers)
        self.skips = tuple(skips)
        self.nf_xyz = mapping_sizes[0] if args.mapping else 0
        self.nf_dir = mapping_sizes[1] if args.mapping else 0
        self.viewdir = bool(args.input_viewdir)
        self.normal = args.normal
        self.roughness = bool(args.roughness)
        self.RPV = bool(args.funcM or args.funcF or args.funcH)
        self.MultiBRDF = bool(args.MultiBRDF)
        self.sun_v = args.sun_v
        self.beta = bool(getattr(args, "beta", False))

    def parameters(self):
        return list(self.p.values())

    def lin(self, name, x):
        return Fnn.linear(x, self.p[name + ".weight"], self.p[name + ".bias"])

    def trunk(self, x):
        enc = fourier(x, self.nf_xyz)
        h = enc
        for i in range(self.layers):
            if i in self.skips:
                h = torch.cat([enc, h], -1)
            h = torch.sin((30.0 if i == 0 else 1.0) * self.lin(f"fc_net.{2 * i}", h))
        return h

    def head(self, name, feats):
        return torch.sigmoid(self.lin(name + ".2", torch.sin(self.lin(name + ".0", feats))))

    def forward(self, x, d=None, sigma_only=False, apply_brdf=False, apply_theta=False,
                nr_an=False, nr_lr=False, t=None):
        """Per-point outputs as a dict (the reference packs them into channels, :694-757)."""
        if nr_an:
            x = x if x.requires_grad else x.detach().requires_grad_(True)
        with torch.enable_grad() if nr_an else _nullctx():
            h = self.trunk(x)
            sigma = Fnn.softplus(self.lin("sigma_from_xyz.0", h))
        out = {"sigma": sigma}
        if sigma_only:
            return out
        feats = self.lin("feats_from_xyz", h)
        rgb_in = torch.cat([feats, fourier(d, self.nf_dir)], -1) if self.viewdir else feats
        out["albedo"] = torch.sigmoid(self.lin("rgb_from_xyzdir.2", torch.sin(self.lin("rgb_from_xyzdir.0", rgb_in))))
        if self.beta:                                   # spsbrdfnerf.py:708-711: softplus(beta_from_xyz([features | t]))
            hid = torch.sin(self.lin("beta_from_xyz.0", torch.cat([feats, t.to(feats.dtype)], -1)))
            out["beta"] = Fnn.softplus(self.lin("beta_from_xyz.2", hid))
        if nr_an:
            with torch.enable_grad():
                keep = torch.is_grad_enabled()
                (g,) = torch.autograd.grad(sigma, x, torch.ones_like(sigma), create_graph=True, retain_graph=True)
            out["normal_an"] = -l2_normalize(g)
        if nr_lr:
            out["normal_lr"] = -l2_normalize(self.lin("grad_from_xyz", h))
        a = self.args
        tile3 = lambda t: t.repeat(1, 3) if t.shape[1] == 1 else t
        if apply_brdf:
            if self.roughness:
                out["roughness"] = self.head("roughness_from_xyz", feats)
            elif self.RPV:
                if a.funcM:
                    out["rpv_k"] = tile3((self.head("k_from_xyz", feats) - 0.5) * 2 + 1)
                if a.funcF:
                    out["rpv_theta"] = tile3((self.head("theta_rpv_from_xyz", feats) - 0.5) * 2)
                if a.funcH:
                    out["rpv_rhoc"] = tile3(self.head("rhoc_from_xyz", feats))
            else:
                if a.b:
                    out["hpk_b"] = tile3(self.head("b_from_xyz", feats))
                if a.c:
                    out["hpk_c"] = tile3(self.head("c_from_xyz", feats))
                if apply_theta and a.theta:
                    out["hpk_theta"] = self.head("theta_from_xyz", feats) * (np.pi * 30.0 / 180.0)
        return out


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def volume_weights(z, sigma, noise, noise_std):
    """cal_weight (spsbrdfnerf.py:50-69). z,sigma,noise: (N,S)."""
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1)
    alpha = 1 - torch.exp(-delta * torch.relu(sigma + noise * noise_std))
    shifted = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1)
    trans = torch.cumprod(shifted, -1)[:, :-1]
    w = alpha * trans
    return alpha, trans, w, (w * z).sum(-1)


def _points(rays_o, rays_d, z):
    return rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(2)


def _run_model(model, xyz, rays_d, chunk, rays_t=None, **kw):
    """The reference evaluates the MLP in chunks of `args.chunk` points (spsbrdfnerf.py:119-125);
    results are independent of the chunking, the oracle keeps it only to bound memory."""
    n, s = xyz.shape[:2]
    pts = xyz.reshape(-1, 3)
    dirs = None if rays_d is None else torch.repeat_interleave(rays_d, s, dim=0)
    ts_ = None if rays_t is None else torch.repeat_interleave(rays_t, s, dim=0)          # spsbrdfnerf.py:98
    outs = []
    for i in range(0, pts.shape[0], chunk):
        outs.append(model.forward(pts[i:i + chunk], None if dirs is None else dirs[i:i + chunk],
                                  t=None if ts_ is None else ts_[i:i + chunk], **kw))
    return {k: torch.cat([o[k] for o in outs], 0).reshape(n, s, -1) for k in outs[0]}


def shade(model, args, per_pt, z, rays_d, sun_d, noise, apply_brdf, apply_theta, cos_irra_on,
          sun_res=None, sort_idx=None, z_unsort=None):
    """`inference` after the MLP (spsbrdfnerf.py:138-416): compositing, irradiance, BRDF, result dict."""
    N, S = z.shape
    sig = per_pt["sigma"].reshape(N, S)
    albedo = per_pt["albedo"]
    alpha, trans, w, depth = volume_weights(z, sig, noise, args.noise_std)
    wv = w.unsqueeze(-1)
    res = {"sigmas": sig.unsqueeze(-1), "albedo": albedo,
           "albedo_accu": (wv * albedo).sum(-2).clamp(0.0, 1.0), "depth": depth, "alphas": alpha,
           "weights": w, "transparency": trans, "z_vals": z}
    sun_res = sun_res or {}
    apply_sun_v = model.sun_v == "analystic" and "sun" in sun_res
    if "sun" in sun_res:
        res["sun"] = sun_res["sun"]
        res["weights_sc"] = sun_res["weights_sc"]
    if sort_idx is not None:
        res["sort_idx"] = sort_idx
    if z_unsort is not None:
        res["z_vals_unsort"] = z_unsort
    if "beta" in per_pt:
        res["beta"] = per_pt["beta"]                             # :225-226
    normal = None
    if "normal_an" in per_pt:
        res["normal_an"] = normal = per_pt["normal_an"]
    if "normal_lr" in per_pt:
        res["normal_lr"] = normal = per_pt["normal_lr"]        # learned wins when both exist (:236-239)
    view = -rays_d
    if normal is not None:
        n_s = l2_normalize((wv * normal).sum(-2))
        res["nr_vw"] = (n_s * view).sum(-1).reshape(N, 1, 1)
        res["nr_sun"] = (n_s * sun_d).sum(-1).reshape(N, 1, 1)
        res["hpk_scl"] = 1.0 / (args.hpk_scl * (res["nr_vw"] + res["nr_sun"]))
    irr = torch.ones_like(albedo)
    if cos_irra_on and normal is not None:
        irr = irr * sun_d[:, 2].abs().reshape(N, 1, 1)          # up-vector . sun  (:260-264)
    elif apply_sun_v:
        irr = sun_res["sun"].repeat(1, 1, 3)
    alb_p = albedo * (1 + 2 * RGB_PAD) - RGB_PAD
    res["rgb"] = (wv * alb_p * irr).sum(-2).clamp(0.0, 1.0)
    albedo_s = (wv * alb_p).sum(-2)
    extra = [k for k in per_pt if k not in ("sigma", "albedo")]
    if not extra:
        return res, "Lambertian"                                  # early return (:281-282)

    brdf_type = "Lambertian"
    a = args
    M = model.MultiBRDF
    sun_pt = torch.repeat_interleave(sun_d, S, dim=0)
    view_pt = torch.repeat_interleave(view, S, dim=0)
    acc = lambda t: (wv * t).sum(-2)
    flat = lambda t: t.reshape(N * S, -1)
    bval = None
    aux = {}
    if model.roughness and apply_brdf:
        brdf_type = "Microfacet"
        rough = per_pt["roughness"]
        if M:
            tup = brdf.microfacet(sun_pt, view_pt, flat(normal), flat(albedo), flat(rough), f0=a.fresnel_f0)
        else:
            tup = brdf.microfacet(sun_d, view, n_s, albedo_s, (w * rough.reshape(N, S)).sum(-1, keepdim=True),
                                  f0=a.fresnel_f0)
        bval = tup[1]
        aux = dict(zip(("glossy", "brdf", "f", "g", "d", "l_dot_n", "v_dot_n", "halfvec", "n_h"), tup))
    elif model.RPV and apply_brdf:
        brdf_type = "RPV"
        if M:
            k = flat(per_pt["rpv_k"]) if a.funcM else None
            th = flat(per_pt["rpv_theta"]) if a.funcF else None
            rc = flat(albedo) if a.funcH == 2 else (flat(per_pt["rpv_rhoc"]) if a.funcH else None)
            bval = brdf.rpv(sun_pt, view_pt, flat(normal), flat(albedo), k, th, rc)[0]
        else:
            k = acc(per_pt["rpv_k"])                      # unconditional in the reference (:314)
            th = acc(per_pt["rpv_theta"]) if a.funcF else None
            rc = albedo_s if a.funcH == 2 else (acc(per_pt["rpv_rhoc"]) if a.funcH else None)
            bval = brdf.rpv(sun_d, view, n_s, albedo_s, k, th, rc)[0]
    elif (apply_brdf and a.b) or a.shell_hapke > 0:
        brdf_type = "Hapke"
        has_b = bool(apply_brdf and a.b)
        has_c = bool(apply_brdf and a.c)
        has_t = bool(apply_theta and a.theta)
        if M:
            tup = brdf.hapke(sun_pt, view_pt, flat(normal), flat(albedo),
                             flat(per_pt["hpk_b"]) if has_b else None,
                             flat(per_pt["hpk_c"]) if has_c else None,
                             per_pt["hpk_theta"].reshape(-1) if has_t else None,
                             hpk_scl=a.hpk_scl, shell_hapke=a.shell_hapke)
        else:
            tup = brdf.hapke(sun_d, view, n_s, albedo_s,
                             acc(per_pt["hpk_b"]) if has_b else None,
                             acc(per_pt["hpk_c"]) if has_c else None,
                             (w * per_pt["hpk_theta"].reshape(N, S)).sum(-1) if has_t else None,
                             hpk_scl=a.hpk_scl, shell_hapke=a.shell_hapke)
        bval = tup[0]
        aux = dict(zip(("brdf", "hpk_P", "hpk_B", "hpk_Hi", "hpk_Hv", "hpk_ShadFunc", "hpk_ci", "hpk_cv"), tup))

    if apply_brdf or a.shell_hapke > 0:
        if M:
            bb = bval.reshape(N, S, 3) * (1 + 2 * RGB_PAD) - RGB_PAD
            rgb = (wv * bb * irr).sum(-2)
        else:
            rgb = irr[:, -1, :] * bval.reshape(N, 3)
        res["rgb"] = rgb.clamp(0.0, 1.0)
    res["irradiance"] = irr
    if apply_brdf:
        Sb = S if M else 1
        if model.roughness:
            res["roughness"] = per_pt["roughness"]
            for k_ in ("glossy", "f", "g", "d", "l_dot_n", "v_dot_n", "n_h"):
                res[k_] = aux[k_].reshape(N, Sb, 1)
            res["brdf"] = aux["brdf"].reshape(N, Sb, 3)
            res["halfvec"] = aux["halfvec"].reshape(N, Sb, 3)
        elif model.RPV:
            for k_ in ("rpv_k", "rpv_theta", "rpv_rhoc"):
                if k_ in per_pt:
                    res[k_] = per_pt[k_]
        elif a.b or a.shell_hapke > 0:
            res["brdf"] = aux["brdf"].reshape(N, Sb, 3)
            res["hpk_P"] = aux["hpk_P"].reshape(N, Sb, 3)
            res["hpk_Hi"] = aux["hpk_Hi"].reshape(N, Sb, 3)
            res["hpk_Hv"] = aux["hpk_Hi"].reshape(N, Sb, 3)     # reference quirk: Hv key holds Hi (:387)
            res["hpk_ci"] = aux["hpk_ci"].reshape(N, Sb, 1)
            res["hpk_cv"] = aux["hpk_cv"].reshape(N, Sb, 1)
            res["hpk_ShadFunc"] = aux["hpk_ShadFunc"].reshape(N, Sb, 1)
            for k_ in ("hpk_b", "hpk_c", "hpk_theta"):
                if k_ in per_pt:
                    res[k_] = per_pt[k_]
    res["rays_d"] = view.reshape(N, 1, 3)
    res["sun_d"] = sun_d.reshape(N, 1, 3)
    return res, brdf_type


def render_rays(model: OracleModel, args, rays, draws: Draws, mode="test", valid_depth=None,
                target_depths=None, target_std=None, apply_brdf=False, bTestNormal=False,
                bTestSun_v=False, gsam_only=False, apply_theta=False, cos_irra_on=False, rays_t=None):
    """rendering.py:168-291 for variant 'spsbrdf-nerf', guided_samples > 0, n_importance == 0.
    Returns (dict with '_coarse' keys, brdf_type, extras) — extras carries pass-1 tensors for tests."""
    dt = model.dtype
    rays = rays.to(dt)
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6], rays[:, 7]
    sun_d = rays[:, 8:11] if args.data == "sat" else torch.ones_like(o)
    S1, G = int(args.n_samples), int(args.guided_samples)
    d_range = float(args.std_range)
    if G <= 0 or G == 2:
        raise NotImplementedError("guided_samples in {<=0, 2} are reference defect paths (SURVEY App. C.5)")
    t_vals, gauss = sampler_np.tables(S1, d_range)
    npf = lambda t: t.detach().to(torch.float32).numpy()
    # pass 1: stratified samples, sigma only
    z1 = torch.from_numpy(sampler_np.stratified_z(npf(near), npf(far), t_vals, npf(draws.u_strat))).to(dt)
    pp1 = _run_model(model, _points(o, d, z1), d, args.chunk, sigma_only=True)
    a1, T1, w1, depth1 = volume_weights(z1, pp1["sigma"].reshape(z1.shape), draws.noise1.to(dt), args.noise_std)
    extras = {"z1": z1, "sigma1": pp1["sigma"].reshape(z1.shape), "weights1": w1, "depth1": depth1}
    # optional sun-visibility march (rendering.py:244-259)
    sun_res = {}
    if (model.sun_v == "analystic" and apply_brdf) or bTestSun_v:
        surf = o + d * depth1.detach().unsqueeze(-1)
        far_sun = depth1.detach().clone()
        if abs(float(sun_d[0, 2])) > 1e-5:
            far_sun = torch.abs(d[0, 2] / sun_d[0, 2]) * far_sun
        n_sun = G if gsam_only else S1
        t_sun, _ = sampler_np.tables(n_sun, d_range)
        zs = torch.from_numpy(sampler_np.stratified_z(npf(far_sun * 0.01), npf(far_sun), t_sun, npf(draws.u_sun))).to(dt)
        pps = _run_model(model, surf.unsqueeze(1) + sun_d.unsqueeze(1) * zs.unsqueeze(2), sun_d, args.chunk, sigma_only=True)
        _, Ts, ws, _ = volume_weights(zs, pps["sigma"].reshape(zs.shape), draws.noise_sun.to(dt), args.noise_std)
        sun_res = {"sun": Ts.unsqueeze(-1).detach(), "weights_sc": ws.detach()}
    # guided samples + merge
    use_gt = mode == "train" and valid_depth is not None
    z2 = sampler_np.guided_z(
        npf(z1), npf(depth1), npf(w1), float(near[0]), float(far[0]), d_range, t_vals, gauss, npf(draws.u_pred),
        valid_depth=None if not use_gt else valid_depth.numpy(),
        gt_depth=None if not use_gt else npf(target_depths[:, 0]),
        gt_std=None if not use_gt else npf(target_std),
        u_gt=None if not use_gt else npf(draws.u_gt))
    if gsam_only:
        z = torch.from_numpy(np.sort(z2, -1)).to(dt)
        z_unsort, idx = z, None
    else:
        zs_, idx_, un_ = sampler_np.merge_sorted(npf(z1), z2)
        z, idx, z_unsort = torch.from_numpy(zs_).to(dt), torch.from_numpy(idx_), torch.from_numpy(un_).to(dt)
    extras["z2"] = torch.from_numpy(np.sort(z2, -1))
    nr_an = model.normal in ("analystic", "analystic_learned") or bTestNormal
    nr_lr = model.normal in ("learned", "analystic_learned")
    pp2 = _run_model(model, _points(o, d, z), d, args.chunk, rays_t=rays_t, apply_brdf=apply_brdf, apply_theta=apply_theta,
                     nr_an=nr_an, nr_lr=nr_lr)
    res, brdf_type = shade(model, args, pp2, z, d, sun_d, draws.noise2.to(dt), apply_brdf, apply_theta,
                           cos_irra_on, sun_res=sun_res, sort_idx=idx, z_unsort=z_unsort)
    return {f"{k}_coarse": v for k, v in res.items()}, brdf_type, extras
