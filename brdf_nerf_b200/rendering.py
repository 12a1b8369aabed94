"""render_rays — drop-in for the reference's `rendering.render_rays` (rendering.py:168-334) for
`args.model == "spsbrdf-nerf"`, executed by the sm_100a kernels of libbrdfnerf_b200.

Pipeline of one call (N rays, S1 = n_samples, G = guided_samples, S = S1+G or G with gsam_only):
  K-A1 stratified z            bn_sample_stratified
  K-B  sigma-only MLP pass     bn_mlp_forward(BN_MLP_SIGMA_ONLY)        (no autograd tape: the
  K-C  weights / depth         bn_composite_sigma                        reference detaches pass 1)
  [sun-visibility march: stratified z along the sun ray + sigma-only MLP + weights]
  K-A2 guided z + merge        bn_sample_guided, bn_merge_samples
  K-B  full MLP pass           bn_mlp_forward (+ bn_mlp_normals_forward for analytic normals)
  K-C  compositing + shading   [bn_brdf_points_forward] bn_composite_forward, bn_shade_rays_forward
and the mirrored backward chain when gradients are required.  Nothing in here synchronises with
the host or inspects device values, so a whole call is CUDA-graph capturable.

Differences from the reference that are visible to a caller are listed in DESIGN.md
("Semantics mirrored / not mirrored"); the result-dict keys and shapes follow SURVEY.md App. D.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops


@dataclasses.dataclass
class Draws:
    """Random draws of one call; pass as `_draws=` to make a call deterministic (parity tests).
    Shapes: u_strat (N,S1), noise1 (N,S1), u_pred (N,G), noise2 (N,S), u_gt (N,G) [row r is used only
    when valid_depth[r] > 0], u_sun / noise_sun (N,S')."""
    u_strat: torch.Tensor
    u_pred: torch.Tensor
    noise1: Optional[torch.Tensor] = None
    noise2: Optional[torch.Tensor] = None
    u_gt: Optional[torch.Tensor] = None
    u_sun: Optional[torch.Tensor] = None
    noise_sun: Optional[torch.Tensor] = None

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device=device, dtype=torch.float32).contiguous()
        return Draws(**{f.name: mv(getattr(self, f.name)) for f in dataclasses.fields(self)})


@dataclasses.dataclass
class _State:
    """Everything the backward chain needs."""
    n: int = 0
    s: int = 0
    pitch: int = 0
    flags: int = 0
    cfg: Optional[L.ShadeCfg] = None
    rays: Optional[torch.Tensor] = None
    z: Optional[torch.Tensor] = None
    packed: Optional[torch.Tensor] = None
    packed_rows: Optional[torch.Tensor] = None      # MLP row order when the trunk is shared between the passes
    idx: Optional[torch.Tensor] = None
    s1: int = 0
    lazy: bool = False                              # packed was never materialised: compositing works on packed_rows through idx
    noise: Optional[torch.Tensor] = None
    noise_std: float = 0.0
    irr: Optional[torch.Tensor] = None
    irr_last: Optional[torch.Tensor] = None
    alpha: Optional[torch.Tensor] = None
    trans: Optional[torch.Tensor] = None
    weights: Optional[torch.Tensor] = None
    wsum: Optional[torch.Tensor] = None
    acc: Optional[torch.Tensor] = None
    acc_irr: Optional[torch.Tensor] = None
    ws: Optional[torch.Tensor] = None
    multi: bool = False
    normal_an: bool = False


def _call_brdf_type(model, args, apply_brdf: bool) -> int:
    """Which BRDF `inference` evaluates for this call (spsbrdfnerf.py:286-346)."""
    if model.roughness == True and apply_brdf:                         # noqa: E712
        return L.BN_BRDF_MICROFACET
    if model.RPV and apply_brdf:
        return L.BN_BRDF_RPV
    if (apply_brdf and args.b == True) or args.shell_hapke > 0:        # noqa: E712
        return L.BN_BRDF_HAPKE
    return L.BN_BRDF_NONE


def _forward(model, args, rays: torch.Tensor, draws: Optional[Draws], mode: str, valid_depth, target_depths,
             target_std, apply_brdf: bool, bTestNormal: bool, bTestSun_v: bool, gsam_only: bool, apply_theta: bool,
             cos_irra_on: bool, train: bool, debug_nan: bool = False, own_ws: bool = False, sync: bool = True,
             reference_rng: bool = False, rays_t: Optional[torch.Tensor] = None, lazy_packed: bool = False,
             defer_kc: bool = False):
    """`lazy_packed` (Trainer, plain Lambertian stage): nobody reads the depth-ordered per-sample rows, so they are never
    materialised — compositing gathers the MLP's rows through sort_idx and its backward scatters the gradients the same way
    (no bn_permute_samples launch in either direction); `outs["packed"]` is None then.
    `defer_kc` (with lazy_packed, 128-sample rays, noise_std 0): compositing / shading are NOT run here; the caller finishes the
    step with `_lambertian_loss_backward` (one launch for compositing + colour + loss + their backward); `outs["deferred"]`.
    `own_ws`: the MLP workspace is allocated for this call and owned by the returned state (autograd bridge: several
    forwards may be alive before their backwards run); otherwise the model's cached per-tag buffer is reused.
    `sync=False`: the caller (Trainer) has refreshed the packed weight copies itself.
    `reference_rng`: draw from torch's generator exactly like the reference does: the sigma-noise
    `randn`s are drawn even when `noise_std == 0` (spsbrdfnerf.py:58) and the ground-truth guided draw has the
    data-dependent shape (n_valid, G) (rendering.py:144) — one host sync; not CUDA-graph capturable."""
    if args.model != "spsbrdf-nerf":
        raise NotImplementedError("only --model spsbrdf-nerf is implemented (BASELINE north star)")
    if args.n_importance > 0:
        raise NotImplementedError("n_importance > 0 (fine model) is not part of the spsbrdf-nerf recipe")
    S1, G = int(args.n_samples), int(args.guided_samples)
    if G <= 0 or G == 2:
        raise NotImplementedError("guided_samples <= 0 / == 2 are reference defect paths (SURVEY App. C.5)")
    if args.sc_lambda > 0:
        raise NotImplementedError("sc_lambda > 0 calls an undefined function in the reference (App. C.4)")
    dev = rays.device
    if not rays.is_cuda:
        raise L.BnError("render_rays needs CUDA rays: there is no CPU path")
    rays = rays.contiguous().float()
    N = rays.shape[0]
    d_range = float(args.std_range)
    noise_std = float(args.noise_std)
    S = G if gsam_only else S1 + G
    use_gt = mode == "train" and valid_depth is not None
    want_sun = (model.sun_v == "analystic" and apply_brdf) or bTestSun_v
    S_sun = G if gsam_only else S1
    if draws is None:
        rnd = lambda *sh: torch.rand(sh, device=dev, dtype=torch.float32)
        if reference_rng:
            def rndn(*sh):                       # drawn (and the generator advanced) even when the noise is switched off
                t = torch.randn(sh, device=dev, dtype=torch.float32)
                return t if noise_std != 0.0 else None

            def rnd_gt():                        # (n_valid, G) draws scattered to the valid rows, in row order
                valid = valid_depth.to(dev).reshape(-1) > 0
                u = torch.zeros((N, G), device=dev, dtype=torch.float32)
                u[valid] = torch.rand((int(valid.sum().item()), G), device=dev, dtype=torch.float32)
                return u
        else:
            # one generator launch for all the uniforms of the call (the order inside the block is the reference's draw order)
            widths = [S1] + ([S_sun] if want_sun else []) + [G] + ([G] if use_gt else [])
            block = torch.rand((N * sum(widths),), device=dev, dtype=torch.float32)
            chunks, off = [], 0
            for wd in widths:
                chunks.append(block[off:off + N * wd].view(N, wd))
                off += N * wd
            chunks = iter(chunks)
            rnd = lambda *sh: next(chunks)
            rndn = lambda *sh: torch.randn(sh, device=dev, dtype=torch.float32) if noise_std != 0.0 else None
            rnd_gt = lambda: rnd(N, G)
        # keyword arguments are evaluated left to right: the order below is the reference's draw order (App. B)
        draws = Draws(u_strat=rnd(N, S1), noise1=rndn(N, S1),
                      u_sun=rnd(N, S_sun) if want_sun else None, noise_sun=rndn(N, S_sun) if want_sun else None,
                      u_pred=rnd(N, G), u_gt=rnd_gt() if use_gt else None, noise2=rndn(N, S))
    else:
        draws = draws.to(dev)
    if sync:
        model.sync_weights()
    t_vals, gauss = ops.sampler_tables(S1, d_range, dev)
    t_g, gauss_g = ops.sampler_tables(G, d_range, dev)

    origins, dirs, sun = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    nr_an = model.normal in ("analystic_learned", "analystic") or bTestNormal
    nr_lr = model.normal in ("analystic_learned", "learned")
    flags = model.mlp_flags(apply_brdf=apply_brdf, apply_theta=apply_theta, nr_an_on=nr_an, nr_lr_on=nr_lr, train=train)
    # The stratified points are evaluated by BOTH passes of the reference (rendering.py:225 and :274, same weights,
    # same points).  Here their trunk runs once: pass 1 keeps the activations in the first N*S1 rows of the pass-2
    # workspace, pass 2 adds the guided points as rows N*S1.. and the heads run over all rows; sort_idx is applied
    # to the packed rows afterwards (bn_permute_samples).  Needs 128-row aligned blocks.
    share_trunk = ((not gsam_only) and (N * S1) % 128 == 0 and not getattr(model, "no_trunk_sharing", False)
                   and not os.environ.get("BN_NO_TRUNK_SHARING"))     # env: A/B timing aid
    # ---- pass 1: stratified samples, sigma only (no tape: the reference detaches it, rendering.py:262)
    z1 = ops.sample_stratified(rays[:, 6], rays[:, 7], 11, t_vals, draws.u_strat)
    sigma1 = torch.empty((N, S1), dtype=torch.float32, device=dev)
    ws = None
    if share_trunk:
        ws = model.workspace(N * S, flags, tag=None if own_ws else ("ws_train" if train else "ws_full"))
        ops.mlp_trunk_forward(model, origins, 11, dirs, 11, z1, flags, N * S, 0, sigma1, ws)
    if want_sun or not share_trunk:
        ws1 = model.workspace(N * max(S1, S_sun), L.MLP_SIGMA_ONLY, tag="ws_sigma")
    if not share_trunk:
        ops.mlp_forward(model, origins, 11, dirs, 11, z1, L.MLP_SIGMA_ONLY, sigma1, 1, ws1)
    # ---- ground-truth depth override of the guided samples (GenerateGuidedSamples, rendering.py:132-147)
    gt_depth = gt_std = vd = None
    gt_stride = 1
    if use_gt:
        vd = valid_depth.to(device=dev, dtype=torch.int64).contiguous()
        td = target_depths.to(device=dev, dtype=torch.float32).contiguous()
        gt_depth, gt_stride = td, td.shape[1] if td.dim() == 2 else 1
        gt_std = target_std.to(device=dev, dtype=torch.float32).contiguous()
    guided_kw = dict(valid_depth=vd, gt_depth=gt_depth, gt_depth_stride=gt_stride, gt_std=gt_std,
                     u_gt=draws.u_gt if use_gt else None, want_std=debug_nan)
    # compositing of the stratified densities, guided samples around the predicted (or ground-truth) depth, merge: one launch
    # (bn_coarse_to_fine) unless only the guided samples are kept (gsam_only: a plain sort instead of the merge)
    std1 = None
    if not gsam_only:
        w1, depth1, std1, z2, z, idx, z_unsort = ops.coarse_to_fine(z1, sigma1, draws.noise1, noise_std, t_g, gauss_g, draws.u_pred,
                                                                    rays[0:1, 6], rays[0:1, 7], d_range, **guided_kw)
    else:
        _, _, w1, depth1, _ = ops.composite_sigma(z1, sigma1, draws.noise1, noise_std)
        z2 = ops.sample_guided(z1, depth1, w1, t_g, gauss_g, draws.u_pred, rays[0:1, 6], rays[0:1, 7], d_range, **guided_kw)
        if debug_nan:
            z2, std1 = z2
        z, idx, z_unsort = z2, None, z2
    nan_counts = None
    if debug_nan:       # check_nan(pred_depth / pred_weight / sampling_std) of rendering.py:121-123, counted on the device
        nan_counts = (ops.count_nan([depth1, w1, std1]), (depth1.numel(), w1.numel(), std1.numel()))

    # ---- optional sun-visibility march from the predicted surface (rendering.py:244-259)
    sun_res = {}
    if want_sun:
        surf = (origins + dirs * depth1.unsqueeze(-1)).contiguous()
        ratio = torch.where(sun[0, 2].abs() > 1e-5, (dirs[0, 2] / sun[0, 2]).abs(), torch.ones((), device=dev))
        far_sun = (ratio * depth1).contiguous()
        near_sun = (far_sun * 0.01).contiguous()
        t_s, _ = ops.sampler_tables(S_sun, d_range, dev)
        z_sun = ops.sample_stratified(near_sun, far_sun, 1, t_s, draws.u_sun)
        sig_sun = torch.empty((N, S_sun), dtype=torch.float32, device=dev)
        sun_c = sun.contiguous()
        ops.mlp_forward(model, surf, 3, sun_c, 3, z_sun, L.MLP_SIGMA_ONLY, sig_sun, 1, ws1)
        _, T_sun, w_sun, _, _ = ops.composite_sigma(z_sun, sig_sun, draws.noise_sun, noise_std, want_all=True)
        sun_res = {"sun": T_sun.unsqueeze(-1), "weights_sc": w_sun}

    # ---- pass 2: full model
    C = model.out_channels(flags)
    brdf_type = _call_brdf_type(model, args, apply_brdf)
    if C == 4:
        brdf_type = L.BN_BRDF_NONE          # `idx == 4` early return (spsbrdfnerf.py:281-282)
    has_normal = nr_an or nr_lr
    if brdf_type != L.BN_BRDF_NONE and not has_normal:
        raise RuntimeError("a BRDF needs a normal (normal != 'none'); the reference fails here as well")
    if apply_brdf and brdf_type == L.BN_BRDF_NONE and has_normal:
        raise RuntimeError("apply_brdf=True without any BRDF head: unbound `brdf` in the reference (App. C.6)")
    multi = bool(model.MultiBRDF) and brdf_type != L.BN_BRDF_NONE
    pitch = C + (3 if multi else 0)
    packed_rows = None
    has_beta = bool(flags & L.MLP_BETA)
    if has_beta:
        if rays_t is None:
            raise ValueError("model.beta is True: render_rays needs `ts` and models['t'] (rendering.py:228-229)")
        rays_t = rays_t.detach().to(device=dev, dtype=torch.float32).contiguous()
    if share_trunk:
        packed_rows = torch.empty((N * S, pitch), dtype=torch.float32, device=dev)
        ops.mlp_trunk_forward(model, origins, 11, dirs, 11, z2, flags, N * S, N * S1, None, ws)
        if has_beta:         # the rays' time embedding next to the features of their stratified and of their guided rows
            ops.mlp_write_t(model, rays_t, N, S1, flags, N * S, 0, ws)
            ops.mlp_write_t(model, rays_t, N, G, flags, N * S, N * S1, ws)
        ops.mlp_heads_forward(model, N * S, flags, packed_rows, pitch, ws)
        if nr_an:
            ops.mlp_normals_forward(model, packed_rows, pitch, N, S, flags, ws)
        lazy = lazy_packed and not has_normal and not multi and not has_beta and brdf_type == L.BN_BRDF_NONE
        packed = None if lazy else ops.permute_samples(packed_rows, idx, N, S1, G, pitch, scatter=False)
    else:
        lazy = False
        packed = torch.empty((N, S, pitch), dtype=torch.float32, device=dev)
        ws = model.workspace(N * S, flags, tag=None if own_ws else ("ws_train" if train else "ws_full"))
        if has_beta:
            ops.mlp_write_t(model, rays_t, N, S, flags, N * S, 0, ws)
        ops.mlp_forward(model, origins, 11, dirs, 11, z, flags, packed, pitch, ws)
        if nr_an:
            ops.mlp_normals_forward(model, packed, pitch, N, S, flags, ws)

    cfg = L.ShadeCfg()
    cfg.n_channels = pitch
    ch = 4 + int(has_beta)                     # the uncertainty channel sits between sigma and the normals (:156-158)
    cfg.normal_ch = -1
    if nr_an:
        cfg.normal_ch = ch; ch += 3
    if nr_lr:
        cfg.normal_ch = ch; ch += 3            # learned wins when both exist (spsbrdfnerf.py:236-239)
    cfg.param_ch = ch if C > ch else -1
    cfg.brdf_ch = C if multi else -1
    cfg.brdf_type = brdf_type
    cfg.funcM, cfg.funcF = int(args.funcM == True), int(args.funcF == True)    # noqa: E712
    cfg.funcH = int(args.funcH) if args.funcH in (1, 2, True) else 0
    cfg.hapke_b = int(bool(apply_brdf and args.b == True))                      # noqa: E712
    cfg.hapke_c = int(bool(apply_brdf and args.c == True))                      # noqa: E712
    cfg.hapke_theta = int(bool(apply_brdf and apply_theta and args.theta == True))   # noqa: E712
    cfg.shell_hapke = int(args.shell_hapke)
    cfg.multi_brdf = int(multi)
    cfg.hpk_scl, cfg.fresnel_f0 = float(args.hpk_scl), float(args.fresnel_f0)
    apply_sun_v = model.sun_v == "analystic" and "sun" in sun_res
    irr = irr_last = None
    if cos_irra_on and has_normal:
        cfg.irr_mode = L.BN_IRR_COS
    elif apply_sun_v:
        if S_sun != S:
            raise RuntimeError("sun_v='analystic' without cos_irra_on needs gsam_only (shape mismatch in the "
                               "reference, SURVEY App. C.3)")
        cfg.irr_mode = L.BN_IRR_SUNVIS
        irr = sun_res["sun"].reshape(N, S).contiguous()
        irr_last = irr[:, -1].contiguous()
    else:
        cfg.irr_mode = L.BN_IRR_ONES

    if lazy and defer_kc and S == 128 and pitch == 4 and noise_std == 0.0 and cfg.irr_mode == L.BN_IRR_ONES:
        st = _State(n=N, s=S, pitch=pitch, flags=flags, cfg=cfg, rays=rays, z=z, packed=None, packed_rows=packed_rows, idx=idx,
                    s1=S1, lazy=True, noise=None, noise_std=0.0, ws=ws, multi=False, normal_an=False)
        outs = dict(deferred=True, packed=None, z=z, z_unsort=z_unsort, idx=idx, sun_res=sun_res, C=C, nr_an=False, nr_lr=False,
                    has_beta=False, brdf_type=brdf_type,
                    extras=dict(z1=z1, z2=z2, sigma1=sigma1, weights1=w1, depth1=depth1, nan_counts=nan_counts))
        return outs, st
    aux_pts = None
    if multi:
        aux_pts = ops.brdf_points_forward(cfg, rays, packed, want_aux=not train)
    if lazy:
        alpha, trans, w, depth, wsum, acc, acc_irr = ops.composite_forward(z, packed_rows, draws.noise2, noise_std, irr,
                                                                           sort_idx=idx, n_stratified=S1)
    else:
        alpha, trans, w, depth, wsum, acc, acc_irr = ops.composite_forward(z, packed, draws.noise2, noise_std, irr)
    sh = ops.shade_rays_forward(cfg, rays, acc, wsum, acc_irr, irr_last, want_normal=has_normal,
                                want_brdf=brdf_type != L.BN_BRDF_NONE and not multi)

    st = _State(n=N, s=S, pitch=pitch, flags=flags, cfg=cfg, rays=rays, z=z, packed=packed, packed_rows=packed_rows,
                idx=idx, s1=S1, lazy=lazy,
                noise=draws.noise2 if noise_std != 0.0 else None, noise_std=noise_std, irr=irr, irr_last=irr_last,
                alpha=alpha, trans=trans, weights=w, wsum=wsum, acc=acc, acc_irr=acc_irr, ws=ws, multi=multi,
                normal_an=nr_an)
    outs = dict(rgb=sh["rgb"], depth=depth, weights=w, packed=packed, alpha=alpha, trans=trans, z=z,
                z_unsort=z_unsort, idx=idx, shade=sh, sun_res=sun_res, aux_pts=aux_pts, C=C, nr_an=nr_an, nr_lr=nr_lr,
                has_beta=has_beta,
                brdf_type=brdf_type, extras=dict(z1=z1, z2=z2, sigma1=sigma1, weights1=w1, depth1=depth1, nan_counts=nan_counts))
    return outs, st


def _lambertian_loss_backward(model, st: _State, args, target_rgb, valid_depth, target_depths, target_std, use_depth: bool,
                              g_params: torch.Tensor, want_outputs: bool = False):
    """Second half of a `defer_kc` forward: compositing + Lambertian colour + colour / depth loss + their backward in one launch
    (bn_lambertian_render_loss), then the MLP backward.  Returns the loss (1,) [, rgb (N,3), depth (N)]."""
    use_ds = use_depth and valid_depth is not None
    loss, g_rows, rgb, depth = ops.lambertian_render_loss(
        st.z, st.packed_rows, st.idx, st.s1, target_rgb, float(args.lambda_rgb), valid_depth=valid_depth if use_ds else None,
        target_depths=target_depths, target_std=target_std, lambda_ds=float(args.ds_lambda) if use_ds else 0.0,
        use_all_depth=bool(getattr(args, "usealldepth", False)), no_weights=bool(getattr(args, "ds_noweights", False)),
        want_outputs=want_outputs)
    ops.mlp_backward(model, st.packed_rows, g_rows, st.pitch, st.n, st.s, st.flags, g_params, st.ws)
    return (loss, rgb, depth) if want_outputs else loss


def _backward(model, st: _State, g_rgb, g_depth, g_weights, g_packed, g_params: torch.Tensor):
    """Backward chain; accumulates parameter gradients into the flat fp32 buffer `g_params`."""
    dev = st.rays.device
    if g_rgb is None:
        g_rgb = torch.zeros((st.n, 3), dtype=torch.float32, device=dev)
    g_acc, g_wsum, g_acc_irr = ops.shade_rays_backward(st.cfg, st.rays, st.acc, st.wsum, st.acc_irr, st.irr_last,
                                                       g_rgb.contiguous())
    if st.lazy:           # gradient rows come out in the MLP's row order directly
        if g_packed is not None:
            raise RuntimeError("lazy_packed render: there are no depth-ordered per-sample rows to take a gradient of")
        gp = ops.composite_backward(st.z, st.packed_rows, st.noise, st.noise_std, st.irr, st.alpha, st.trans, st.weights,
                                    g_acc, g_acc_irr, None if g_depth is None else g_depth.contiguous(), g_wsum,
                                    None if g_weights is None else g_weights.contiguous(), None,
                                    sort_idx=st.idx, n_stratified=st.s1)
        ops.mlp_backward(model, st.packed_rows, gp, st.pitch, st.n, st.s, st.flags, g_params, st.ws)
        return
    gp = ops.composite_backward(st.z, st.packed, st.noise, st.noise_std, st.irr, st.alpha, st.trans, st.weights,
                                g_acc, g_acc_irr, None if g_depth is None else g_depth.contiguous(), g_wsum,
                                None if g_weights is None else g_weights.contiguous(),
                                None if g_packed is None else g_packed.contiguous())
    if st.multi:
        ops.brdf_points_backward(st.cfg, st.rays, st.packed, gp)
    out = st.packed
    if st.packed_rows is not None:            # gradients back into the MLP's row order (inverse of sort_idx)
        gp = ops.permute_samples(gp, st.idx, st.n, st.s1, st.s - st.s1, st.pitch, scatter=True)
        out = st.packed_rows
    if st.normal_an:
        ops.mlp_normals_backward(model, out, gp, st.pitch, st.n, st.s, st.flags, g_params, st.ws)
    ops.mlp_backward(model, out, gp, st.pitch, st.n, st.s, st.flags, g_params, st.ws)


class _RenderFunction(torch.autograd.Function):
    """Autograd bridge: parameters in, (rgb, depth, weights, packed) out."""

    @staticmethod
    def forward(ctx, model, args, rays, draws, kw, holder, *params):
        # own_ws: this call's activations stay alive (in ctx.st) until ITS backward has run, however many other forwards
        # (chunks, other models, other streams) happen in between; `holder` hands the non-tensor results back to the caller
        outs, st = _forward(model, args, rays, draws, train=True, own_ws=True, **kw)
        ctx.model, ctx.st = model, st
        holder["outs"] = outs
        ctx.mark_non_differentiable(outs["alpha"], outs["trans"], outs["z"])
        return outs["rgb"], outs["depth"], outs["weights"], outs["packed"], outs["alpha"], outs["trans"], outs["z"]

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_weights, g_packed, *_):
        model, st = ctx.model, ctx.st
        flat = torch.zeros_like(model.flat_params)
        _backward(model, st, g_rgb, g_depth, g_weights, g_packed, flat)
        grads = model.grad_views(flat)
        return (None, None, None, None, None, None, *grads)


def _assemble(model, args, rays, outs, rgb, depth, weights, packed, apply_brdf, apply_theta) -> Dict[str, torch.Tensor]:
    """Result dict of `inference` (spsbrdfnerf.py:202-226,230-255,357-416); keys per SURVEY App. D."""
    N, S = outs["z"].shape
    C = outs["C"]
    sh = outs["shade"]
    res = {"sigmas": packed[..., 3:4], "albedo": packed[..., 0:3], "albedo_accu": sh["albedo_accu"], "depth": depth,
           "alphas": outs["alpha"], "weights": weights, "transparency": outs["trans"], "z_vals": outs["z"]}
    if "sun" in outs["sun_res"]:
        res["sun"] = outs["sun_res"]["sun"]
        res["weights_sc"] = outs["sun_res"]["weights_sc"]
    if outs["idx"] is not None:
        res["sort_idx"] = outs["idx"]
    res["z_vals_unsort"] = outs["z_unsort"]
    ch = 4
    if outs.get("has_beta"):
        res["beta"] = packed[..., ch:ch + 1]; ch += 1          # spsbrdfnerf.py:225-226
    if outs["nr_an"]:
        res["normal_an"] = packed[..., ch:ch + 3]; ch += 3
    if outs["nr_lr"]:
        res["normal_lr"] = packed[..., ch:ch + 3]; ch += 3
    has_normal = outs["nr_an"] or outs["nr_lr"]
    if has_normal:
        res["nr_vw"] = sh["nr_vw"].reshape(N, 1, 1)
        res["nr_sun"] = sh["nr_sun"].reshape(N, 1, 1)
        res["hpk_scl"] = sh["hpk_scl"].reshape(N, 1, 1)
    res["rgb"] = rgb
    if C == 4:
        return res
    sun_d = rays[:, 8:11]
    if has_normal and outs["shade"] is not None:
        pass
    # irradiance tensor of the reference (N,S,3): a broadcast view here, never materialised by a kernel
    cfg_irr = outs["irr_mode"]
    if cfg_irr == L.BN_IRR_COS:
        res["irradiance"] = sun_d[:, 2].abs().reshape(N, 1, 1).expand(N, S, 3)
    elif cfg_irr == L.BN_IRR_SUNVIS:
        res["irradiance"] = outs["sun_res"]["sun"].expand(N, S, 3)
    else:
        res["irradiance"] = torch.ones((1, 1, 1), device=rays.device).expand(N, S, 3)
    bt = outs["brdf_type"]
    multi = bool(model.MultiBRDF) and bt != L.BN_BRDF_NONE
    Sb = S if multi else 1
    if apply_brdf:
        if multi:
            aux = outs["aux_pts"]                       # (N,S,3,8) or None when training
            brdf = packed[..., C:C + 3]
        else:
            aux = sh["aux"].reshape(N, 1, 3, 8) if sh["aux"] is not None else None
            brdf = sh["brdf"].reshape(N, 1, 3) if sh["brdf"] is not None else None
        if model.roughness == True:                                    # noqa: E712
            res["roughness"] = packed[..., ch:ch + 1]
            if aux is not None:
                for i, k in enumerate(("glossy", "f", "g", "d", "l_dot_n", "v_dot_n", "n_h")):
                    res[k] = aux[:, :, 0, i].reshape(N, Sb, 1)
                # half vector: normalised l + v (per ray constant)
                l = torch.nn.functional.normalize(sun_d, dim=-1, eps=1e-6)
                v = torch.nn.functional.normalize(-rays[:, 3:6], dim=-1, eps=1e-6)
                res["halfvec"] = torch.nn.functional.normalize(l + v, dim=-1, eps=1e-6).reshape(N, 1, 3).expand(N, Sb, 3)
            if brdf is not None:
                res["brdf"] = brdf
        elif model.RPV:
            for flag, key in ((args.funcM, "rpv_k"), (args.funcF, "rpv_theta"), (args.funcH, "rpv_rhoc")):
                if flag == True:                                       # noqa: E712
                    res[key] = packed[..., ch:ch + 3]; ch += 3
        elif args.b == True or args.shell_hapke > 0:                   # noqa: E712
            if brdf is not None:
                res["brdf"] = brdf
            if aux is not None:
                res["hpk_P"] = aux[..., 0].reshape(N, Sb, 3)
                res["hpk_Hi"] = aux[..., 1].reshape(N, Sb, 3)
                res["hpk_Hv"] = aux[..., 1].reshape(N, Sb, 3)        # reference stores Hi under this key (:387)
                res["hpk_ci"] = aux[:, :, 0, 3].reshape(N, Sb, 1)
                res["hpk_cv"] = aux[:, :, 0, 4].reshape(N, Sb, 1)
                res["hpk_ShadFunc"] = aux[:, :, 0, 5].reshape(N, Sb, 1)
            if args.b == True:                                         # noqa: E712
                res["hpk_b"] = packed[..., ch:ch + 3]; ch += 3
            if args.c == True:                                         # noqa: E712
                res["hpk_c"] = packed[..., ch:ch + 3]; ch += 3
            if apply_theta and args.theta == True:                     # noqa: E712
                res["hpk_theta"] = packed[..., ch:ch + 1]; ch += 1
    res["rays_d"] = (-rays[:, 3:6]).reshape(N, 1, 3)
    res["sun_d"] = sun_d.reshape(N, 1, 3)
    return res


def render_rays(models, args, rays, ts, mode="test", valid_depth=None, target_depths=None, target_std=None,
                apply_brdf=False, print_debuginfo=False, bTestNormal=False, bTestSun_v=False, gsam_only=False,
                rows=None, cols=None, percent=0, apply_theta=False, cos_irra_on=False, _draws: Optional[Draws] = None,
                _return_extras=False, _reference_rng: bool = False):
    """Same signature / return as the reference (rendering.py:168,334): (dict with `_coarse` keys, brdf_type).
    `_draws` (extension) injects the random draws for deterministic parity runs; `_reference_rng=True` (extension) consumes
    torch's CUDA generator in exactly the reference's order and shapes (see `_forward`), so that a seeded reference run on
    the same device sees the same draws."""
    model = models["coarse"]
    rays_t = None
    if getattr(args, "beta", False) == True and ts is not None:          # noqa: E712  rendering.py:228-229
        with torch.no_grad():         # the embedding is an input here: no gradient flows back into models['t']
            rays_t = models["t"](ts.to(rays.device).reshape(-1).long())
    kw = dict(mode=mode, valid_depth=valid_depth, target_depths=target_depths, target_std=target_std, rays_t=rays_t,
              apply_brdf=bool(apply_brdf), bTestNormal=bool(bTestNormal), bTestSun_v=bool(bTestSun_v),
              gsam_only=bool(gsam_only), apply_theta=bool(apply_theta), cos_irra_on=bool(cos_irra_on),
              reference_rng=bool(_reference_rng))
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    if need_grad:
        holder = {}
        rgb, depth, weights, packed, alpha, trans, z = _RenderFunction.apply(model, args, rays, _draws, kw, holder,
                                                                             *model.parameters())
        outs = holder["outs"]
    else:
        outs, _ = _forward(model, args, rays, _draws, train=False, debug_nan=bool(print_debuginfo), **kw)
        rgb, depth, weights, packed = outs["rgb"], outs["depth"], outs["weights"], outs["packed"]
        _report_nans(outs["extras"].get("nan_counts"))
    outs["irr_mode"] = _irr_mode_of(outs, model, cos_irra_on)
    res = _assemble(model, args, rays.float(), outs, rgb, depth, weights, packed, bool(apply_brdf), bool(apply_theta))
    if rows is not None and cols is not None:
        N = rays.shape[0]
        r_, c_ = rows.reshape(N).float(), cols.reshape(N).float()
        res["ref_sphere"] = torch.stack([c_, -r_, torch.sqrt(torch.abs(1 - r_ * r_ - c_ * c_))], -1).reshape(N, 1, 3)
    names = {L.BN_BRDF_NONE: "Lambertian", L.BN_BRDF_MICROFACET: "Microfacet", L.BN_BRDF_RPV: "RPV", L.BN_BRDF_HAPKE: "Hapke"}
    result = {f"{k}_coarse": v for k, v in res.items()}
    if _return_extras:
        return result, names[outs["brdf_type"]], outs["extras"]
    return result, names[outs["brdf_type"]]


def _report_nans(nan_counts):
    """`print_debuginfo=True`: the lines train_utils.check_nan prints for pred_depth / pred_weight / sampling_std
    (rendering.py:121-123) — one read-back of three device counters instead of three isnan().sum() round trips."""
    if nan_counts is None:
        return
    counts, totals = nan_counts
    for name, c, t in zip(("pred_depth", "pred_weight", "sampling_std"), counts.cpu().tolist(), totals):
        print("----nan nb in {}, val_in: {:.0f} / {:.0f}".format(name, c, t))


def _irr_mode_of(outs, model, cos_irra_on):
    if cos_irra_on and (outs["nr_an"] or outs["nr_lr"]):
        return L.BN_IRR_COS
    if model.sun_v == "analystic" and "sun" in outs["sun_res"]:
        return L.BN_IRR_SUNVIS
    return L.BN_IRR_ONES
