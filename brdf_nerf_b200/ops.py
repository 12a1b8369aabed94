"""Thin Python wrappers over the C ABI: they allocate outputs with torch and enqueue the kernels on
torch's current CUDA stream.  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L

_TABLES: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def sampler_tables(n: int, d_range: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """t_vals = linspace(0,1,n) and the (n-1)-bin Gaussian window, built with torch on the CPU exactly
    as the reference does (rendering.py:63,68-69,151) and cached on the device; the kernels take them
    as inputs because torch.linspace / torch.exp have their own rounding."""
    key = (n, float(d_range), str(device))
    if key not in _TABLES:
        t = torch.linspace(0.0, 1.0, steps=n)
        x = torch.linspace(-d_range, d_range, steps=n - 1)
        g = 1.0 / math.sqrt(2 * math.pi) * torch.exp(-0.5 * x.pow(2))
        _TABLES[key] = (t.to(device), g.to(device))
    return _TABLES[key]


def sample_stratified(near: torch.Tensor, far: torch.Tensor, stride: int, t_vals, u) -> torch.Tensor:
    """near/far: 1-element-offset views into a row-major buffer read with `stride` floats per ray."""
    n, s = u.shape
    z = torch.empty_like(u)
    L.check(L.load().bn_sample_stratified(C.c_void_p(near.data_ptr()), C.c_void_p(far.data_ptr()), stride,
                                          L.ptr(t_vals), L.ptr(u), L.ptr(z), n, s, L.stream_ptr()))
    return z


def sample_guided(z1, depth, weights, t_vals, gauss_w, u_pred, near0, far0, d_range,
                  valid_depth=None, gt_depth=None, gt_depth_stride=1, gt_std=None, u_gt=None, want_std=False):
    n, s1 = z1.shape
    g = u_pred.shape[1]
    z2 = torch.empty((n, g), dtype=torch.float32, device=z1.device)
    std = torch.empty(n, dtype=torch.float32, device=z1.device) if want_std else None
    L.check(L.load().bn_sample_guided(
        L.ptr(z1), L.ptr(depth), L.ptr(weights), L.ptr(t_vals), L.ptr(gauss_w), L.ptr(u_pred),
        C.c_void_p(near0.data_ptr()), C.c_void_p(far0.data_ptr()), float(d_range),
        L.ptr(valid_depth, torch.int64), None if gt_depth is None else C.c_void_p(gt_depth.data_ptr()),
        gt_depth_stride, L.ptr(gt_std), L.ptr(u_gt), L.ptr(z2), L.ptr(std), n, s1, g, L.stream_ptr()))
    return (z2, std) if want_std else z2


def count_nan(tensors, counters=None):
    """Accumulates the NaN count of each tensor into one slot of an int32 device buffer (allocated zeroed when None) without
    synchronising: the device-side stand-in for train_utils.check_nan's torch.isnan(x).sum() (train_utils.py:61-78)."""
    tensors = list(tensors)
    if counters is None:
        counters = torch.zeros(len(tensors), dtype=torch.int32, device=tensors[0].device)
    for i, t in enumerate(tensors):
        t = t.contiguous()
        if t.numel():
            L.check(L.load().bn_count_nan(L.ptr(t), t.numel(), C.c_void_p(counters.data_ptr() + 4 * i), L.stream_ptr()))
    return counters


def merge_samples(z1, z2):
    n, s1 = z1.shape
    g = z2.shape[1]
    z = torch.empty((n, s1 + g), dtype=torch.float32, device=z1.device)
    idx = torch.empty((n, s1 + g), dtype=torch.int64, device=z1.device)
    unsort = torch.empty_like(z)
    L.check(L.load().bn_merge_samples(L.ptr(z1), L.ptr(z2), L.ptr(z), L.ptr(idx, torch.int64), L.ptr(unsort),
                                      n, s1, g, L.stream_ptr()))
    return z, idx, unsort


def sort_rows(x):
    out = torch.empty_like(x)
    L.check(L.load().bn_sort_rows(L.ptr(x), L.ptr(out), x.shape[0], x.shape[1], L.stream_ptr()))
    return out


def coarse_to_fine(z1, sigma1, noise1, noise_std, t_g, gauss_g, u_pred, near0, far0, d_range, valid_depth=None, gt_depth=None,
                   gt_depth_stride=1, gt_std=None, u_gt=None, want_std=False):
    """composite_sigma + sample_guided + merge_samples in one launch (bn_coarse_to_fine); same results, bit for bit."""
    n, s1 = z1.shape
    g = u_pred.shape[1]
    dev = z1.device
    w1 = torch.empty_like(z1)
    depth1 = torch.empty(n, dtype=torch.float32, device=dev)
    std1 = torch.empty(n, dtype=torch.float32, device=dev) if want_std else None
    z2 = torch.empty((n, g), dtype=torch.float32, device=dev)
    z = torch.empty((n, s1 + g), dtype=torch.float32, device=dev)
    idx = torch.empty((n, s1 + g), dtype=torch.int64, device=dev)
    unsort = torch.empty((n, s1 + g), dtype=torch.float32, device=dev)
    L.check(L.load().bn_coarse_to_fine(
        L.ptr(z1), L.ptr(sigma1), L.ptr(noise1), float(noise_std), L.ptr(t_g), L.ptr(gauss_g), L.ptr(u_pred),
        C.c_void_p(near0.data_ptr()), C.c_void_p(far0.data_ptr()), float(d_range), L.ptr(valid_depth, torch.int64),
        None if gt_depth is None else C.c_void_p(gt_depth.data_ptr()), int(gt_depth_stride), L.ptr(gt_std), L.ptr(u_gt),
        L.ptr(w1), L.ptr(depth1), L.ptr(std1), L.ptr(z2), L.ptr(z), L.ptr(idx, torch.int64), L.ptr(unsort), n, s1, g,
        L.stream_ptr()))
    return w1, depth1, std1, z2, z, idx, unsort


def composite_sigma(z, sigma, noise, noise_std, want_all=False, want_std=False):
    n, s = z.shape
    dev = z.device
    w = torch.empty_like(z)
    depth = torch.empty(n, dtype=torch.float32, device=dev)
    alpha = torch.empty_like(z) if want_all else None
    trans = torch.empty_like(z) if want_all else None
    std = torch.empty(n, dtype=torch.float32, device=dev) if want_std else None
    L.check(L.load().bn_composite_sigma(L.ptr(z), L.ptr(sigma), L.ptr(noise), float(noise_std), L.ptr(alpha),
                                        L.ptr(trans), L.ptr(w), L.ptr(depth), L.ptr(std), n, s, L.stream_ptr()))
    return alpha, trans, w, depth, std


def composite_forward(z, packed, noise, noise_std, irr=None, sort_idx=None, n_stratified=0):
    """`sort_idx` (N,S) int64: `packed` is the MLP's [N*S, C] rows in generation order and is gathered through the index."""
    n, s = z.shape
    c = packed.shape[-1]
    dev = z.device
    alpha, trans, w = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
    depth = torch.empty(n, dtype=torch.float32, device=dev)
    wsum = torch.empty(n, dtype=torch.float32, device=dev)
    acc = torch.empty((n, c), dtype=torch.float32, device=dev)
    acc_irr = torch.empty((n, 4), dtype=torch.float32, device=dev) if irr is not None else None
    L.check(L.load().bn_composite_forward(L.ptr(z), L.ptr(packed), c, 3, L.ptr(noise), float(noise_std), L.ptr(irr),
                                          L.ptr(alpha), L.ptr(trans), L.ptr(w), L.ptr(depth), L.ptr(wsum), L.ptr(acc),
                                          L.ptr(acc_irr), n, s, L.ptr(sort_idx, torch.int64), int(n_stratified),
                                          L.stream_ptr()))
    return alpha, trans, w, depth, wsum, acc, acc_irr


def composite_backward(z, packed, noise, noise_std, irr, alpha, trans, w, g_acc, g_acc_irr, g_depth, g_wsum,
                       g_weights, g_packed_direct, sort_idx=None, n_stratified=0):
    """With `sort_idx`, `packed` and the returned gradient are [N*S, C] rows in the MLP's generation order."""
    n, s = z.shape
    c = packed.shape[-1]
    g_packed = torch.empty_like(packed)
    L.check(L.load().bn_composite_backward(
        L.ptr(z), L.ptr(packed), c, 3, L.ptr(noise), float(noise_std), L.ptr(irr), L.ptr(alpha), L.ptr(trans), L.ptr(w),
        L.ptr(g_acc), L.ptr(g_acc_irr), L.ptr(g_depth), L.ptr(g_wsum), L.ptr(g_weights), L.ptr(g_packed_direct),
        L.ptr(g_packed), n, s, L.ptr(sort_idx, torch.int64), int(n_stratified), L.stream_ptr()))
    return g_packed


def shade_rays_forward(cfg: L.ShadeCfg, rays, acc, wsum, acc_irr, irr_last, want_normal, want_brdf):
    n = rays.shape[0]
    dev = rays.device
    f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
    rgb, albedo_accu = f(n, 3), f(n, 3)
    normal_s = f(n, 3) if want_normal else None
    nr_vw = f(n) if want_normal else None
    nr_sun = f(n) if want_normal else None
    hpk = f(n) if want_normal else None
    brdf = f(n, 3) if want_brdf else None
    aux = torch.zeros((n, 3, 8), dtype=torch.float32, device=dev) if want_brdf else None
    L.check(L.load().bn_shade_rays_forward(C.byref(cfg), L.ptr(rays), L.ptr(acc), L.ptr(wsum), L.ptr(acc_irr),
                                           L.ptr(irr_last), L.ptr(rgb), L.ptr(albedo_accu), L.ptr(normal_s), L.ptr(nr_vw),
                                           L.ptr(nr_sun), L.ptr(hpk), L.ptr(brdf), L.ptr(aux), n, L.stream_ptr()))
    return dict(rgb=rgb, albedo_accu=albedo_accu, normal_s=normal_s, nr_vw=nr_vw, nr_sun=nr_sun, hpk_scl=hpk,
                brdf=brdf, aux=aux)


def shade_rays_backward(cfg: L.ShadeCfg, rays, acc, wsum, acc_irr, irr_last, g_rgb):
    n, c = acc.shape
    dev = rays.device
    g_acc = torch.empty((n, c), dtype=torch.float32, device=dev)
    g_wsum = torch.empty(n, dtype=torch.float32, device=dev)
    g_acc_irr = torch.zeros((n, 4), dtype=torch.float32, device=dev) if acc_irr is not None else None
    L.check(L.load().bn_shade_rays_backward(C.byref(cfg), L.ptr(rays), L.ptr(acc), L.ptr(wsum), L.ptr(acc_irr),
                                            L.ptr(irr_last), L.ptr(g_rgb), L.ptr(g_acc), L.ptr(g_wsum), L.ptr(g_acc_irr),
                                            n, L.stream_ptr()))
    return g_acc, g_wsum, g_acc_irr


def brdf_points_forward(cfg: L.ShadeCfg, rays, packed, want_aux=False):
    n, s, _ = packed.shape
    aux = torch.zeros((n, s, 3, 8), dtype=torch.float32, device=rays.device) if want_aux else None
    L.check(L.load().bn_brdf_points_forward(C.byref(cfg), L.ptr(rays), L.ptr(packed), L.ptr(aux), n, s, L.stream_ptr()))
    return aux


def brdf_points_backward(cfg: L.ShadeCfg, rays, packed, g_packed):
    n, s, _ = packed.shape
    L.check(L.load().bn_brdf_points_backward(C.byref(cfg), L.ptr(rays), L.ptr(packed), L.ptr(g_packed), n, s, L.stream_ptr()))


def mlp_forward(model, origins, o_stride, dirs, d_stride, z, flags, out, pitch, ws):
    """origins / dirs: tensors whose data_ptr is the first coordinate of ray 0."""
    n, s = z.shape
    L.check(L.load().bn_mlp_forward(model.handle(), L.ptr(model.flat_params), C.c_void_p(origins.data_ptr()), o_stride,
                                    C.c_void_p(dirs.data_ptr()), d_stride, L.ptr(z), n, s, flags, L.ptr(out), pitch,
                                    C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def mlp_trunk_forward(model, origins, o_stride, dirs, d_stride, z, flags, total_points, row0, sigma_out, ws):
    """PE + trunk of the points of `z` into rows [row0, row0 + z.numel()) of a workspace sized for total_points."""
    n, s = z.shape
    L.check(L.load().bn_mlp_trunk_forward(model.handle(), L.ptr(model.flat_params), C.c_void_p(origins.data_ptr()), o_stride,
                                          C.c_void_p(dirs.data_ptr()), d_stride, L.ptr(z), n, s, flags, int(total_points),
                                          int(row0), L.ptr(sigma_out), C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def mlp_write_t(model, t_rows, n, s, flags, total_points, row0, ws):
    """Time embedding of n rays (t_rows (n, t_dims) fp32) into the workspace rows of their n*s points (bn_mlp_write_t)."""
    if t_rows is None:
        raise L.BnError("this model has a beta head: the time embedding of the rays (models['t'](ts)) is required")
    if t_rows.shape != (n, model.t_embedding_dims):
        raise L.BnError(f"time embedding must be ({n}, {model.t_embedding_dims}), got {tuple(t_rows.shape)}")
    L.check(L.load().bn_mlp_write_t(model.handle(), L.ptr(t_rows), n, s, flags, int(total_points), int(row0),
                                    C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def mlp_heads_forward(model, total_points, flags, out, pitch, ws):
    L.check(L.load().bn_mlp_heads_forward(model.handle(), L.ptr(model.flat_params), int(total_points), flags, L.ptr(out), pitch,
                                          C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def permute_samples(src, idx, n, s1, g, pitch, scatter: bool):
    """sort_idx applied to per-point rows: MLP row order ([N][S1] block, [N][G] block) <-> depth order (N,S1+G)."""
    dst = torch.empty((n, s1 + g, pitch) if not scatter else (n * (s1 + g), pitch), dtype=torch.float32, device=src.device)
    L.check(L.load().bn_permute_samples(L.ptr(src), L.ptr(idx, torch.int64), L.ptr(dst), n, s1, g, pitch, int(bool(scatter)),
                                        L.stream_ptr()))
    return dst


def mlp_backward(model, out, g_out, pitch, n, s, flags, g_params, ws):
    L.check(L.load().bn_mlp_backward(model.handle(), L.ptr(model.flat_params), L.ptr(out), L.ptr(g_out), pitch, n, s, flags,
                                     L.ptr(g_params), C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def mlp_normals_forward(model, out, pitch, n, s, flags, ws):
    L.check(L.load().bn_mlp_normals_forward(model.handle(), L.ptr(model.flat_params), L.ptr(out), pitch, n, s, flags,
                                            4 + int(bool(flags & L.MLP_BETA)),        # the beta channel precedes the normals
                                            C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))


def mlp_normals_backward(model, out, g_out, pitch, n, s, flags, g_params, ws):
    L.check(L.load().bn_mlp_normals_backward(model.handle(), L.ptr(model.flat_params), L.ptr(out), L.ptr(g_out), pitch, n, s,
                                             flags, 4 + int(bool(flags & L.MLP_BETA)), L.ptr(g_params), C.c_void_p(ws.data_ptr()), ws.numel(),
                                             L.stream_ptr()))


def adam_step(params, grads, m, v, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    L.check(L.load().bn_adam_step(L.ptr(params), L.ptr(grads), L.ptr(m), L.ptr(v), params.numel(), float(lr),
                                  float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                  float(grad_scale), L.stream_ptr()))


def adam_step_graph(params, grads, m, v, state, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """Adam whose learning rate and step counter live on the device (`state` = [lr, step, -, -] fp32): capturable."""
    L.check(L.load().bn_adam_step_graph(L.ptr(params), L.ptr(grads), L.ptr(m), L.ptr(v), params.numel(), L.ptr(state),
                                        float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(grad_scale),
                                        L.stream_ptr()))


def loss_color_depth(rgb, target_rgb, lambda_rgb, depth=None, z=None, weights=None, valid_depth=None, target_depths=None,
                     target_std=None, lambda_ds=0.0, use_all_depth=False, no_weights=False):
    """SNerfLoss + DepthLoss(subset=True) fused with their gradients (reference metrics.py:39-61, 82-161).
    Returns (loss (1,), g_rgb (N,3), g_depth (N) or None).  target_depths is the reference's (N,2) [depth, weight]."""
    n = rgb.shape[0]
    dev = rgb.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    g_rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    use_ds = valid_depth is not None and float(lambda_ds) > 0
    g_depth = torch.empty(n, dtype=torch.float32, device=dev) if use_ds else None
    td = tw = None
    stride = 1
    if use_ds:
        tdm = target_depths.contiguous().float()
        stride = tdm.shape[1] if tdm.dim() == 2 else 1
        td = tdm
        tw = None if (no_weights or tdm.dim() != 2) else tdm[:, 1]
        valid_depth = valid_depth.to(torch.int64).contiguous()
    s = z.shape[1] if z is not None else 1
    L.check(L.load().bn_loss_color_depth(
        L.ptr(rgb.contiguous()), L.ptr(target_rgb.contiguous()), L.ptr(depth), L.ptr(z), L.ptr(weights),
        L.ptr(valid_depth, torch.int64) if use_ds else None, L.ptr(td), (C.c_void_p(tw.data_ptr()) if tw is not None else None), stride,
        L.ptr(target_std.contiguous().float()) if use_ds else None, float(lambda_rgb), float(lambda_ds), int(bool(use_all_depth)),
        L.ptr(loss), L.ptr(g_rgb), L.ptr(g_depth), n, s, L.stream_ptr()))
    return loss, g_rgb, g_depth


def lambertian_render_loss(z, rows, sort_idx, n_stratified, target_rgb, lambda_rgb, valid_depth=None, target_depths=None,
                           target_std=None, lambda_ds=0.0, use_all_depth=False, no_weights=False, want_outputs=False):
    """Compositing + Lambertian colour + colour / depth loss + their backward in one launch (bn_lambertian_render_loss).
    Returns (loss (1,), g_rows [N*S, 4], rgb (N,3) or None, depth (N) or None)."""
    n, s = z.shape
    dev = z.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    g_rows = torch.empty_like(rows)
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_outputs else None
    depth = torch.empty(n, dtype=torch.float32, device=dev) if want_outputs else None
    use_ds = valid_depth is not None and float(lambda_ds) > 0
    td = tw = None
    stride = 1
    if use_ds:
        tdm = target_depths.contiguous().float()
        stride = tdm.shape[1] if tdm.dim() == 2 else 1
        td = tdm
        tw = None if (no_weights or tdm.dim() != 2) else tdm[:, 1]
        valid_depth = valid_depth.to(torch.int64).contiguous()
    L.check(L.load().bn_lambertian_render_loss(
        L.ptr(z), L.ptr(rows), L.ptr(sort_idx, torch.int64), int(n_stratified), L.ptr(target_rgb.contiguous()),
        L.ptr(valid_depth, torch.int64) if use_ds else None, L.ptr(td), (C.c_void_p(tw.data_ptr()) if tw is not None else None),
        stride, L.ptr(target_std.contiguous().float()) if use_ds else None, float(lambda_rgb), float(lambda_ds) if use_ds else 0.0,
        int(bool(use_all_depth)), L.ptr(loss), L.ptr(g_rows), L.ptr(rgb), L.ptr(depth), n, s, L.stream_ptr()))
    return loss, g_rows, rgb, depth


def loss_regularizers(loss, weights, z, depth, packed, rays, normal_an_ch=-1, lambda_nr_an=0.0, normal_lr_ch=-1,
                      lambda_nr_lr=0.0, lambda_hs=0.0, g_depth=None, want_bad_count=False):
    """NormalRegLoss / HardSurfaceLoss (metrics.py:179-216, 263-290) fused with their gradients.  ADDS to `loss` (1,) and
    to `g_depth` (N); returns (g_weights (N,S), g_packed (N,S,pitch) or None, g_depth, bad_count (2,) or None)."""
    n, s = weights.shape
    dev = weights.device
    pitch = packed.shape[-1]
    nr = (normal_an_ch >= 0 and lambda_nr_an != 0.0) or (normal_lr_ch >= 0 and lambda_nr_lr != 0.0)
    g_weights = torch.empty_like(weights)
    g_packed = torch.empty((n, s, pitch), dtype=torch.float32, device=dev) if nr else None
    if lambda_hs != 0.0 and g_depth is None:
        g_depth = torch.zeros(n, dtype=torch.float32, device=dev)
    bad = torch.zeros(2, dtype=torch.float32, device=dev) if want_bad_count else None
    L.check(L.load().bn_loss_regularizers(L.ptr(weights), L.ptr(z), L.ptr(depth), L.ptr(packed), pitch, int(normal_an_ch),
                                          float(lambda_nr_an), int(normal_lr_ch), float(lambda_nr_lr), L.ptr(rays),
                                          float(lambda_hs), L.ptr(loss), L.ptr(g_weights), L.ptr(g_packed), L.ptr(g_depth),
                                          L.ptr(bad), n, s, L.stream_ptr()))
    return g_weights, g_packed, g_depth, bad
