"""ORACLE (test infrastructure only — never imported by the product path).

Restatement of the two losses the Lambertian / depth-supervised training step uses.
  * SNerfLoss (lambda_sc == 0 on this path, opt.py:340-341)   metrics.py:39-61
  * DepthLoss, subset=True, GNLL=False                          metrics.py:82-161
and of the two optional regularisers of the BRDF stage (main.py:269-299):
  * NormalRegLoss (normals facing away from the camera)          metrics.py:179-216
  * HardSurfaceLoss (variance of the samples around the depth)   metrics.py:263-290, train_utils.py:38-39
"""
from __future__ import annotations

import torch


def color_loss(res, rgbs, lambda_rgb=1.0):
    return lambda_rgb * torch.mean((res["rgb_coarse"] - rgbs) ** 2)


def depth_loss(res, target_depth, target_weight, valid_depth, target_std, lambda_ds, usealldepth=False):
    """metrics.py:98-161. Selected rays = valid & (|d-d*| > std* or pred_std > std*); value is
    lambda_ds/3 * mean_sel( n_sel/n_batch * weight * (d-d*)^2 ); zero when nothing is selected."""
    v = valid_depth > 0
    z, d, w = res["z_vals_coarse"][v], res["depth_coarse"][v], res["weights_coarse"][v]
    if d.shape[0] == 0:
        return torch.zeros((), dtype=res["depth_coarse"].dtype)
    tw, td, ts = target_weight[v], target_depth[v], target_std[v]
    pred_std = torch.sqrt((((z - d.unsqueeze(-1)) ** 2) * w).sum(-1))
    sel = torch.ones_like(d, dtype=torch.bool) if usealldepth else (((d - td).abs() - ts) > 0) | (ts < pred_std)
    if int(sel.sum()) == 0:
        return torch.zeros((), dtype=res["depth_coarse"].dtype)
    frac = float(sel.sum()) / float(valid_depth.shape[0])
    per_ray = frac * tw[sel] * (d[sel] - td[sel]) ** 2
    return (lambda_ds / 3.0) * per_ray.mean()


def normal_reg_loss(res, lambda_nr_reg, keyword="normal_an"):
    """metrics.py:187-214.  n.v with v = rays_d_coarse (= -d, facing the camera), clipped at 0 from above, squared,
    weighted by the compositing weights and SUMMED over every ray and sample (the reference flattens both tensors
    before `.sum(dim=-1)`, so its `torch.mean` acts on a scalar).  Returns (loss, % of normals with n.v < 0)."""
    normal = res[f"{keyword}_coarse"].reshape(-1, 3)
    weights = res["weights_coarse"].reshape(-1)
    view = res["rays_d_coarse"].reshape(-1, 3)
    rep = normal.shape[0] // view.shape[0]
    n_dot_v = (normal * torch.repeat_interleave(view, rep, dim=0)).sum(-1)
    perc = 100.0 * float((n_dot_v < 0).sum()) / n_dot_v.numel()
    return lambda_nr_reg * (weights * torch.clamp(n_dot_v, max=0.0) ** 2).sum(), perc


def hard_surface_loss(res, lambda_hs):
    """metrics.py:270-288 with calc_depth_std_2 (train_utils.py:38-39): lambda_hs * mean_r sum_s (z - depth_r)^2 w."""
    var = ((res["z_vals_coarse"] - res["depth_coarse"].unsqueeze(-1)) ** 2 * res["weights_coarse"]).sum(-1)
    return lambda_hs * var.mean()


def train_loss(res, batch, args):
    loss = color_loss(res, batch.rgbs.to(res["rgb_coarse"].dtype), args.lambda_rgb)
    if args.ds_lambda > 0 and batch.valid_depth is not None:
        dt = res["depth_coarse"].dtype
        loss = loss + depth_loss(res, batch.target_depths[:, 0].to(dt), batch.target_depths[:, 1].to(dt),
                                 batch.valid_depth, batch.target_std.to(dt), args.ds_lambda,
                                 usealldepth=bool(args.usealldepth))
    return loss
