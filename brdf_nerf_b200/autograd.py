"""torch.autograd bridge for direct use of the module (`SpSBRDFNeRF.forward` on a batch of points).
The ray pipeline has its own bridge in rendering.py (`_RenderFunction`)."""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops


class PointsFunction(torch.autograd.Function):
    """(B,3) points -> (B,1) sigma or packed (B,C) (reference spsbrdfnerf.py:662-757).
    A point is evaluated as a one-sample ray with origin x and z = 0 (x + d*0 is exact); `dirs` (B,3) is the view
    direction of each point for models built with input_viewdir=1 (None otherwise)."""

    @staticmethod
    def forward(ctx, model, xyz, dirs, t_rows, sigma_only, apply_brdf, apply_theta, nr_an_on, nr_lr_on, need_grad, *params):
        if not xyz.is_cuda:
            raise L.BnError("SpSBRDFNeRF.forward needs CUDA tensors: there is no CPU path")
        xyz = xyz.detach().float().contiguous()
        dirs = xyz if dirs is None else dirs.detach().float().contiguous()
        B = xyz.shape[0]
        model.sync_weights()
        z = torch.zeros((B, 1), dtype=torch.float32, device=xyz.device)
        full_for_grad = sigma_only and need_grad
        if sigma_only and not full_for_grad:
            flags = L.MLP_SIGMA_ONLY
            out = torch.empty((B, 1), dtype=torch.float32, device=xyz.device)
            ops.mlp_forward(model, xyz, 3, dirs, 3, z, flags, out, 1, model.workspace(B, flags, "ws_points"))
            return out
        flags = model.mlp_flags(apply_brdf=apply_brdf, apply_theta=apply_theta, nr_an_on=nr_an_on and not sigma_only,
                                nr_lr_on=nr_lr_on and not sigma_only, train=need_grad)
        if sigma_only:
            flags &= ~L.MLP_BETA                    # the density does not depend on the time embedding
        C = model.out_channels(flags)
        out = torch.empty((B, C), dtype=torch.float32, device=xyz.device)
        # a call that will be differentiated owns its workspace (kept alive by ctx): the reference calls the module once per
        # chunk and runs backward afterwards, so a shared buffer would be overwritten before it is read
        ws = model.workspace(B, flags, None if need_grad else "ws_points")
        if flags & L.MLP_BETA:                       # every point is its own one-sample ray: t row i belongs to point i
            ops.mlp_write_t(model, t_rows, B, 1, flags, B, 0, ws)
        ops.mlp_forward(model, xyz, 3, dirs, 3, z, flags, out, C, ws)
        if flags & L.MLP_NORMAL_AN:
            ops.mlp_normals_forward(model, out, C, B, 1, flags, ws)
        ctx.model, ctx.flags, ctx.B, ctx.C, ctx.sigma_col = model, flags, B, C, full_for_grad
        ctx.ws = ws
        ctx.save_for_backward(out)
        return out[:, 3:4].clone() if full_for_grad else out

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        (out,) = ctx.saved_tensors
        if ctx.sigma_col:
            g_out = torch.zeros_like(out)
            g_out[:, 3:4] = g
        else:
            g_out = g.contiguous()
        flat = torch.zeros_like(model.flat_params)
        if ctx.flags & L.MLP_NORMAL_AN:
            ops.mlp_normals_backward(model, out, g_out, ctx.C, ctx.B, 1, ctx.flags, flat, ctx.ws)
        ops.mlp_backward(model, out, g_out, ctx.C, ctx.B, 1, ctx.flags, flat, ctx.ws)
        grads = model.grad_views(flat)
        return (None, None, None, None, None, None, None, None, None, None, *grads)
