"""ORACLE (test infrastructure only — never imported by the product path).

One optimisation step of the UNMODIFIED reference on a chosen torch device: the reference's own `render_rays`
(rendering.py:168-334), its own loss classes (metrics.py: `load_loss` -> SNerfLoss, `DepthLoss(subset=True)`) composed the
way `NeRF_pl.training_step` composes them (main.py:194-268) and `torch.optim.Adam(lr=args.lr)` over
`filter(requires_grad)` (main.py:147-150).  Only the Lightning glue around those calls is missing (pytorch_lightning is not
installed): this file is that glue, nothing of the path is restated here.

Used by bench.py:
  * `--impl reference`      -> device 'cpu', all host threads (the reference arm, `cpu_baseline.kind = "reference"`)
  * `reference_cuda_eager`  -> device 'cuda' (the same code through torch-CUDA eager kernels on the same B200:
                               the same-box comparator SURVEY §8d asks for; allow_tf32 = False, strict fp32)
"""
from __future__ import annotations

import contextlib
import io
import time

import torch

from . import ref_harness as RH


class ReferenceStep:
    def __init__(self, args, device="cpu", seed=0):
        self.args = args
        self.device = torch.device(device)
        self.rendering, _, self.metrics = RH.load()
        self.model = RH.build_model(args, seed=seed).to(self.device)
        self.models = {"coarse": self.model}
        with contextlib.redirect_stdout(io.StringIO()):
            self.loss = self.metrics.load_loss(args)                       # main.py:57
            self.depth_loss = None
            if float(args.ds_lambda) > 0:                                  # main.py:65-68
                self.depth_loss = self.metrics.DepthLoss(lambda_ds=args.ds_lambda, GNLL=args.GNLL, usealldepth=args.usealldepth,
                                                         margin=args.margin, stdscale=args.stdscale, subset=True)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.opt = torch.optim.Adam(params, lr=args.lr, weight_decay=0)    # main.py:147-150

    def step(self, batch, apply_brdf=False, apply_theta=False, cos_irra_on=False, gsam_only=False):
        """batch: brdf_nerf_b200.synth.RayBatch on self.device.  Returns the loss tensor."""
        a = self.args
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res, _ = self.rendering.render_rays(self.models, a, batch.rays, None, mode="train", valid_depth=batch.valid_depth,
                                                target_depths=batch.target_depths, target_std=batch.target_std,
                                                apply_brdf=apply_brdf, print_debuginfo=False, bTestNormal=False,
                                                gsam_only=gsam_only, apply_theta=apply_theta, cos_irra_on=cos_irra_on)
            loss, _ = self.loss(res, batch.rgbs)
            if self.depth_loss is not None and batch.valid_depth is not None:
                ld, _ = self.depth_loss(res, batch.target_depths[:, 0], batch.target_depths[:, 1],
                                        target_valid_depth=batch.valid_depth, target_std=batch.target_std)
                loss = loss + ld
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    @torch.no_grad()
    def render(self, rays, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            return self.rendering.render_rays(self.models, self.args, rays, None, **kw)


def rate(args, batch, steps, warmup, device="cpu", step_kw=None):
    """rays/s of `steps` reference steps after `warmup` (wall clock around device-synchronised steps)."""
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = False           # strict fp32 (the reference never sets it)
        torch.backends.cudnn.allow_tf32 = False
    rs = ReferenceStep(args, device=device)
    b = batch.to(dev)
    times = []
    for i in range(warmup + steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        rs.step(b, **(step_kw or {}))
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return b.rays.shape[0] / (ms / 1e3), ms
