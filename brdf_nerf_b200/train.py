"""Minimal training-step driver for the hot path (replaces the Lightning glue of the reference,
main.py:194-353 + 147-168): forward through the CUDA pipeline, loss gradients,
backward kernels writing straight into the flat fp32 gradient bucket, ONE all-reduce of that bucket
over NCCL when world_size > 1, fused Adam on the flat buffers.

No autograd tape is built: the step calls the forward / backward kernel chains directly
(`rendering._forward` / `rendering._backward`).  Nothing in a step synchronises with the host, so it
can be captured in a CUDA graph (`use_graph=True`).

Losses (reference metrics.py:39-61 SNerfLoss with lambda_sc = 0, metrics.py:82-161 DepthLoss with
subset=True, GNLL=False):
    L = lambda_rgb * mean((rgb - target)^2)
      + [step < ds_drop] (lambda_ds / 3) * mean_sel( n_sel/N * w_i (d_i - d*_i)^2 )
The depth term's normaliser n_sel/N * 1/n_sel collapses to 1/N, so the selection mask is applied as
a 0/1 weight without any host round trip (the reference syncs three times to build index lists).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import ops
from . import rendering as R
from .synth import RayBatch


def loss_and_grads(args, outs, st, batch: RayBatch, use_depth: bool):
    """Returns (loss (1,) tensor, g_rgb (N,3), g_depth (N) or None): one fused launch (csrc/loss.cu)."""
    use_ds = use_depth and batch.valid_depth is not None
    return ops.loss_color_depth(outs["rgb"], batch.rgbs, float(args.lambda_rgb), depth=outs["depth"], z=outs["z"],
                                weights=outs["weights"], valid_depth=batch.valid_depth if use_ds else None,
                                target_depths=batch.target_depths, target_std=batch.target_std,
                                lambda_ds=float(args.ds_lambda) if use_ds else 0.0,
                                use_all_depth=bool(getattr(args, "usealldepth", False)),
                                no_weights=bool(getattr(args, "ds_noweights", False)))


def allreduce_grads_(flat_grads: torch.Tensor, world_size: int, group=None) -> float:
    """Data-parallel gradient exchange: ONE all-reduce (sum) of the flat fp32 bucket over
    NCCL/NVLink (gloo in the CPU tests).  Returns the scale (1/world) the optimizer must apply — the
    averaging is folded into the fused Adam kernel instead of a separate pass over the bucket.
    Matches what DDP does in the reference (Lightning `gpus > 1`, main.py:720-731): per-rank mean loss,
    gradients averaged over ranks."""
    if world_size > 1:
        torch.distributed.all_reduce(flat_grads, op=torch.distributed.ReduceOp.SUM, group=group)
    return 1.0 / world_size


class Trainer:
    def __init__(self, model, args, lr: Optional[float] = None, world_size: int = 1, process_group=None,
                 use_graph: bool = False):
        self.model, self.args = model, args
        self.lr = float(args.lr if lr is None else lr)
        self.world = int(world_size)
        self.pg = process_group
        self.step_count = 0
        self.use_graph = use_graph
        flat = model.flat_params
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self._graph = None
        self._static: Optional[RayBatch] = None
        self._loss = None
        self._kw = None
        self.graph_launches = 0
        self.use_depth_loss = True          # switched off by the schedule after ds_drop (main.py:248)

    # one optimisation step; `batch` tensors must already live on the model's device
    def _step_impl(self, batch: RayBatch, draws, kw):
        model, args = self.model, self.args
        outs, st = R._forward(model, args, batch.rays, draws, train=True, mode="train",
                              valid_depth=batch.valid_depth, target_depths=batch.target_depths,
                              target_std=batch.target_std, **kw)
        use_depth = float(args.ds_lambda) > 0 and self.use_depth_loss
        loss, g_rgb, g_depth = loss_and_grads(args, outs, st, batch, use_depth)
        grads = model.flat_grads
        grads.zero_()
        R._backward(model, st, g_rgb, g_depth, None, None, grads)
        return loss

    def _reduce_and_update(self):
        model = self.model
        scale = allreduce_grads_(model.flat_grads, self.world, self.pg)
        self.step_count += 1
        ops.adam_step(model.flat_params, model.flat_grads, self.m, self.v, self.lr, self.step_count,
                      grad_scale=scale)
        model._synced_version = -1            # the packed bf16 copies are stale now

    def step(self, batch: RayBatch, draws=None, apply_brdf=False, apply_theta=False, cos_irra_on=False,
             gsam_only=False):
        kw = dict(apply_brdf=apply_brdf, bTestNormal=False, bTestSun_v=False, gsam_only=gsam_only,
                  apply_theta=apply_theta, cos_irra_on=cos_irra_on)
        if not self.use_graph or draws is not None:
            loss = self._step_impl(batch, draws, kw)
            self._reduce_and_update()
            return loss
        return self._graph_step(batch, kw)

    # ---- CUDA-graph path: static input buffers, forward+backward captured once, replayed per step
    def _graph_step(self, batch: RayBatch, kw):
        kw = dict(kw, _use_depth=self.use_depth_loss)
        if self._graph is None or self._kw != kw:
            self._static = RayBatch(*[None if t is None else t.clone() for t in
                                      (batch.rays, batch.rgbs, batch.valid_depth, batch.target_depths, batch.target_std)])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up: fills table / workspace caches
                rkw = {k: v for k, v in kw.items() if not k.startswith("_")}
                for _ in range(2):
                    self.model.sync_weights(force=True)
                    self._step_impl(self._static, None, rkw)
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            lib = L.load()
            lc0 = lib.bn_launch_count()
            with torch.cuda.graph(self._graph):
                self.model.sync_weights(force=True)
                self._loss = self._step_impl(self._static, None, rkw)
            self.graph_launches = int(lib.bn_launch_count() - lc0)     # library kernels replayed by every graph launch
            self._kw = dict(kw)
        for dst, src in zip((self._static.rays, self._static.rgbs, self._static.valid_depth,
                             self._static.target_depths, self._static.target_std),
                            (batch.rays, batch.rgbs, batch.valid_depth, batch.target_depths, batch.target_std)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self._reduce_and_update()
        return self._loss


class TrainLoop:
    """The reference's training loop around the hot path (NeRF_pl.training_step + DataLoader + StepLR,
    main.py:147-246) without Lightning: device-resident ray pool -> schedule flags -> Trainer.step.
    One call of `step()` = one optimisation step; nothing synchronises with the host."""

    def __init__(self, model, args, pool: RayBatch, world_size: int = 1, rank: int = 0, process_group=None,
                 use_graph: bool = True, seed: int = 0):
        from .schedule import DeviceRayPool, Schedule
        self.args = args
        self.trainer = Trainer(model, args, world_size=world_size, process_group=process_group, use_graph=use_graph)
        self.feed = DeviceRayPool(pool, int(args.batch_size), rank=rank, world_size=world_size, seed=seed)
        self.schedule = Schedule(args, dataset_len=pool.rays.shape[0], world_size=world_size)
        if float(args.noise_std) != 0.0 and use_graph:
            raise ValueError("noise_std decays every step (main.py:246): run with use_graph=False when it is non-zero")

    def step(self):
        f = self.schedule.next()
        self.trainer.lr = f.lr
        self.trainer.use_depth_loss = f.use_depth_loss
        self.args.noise_std = f.noise_std
        batch = self.feed.next_batch()
        return self.trainer.step(batch, apply_brdf=f.apply_brdf, apply_theta=f.apply_theta, cos_irra_on=f.cos_irra_on,
                                 gsam_only=f.gsam_only)
