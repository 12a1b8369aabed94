"""Step-fraction schedules and the device-resident ray feed of the training loop.

Host-side restatement of the bookkeeping `NeRF_pl.training_step` does around the hot path (reference
main.py:57-68, 194-246, 147-168; train_utils.py:144-158).  Nothing here touches the device except the
ray pool, whose per-step batch is ONE gather per tensor on the GPU instead of the reference's per-ray
Python `__getitem__` dicts collated by four DataLoader workers (datasets/satellite_rgb_dep.py:708-716,
main.py:170-176).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from .synth import RayBatch


@dataclasses.dataclass
class StepFlags:
    """What one training step passes to render_rays / the losses."""
    gsam_only: bool
    apply_brdf: bool
    apply_theta: bool
    cos_irra_on: bool
    use_depth_loss: bool
    use_normal_reg: bool
    lr: float
    noise_std: float


class Schedule:
    """Thresholds are fractions of `max_train_steps`, rounded like the reference (np.round, main.py:59-68);
    a switch turns on when `train_steps > threshold` (main.py:202-210), the depth loss stays on while
    `train_steps < ds_drop` (main.py:248); `train_steps` advances by the number of GPUs per step (main.py:196);
    `noise_std` decays by 0.9 per step (main.py:246); the learning rate is StepLR(step_size=1 epoch, gamma=0.9)
    (train_utils.py:153-155) with epoch = train_steps // (dataset_len // batch_size) (train_utils.py:117-118)."""

    def __init__(self, args, dataset_len: int, world_size: int = 1):
        m = float(args.max_train_steps)
        self.brdf_on = float(np.round(args.brdf_on * m))
        self.nrrg_on = float(np.round(args.nrrg_on * m))
        self.gsam_only_on = float(np.round(args.gsam_only_on * m))
        self.cos_irra_on = float(np.round(args.cos_irra_on * m))
        self.depth = float(args.ds_lambda) > 0
        self.ds_drop = float(np.round(args.ds_drop * m)) if self.depth else 0.0
        self.base_lr = float(args.lr)
        self.batch_size = int(args.batch_size)
        self.dataset_len = int(dataset_len)
        self.world = int(world_size)
        self.noise_std = float(args.noise_std)
        self.train_steps = 0

    def epoch(self, train_steps: Optional[int] = None) -> int:
        t = self.train_steps if train_steps is None else train_steps
        return int(t // max(self.dataset_len // self.batch_size, 1))

    def next(self) -> StepFlags:
        """Advance by one optimisation step and return its flags."""
        self.train_steps += self.world
        t = self.train_steps
        flags = StepFlags(gsam_only=t > self.gsam_only_on, apply_brdf=t > self.brdf_on, apply_theta=t > self.brdf_on * 2,
                          cos_irra_on=t > self.cos_irra_on, use_depth_loss=self.depth and t < self.ds_drop,
                          use_normal_reg=t > self.nrrg_on, lr=self.base_lr * (0.9 ** self.epoch(t)), noise_std=self.noise_std)
        self.noise_std *= 0.9
        return flags


class DeviceRayPool:
    """All training rays resident on the GPU: `all_rays (R,11)`, `all_rgbs (R,3)` and, with depth supervision,
    `all_valid_depth (R)`, `all_depths (R,2)`, `all_depth_stds (R)` (the tensors SatelliteRGBDEPDataset keeps on
    the host, satellite_rgb_dep.py:390-548).  `next_batch()` walks a device-side random permutation that is
    redrawn every epoch (DataLoader(shuffle=True) semantics, main.py:170-176); with `world_size > 1` every rank
    takes its own stride of the permutation (DistributedSampler semantics)."""

    def __init__(self, pool: RayBatch, batch_size: int, rank: int = 0, world_size: int = 1, seed: int = 0):
        if not pool.rays.is_cuda:
            raise ValueError("DeviceRayPool needs the pool on a CUDA device")
        self.pool, self.bs, self.rank, self.world = pool, int(batch_size), int(rank), int(world_size)
        self.gen = torch.Generator(device=pool.rays.device)
        self.gen.manual_seed(seed)
        self.n = pool.rays.shape[0]
        self.epoch = -1
        self._perm = None
        self._pos = 0

    def _reshuffle(self):
        self.epoch += 1
        perm = torch.randperm(self.n, device=self.pool.rays.device, generator=self.gen)
        self._perm = perm[self.rank::self.world]
        self._pos = 0

    def next_batch(self) -> RayBatch:
        if self._perm is None or self._pos + self.bs > self._perm.shape[0]:
            self._reshuffle()
        idx = self._perm[self._pos:self._pos + self.bs]
        self._pos += self.bs
        take = lambda t: None if t is None else t.index_select(0, idx)
        p = self.pool
        return RayBatch(take(p.rays), take(p.rgbs), take(p.valid_depth), take(p.target_depths), take(p.target_std))
