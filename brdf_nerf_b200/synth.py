"""Deterministic synthetic satellite-ray batches (SURVEY.md §8d "Synthetic inputs").

The ray record is the reference's `(N, 11)` fp32 row `[o(3), d(3), near, far, sun_d(3)]`
(reference `datasets/satellite_rgb_dep.py:311-322,390,550-559`), normalised scene in [-1, 1]^3.
Everything is generated on the CPU with an explicit `torch.Generator` so that the GPU run, the
CPU oracle and the golden fixtures all see identical bits.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

RAY_SEED = 20240912

_OFF_NADIR_DEG = (5.0, 15.0, 25.0)
_VIEW_AZ_DEG = (30.0, 150.0, 270.0)
_SUN_EL_DEG = (60.0, 50.0, 40.0)
_SUN_AZ_DEG = (140.0, 150.0, 160.0)
_SLAB_TOP = 0.3
_SLAB_THICKNESS = 0.6


@dataclass
class RayBatch:
    rays: torch.Tensor                 # (N, 11) fp32
    rgbs: torch.Tensor                 # (N, 3)  fp32 targets
    valid_depth: Optional[torch.Tensor] = None     # (N,) int64
    target_depths: Optional[torch.Tensor] = None   # (N, 2) [depth, correlation weight]
    target_std: Optional[torch.Tensor] = None      # (N,)
    flat = None                                    # set by packed(): the single buffer behind all tensors
    _ready = None                                  # set by Trainer.prefetch(): events guarding a staging batch
    _free = None

    def to(self, device, non_blocking=False):
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return RayBatch(mv(self.rays), mv(self.rgbs), mv(self.valid_depth),
                        mv(self.target_depths), mv(self.target_std))

    def pin(self):
        pn = lambda t: None if t is None else t.pin_memory()
        return RayBatch(pn(self.rays), pn(self.rgbs), pn(self.valid_depth),
                        pn(self.target_depths), pn(self.target_std))

    def packed(self, device=None, pin=False) -> "RayBatch":
        """The same batch with all its tensors laid out back to back in ONE buffer (`.flat`, bytes), so that a host
        batch reaches the device with a single copy (`dst.flat.copy_(src.flat)`) instead of one per tensor."""
        ts = [self.rays, self.rgbs, self.valid_depth, self.target_depths, self.target_std]
        dev = ts[0].device if device is None else torch.device(device)
        offs, total = [], 0
        for t in ts:
            offs.append(total)
            if t is not None:
                total += -(-t.numel() * t.element_size() // 16) * 16
        flat = torch.empty(total, dtype=torch.uint8, device=dev)
        if pin:
            flat = flat.pin_memory()
        views = []
        for t, o in zip(ts, offs):
            if t is None:
                views.append(None)
                continue
            v = flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            v.copy_(t)
            views.append(v)
        out = RayBatch(*views)
        out.flat = flat
        return out

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.rays, self.rgbs, self.valid_depth, self.target_depths, self.target_std)
                   if t is not None)

    def shard(self, rank: int, world: int) -> "RayBatch":
        """Contiguous ray shard of rank `rank`."""
        n = self.rays.shape[0]
        per = n // world
        sl = slice(rank * per, (rank + 1) * per)
        cut = lambda t: None if t is None else t[sl].contiguous()
        return RayBatch(cut(self.rays), cut(self.rgbs), cut(self.valid_depth),
                        cut(self.target_depths), cut(self.target_std))


def _view_and_sun(v: int):
    th, az = math.radians(_OFF_NADIR_DEG[v]), math.radians(_VIEW_AZ_DEG[v])
    d = torch.tensor([math.sin(th) * math.sin(az), math.sin(th) * math.cos(az), -math.cos(th)],
                     dtype=torch.float64)
    el, saz = math.radians(_SUN_EL_DEG[v]), math.radians(_SUN_AZ_DEG[v])
    # sun direction formula: reference satellite_rgb_dep.py:571-573
    s = torch.tensor([math.sin(saz) * math.cos(el), math.cos(saz) * math.cos(el), math.sin(el)],
                     dtype=torch.float64)
    return d, s


def _surface_depth(o: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """Ray / height-field z = -0.1 + 0.1 sin(3x) cos(2y) intersection distance (fp64 bisection)."""
    lo = torch.zeros(o.shape[0], dtype=torch.float64)
    hi = (_SLAB_THICKNESS / d[:, 2].abs())
    f = lambda t: (o[:, 2] + d[:, 2] * t) - (-0.1 + 0.1 * torch.sin(3 * (o[:, 0] + d[:, 0] * t))
                                             * torch.cos(2 * (o[:, 1] + d[:, 1] * t)))
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        above = f(mid) > 0
        lo = torch.where(above, mid, lo)
        hi = torch.where(above, hi, mid)
    return 0.5 * (lo + hi)


def make_rays(n: int, seed: int = RAY_SEED, depth_supervision: bool = False,
              zero_std: bool = False, stdscale: float = 1.0, margin: float = 1e-4,
              single_view: Optional[int] = None) -> RayBatch:
    """`n` synthetic pushbroom rays over three views (or one, for tile inference)."""
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g, dtype=torch.float64) * 2 - 1
    jitter = torch.randn(n, 3, generator=g, dtype=torch.float64) * 1e-3
    rgbs = torch.rand(n, 3, generator=g, dtype=torch.float32)
    view = torch.arange(n) % 3 if single_view is None else torch.full((n,), single_view)
    dv = torch.stack([_view_and_sun(v)[0] for v in range(3)])[view]
    sv = torch.stack([_view_and_sun(v)[1] for v in range(3)])[view]
    d = dv + jitter
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.cat([xy, torch.full((n, 1), _SLAB_TOP, dtype=torch.float64)], -1)
    near = torch.zeros(n, 1, dtype=torch.float64)        # satellite_rgb_dep.py:73
    far = _SLAB_THICKNESS / d[:, 2:3].abs()
    rays = torch.cat([o, d, near, far, sv], -1).to(torch.float32).contiguous()
    batch = RayBatch(rays=rays, rgbs=rgbs)
    if depth_supervision:
        valid = (torch.rand(n, generator=g) < 0.7).to(torch.int64)
        corr = torch.rand(n, generator=g, dtype=torch.float64) * 0.5 + 0.5
        depth = _surface_depth(o, d)
        batch.valid_depth = valid
        batch.target_depths = torch.stack([depth, corr], -1).to(torch.float32).contiguous()
        std = torch.zeros(n, dtype=torch.float64) if zero_std else stdscale * (1 - corr) + margin
        batch.target_std = std.to(torch.float32).contiguous()
    return batch


def make_tile_rays(h: int, w: int, view: int = 0) -> torch.Tensor:
    """Row-major pixel grid of one view for full-tile inference (cfg 5); returns (h*w, 11)."""
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h, dtype=torch.float64),
                            torch.linspace(-1, 1, w, dtype=torch.float64), indexing="ij")
    d0, s0 = _view_and_sun(view)
    n = h * w
    o = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.full((n,), _SLAB_TOP, dtype=torch.float64)], -1)
    d = d0.expand(n, 3)
    far = torch.full((n, 1), _SLAB_THICKNESS / abs(float(d0[2])), dtype=torch.float64)
    rays = torch.cat([o, d, torch.zeros(n, 1, dtype=torch.float64), far, s0.expand(n, 3)], -1)
    return rays.to(torch.float32).contiguous()


def tile_surface_depth(rays: torch.Tensor) -> torch.Tensor:
    """Depth (distance along each ray, float32) at which the rays of a tile meet the synthetic height field
    z = -0.1 + 0.1 sin(3x) cos(2y): stands in for the rendered `depth_coarse` of a tile in the DSM tests / benchmarks."""
    r = rays.to(torch.float64).cpu()
    return _surface_depth(r[:, 0:3], r[:, 3:6]).to(torch.float32)


# georeferencing of the synthetic scene (stands in for scene.loc, satellite_rgb_dep.py:163-165): a 2048-pixel tile at a
# ground sampling distance of ~0.3 m spans 2 * 307.2 m; UTM-sized offsets so that float32 / float64 effects are realistic
SCENE_RANGE = 307.2
SCENE_CENTER = (368123.4, 3459876.5, 35.2)


def synthetic_rpc_dict(view: int = 0) -> dict:
    """A well-conditioned RPC00B camera model ("rpcm" dict layout, as in the scene JSON files) of a 2048 x 2048 pushbroom image
    over Jacksonville-like coordinates (the DFC2019 area the reference trains on): near-affine numerators with small second /
    third order terms, denominators close to 1.  Polynomial variables: x = lat (index 2), y = lon (index 1), z = alt (index 3)."""
    tilt = (0.06, -0.11, 0.17)[view % 3]
    num_c, den_c, num_r, den_r = [0.0] * 20, [0.0] * 20, [0.0] * 20, [0.0] * 20
    num_c[0], num_c[1], num_c[2], num_c[3] = 0.002, 1.01, 0.015, tilt
    num_c[4], num_c[7], num_c[8], num_c[5], num_c[11] = 3e-3, -1.5e-3, 8e-4, 2e-3, 4e-4
    den_c[0], den_c[1], den_c[2], den_c[3], den_c[8] = 1.0, 1.2e-3, -8e-4, 5e-4, 2e-4
    num_r[0], num_r[1], num_r[2], num_r[3] = -0.003, 0.02, -0.99, 0.5 * tilt + 0.04
    num_r[4], num_r[7], num_r[8], num_r[6], num_r[15] = -2e-3, 9e-4, 1.1e-3, -1.5e-3, -3e-4
    den_r[0], den_r[1], den_r[2], den_r[3], den_r[7] = 1.0, -9e-4, 1.1e-3, -4e-4, 1.5e-4
    return dict(row_offset=1023.5, col_offset=1023.5, lat_offset=30.3105, lon_offset=-81.6632, alt_offset=5.0,
                row_scale=1024.0, col_scale=1024.0, lat_scale=0.0031, lon_scale=0.0036, alt_scale=120.0,
                row_num=num_r, row_den=den_r, col_num=num_c, col_den=den_c)
