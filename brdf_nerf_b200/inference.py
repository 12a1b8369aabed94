"""Callers on the inference side of the hot path: checkpoint hand-off by key prefix and chunked /
ray-sharded tile rendering.

  extract_model_state_dict / load_ckpt   reference eval.py:26-54 (Lightning checkpoints: keys `nerf_coarse.<param>`)
  warm_start                             reference main.py:96-104 (stage-2 BRDF training starts from the stage-1 trunk)
  batched_inference                      reference eval.py:56-76 (render_rays over chunks of `args.chunk` rays, concatenated)
  render_tile_to_dsm                     render_tile -> depth -> DSM without gathering the depth image (eval.py:153-182)
  tile_shards / render_tile              ray-sharded full-tile inference: every rank renders whole chunks, no collective
                                         on the data path; an optional all_gather assembles the image

The chunk size is part of the result: the reference clamps the guided samples with the FIRST ray of each chunk
(`near[0,0]`, `far[0,0]`, `rays_d[0,2]/sun_d[0,2]`, rendering.py:133,144,247-248), so shards are cut on
chunk boundaries and a sharded render equals the single-process one chunk for chunk.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from .rendering import render_rays


def extract_model_state_dict(ckpt, model_name="model", prefixes_to_ignore: Iterable[str] = (), drop_len=-1):
    """eval.py:26-47.  `ckpt` is a path or an already loaded checkpoint / state dict."""
    checkpoint = torch.load(ckpt, map_location=torch.device("cpu")) if isinstance(ckpt, (str, bytes)) else ckpt
    if "state_dict" in checkpoint:                     # pytorch-lightning checkpoint
        checkpoint = checkpoint["state_dict"]
    out, loaded = {}, []
    for k, v in checkpoint.items():
        if not k.startswith(model_name):
            continue
        if drop_len < 0:
            drop_len = len(model_name)
        k = k[drop_len + 1:]
        if any(k.startswith(p) for p in prefixes_to_ignore):
            continue
        loaded.append(k)
        out[k] = v
    return out, " ".join(loaded)


def load_ckpt(model, ckpt, model_name="model", prefixes_to_ignore: Iterable[str] = (), drop_len=-1) -> List[str]:
    """eval.py:49-54: overwrite the parameters whose checkpoint key starts with `model_name`; returns the loaded keys.
    The parameters stay views of the module's flat buffer (load_state_dict copies in place)."""
    model_dict = model.state_dict()
    part, loaded = extract_model_state_dict(ckpt, model_name, prefixes_to_ignore, drop_len)
    model_dict.update(part)
    model.load_state_dict(model_dict)
    return loaded.split()


def warm_start(model, ckpt, args) -> List[str]:
    """main.py:96-104: a stage-2 (BRDF) model takes the trunk, the density head, the feature layer and — unless it is a
    Hapke model (`args.b`) — the colour head of a stage-1 checkpoint; every BRDF head keeps its fresh initialisation."""
    drop = len("nerf_coarse")
    loaded = []
    for part in ("fc_net", "sigma_from_xyz", "feats_from_xyz"):
        loaded += load_ckpt(model, ckpt, model_name=f"nerf_coarse.{part}", drop_len=drop)
    if args.b != True:                                 # noqa: E712 (as the reference)
        loaded += load_ckpt(model, ckpt, model_name="nerf_coarse.rgb_from_xyzdir", drop_len=drop)
    return loaded


@torch.no_grad()
def batched_inference(models, rays, ts, args, apply_brdf=False, cos_irra_on=False, **render_kw):
    """eval.py:56-76: render_rays over consecutive chunks of `args.chunk` rays; per-key concatenation."""
    chunk = int(args.chunk)
    results = defaultdict(list)
    brdf_type = None
    for i in range(0, rays.shape[0], chunk):
        res, brdf_type = render_rays(models, args, rays[i:i + chunk], None if ts is None else ts[i:i + chunk],
                                     apply_brdf=apply_brdf, cos_irra_on=cos_irra_on, **render_kw)
        for k, v in res.items():
            results[k].append(v)
    out = {k: (None if v[0] is None else torch.cat(v, 0)) for k, v in results.items()}
    return out, brdf_type


def tile_shards(n_rays: int, chunk: int, world_size: int) -> List[Tuple[int, int]]:
    """[start, end) ray ranges, one per rank: contiguous runs of whole chunks, as even as the chunk count allows."""
    n_chunks = -(-n_rays // chunk)
    base, extra = divmod(n_chunks, world_size)
    out, c0 = [], 0
    for r in range(world_size):
        c1 = c0 + base + (1 if r < extra else 0)
        out.append((min(c0 * chunk, n_rays), min(c1 * chunk, n_rays)))
        c0 = c1
    return out


@torch.no_grad()
def render_tile(models, rays, args, rank=0, world_size=1, keys: Optional[Iterable[str]] = None, gather=False,
                group=None, **kw) -> Tuple[Dict[str, torch.Tensor], str]:
    """Full-tile inference sharded by rays (BASELINE configs[4]): this rank renders its run of chunks of the row-major
    pixel list `rays` (H*W, 11).  `keys` keeps only those result keys (e.g. rgb_coarse, depth_coarse).  With `gather`
    the per-rank results are all-gathered into full (H*W, ...) tensors on every rank — the only collective, and off the
    compute path; otherwise each rank returns its own slice."""
    lo, hi = tile_shards(rays.shape[0], int(args.chunk), world_size)[rank]
    if hi > lo:
        res, brdf_type = batched_inference(models, rays[lo:hi], None, args, **kw)
    elif gather and world_size > 1 and rays.shape[0] > 0:
        # fewer chunks than ranks: this rank still has to join every all_gather below with a zero-length part of the right
        # key set / trailing shape / dtype (and report the same brdf_type) — one dummy ray tells it all three
        res, brdf_type = batched_inference(models, rays[0:1], None, args, **kw)
        res = {k: v[0:0] for k, v in res.items()}
    else:
        res, brdf_type = {}, None
    if keys is not None:
        res = {k: res[k] for k in keys if k in res}
    if not gather or world_size == 1:
        return res, brdf_type
    sizes = [b - a for a, b in tile_shards(rays.shape[0], int(args.chunk), world_size)]
    full = {}
    for k in sorted(res):
        v = res[k].contiguous()
        parts = [torch.empty((s,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for s in sizes]
        torch.distributed.all_gather(parts, v, group=group)
        full[k] = torch.cat(parts, 0)
    return full, brdf_type


@torch.no_grad()
def render_tile_to_dsm(models, rays, args, georef, rank=0, world_size=1, group=None, roi_txt=None, **kw):
    """The evaluation product of a tile (reference eval.py:153-182: render -> depth -> get_dsm_from_nerf_prediction) with
    the rays sharded over ranks: each rank renders its run of chunks (`render_tile`, no gather), turns ITS depths into
    points and accumulators (`brdf_nerf_b200.dsm`), and the ranks all-reduce the raster bounds and the accumulators — a few
    MB instead of the tile's depth image.  Returns (dsm (ysize, xsize, 1) float32 on every rank, DsmGrid, this rank's result
    dict).  `georef` is a `dsm.DsmGeoref`."""
    res, _ = render_tile(models, rays, args, rank=rank, world_size=world_size, gather=False, **kw)
    lo, hi = tile_shards(rays.shape[0], int(args.chunk), world_size)[rank]
    depth = res["depth_coarse"] if hi > lo else rays.new_zeros(0)
    if world_size == 1:
        dsm, grid = georef.get_dsm_from_nerf_prediction(rays[lo:hi], depth, roi_txt=roi_txt, return_grid=True)
    else:
        dsm, grid = georef.get_dsm_from_nerf_prediction_sharded(rays[lo:hi], depth, group=group, roi_txt=roi_txt, return_grid=True)
    return dsm, grid, res
