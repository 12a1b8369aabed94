"""Throughput of the other BASELINE.json configurations on ONE B200 (the default bench.py line is configs[1]):
  cfg3  RPV111 + analytic normals (second-order backward), cos_irra_on, 1024 rays/step
  cfg4  Hapke b,c,theta / microfacet, 8192 rays/step
  cfg5  inference of RGB + depth + normals + albedo, chunks of 5120 rays (reference eval.py:56-76), rays/s
Prints one JSON line per configuration.
    python scripts/bench_configs.py [--steps K]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.rendering import render_rays  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_cfg(name, cfg, n_rays, steps, **kw):
    dev = torch.device("cuda:0")
    args = named_config(cfg)
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(dev)
    tr = Trainer(model, args, use_graph=True)
    batch = make_rays(n_rays, depth_supervision=False).to(dev)
    ms = timed(lambda: tr.step(batch, **kw), steps)
    print(json.dumps({"config": name, "model": cfg, "rays_per_step": n_rays, "ms_per_step": ms, "rays_per_s": n_rays / ms * 1e3,
                      "kwargs": kw, "cuda_graph": True}), flush=True)
    del tr, model
    torch.cuda.empty_cache()


def infer_cfg(steps):
    dev = torch.device("cuda:0")
    args = named_config("rpv111")
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(dev)
    chunk = 5120
    rays = make_rays(chunk).rays.to(dev)

    def fn():
        with torch.no_grad():
            res, _ = render_rays({"coarse": model}, args, rays, None, mode="test", apply_brdf=True, cos_irra_on=True)
        return res["rgb_coarse"], res["depth_coarse"]

    ms = timed(fn, steps)
    rps = chunk / ms * 1e3
    print(json.dumps({"config": "cfg5 inference RGB+depth+normals+albedo (RPV111, analytic normals)", "chunk_rays": chunk,
                      "ms_per_chunk": ms, "rays_per_s": rps, "tile_2048x2048_seconds_1gpu": 2048 * 2048 / rps,
                      "note": "820 chunks per 2048^2 tile; ray-sharded over 8 GPUs with no collective"}), flush=True)


def tile_cfg(side):
    """configs[4]: one side x side tile (row-major pixel rays) rendered by render_tile, ray-sharded over the ranks of a
    torchrun launch (no collective on the data path; the timing is the slowest rank's, taken on the device)."""
    from brdf_nerf_b200.inference import render_tile, tile_shards
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    args = named_config("rpv111")
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(dev)
    n = side * side
    lo, hi = tile_shards(n, args.chunk, world)[rank]
    rays = make_rays(hi - lo, seed=rank).rays.to(dev)          # this rank's pixels (synthetic rays; only the count matters)

    class _View:                                                # render_tile slices [lo:hi] of the tile's ray list
        shape = (n, 11)

        def __getitem__(self, sl):
            return rays
    keys = ("rgb_coarse", "depth_coarse", "albedo_accu_coarse", "nr_vw_coarse")
    fn = lambda: render_tile({"coarse": model}, _View(), args, rank=rank, world_size=world, keys=keys,
                             apply_brdf=True, cos_irra_on=True)
    fn()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, _ = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"config": f"cfg5 tile {side}x{side} RGB+depth+normals+albedo, ray-sharded over {world} GPU(s)",
                          "n_gpus": world, "rays": n, "ms": ms, "rays_per_s": n / ms * 1e3,
                          "tile_2048x2048_seconds": 2048 * 2048 / (n / ms * 1e3), "chunk": int(args.chunk)}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tile", type=int, default=0, help="render one tile of this side with render_tile (torchrun-aware) and exit")
    o = ap.parse_args()
    if o.tile:
        tile_cfg(o.tile)
        return
    train_cfg("cfg3 BRDF stage RPV111 + analytic normals + cos_irra_on", "rpv111", 1024, o.steps, apply_brdf=True, cos_irra_on=True)
    train_cfg("cfg4 Hapke b,c,theta, 8192 rays", "hapke_bct", 8192, max(3, o.steps // 3), apply_brdf=True, apply_theta=True, cos_irra_on=True)
    train_cfg("cfg4 microfacet, 8192 rays", "microfacet", 8192, max(3, o.steps // 3), apply_brdf=True, cos_irra_on=True)
    infer_cfg(o.steps)


if __name__ == "__main__":
    main()
