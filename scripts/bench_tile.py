"""BASELINE configs[4] end to end: one side x side tile (row-major pixel rays of one view) rendered with RGB + depth + normals +
albedo by the RPV111 / analytic-normal model, ray-sharded over the ranks of a torchrun launch (whole chunks per rank, no
collective on the render path), then depth -> DSM with the ranks all-reducing raster bounds and accumulators
(`inference.render_tile_to_dsm`).  Times are the slowest rank's, taken on the device.  Rank 0 prints one JSON line.
    python scripts/bench_tile.py [--tile 2048]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 scripts/bench_tile.py --tile 2048"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.dsm import DsmGeoref  # noqa: E402
from brdf_nerf_b200.inference import render_tile_to_dsm, tile_shards  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import SCENE_CENTER, SCENE_RANGE, make_tile_rays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=2048)
    o = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    args = named_config("rpv111")
    torch.manual_seed(0)
    models = {"coarse": load_model(args, precision="bf16").to(dev)}
    side = o.tile
    n = side * side
    rays = make_tile_rays(side, side, view=0).to(dev)
    geo = DsmGeoref(SCENE_RANGE * side / 2048, SCENE_CENTER)
    kw = dict(apply_brdf=True, cos_irra_on=True)
    # warm-up on a few chunks per rank (kernel / workspace caches, NCCL communicator)
    warm = rays[:int(args.chunk) * 2 * world]
    render_tile_to_dsm(models, warm, args, geo, rank=rank, world_size=world, **kw)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    dsm, grid, res = render_tile_to_dsm(models, rays, args, geo, rank=rank, world_size=world, **kw)
    e[1].record()
    torch.cuda.synchronize()
    lo, hi = tile_shards(n, int(args.chunk), world)[rank]
    # DSM part alone (the depths are there now)
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if world > 1:
        torch.distributed.barrier()
    e2[0].record()
    if world > 1:
        geo.get_dsm_from_nerf_prediction_sharded(rays[lo:hi], res["depth_coarse"])
    else:
        geo.get_dsm_from_nerf_prediction(rays[lo:hi], res["depth_coarse"])
    e2[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]), e2[0].elapsed_time(e2[1])], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        ms, ms_dsm = float(t[0]), float(t[1])
        keys = sorted(k for k in res if k.startswith(("rgb", "depth", "albedo_accu", "nr_vw", "normal")))
        print(json.dumps({"config": f"configs[4]: {side}x{side} tile, RGB + depth + normals + albedo (RPV111, analytic normals) -> DSM, "
                                    f"ray-sharded over {world} GPU(s)",
                          "n_gpus": world, "rays": n, "chunk": int(args.chunk), "ms_tile_to_dsm": ms, "ms_dsm_part": ms_dsm,
                          "rays_per_s": n / ms * 1e3, "raster": [grid.ysize, grid.xsize],
                          "dsm_cells_with_data": int((~torch.isnan(dsm)).sum().item()), "result_keys": keys,
                          "reference_cpu": "258-930 rays/s on the host cores (SURVEY 8d, bench.py cpu_baseline): hours per tile"}),
              flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
