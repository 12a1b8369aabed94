"""Multi-GPU check (test infrastructure; not collected by pytest, which the driver runs on one GPU).
Ray-sharded tile -> DSM over NCCL: every rank owns a contiguous pixel block of the tile, builds its
part of the cloud, and the ranks all-reduce the raster bounds and the (sum, count) accumulators; every rank ends with the
full raster.  Rank 0 checks it against the unsharded GPU result and the CPU oracle and prints one JSON line.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_check_dsm_sharded.py [H W]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import dsm as PD  # noqa: E402
from brdf_nerf_b200.synth import SCENE_CENTER, make_tile_rays, tile_surface_depth  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 1024)
    scene_range = 307.2 * w / 2048
    rays = make_tile_rays(h, w, view=0)
    depth = tile_surface_depth(rays)
    geo = PD.DsmGeoref(scene_range, SCENE_CENTER)
    per = (h * w + world - 1) // world
    sl = slice(rank * per, min((rank + 1) * per, h * w))
    my_rays, my_depth = rays[sl].to(dev), depth[sl].to(dev)
    for _ in range(2):
        dsm, grid = geo.get_dsm_from_nerf_prediction_sharded(my_rays, my_depth, return_grid=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        dsm, grid = geo.get_dsm_from_nerf_prediction_sharded(my_rays, my_depth, return_grid=True)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank must hold the same raster
    ref = dsm.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(ref.nan_to_num(-1e9), dsm.nan_to_num(-1e9)))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        from oracle import dsm_np as D
        full, g1 = geo.get_dsm_from_nerf_prediction(rays.to(dev), depth.to(dev), return_grid=True)
        want, og = D.dsm_from_nerf_prediction(rays.numpy(), depth.numpy(), scene_range, SCENE_CENTER)
        got = dsm.cpu().numpy()
        out = {"world": world, "tile": [h, w], "raster": [grid.ysize, grid.xsize], "ms_per_call_max_over_ranks": float(t[0]),
               "grid_equals_unsharded": g1 == grid, "grid_equals_oracle": (grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize) == og,
               "all_ranks_same_raster": bool(same.item()),
               "nan_mask_equals_oracle": bool(np.array_equal(np.isnan(got), np.isnan(want))),
               "max_abs_vs_oracle_m": float(np.nanmax(np.abs(got - want))),
               "max_abs_vs_unsharded_m": float((dsm - full).abs().nan_to_num(0).max().item())}
        out["ok"] = bool(out["grid_equals_unsharded"] and out["grid_equals_oracle"] and out["all_ranks_same_raster"]
                         and out["nan_mask_equals_oracle"] and out["max_abs_vs_oracle_m"] <= 1e-3)
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
