"""One launch of every tile-product kernel at full size (2048 x 2048), for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'dsm_|rpc_' -o gpurun_out/prof_tile python scripts/prof_dsm.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import dsm as PD  # noqa: E402
from brdf_nerf_b200 import georays as PG  # noqa: E402
from brdf_nerf_b200.synth import SCENE_CENTER, SCENE_RANGE, make_tile_rays, synthetic_rpc_dict, tile_surface_depth  # noqa: E402

h = w = 2048
dev = torch.device("cuda:0")
rays = make_tile_rays(h, w, view=0)
depth = tile_surface_depth(rays)
geo = PD.DsmGeoref(SCENE_RANGE, SCENE_CENTER)
rd, dd = rays.to(dev), depth.to(dev)
cloud, pts, bounds = geo._points(rd, dd, True, True)
grid = PD.grid_from_bounds(*bounds.cpu().tolist(), 0.5)
PD.rasterize_cloud(cloud, grid)
PD.normals_from_points(pts.view(h, w, 3))
rpc = PG.RPCModel.from_dict(synthetic_rpc_dict(0))
PG.image_rays(rpc, h, w, -25.0, 95.0, "utm", (436200.0, 3353400.0, 30.0), 400.0, 62.5, 148.0, device=dev, check=False)
torch.cuda.synchronize()
print("ok")
