"""ORACLE (test infrastructure only).  Generates tests/golden/dsm_tile.npz: a small synthetic tile (64 x 48 rays), its
"rendered" depth, and
  * ref_east / ref_north / ref_alt : outputs of the LIVE reference's get_latlonalt_from_nerf_prediction
    (datasets/satellite_rgb_dep.py:601-634, cs='utm'), float64;
  * ref_normals                    : the LIVE reference's calc_normal_from_pts3d (sat_utils.py:16-50) on the float32 point
    image, i.e. calc_normal_from_depth_v2 (satellite_rgb_dep.py:578-585);
  * grid                           : (xoff, yoff, resolution, xsize, ysize) of satellite_rgb_dep.py:665-671;
  * restated_raster / restated_count : the C restatement of plyflatten (oracle/plyflatten_restated.c).  plyflatten itself
    is an absent third-party package: these two arrays are NOT reference outputs (parity unpinned for the rasteriser).
Run here (the reference tree is not available on the GPU box):   python -m oracle.make_golden_dsm
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from brdf_nerf_b200.synth import SCENE_CENTER, make_tile_rays, tile_surface_depth   # noqa: E402
from oracle import dsm_np as D                                                       # noqa: E402
from oracle import ref_harness as RH                                                 # noqa: E402

H, W = 48, 64
SCENE_RANGE_SMALL = 12.0        # 64 pixels over 24 m: ~0.38 m ground sampling distance, as in the full-size tile


def main():
    rays = make_tile_rays(H, W, view=1)
    depth = tile_surface_depth(rays)
    e, n, a = RH.ref_latlonalt(rays, depth, SCENE_RANGE_SMALL, SCENE_CENTER)
    pts = torch.from_numpy(np.vstack([e, n, a]).T).type(torch.FloatTensor)
    normals = RH.ref_normals_from_pts3d(pts.reshape(H, W, 3)).reshape(-1, 3).numpy()
    grid = D.dsm_grid(e, n, 0.5)
    raster, count = D.plyflatten(np.vstack([e, n, a]).T, *grid, radius=1, sigma=float("inf"), return_count=True)
    path = os.path.join(ROOT, "tests", "golden", "dsm_tile.npz")
    np.savez_compressed(path, rays=rays.numpy(), depth=depth.numpy(), scene_range=np.float64(SCENE_RANGE_SMALL),
                        center=np.asarray(SCENE_CENTER, np.float64), hw=np.asarray([H, W]), ref_east=e, ref_north=n, ref_alt=a,
                        ref_normals=normals, grid=np.asarray(grid, np.float64), restated_raster=raster, restated_count=count)
    print(f"dsm_tile: grid {grid}, {np.isnan(raster).mean() * 100:.1f}% empty cells, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
