"""Tile inference -> DSM micro-benchmark on one B200 (BASELINE config 5 size: 2048 x 2048 rays):
per-kernel time and algorithmic GB/s of bn_dsm_points / bn_dsm_rasterize / bn_dsm_normals_from_points and the end-to-end
depth -> DSM call (including its one 32-byte host round trip).  The reference's CPU path is timed beside it by bench.py
(`tile_products` leg), the only place outside tests/ that may execute the oracle.
    python scripts/bench_dsm.py [H W]       prints one JSON line"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import dsm as PD  # noqa: E402
from brdf_nerf_b200.synth import SCENE_CENTER, SCENE_RANGE, make_tile_rays, tile_surface_depth  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run(h=2048, w=2048):
    dev = torch.device("cuda:0")
    n = h * w
    rays = make_tile_rays(h, w, view=0)
    depth = tile_surface_depth(rays)
    geo = PD.DsmGeoref(SCENE_RANGE, SCENE_CENTER)
    rd, dd = rays.to(dev), depth.to(dev)
    peak = hbm_peak()
    cloud, pts, bounds = geo._points(rd, dd, True, True)
    grid = PD.grid_from_bounds(*bounds.cpu().tolist(), 0.5)
    cells = grid.xsize * grid.ysize
    apron = (grid.xsize + 2) * (grid.ysize + 2)
    out = {"rays": n, "raster": [grid.ysize, grid.xsize], "hbm_peak_gbs": peak}
    # algorithmic bytes: points = ray record 44 + depth 4 in, 24 (f64 point) + 12 (f32 point) out;
    # rasterise = 24 in per point + 12 B accumulator per apron cell zeroed, read once, + 4 B raster out per cell
    # (the two atomics per point resolve in L2); normals = 12 in + 12 out per pixel
    legs = {
        "dsm_points (cloud f64 + points f32 + bounds)": (lambda: geo._points(rd, dd, True, True), n * (48 + 36)),
        "dsm_rasterize (memset + scatter + box finalize)": (lambda: PD.rasterize_cloud(cloud, grid), n * 24 + apron * 24 + cells * 4),
        "dsm_normals_from_points": (lambda: PD.normals_from_points(pts.view(h, w, 3)), n * 24),
    }
    for name, (fn, nbytes) in legs.items():
        us = timeit(fn)
        out[name] = {"us": us, "GB/s": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak, "bytes": nbytes}
    us = timeit(lambda: geo.get_dsm_from_nerf_prediction(rd, dd), iters=10)
    out["depth -> DSM end to end (incl. 32 B host round trip)"] = {"us": us, "Mrays/s": n / us}
    return out


if __name__ == "__main__":
    hw = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 2048)
    print(json.dumps(run(*hw)))
