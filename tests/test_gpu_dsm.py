"""GPU parity of the tile-inference -> DSM kernels (csrc/dsm.cu, through the C ABI) against the oracle and the golden
fixture (live-reference cloud and normals).  Bars: float64 point cloud, float32 point image, bounds, raster grid and the
count image are bit/integer-exact; raster heights <= 1e-3 m (float32 running mean in the reference's rasteriser vs float64
sums here; measured ~1e-5); normals <= 1e-5."""
import math
import os

import numpy as np
import pytest
import torch

from brdf_nerf_b200 import dsm as PD
from brdf_nerf_b200.synth import SCENE_CENTER, SCENE_RANGE, make_tile_rays, tile_surface_depth
from oracle import dsm_np as D

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsm_tile.npz")
RASTER_TOL = 1e-3


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _geo(g):
    return PD.DsmGeoref(float(g["scene_range"]), g["center"])


def test_cloud_bit_exact_vs_reference_golden(cuda, g):
    e, n, a = _geo(g).get_latlonalt_from_nerf_prediction(torch.from_numpy(g["rays"]).to(cuda), torch.from_numpy(g["depth"]).to(cuda))
    assert e.dtype == torch.float64
    assert np.array_equal(e.cpu().numpy(), g["ref_east"]) and np.array_equal(n.cpu().numpy(), g["ref_north"])
    assert np.array_equal(a.cpu().numpy(), g["ref_alt"])


@pytest.mark.parametrize("n", [1, 31, 256, 257, 4099])
def test_cloud_points_bounds_vs_oracle_ragged(cuda, n):
    gen = torch.Generator().manual_seed(n)
    rays = torch.rand(n, 11, generator=gen) * 2 - 1
    depth = torch.rand(n, 1, generator=gen) * 1.5                       # (n, 1): the shape render_rays' callers pass
    geo = PD.DsmGeoref(SCENE_RANGE, SCENE_CENTER)
    cloud, pts, bounds = geo._points(rays.to(cuda), depth.to(cuda), True, True)
    e, no, a = D.latlonalt_from_nerf_prediction(rays.numpy(), depth.numpy().reshape(-1), SCENE_RANGE, SCENE_CENTER)
    want = np.vstack([e, no, a]).T
    assert np.array_equal(cloud.cpu().numpy(), want)
    assert np.array_equal(pts.cpu().numpy(), want.astype(np.float32))
    assert bounds.cpu().tolist() == [e.min(), e.max(), no.min(), no.max()]


def test_dsm_vs_restated_rasteriser_golden(cuda, g):
    dsm, grid = _geo(g).get_dsm_from_nerf_prediction(torch.from_numpy(g["rays"]).to(cuda), torch.from_numpy(g["depth"]).to(cuda),
                                                     return_grid=True)
    assert [grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize] == list(g["grid"])
    want = g["restated_raster"]
    got = dsm.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want)) <= RASTER_TOL


@pytest.mark.parametrize("sigma,radius", [(float("inf"), 1), (float("inf"), 2), (0.35, 1), (1.0, 2)])
def test_rasterize_counts_and_heights_vs_oracle(cuda, g, sigma, radius):
    cloud = np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T
    grid = PD.DsmGrid(float(g["grid"][0]) + 3.0, float(g["grid"][1]) - 2.5, 0.5, 30, 31)     # smaller than the cloud
    want, wc = D.plyflatten(cloud, grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize, radius=radius, sigma=sigma,
                            return_count=True)
    got, gc = PD.rasterize_cloud(torch.from_numpy(cloud).to(cuda), grid, radius=radius, sigma=sigma, return_count=True)
    got, gc = got.cpu().numpy(), gc.cpu().numpy()
    if math.isinf(sigma):
        assert np.array_equal(gc, wc)                                   # integer counts: exact
    else:
        assert np.allclose(gc, wc, rtol=2e-5, atol=1e-6)                # float32 weight sums in a different order
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want)) <= RASTER_TOL


def test_roi_grid_and_all_points_outside(cuda, g):
    geo = _geo(g)
    rays, depth = torch.from_numpy(g["rays"]).to(cuda), torch.from_numpy(g["depth"]).to(cuda)
    roi = [float(g["grid"][0]) + 2.0, float(g["grid"][1]) - 20.0, 24, 0.5]       # (xoff, yoff of the LOWER edge, size, res)
    dsm, grid = geo.get_dsm_from_nerf_prediction(rays, depth, roi_txt=roi, return_grid=True)
    want, og = D.dsm_from_nerf_prediction(g["rays"], g["depth"], float(g["scene_range"]), g["center"], roi=roi)
    assert (grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize) == og
    got = dsm.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.nanmax(np.abs(got - want)) <= RASTER_TOL
    far = geo.get_dsm_from_nerf_prediction(rays, depth, roi_txt=[0.0, 0.0, 16, 0.5])
    assert torch.isnan(far).all()                                        # nothing lands: all no-data, no crash


def test_accumulate_in_two_halves_equals_one_pass(cuda, g):
    """The sharded-tile rule on one GPU: accumulating two pixel blocks into one workspace == rasterising the whole cloud."""
    cloud = torch.from_numpy(np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T).to(cuda)
    grid = PD.DsmGrid(*[float(v) for v in g["grid"][:3]], int(g["grid"][3]), int(g["grid"][4]))
    half = cloud.shape[0] // 2
    ws = PD.accumulate_cloud(cloud[:half], grid)
    PD.accumulate_cloud(cloud[half:], grid, workspace=ws)
    two, c2 = PD.finalize_raster(ws, grid, return_count=True)
    one, c1 = PD.rasterize_cloud(cloud, grid, return_count=True)
    assert torch.equal(c1, c2)
    assert torch.equal(torch.isnan(one), torch.isnan(two))
    assert (one - two).abs().nan_to_num(0).max().item() <= 1e-5
    assert np.array_equal(c1.cpu().numpy(), g["restated_count"])


def test_normals_vs_reference_golden(cuda, g):
    h, w = (int(v) for v in g["hw"])
    nr, valid = _geo(g).calc_normal_from_depth_v2(torch.from_numpy(g["rays"]).to(cuda), torch.from_numpy(g["depth"]).to(cuda), h, w)
    got = nr.cpu().numpy()
    assert got.shape == (h * w, 3) and valid.shape == (h * w,) and bool((valid == 1).all())
    assert np.abs(got - g["ref_normals"]).max() <= 1e-5
    assert not got.reshape(h, w, 3)[0].any() and not got.reshape(h, w, 3)[:, 0].any()


@pytest.mark.parametrize("hw", [(3, 3), (2, 7), (1, 1), (40, 33)])
def test_normals_small_and_degenerate_images(cuda, hw):
    h, w = hw
    gen = torch.Generator().manual_seed(h * 100 + w)
    pts = torch.rand(h, w, 3, generator=gen)
    pts[..., 0] += torch.arange(w)[None, :] * 0.5
    pts[..., 1] -= torch.arange(h)[:, None] * 0.5
    if h >= 3 and w >= 3:
        pts[1, 1] = pts[1, 2]                                            # a zero-length neighbour difference
    want = D.calc_normal_from_pts3d(pts) if (h >= 3 and w >= 3) else torch.zeros_like(pts)
    got = PD.normals_from_points(pts.to(cuda)).cpu()
    assert (got - want).abs().max().item() <= 1e-5


def test_full_tile_2048_properties_and_full_size_oracle(cuda):
    """BASELINE config 5 size: the 2048 x 2048 tile (4.19 M rays).  Size-independent properties (count mass, weighted-height
    mass, determinism of the count image) and, because the C restatement rasterises 4 M points in about a second, the
    direct comparison at full size too."""
    h = w = 2048
    rays = make_tile_rays(h, w, view=0)
    depth = tile_surface_depth(rays)
    geo = PD.DsmGeoref(SCENE_RANGE, SCENE_CENTER)
    rd, dd = rays.to(cuda), depth.to(cuda)
    cloud, _, bounds = geo._points(rd, dd, False, True)
    grid = PD.grid_from_bounds(*bounds.cpu().tolist(), 0.5)
    dsm, cnt = PD.rasterize_cloud(cloud, grid, return_count=True)
    # count mass: every point adds 1 to each in-raster cell of its 3 x 3 window
    i = torch.floor((cloud[:, 0] - grid.xoff) / grid.resolution).long()
    j = torch.floor((-cloud[:, 1] - (-grid.yoff)) / grid.resolution).long()
    nx = (torch.clamp(i + 1, max=grid.xsize - 1) - torch.clamp(i - 1, min=0) + 1).clamp(min=0)
    ny = (torch.clamp(j + 1, max=grid.ysize - 1) - torch.clamp(j - 1, min=0) + 1).clamp(min=0)
    assert int(cnt.double().sum().item()) == int((nx * ny).sum().item())
    # weighted-height mass: sum(raster * count) == sum over points of alt * (cells it reached)
    mass = (dsm[..., 0].double().nan_to_num(0) * cnt.double()).sum().item()
    want_mass = (cloud[:, 2].float().double() * (nx * ny).double()).sum().item()
    assert abs(mass - want_mass) <= 1e-6 * abs(want_mass)
    dsm2, cnt2 = PD.rasterize_cloud(cloud, grid, return_count=True)
    assert torch.equal(cnt, cnt2) and (dsm - dsm2).abs().nan_to_num(0).max().item() <= 1e-5
    # full-size oracle
    e, n, a = D.latlonalt_from_nerf_prediction(rays.numpy(), depth.numpy(), SCENE_RANGE, SCENE_CENTER)
    assert np.array_equal(cloud.cpu().numpy(), np.vstack([e, n, a]).T)
    og = D.dsm_grid(e, n, 0.5)
    assert (grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize) == og
    want, wc = D.plyflatten(np.vstack([e, n, a]).T, *og, return_count=True)
    assert np.array_equal(cnt.cpu().numpy(), wc)
    got = dsm.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.nanmax(np.abs(got - want)) <= RASTER_TOL


def test_render_tile_to_dsm_end_to_end(cuda):
    """render_tile -> depth -> DSM through the real renderer (random-init weights): the DSM of the rendered depths equals the
    oracle's DSM of the same depths; an empty ray shard contributes neutral bounds and zero accumulators."""
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.inference import render_tile_to_dsm
    from brdf_nerf_b200.models import load_model
    args = named_config("lambertian", chunk=1024)
    h, w = 40, 48
    rays = make_tile_rays(h, w, view=0).to(cuda)
    torch.manual_seed(0)
    models = {"coarse": load_model(args, precision="bf16").to(cuda)}
    geo = PD.DsmGeoref(10.0, SCENE_CENTER)
    torch.manual_seed(5)
    dsm, grid, res = render_tile_to_dsm(models, rays, args, geo)
    depth = res["depth_coarse"]
    assert depth.shape == (h * w,) and dsm.shape == (grid.ysize, grid.xsize, 1)
    want, og = D.dsm_from_nerf_prediction(rays.cpu().numpy(), depth.cpu().numpy(), 10.0, SCENE_CENTER)
    assert (grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize) == og
    got = dsm.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.nanmax(np.abs(got - want)) <= RASTER_TOL
    # empty shard
    cloud, pts, bounds = geo._points(rays[:0], depth[:0], True, True)
    assert cloud.shape == (0, 3) and pts.shape == (0, 3) and bounds.tolist() == [math.inf, -math.inf, math.inf, -math.inf]
    ws = PD.accumulate_cloud(cloud, grid)
    assert not ws.any()
    full_ws = PD.accumulate_cloud(geo._points(rays, depth, False, False)[0], grid)
    before = full_ws.clone()
    assert PD.accumulate_cloud(cloud, grid, workspace=full_ws) is full_ws and torch.equal(before, full_ws)


def test_cloud_ecef_scene_vs_oracle(cuda):
    """cs == 'ecef': geocentric scene coordinates -> ecef_to_latlon_custom -> UTM of the first point's zone.  CUDA libm vs
    numpy differ in the last bits of the trigonometry: float64 results within 1e-6 m (measured ~1e-9)."""
    from oracle import georays_np as G
    n = 3000
    gen = torch.Generator().manual_seed(4)
    rays = torch.rand(n, 11, generator=gen) * 2 - 1
    depth = torch.rand(n, generator=gen)
    cx, cy, cz = (float(v) for v in G.latlon_to_ecef_custom(np.float64(30.31), np.float64(-81.66), np.float64(20.0)))
    geo = PD.DsmGeoref(300.0, (cx, cy, cz), cs="ecef")
    e, no, a = geo.get_latlonalt_from_nerf_prediction(rays.to(cuda), depth.to(cuda))
    we, wn, wa = D.latlonalt_from_nerf_prediction(rays.numpy(), depth.numpy(), 300.0, (cx, cy, cz), cs="ecef")
    assert geo._first_point_zone(rays.to(cuda), depth.to(cuda)) == 17
    assert np.abs(e.cpu().numpy() - we).max() < 1e-6 and np.abs(no.cpu().numpy() - wn).max() < 1e-6
    assert np.abs(a.cpu().numpy() - wa).max() < 1e-6
    assert 4.3e5 < we.mean() < 4.4e5 and 3.35e6 < wn.mean() < 3.36e6          # Jacksonville, UTM 17N
