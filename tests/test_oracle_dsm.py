"""CPU: the DSM oracle (oracle/dsm_np.py + oracle/plyflatten_restated.c) against the committed golden fixture
(tests/golden/dsm_tile.npz: live-reference point cloud and normals), the C restatement of the rasteriser against its
pure-Python twin, the host-side grid logic of the product, the scatter/box-filter factorisation the CUDA kernels rely on,
the two-rank accumulator all-reduce rule (gloo), and argument validation through the C ABI (no GPU needed)."""
import ctypes
import math
import os
import socket

import numpy as np
import pytest
import torch

from brdf_nerf_b200 import dsm as PD
from oracle import dsm_np as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsm_tile.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def test_oracle_cloud_matches_reference_golden_bit_exact(g):
    e, n, a = D.latlonalt_from_nerf_prediction(g["rays"], g["depth"], float(g["scene_range"]), g["center"])
    assert np.array_equal(e, g["ref_east"]) and np.array_equal(n, g["ref_north"]) and np.array_equal(a, g["ref_alt"])


def test_oracle_normals_match_reference_golden_bit_exact(g):
    h, w = (int(v) for v in g["hw"])
    nr = D.normal_from_depth_v2(g["rays"], g["depth"], h, w, float(g["scene_range"]), g["center"]).numpy()
    assert np.array_equal(nr, g["ref_normals"])
    inner = nr.reshape(h, w, 3)[1:-1, 1:-1]
    assert np.allclose(np.linalg.norm(inner, axis=-1), 1.0, atol=1e-5) and (np.abs(inner[..., 2]) > 0.5).all()   # (sign: rows of the synthetic tile run north)
    assert not nr.reshape(h, w, 3)[0].any() and not nr.reshape(h, w, 3)[:, -1].any()        # border stays zero


def test_grid_and_restated_raster_match_golden(g):
    grid = D.dsm_grid(g["ref_east"], g["ref_north"], 0.5)
    assert list(grid) == list(g["grid"])
    cloud = np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T
    raster, cnt = D.plyflatten(cloud, *grid, return_count=True)
    assert raster.shape == (grid[4], grid[3], 1) and raster.dtype == np.float32
    assert np.array_equal(cnt, g["restated_count"]) and np.array_equal(raster, g["restated_raster"], equal_nan=True)
    assert np.array_equal(np.isnan(raster[..., 0]), cnt == 0)


@pytest.mark.parametrize("sigma", [float("inf"), 0.35])
def test_c_restatement_equals_python_loop(sigma):
    rng = np.random.default_rng(5)
    cloud = np.stack([rng.uniform(100.0, 108.0, 400), rng.uniform(-50.0, -44.0, 400), rng.normal(30.0, 5.0, 400)], 1)
    grid = (99.5, -43.5, 0.5, 15, 12)            # deliberately smaller than the cloud: points fall outside
    rc, cc = D.plyflatten(cloud, *grid, radius=1, sigma=sigma, return_count=True)
    rp, cp = D.plyflatten_py(cloud, *grid, radius=1, sigma=sigma)
    if math.isinf(sigma):
        assert np.array_equal(cc, cp) and np.array_equal(rc, rp, equal_nan=True)
    else:                                          # libm expf vs numpy exp: last-bit differences
        assert np.allclose(cc, cp, rtol=1e-5, atol=1e-6) and np.allclose(rc, rp, rtol=1e-5, atol=1e-4, equal_nan=True)


def test_empty_and_ragged_inputs():
    grid = (0.0, 4.0, 0.5, 8, 8)
    r, c = D.plyflatten(np.zeros((0, 3)), *grid, return_count=True)          # empty cloud: all NaN
    assert np.isnan(r).all() and not c.any()
    r, c = D.plyflatten(np.array([[100.0, 100.0, 5.0]]), *grid, return_count=True)      # far outside: nothing lands
    assert np.isnan(r).all()
    r, c = D.plyflatten(np.array([[-0.1, 4.1, 7.0]]), *grid, return_count=True)          # own cell outside, neighbour inside
    assert c[0, 0] == 1 and r[0, 0, 0] == 7.0 and c.sum() == 1


def test_product_grid_logic_equals_oracle():
    rng = np.random.default_rng(11)
    for _ in range(200):
        x0, y0 = rng.uniform(-1e6, 1e6), rng.uniform(-4e6, 4e6)
        e = x0 + rng.uniform(0, 700, 50)
        n = y0 + rng.uniform(0, 700, 50)
        want = D.dsm_grid(e, n, 0.5)
        got = PD.grid_from_bounds(e.min(), e.max(), n.min(), n.max(), 0.5)
        assert (got.xoff, got.yoff, got.resolution, got.xsize, got.ysize) == want
    roi = [368000.0, 3459000.0, 256, 0.5]
    want = D.dsm_grid(None, None, roi=roi)
    got = PD.grid_from_roi(roi)
    assert (got.xoff, got.yoff, got.resolution, got.xsize, got.ysize) == want


def _emulate_cuda_box_path(cloud, grid, radius=1):
    """numpy statement of what dsm_scatter_kernel + dsm_finalize_kernel do for sigma == inf: histogram the points' own cells
    on an apron-extended grid, then box-sum (2 radius + 1)^2."""
    xoff, yoff, res, xs, ys = grid
    i = np.floor((cloud[:, 0] - xoff) / res).astype(np.int64) + radius
    j = np.floor((-cloud[:, 1] - (-yoff)) / res).astype(np.int64) + radius
    gw, gh = xs + 2 * radius, ys + 2 * radius
    ok = (i >= 0) & (j >= 0) & (i < gw) & (j < gh)
    s = np.zeros((gh, gw)); c = np.zeros((gh, gw), np.float32)
    np.add.at(s, (j[ok], i[ok]), cloud[ok, 2].astype(np.float32).astype(np.float64))
    np.add.at(c, (j[ok], i[ok]), 1.0)
    S = sum(s[dj:dj + ys, di:di + xs] for dj in range(2 * radius + 1) for di in range(2 * radius + 1))
    Cn = sum(c[dj:dj + ys, di:di + xs] for dj in range(2 * radius + 1) for di in range(2 * radius + 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        out = np.where(Cn > 0, S / Cn, np.nan).astype(np.float32)
    return out, Cn.astype(np.float32), (s, c)


@pytest.mark.parametrize("grid_kind", ["from_bounds", "roi_smaller_than_cloud"])
def test_scatter_then_box_filter_equals_rasteriser(g, grid_kind):
    """The factorisation behind the CUDA kernels: count image identical, heights within the float32 running-mean error."""
    cloud = np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T
    grid = tuple(g["grid"][:3]) + (int(g["grid"][3]), int(g["grid"][4]))
    if grid_kind == "roi_smaller_than_cloud":
        grid = (grid[0] + 3.0, grid[1] - 2.5, 0.5, 30, 31)
    want, wc = D.plyflatten(cloud, *grid, return_count=True)
    got, gc, _ = _emulate_cuda_box_path(cloud, grid)
    assert np.array_equal(gc, wc)
    assert np.array_equal(np.isnan(got), np.isnan(want[..., 0]))
    assert np.nanmax(np.abs(got - want[..., 0])) <= 1e-4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(GOLDEN))
    cloud = np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T
    per = cloud.shape[0] // world
    mine = cloud[rank * per:(rank + 1) * per]                       # contiguous pixel block of this rank
    bounds = torch.tensor([mine[:, 0].min(), mine[:, 0].max(), mine[:, 1].min(), mine[:, 1].max()], dtype=torch.float64)
    b = PD.reduce_bounds(bounds).tolist()                           # product code: the MAX all-reduce of the signed bounds
    grid = PD.grid_from_bounds(*b, 0.5)
    _, _, (s, c) = _emulate_cuda_box_path(mine, (grid.xoff, grid.yoff, grid.resolution, grid.xsize, grid.ysize))
    ws = torch.cat([torch.from_numpy(s.reshape(-1)).view(torch.uint8), torch.from_numpy(c.reshape(-1)).view(torch.uint8)])
    sums, counts = PD.workspace_views(ws)                           # product code: the two all-reduce operands
    dist.all_reduce(sums)
    dist.all_reduce(counts)
    if rank == 0:
        q.put((b, sums.numpy().copy(), counts.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_accumulators_allreduce_to_the_full_tile(g):
    """world_size 2 over gloo: bounds all-reduce + accumulator all-reduce of two pixel blocks == the whole tile."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    b, sums, counts = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    cloud = np.vstack([g["ref_east"], g["ref_north"], g["ref_alt"]]).T
    assert b == [cloud[:, 0].min(), cloud[:, 0].max(), cloud[:, 1].min(), cloud[:, 1].max()]
    grid = tuple(g["grid"][:3]) + (int(g["grid"][3]), int(g["grid"][4]))
    _, _, (s, c) = _emulate_cuda_box_path(cloud, grid)
    assert np.array_equal(counts, c.reshape(-1)) and np.allclose(sums, s.reshape(-1), rtol=1e-12, atol=0)


def test_dsm_abi_argument_validation():
    """Through the C ABI without a GPU: bad arguments are refused before any CUDA call."""
    from brdf_nerf_b200 import _lib
    lib = _lib.load()
    assert lib.bn_dsm_points(None, 11, None, 10, 1.0, 0.0, 0.0, 0.0, 1, 0, None, None, None, None, None) == -1
    assert b"null pointer" in lib.bn_last_error()
    assert lib.bn_dsm_workspace_bytes(100, 50, 1, float("inf")) == 102 * 52 * 12
    assert lib.bn_dsm_workspace_bytes(100, 50, 1, 0.5) == 100 * 50 * 12
    assert lib.bn_dsm_workspace_bytes(0, 50, 1, 0.5) == 0
    assert lib.bn_dsm_rasterize(None, 3, 2, 10, 0.0, 0.0, 0.5, 4, 4, 1, float("inf"), None, None, None, 0, None) == -1
    assert lib.bn_dsm_normals_from_points(None, 4, 4, None, None) == -1
    with pytest.raises(_lib.BnError):
        PD.DsmGeoref(1.0, (0, 0, 0)).get_latlonalt_from_nerf_prediction(torch.zeros(4, 11), torch.zeros(4))
    dummy = ctypes.c_void_p(256)                                          # never dereferenced: validation fails first
    assert lib.bn_dsm_points(dummy, 11, dummy, 10, 1.0, 0.0, 0.0, 0.0, 2, 0, dummy, None, None, None, None) == -1
    assert lib.bn_dsm_points(dummy, 11, dummy, 10, 1.0, 0.0, 0.0, 0.0, 0, 0, dummy, None, None, None, None) == -1   # ecef needs a zone
    with pytest.raises(NotImplementedError, match="defect"):
        PD.DsmGeoref(1.0, (0, 0, 0), cs="ecef").get_dsm_from_nerf_prediction(torch.zeros(4, 11), torch.zeros(4))


def test_box_filter_factorisation_random_grids():
    """Randomised version of the factorisation check: any radius 0..3, grids that cut through the cloud, points exactly on cell
    borders and far outside — count images identical, heights within the float32 running-mean error."""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        radius = int(rng.integers(0, 4))
        res = float(rng.choice([0.25, 0.5, 1.0, 2.0]))
        xs, ys = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        xoff = float(np.round(rng.uniform(-500, 500) / res) * res)
        yoff = float(np.round(rng.uniform(-500, 500) / res) * res)
        n = int(rng.integers(1, 600))
        x = rng.uniform(xoff - 3 * res, xoff + (xs + 3) * res, n)
        y = rng.uniform(yoff - (ys + 3) * res, yoff + 3 * res, n)
        snap = rng.random(n) < 0.2                                       # a fifth of the points exactly on cell borders
        x[snap] = xoff + np.round((x[snap] - xoff) / res) * res
        y[snap] = yoff - np.round((yoff - y[snap]) / res) * res
        cloud = np.stack([x, y, rng.normal(50.0, 20.0, n)], 1)
        grid = (xoff, yoff, res, xs, ys)
        want, wc = D.plyflatten(cloud, *grid, radius=radius, return_count=True)
        got, gc, _ = _emulate_cuda_box_path(cloud, grid, radius=radius)
        assert np.array_equal(gc, wc), (trial, radius, grid)
        assert np.array_equal(np.isnan(got), np.isnan(want[..., 0]))
        if (~np.isnan(got)).any():
            assert np.nanmax(np.abs(got - want[..., 0])) <= 2e-4, (trial, np.nanmax(np.abs(got - want[..., 0])))
