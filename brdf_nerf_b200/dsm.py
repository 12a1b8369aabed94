"""Tile inference -> DSM on the GPU: host-side mirror of the reference's dataset methods that turn the
rendered depth of a tile into the products its users evaluate (`eval.py:153-182`, `main.py:477,621`):

    SatelliteRGBDEPDataset.get_latlonalt_from_nerf_prediction   datasets/satellite_rgb_dep.py:601-634
    SatelliteRGBDEPDataset.get_dsm_from_nerf_prediction         datasets/satellite_rgb_dep.py:636-697
    SatelliteRGBDEPDataset.calc_normal_from_depth_v2            datasets/satellite_rgb_dep.py:578-585
                                                                (-> sat_utils.calc_normal_from_pts3d, sat_utils.py:16-50)

`DsmGeoref` stands in for the dataset object: it carries the three attributes those methods read (`range`, `center`,
`cs`; satellite_rgb_dep.py:138,164-165) and offers the methods under the reference's names with the same argument
meaning.  Inputs and outputs are CUDA tensors: the depth of a 2048x2048 tile never leaves the device (the reference
moves rays and depth to the CPU, builds the cloud in numpy and rasterises it single-threaded in C).  Kernels:
`csrc/dsm.cu` through the C ABI (`bn_dsm_*`); there is no CPU path.

`cs == 'ecef'` (geocentric scene coordinates) is supported for the point cloud: ecef_to_latlon_custom (sat_utils.py:127-146)
followed by the UTM projection of sat_utils.py:148-162 (pyproj in the reference, restated in csrc/geodesy.cuh).  Not mirrored:
the GeoTIFF write (`dsm_path`, rasterio; file I/O is out of scope) and `get_dsm_from_nerf_prediction` with cs='ecef', a
reference defect path (it projects the already projected coordinates a second time).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


@dataclass
class DsmGrid:
    """Raster georeferencing: cell (row j, col i) covers x in [xoff + i res, xoff + (i+1) res), y in (yoff - (j+1) res, yoff - j res]."""
    xoff: float
    yoff: float
    resolution: float
    xsize: int
    ysize: int


def grid_from_bounds(xmin: float, xmax: float, ymin: float, ymax: float, resolution: float = 0.5) -> DsmGrid:
    """satellite_rgb_dep.py:665-671 (numpy float64 scalar arithmetic == Python floats)."""
    xoff = math.floor(xmin / resolution) * resolution
    xsize = int(1 + math.floor((xmax - xoff) / resolution))
    yoff = math.ceil(ymax / resolution) * resolution
    ysize = int(1 - math.floor((ymin - yoff) / resolution))
    return DsmGrid(xoff, yoff, resolution, xsize, ysize)


def grid_from_roi(roi: Sequence[float]) -> DsmGrid:
    """satellite_rgb_dep.py:658-662: roi = the four numbers of `roi_txt` (xoff, yoff, size, resolution)."""
    xoff, yoff = float(roi[0]), float(roi[1])
    size = int(roi[2])
    resolution = float(roi[3])
    return DsmGrid(xoff, yoff + size * resolution, resolution, size, size)


def _check_cloud(cloud: torch.Tensor, grid: DsmGrid):
    if cloud.dim() != 2 or cloud.shape[1] < 3:
        raise ValueError("cloud must be (N, >=3): [x, y, values...]")
    if grid.xsize <= 0 or grid.ysize <= 0:
        raise ValueError(f"empty raster {grid}")


def accumulate_cloud(cloud: torch.Tensor, grid: DsmGrid, radius: int = 1, sigma: float = float("inf"), value_col: int = 2,
                     workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First half of the rasteriser: adds the points' (value, weight) into the accumulator workspace (allocated and zeroed
    when `workspace` is None, else accumulated on top).  The workspace is `cells` float64 sums followed by `cells` float32
    weight sums; summing the workspaces of several ray shards gives the workspace of their union."""
    lib = L.load()
    _check_cloud(cloud, grid)
    cloud = cloud.contiguous()
    nbytes = lib.bn_dsm_workspace_bytes(grid.xsize, grid.ysize, radius, sigma)
    fresh = workspace is None
    if cloud.shape[0] == 0:           # an empty shard contributes nothing
        return torch.zeros(nbytes, dtype=torch.uint8, device=cloud.device) if fresh else workspace
    if fresh:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=cloud.device)
    L.check(lib.bn_dsm_accumulate(L.ptr(cloud, torch.float64), cloud.shape[1], value_col, cloud.shape[0], grid.xoff, grid.yoff,
                                  grid.resolution, grid.xsize, grid.ysize, radius, sigma, 1 if fresh else 0,
                                  L.ptr(workspace, torch.uint8), workspace.numel(), L.stream_ptr()))
    return workspace


def workspace_views(workspace: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(float64 sums, float32 weight sums) views of an accumulator workspace: the two all-reduce operands."""
    cells = workspace.numel() // 12
    return workspace[:cells * 8].view(torch.float64), workspace[cells * 8:cells * 12].view(torch.float32)


def finalize_raster(workspace: torch.Tensor, grid: DsmGrid, radius: int = 1, sigma: float = float("inf"),
                    return_count: bool = False):
    """Second half: accumulators -> (ysize, xsize, 1) float32 raster (NaN where no point fell) [+ weight-sum image]."""
    dev = workspace.device
    raster = torch.empty(grid.ysize, grid.xsize, 1, dtype=torch.float32, device=dev)
    count = torch.empty(grid.ysize, grid.xsize, dtype=torch.float32, device=dev) if return_count else None
    L.check(L.load().bn_dsm_finalize(grid.xsize, grid.ysize, radius, sigma, L.ptr(workspace, torch.uint8), workspace.numel(),
                                     L.ptr(raster), L.ptr(count), L.stream_ptr()))
    return (raster, count) if return_count else raster


def rasterize_cloud(cloud: torch.Tensor, grid: DsmGrid, radius: int = 1, sigma: float = float("inf"),
                    value_col: int = 2, return_count: bool = False):
    """`plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma)` (satellite_rgb_dep.py:680) for one value
    column: cloud (N, >=3) float64 CUDA -> (ysize, xsize, 1) float32 CUDA, NaN where no point fell."""
    lib = L.load()
    _check_cloud(cloud, grid)
    cloud = cloud.contiguous()
    dev = cloud.device
    raster = torch.empty(grid.ysize, grid.xsize, 1, dtype=torch.float32, device=dev)
    count = torch.empty(grid.ysize, grid.xsize, dtype=torch.float32, device=dev) if return_count else None
    nbytes = lib.bn_dsm_workspace_bytes(grid.xsize, grid.ysize, radius, sigma)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.bn_dsm_rasterize(L.ptr(cloud, torch.float64), cloud.shape[1], value_col, cloud.shape[0], grid.xoff, grid.yoff,
                                 grid.resolution, grid.xsize, grid.ysize, radius, sigma, L.ptr(raster), L.ptr(count),
                                 L.ptr(ws, torch.uint8), nbytes, L.stream_ptr()))
    return (raster, count) if return_count else raster


def reduce_bounds(bounds: torch.Tensor, group=None) -> torch.Tensor:
    """[xmin, xmax, ymin, ymax] of this rank's points -> the bounds of all ranks' points (one MAX all-reduce of
    [-xmin, xmax, -ymin, ymax]).  Works on any backend (gloo in the CPU tests)."""
    sign = torch.tensor([-1.0, 1.0, -1.0, 1.0], dtype=bounds.dtype, device=bounds.device)
    v = bounds * sign
    torch.distributed.all_reduce(v, op=torch.distributed.ReduceOp.MAX, group=group)
    return v * sign


def valid_normal_mask(height: int, width: int, device, valid_depth: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The `valid_normal` output of sat_utils.calc_normal_from_pts3d (sat_utils.py:19-24), flattened: ones without a mask;
    with an (H, W) depth-validity image, a pixel below 1e-5 keeps its own value and an interior pixel becomes the product of
    its four neighbours' validities.  Element-wise torch on the device (a few bytes per pixel of mask algebra)."""
    if valid_depth is None:
        return torch.ones(height * width, dtype=torch.float32, device=device)
    vd = valid_depth.to(device).reshape(height, width)
    out = torch.where(vd < 1e-5, vd, torch.ones_like(vd))
    out[1:-1, 1:-1] = vd[2:, 1:-1] * vd[:-2, 1:-1] * vd[1:-1, 2:] * vd[1:-1, :-2]
    return out.flatten()


def normals_from_points(points: torch.Tensor) -> torch.Tensor:
    """sat_utils.calc_normal_from_pts3d(pts3d, valid_depth=None, Flatten=False)[0]: (H, W, 3) float32 -> (H, W, 3)."""
    if points.dim() != 3 or points.shape[-1] != 3:
        raise ValueError("points must be (H, W, 3)")
    points = points.contiguous()
    out = torch.empty_like(points)
    L.check(L.load().bn_dsm_normals_from_points(L.ptr(points), points.shape[0], points.shape[1], L.ptr(out), L.stream_ptr()))
    return out


class DsmGeoref:
    """The georeferencing state of `SatelliteRGBDEPDataset` that the prediction -> DSM methods read.

    scene_range / center: X/Y/Z_scale maximum and X/Y/Z_offset of `scene.loc` (satellite_rgb_dep.py:163-165); they are
    float32 tensors in the reference, so they are rounded to float32 here before entering the float64 arithmetic."""

    def __init__(self, scene_range: float, center: Sequence[float], cs: str = "utm"):
        self.range = float(np.float32(scene_range))
        self.center = tuple(float(np.float32(c)) for c in center)
        if cs not in ("utm", "ecef"):
            raise ValueError(f"cs must be 'utm' or 'ecef', got {cs!r}")
        self.cs = cs

    def _points(self, rays: torch.Tensor, depth: torch.Tensor, want_f32: bool, want_bounds: bool):
        if not rays.is_cuda:
            raise L.BnError("brdf_nerf_b200.dsm needs CUDA tensors (there is no CPU path)")
        rays = rays.contiguous()
        depth = depth.reshape(-1).contiguous()
        n = rays.shape[0]
        if depth.shape[0] != n:
            raise ValueError(f"rays ({n}) and depth ({depth.shape[0]}) disagree")
        dev = rays.device
        if n == 0:                    # an empty ray shard: no points, neutral bounds for the MAX all-reduce
            inf = float("inf")
            return (torch.empty(0, 3, dtype=torch.float64, device=dev),
                    torch.empty(0, 3, dtype=torch.float32, device=dev) if want_f32 else None,
                    torch.tensor([inf, -inf, inf, -inf], dtype=torch.float64, device=dev) if want_bounds else None)
        cloud = torch.empty(n, 3, dtype=torch.float64, device=dev)
        pts = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_f32 else None
        bounds = torch.empty(4, dtype=torch.float64, device=dev) if want_bounds else None
        scratch = torch.empty(4, dtype=torch.int64, device=dev) if want_bounds else None
        zone = self._first_point_zone(rays, depth) if self.cs == "ecef" else 0
        L.check(L.load().bn_dsm_points(L.ptr(rays), rays.shape[1], L.ptr(depth), n, self.range, *self.center,
                                       1 if self.cs == "utm" else 0, zone, L.ptr(cloud, torch.float64), L.ptr(pts), L.ptr(bounds, torch.float64),
                                       L.ptr(scratch, torch.int64), L.stream_ptr()))
        return cloud, pts, bounds

    def _first_point_zone(self, rays: torch.Tensor, depth: torch.Tensor) -> int:
        """cs == 'ecef': the reference projects with the UTM zone of the FIRST point (`utm.latlon_to_zone_number(lats[0],
        lons[0])`, sat_utils.py:155).  One 28-byte read-back, then ecef_to_latlon_custom (sat_utils.py:127-146) in host floats."""
        from .georays import utm_zone_number
        r0 = rays[0, :6].to(torch.float64).cpu().tolist()
        d0 = float(depth[0].to(torch.float64).cpu())
        x, y, z = ((r0[k] + r0[3 + k] * d0) * self.range + self.center[k] for k in range(3))
        a, e = 6378137.0, 8.1819190842622e-2
        b = math.sqrt(a * a * (1 - e * e))
        ep = math.sqrt((a * a - b * b) / (b * b))
        p = math.sqrt(x * x + y * y)
        th = math.atan2(a * z, b * p)
        lon = math.atan2(y, x)
        lat = math.atan2(z + ep * ep * b * math.sin(th) ** 3, p - e * e * a * math.cos(th) ** 3)
        return utm_zone_number(math.degrees(lat), math.degrees(lon))

    def _dsm_cs_check(self):
        if self.cs != "utm":
            raise NotImplementedError(
                "get_dsm_from_nerf_prediction with cs='ecef' is a reference defect path: it feeds the (east, north) that "
                "get_latlonalt_from_nerf_prediction returns into utm_from_latlon a second time as if they were (lat, lon) "
                "(satellite_rgb_dep.py:649-651).  Use get_latlonalt_from_nerf_prediction + rasterize_cloud, or cs='utm'.")

    def get_latlonalt_from_nerf_prediction(self, rays: torch.Tensor, depth: torch.Tensor, bPrint: bool = False
                                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """rays (h*w, 11), depth (h*w[, 1]) -> easts, norths, alts: float64 CUDA vectors of length h*w."""
        cloud, _, _ = self._points(rays, depth, False, False)
        return cloud[:, 0], cloud[:, 1], cloud[:, 2]

    def get_dsm_from_nerf_prediction(self, rays: torch.Tensor, depth: torch.Tensor, dsm_path: Optional[str] = None,
                                     roi_txt=None, return_grid: bool = False):
        """-> dsm (ysize, xsize, 1) float32 CUDA tensor (NaN = no data).  `roi_txt`: path of the ROI text file or its four
        numbers.  With `return_grid` also returns the `DsmGrid` (the affine transform the reference writes into the GeoTIFF
        profile, satellite_rgb_dep.py:694: Affine(res, 0, xoff, 0, -res, yoff))."""
        if dsm_path is not None:
            raise NotImplementedError("writing the GeoTIFF (rasterio) is outside the hot path: save the returned tensor")
        self._dsm_cs_check()
        if roi_txt is not None:
            roi = np.loadtxt(roi_txt) if isinstance(roi_txt, (str, bytes)) else np.asarray(roi_txt, dtype=np.float64)
            cloud, _, _ = self._points(rays, depth, False, False)
            grid = grid_from_roi(roi)
        else:
            cloud, _, bounds = self._points(rays, depth, False, True)
            b = bounds.cpu().tolist()          # the one host round trip: four scalars that size the output raster
            if not all(math.isfinite(v) for v in b):
                raise ValueError("no finite point in the predicted cloud")
            grid = grid_from_bounds(b[0], b[1], b[2], b[3], 0.5)
        dsm = rasterize_cloud(cloud, grid, radius=1, sigma=float("inf"))
        return (dsm, grid) if return_grid else dsm

    def get_dsm_from_nerf_prediction_sharded(self, rays: torch.Tensor, depth: torch.Tensor, group=None, roi_txt=None,
                                             return_grid: bool = False):
        """The same DSM when the tile's rays are sharded over the ranks of `group` (pixel blocks per rank, as
        `inference.render_tile` leaves them): each rank passes ITS rays and depths; the raster grid comes from the
        all-reduced bounds, each rank accumulates its points, the accumulators are summed over NCCL (two all-reduces:
        float64 sums, float32 counts), and every rank finalises the full raster.  The depth image is never gathered."""
        import torch.distributed as dist
        self._dsm_cs_check()
        if roi_txt is not None:
            roi = np.loadtxt(roi_txt) if isinstance(roi_txt, (str, bytes)) else np.asarray(roi_txt, dtype=np.float64)
            cloud, _, _ = self._points(rays, depth, False, False)
            grid = grid_from_roi(roi)
        else:
            cloud, _, bounds = self._points(rays, depth, False, True)
            b = reduce_bounds(bounds, group).cpu().tolist()
            if not all(math.isfinite(v) for v in b):        # same on every rank (the bounds are all-reduced): all raise together
                raise ValueError("no finite point in the predicted cloud")
            grid = grid_from_bounds(b[0], b[1], b[2], b[3], 0.5)
        ws = accumulate_cloud(cloud, grid, radius=1, sigma=float("inf"))
        sums, counts = workspace_views(ws)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dsm = finalize_raster(ws, grid, radius=1, sigma=float("inf"))
        return (dsm, grid) if return_grid else dsm

    def calc_normal_from_depth_v2(self, rays: torch.Tensor, depth: torch.Tensor, height: int, width: int,
                                  valid_depth=None):
        """-> (normals (h*w, 3) float32, valid_normal (h*w,)): satellite_rgb_dep.py:578-585.  The reference's callers pass no
        `valid_depth` (eval.py:434, main.py:477: the mask is all ones); with an (h, w) mask the validity image follows
        sat_utils.py:19-24 (the normals themselves do not depend on it)."""
        _, pts, _ = self._points(rays, depth, True, False)
        normals = normals_from_points(pts.view(height, width, 3)).reshape(-1, 3)
        return normals, valid_normal_mask(height, width, rays.device, valid_depth)
