"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (numpy float64) of the reference's tile-inference -> DSM product path:

    get_latlonalt_from_nerf_prediction   datasets/satellite_rgb_dep.py:601-634   (cs == 'utm', the default: opt.py:252)
    get_dsm_from_nerf_prediction         datasets/satellite_rgb_dep.py:636-697   (grid derivation 657-671, plyflatten call 680)
    calc_normal_from_depth_v2            datasets/satellite_rgb_dep.py:578-585 -> sat_utils.calc_normal_from_pts3d (sat_utils.py:16-50)

Pinned: `latlonalt_from_nerf_prediction` and `calc_normal_from_pts3d` are checked bit-exact against the live reference
methods in tests/test_oracle_vs_reference.py (oracle/ref_harness.load_dataset_module), the grid derivation is a literal
restatement of numpy scalar arithmetic.  UNPINNED: the rasteriser `plyflatten` itself is an absent third-party package
(plyflatten==0.2.0); oracle/plyflatten_restated.c restates its published algorithm (see that file's header) and
`plyflatten_py` below is the same loop in pure Python for small cross-checks.

cs == 'ecef' additionally runs ecef_to_latlon_custom (sat_utils.py:127-146) and pyproj's UTM projection
(sat_utils.py:148-162; pyproj / utm are absent): not restated, the host mirror raises for it.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "plyflatten_restated.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libplyflatten_restated.so")


def ecef_to_latlon_custom(x, y, z):
    """Geocentric metres -> geodetic (lat deg, lon deg, alt m): one Bowring step, same operation order as
    sat_utils.py:127-146 (pinned bit-exact against the live reference).  Note the constants: the eccentricity is given as a
    literal, the semi-minor axis derived from it."""
    a, ecc = 6378137.0, 8.1819190842622e-2
    a2, ecc2 = a ** 2, ecc ** 2
    b = np.sqrt(a2 * (1 - ecc2))
    b2 = b ** 2
    ecc_prime = np.sqrt((a2 - b2) / b2)
    p = np.sqrt((x ** 2) + (y ** 2))
    theta = np.arctan2(a * z, b * p)
    lon = np.arctan2(y, x)
    lat = np.arctan2((z + (ecc_prime ** 2) * b * (np.sin(theta) ** 3)), (p - ecc2 * a * (np.cos(theta) ** 3)))
    prime_vertical = a / (np.sqrt(1 - ecc2 * (np.sin(lat) ** 2)))
    alt = p / np.cos(lat) - prime_vertical
    return lat * 180 / np.pi, lon * 180 / np.pi, alt


def latlonalt_from_nerf_prediction(rays, depth, scene_range, center, cs="utm"):
    """satellite_rgb_dep.py:614-634.  cs == 'utm' (default): pinned.  cs == 'ecef': ecef_to_latlon_custom (pinned) followed by
    sat_utils.utm_from_latlon (pyproj: restated in oracle/georays_np.py, unpinned).  rays (N,11) float32, depth (N,) float32; `scene_range` and
    `center` are the dataset's float32 values (`self.range` 0-dim float32 tensor :165, `self.center` float32 (3,) :164).
    Returns easts, norths, alts as float64 vectors."""
    rays = np.asarray(rays, dtype=np.float32).astype(np.float64)          # :614  rays.double()
    depth = np.asarray(depth, dtype=np.float32).astype(np.float64).reshape(-1, 1)
    xyz_n = rays[:, 0:3] + rays[:, 3:6] * depth                            # :619
    xyz = xyz_n * np.float64(np.float32(scene_range))                      # :622  double tensor * 0-dim float32 tensor
    c = np.asarray(center, dtype=np.float32).astype(np.float64)
    xyz[:, 0] += c[0]                                                      # :623-625
    xyz[:, 1] += c[1]
    xyz[:, 2] += c[2]
    if cs == "ecef":                                                       # :629-631
        from oracle.georays_np import utm_from_latlon
        lats, lons, alts = ecef_to_latlon_custom(xyz[:, 0], xyz[:, 1], xyz[:, 2])
        easts, norths = utm_from_latlon(lats, lons)
        return easts, norths, alts
    return xyz[:, 0].copy(), xyz[:, 1].copy(), xyz[:, 2].copy()            # :632-633 (cs == 'utm')


def dsm_grid(easts, norths, resolution=0.5, roi=None):
    """satellite_rgb_dep.py:657-671.  roi = the four numbers of roi_txt (xoff, yoff, size, resolution) or None.
    Returns (xoff, yoff, resolution, xsize, ysize)."""
    if roi is not None:
        xoff, yoff = float(roi[0]), float(roi[1])                          # :659
        xsize, ysize = int(roi[2]), int(roi[2])                            # :660
        resolution = float(roi[3])                                         # :661
        yoff += ysize * resolution                                         # :662
        return xoff, yoff, resolution, xsize, ysize
    xmin, xmax = float(np.min(easts)), float(np.max(easts))                # :666
    ymin, ymax = float(np.min(norths)), float(np.max(norths))              # :667
    xoff = math.floor(xmin / resolution) * resolution                      # :668
    xsize = int(1 + math.floor((xmax - xoff) / resolution))                # :669
    yoff = math.ceil(ymax / resolution) * resolution                       # :670
    ysize = int(1 - math.floor((ymin - yoff) / resolution))                # :671
    return xoff, yoff, resolution, xsize, ysize


def build_c(force=False) -> str:
    """gcc -O2 -shared the C restatement into oracle/_build/ (git-ignored; travels to the GPU box with the snapshot)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(_OUT_DIR, exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"], check=True)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c())
        lib.plyflatten_restated.restype = C.c_int
        lib.plyflatten_restated.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1, sigma=float("inf"), return_count=False):
    """The call of satellite_rgb_dep.py:680 through the C restatement.  cloud (N, 2+E) float64 -> (ysize, xsize, E) float32."""
    cloud = np.ascontiguousarray(np.asarray(cloud, dtype=np.float64))
    n, e = cloud.shape[0], cloud.shape[1] - 2
    raster = np.empty((ysize, xsize, e), dtype=np.float32)
    cnt = np.empty((ysize, xsize), dtype=np.float32)
    rc = _load().plyflatten_restated(cloud.ctypes.data, n, e, float(xoff), float(yoff), float(resolution), int(xsize),
                                     int(ysize), int(radius), float(sigma), raster.ctypes.data, cnt.ctypes.data)
    if rc != 0:
        raise MemoryError("plyflatten_restated failed")
    return (raster, cnt) if return_count else raster


def plyflatten_py(cloud, xoff, yoff, resolution, xsize, ysize, radius=1, sigma=float("inf")):
    """The same loop in pure Python / numpy float32 scalars (small inputs only): cross-check of the C build."""
    f32 = np.float32
    e = cloud.shape[1] - 2
    avg = np.zeros((ysize, xsize, e), dtype=np.float32)
    cnt = np.zeros((ysize, xsize), dtype=np.float32)
    for p in range(cloud.shape[0]):
        xx, yy = float(cloud[p, 0]), float(cloud[p, 1])
        i = math.floor((xx - xoff) / resolution)
        j = math.floor((-yy - (-yoff)) / resolution)
        for k1 in range(-radius, radius + 1):
            for k2 in range(-radius, radius + 1):
                ii, jj = i + k1, j + k2
                if math.isinf(sigma):
                    w = f32(1.0)
                else:
                    dx = f32(xx - (xoff + resolution * (0.5 + ii)))
                    dy = f32(yy - (yoff - resolution * (0.5 + jj)))
                    d = f32(np.hypot(dx, dy))
                    w = f32(np.exp(-d * d / (f32(2.0) * f32(sigma) * f32(sigma))))
                if ii < 0 or jj < 0 or ii >= xsize or jj >= ysize:
                    continue
                for c in range(e):
                    v = f32(cloud[p, 2 + c])
                    avg[jj, ii, c] = (v * w + cnt[jj, ii] * avg[jj, ii, c]) / (w + cnt[jj, ii])
                cnt[jj, ii] += w
    avg[cnt == 0] = np.nan
    return avg, cnt


def dsm_from_nerf_prediction(rays, depth, scene_range, center, roi=None, return_count=False):
    """get_dsm_from_nerf_prediction (satellite_rgb_dep.py:636-697) without the GeoTIFF write: (ysize, xsize, 1) float32."""
    easts, norths, alts = latlonalt_from_nerf_prediction(rays, depth, scene_range, center)
    cloud = np.vstack([easts, norths, alts]).T                             # :655
    xoff, yoff, res, xsize, ysize = dsm_grid(easts, norths, 0.5, roi)
    out = plyflatten(cloud, xoff, yoff, res, xsize, ysize, radius=1, sigma=float("inf"), return_count=return_count)
    grid = (xoff, yoff, res, xsize, ysize)
    return (out[0], out[1], grid) if return_count else (out, grid)


def _l2n(x):
    """train_utils.l2_normalize (train_utils.py:28-33): x / sqrt(max(sum x^2, eps)), eps = float32 machine epsilon."""
    import torch
    norm = torch.sum(x ** 2, dim=-1, keepdim=True)
    return x / torch.sqrt(torch.maximum(norm, torch.tensor(torch.finfo(torch.float32).eps)))


def calc_normal_from_pts3d(pts3d):
    """sat_utils.calc_normal_from_pts3d(pts3d, valid_depth=None, Flatten=False)[0] (sat_utils.py:16-50), torch CPU float32:
    four cross products of the normalised differences to the S/N/E/W neighbours, each normalised, averaged, normalised;
    the border stays zero.  (The reference calls torch.cross without `dim`: the first axis of size 3, which is the last one
    for images that are not 5 pixels high or wide.)"""
    import torch
    pts3d = torch.as_tensor(pts3d, dtype=torch.float32)
    c = pts3d[1:-1, 1:-1, :]
    south = _l2n(pts3d[2:, 1:-1, :] - c)
    north = _l2n(pts3d[:-2, 1:-1, :] - c)
    east = _l2n(pts3d[1:-1, 2:, :] - c)
    west = _l2n(pts3d[1:-1, :-2, :] - c)
    n1 = _l2n(torch.cross(east, north, dim=-1))
    n2 = _l2n(torch.cross(west, south, dim=-1))
    n3 = _l2n(torch.cross(north, west, dim=-1))
    n4 = _l2n(torch.cross(south, east, dim=-1))
    mean = _l2n((n1 + n2 + n3 + n4) / 4.)
    normals = torch.zeros_like(pts3d)
    normals[1:-1, 1:-1, :] = mean
    return normals


def normal_from_depth_v2(rays, depth, height, width, scene_range, center):
    """calc_normal_from_depth_v2 (satellite_rgb_dep.py:578-585): float64 cloud -> float32 point image -> normals (h*w, 3)."""
    import torch
    e, n, a = latlonalt_from_nerf_prediction(rays, depth, scene_range, center)
    pts = torch.from_numpy(np.vstack([e, n, a]).T).type(torch.FloatTensor)
    return calc_normal_from_pts3d(pts.reshape(height, width, 3)).reshape(-1, 3)
