"""Ray feed geometry on the GPU: RPC camera model + pixel grid -> the ray records `render_rays` consumes.

Host-side mirror of the reference's per-image ray construction (same names, argument meaning and error behaviour):

    get_rays(cols, rows, rpc, min_alt, max_alt, bPrint=False, cs='ecef')   datasets/satellite_rgb_dep.py:23-78
    SatelliteRGBDEPDataset.normalize_rays / get_sun_dirs                    datasets/satellite_rgb_dep.py:550-576
    sat_utils.rescale_rpc                                                   sat_utils.py:90-108
    rpcm.RPCModel(d["rpc"], dict_format="rpcm")                             satellite_rgb_dep.py:246 (third-party `rpcm`)

`RPCModel` here only holds the attributes (it is the "rpcm" dict of the scene JSON); the per-pixel localisation runs in
`csrc/georays.cu` through `bn_rays_from_rpc`.  The one piece of host arithmetic is the UTM zone, which the reference takes
from the FIRST point it projects (`utm.latlon_to_zone_number(lats[0], lons[0])`, sat_utils.py:155): `localize_one` inverts
the RPC for that single pixel in Python floats.  There is no CPU path for the rays themselves.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

_KEYS = ("row_offset", "col_offset", "lat_offset", "lon_offset", "alt_offset",
         "row_scale", "col_scale", "lat_scale", "lon_scale", "alt_scale")
_POLYS = ("row_num", "row_den", "col_num", "col_den")


class BnRpc(C.Structure):
    _fields_ = [(k, C.c_double) for k in _KEYS] + [(k, C.c_double * 20) for k in _POLYS]


@dataclass
class RPCModel:
    row_offset: float
    col_offset: float
    lat_offset: float
    lon_offset: float
    alt_offset: float
    row_scale: float
    col_scale: float
    lat_scale: float
    lon_scale: float
    alt_scale: float
    row_num: List[float] = field(default_factory=list)
    row_den: List[float] = field(default_factory=list)
    col_num: List[float] = field(default_factory=list)
    col_den: List[float] = field(default_factory=list)

    @classmethod
    def from_dict(cls, d, dict_format: str = "rpcm") -> "RPCModel":
        """`rpcm.RPCModel(d, dict_format="rpcm")`: the dict already carries the attribute names."""
        if dict_format != "rpcm":
            raise NotImplementedError("only dict_format='rpcm' (the scene JSON layout) is on the path")
        for k in _POLYS:
            if len(d[k]) != 20:
                raise ValueError(f"{k} must have 20 coefficients")
        return cls(**{k: float(d[k]) for k in _KEYS}, **{k: [float(v) for v in d[k]] for k in _POLYS})

    def as_struct(self) -> BnRpc:
        s = BnRpc()
        for k in _KEYS:
            setattr(s, k, float(getattr(self, k)))
        for k in _POLYS:
            setattr(s, k, (C.c_double * 20)(*[float(v) for v in getattr(self, k)]))
        return s


def rescale_rpc(rpc: RPCModel, alpha: float) -> RPCModel:
    """sat_utils.py:90-108: RPC of the image resized by `alpha`."""
    r = copy.copy(rpc)
    r.row_scale *= float(alpha)
    r.col_scale *= float(alpha)
    r.row_offset *= float(alpha)
    r.col_offset *= float(alpha)
    return r


def _poly(p, x, y, z):
    return (p[0] + (p[1] * y + p[2] * x + p[3] * z) + (p[4] * y * x + p[5] * y * z + p[6] * x * z)
            + (p[7] * y * y + p[8] * x * x + p[9] * z * z) + p[10] * x * y * z + p[11] * y * y * y
            + (p[12] * y * x * x + p[13] * y * z * z + p[14] * y * y * x) + p[15] * x * x * x
            + (p[16] * x * z * z + p[17] * y * y * z + p[18] * x * x * z) + p[19] * z * z * z)


def localize_one(rpc: RPCModel, col: float, row: float, alt: float):
    """(lon, lat) in degrees of ONE pixel at altitude `alt`: the iterative RPC inversion in Python floats, used only to pick
    the UTM zone.  Raises like rpcm when 100 iterations do not converge."""
    ncol, nrow = (col - rpc.col_offset) / rpc.col_scale, (row - rpc.row_offset) / rpc.row_scale
    nalt = (alt - rpc.alt_offset) / rpc.alt_scale
    f = lambda la, lo: (_poly(rpc.col_num, la, lo, nalt) / _poly(rpc.col_den, la, lo, nalt),
                        _poly(rpc.row_num, la, lo, nalt) / _poly(rpc.row_den, la, lo, nalt))
    lon = lat = -1.0
    eps = 2.0
    n = 0
    while True:
        x0, y0 = f(lat, lon)
        if (x0 - ncol) ** 2 + (y0 - nrow) ** 2 < 1e-18:
            break
        if n > 100:
            raise RuntimeError("Max localization iterations (100) exceeded")
        x1, y1 = f(lat, lon + eps)
        x2, y2 = f(lat + eps, lon)
        e1, e2, u = (x1 - x0, y1 - y0), (x2 - x0, y2 - y0), (ncol - x0, nrow - y0)
        lon += (u[0] * e1[0] + u[1] * e1[1]) / (e1[0] ** 2 + e1[1] ** 2) * eps
        lat += (u[0] * e2[0] + u[1] * e2[1]) / (e2[0] ** 2 + e2[1] ** 2) * eps
        eps = .1
        n += 1
    return lon * rpc.lon_scale + rpc.lon_offset, lat * rpc.lat_scale + rpc.lat_offset


def utm_zone_number(latitude: float, longitude: float) -> int:
    """`utm.latlon_to_zone_number` (utm==0.7.0, requirements.txt:19), including the Norway / Svalbard exceptions."""
    if 56 <= latitude < 64 and 3 <= longitude < 12:
        return 32
    if 72 <= latitude <= 84 and longitude >= 0:
        for bound, zone in ((9, 31), (21, 33), (33, 35), (42, 37)):
            if longitude < bound:
                return zone
    return int((longitude + 180) / 6) + 1


ZONE_LETTERS = "CDEFGHJKLMNPQRSTUVWXX"


def utm_zone_letter(latitude: float):
    """`utm.latitude_to_zone_letter` (utm==0.7.0): 8-degree bands C..X from 80 S to 84 N, None outside."""
    if -80 <= latitude <= 84:
        return ZONE_LETTERS[int(latitude + 80) >> 3]
    return None


def get_zone(cols, rows, rpc: RPCModel, min_alt):
    """datasets/satellite_rgb_dep.py:80-85: UTM zone number and band letter of the image's first pixel at `min_alt` (what the
    reference writes into the DSM GeoTIFF's CRS, satellite_rgb_dep.py:150,464,682)."""
    c0, r0 = float(np.asarray(cols).reshape(-1)[0]), float(np.asarray(rows).reshape(-1)[0])
    lon, lat = localize_one(rpc, c0, r0, float(min_alt))
    return utm_zone_number(lat, lon), utm_zone_letter(lat)


def _launch(rpc: RPCModel, cols, rows, n, width, min_alt, max_alt, cs, device, normalize, center, scene_range, sun_dir):
    if cs not in ("ecef", "utm"):
        raise ValueError(f"cs must be 'ecef' or 'utm', got {cs!r}")
    lib = L.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.BnError("brdf_nerf_b200.georays builds rays on a CUDA device (there is no CPU path)")
    zones = (0, 0)
    if cs == "utm":       # the reference derives the zone separately in each of its two utm_from_latlon calls (first point)
        c0 = float(cols.reshape(-1)[0]) if cols is not None else 0.0
        r0 = float(rows.reshape(-1)[0]) if rows is not None else 0.0
        zones = tuple(utm_zone_number(*reversed(localize_one(rpc, c0, r0, alt))) for alt in (max_alt, min_alt))
        if zones[0] != zones[1]:
            raise NotImplementedError("the image's first pixel falls into different UTM zones at min_alt and max_alt")
    stride = 11 if sun_dir is not None else 8
    out = torch.empty(n, stride, dtype=torch.float32, device=dev)
    fail = torch.zeros(1, dtype=torch.int32, device=dev)
    iters = torch.empty(2, dtype=torch.int32, device=dev)
    sun = (C.c_float * 3)(*[float(v) for v in sun_dir]) if sun_dir is not None else None
    c = [float(np.float32(v)) for v in (center if center is not None else (0.0, 0.0, 0.0))]
    s = rpc.as_struct()
    with torch.cuda.device(dev):
        L.check(lib.bn_rays_from_rpc(C.byref(s), L.ptr(cols, torch.float64), L.ptr(rows, torch.float64), n, int(width),
                                     float(min_alt), float(max_alt), 0 if cs == "ecef" else 1, zones[0], 1 if normalize else 0,
                                     c[0], c[1], c[2], float(np.float32(scene_range)), sun, L.ptr(out), stride,
                                     L.ptr(fail, torch.int32), L.ptr(iters, torch.int32), L.stream_ptr()))
    return out, fail


def get_rays(cols, rows, rpc: RPCModel, min_alt, max_alt, bPrint: bool = False, cs: str = "ecef", device="cuda",
             check: bool = True) -> torch.Tensor:
    """(h*w, 8) float32 CUDA tensor [o(3), d(3), near = 0, far] of the pixels (cols[i], rows[i]).  `check` reads the
    convergence counter back (one host sync) and raises like rpcm when a localisation did not converge."""
    dev = torch.device(device)
    cols = torch.as_tensor(np.asarray(cols, dtype=np.float64)).to(dev) if not torch.is_tensor(cols) else cols.to(dev, torch.float64)
    rows = torch.as_tensor(np.asarray(rows, dtype=np.float64)).to(dev) if not torch.is_tensor(rows) else rows.to(dev, torch.float64)
    cols, rows = cols.reshape(-1).contiguous(), rows.reshape(-1).contiguous()
    if cols.numel() != rows.numel() or cols.numel() == 0:
        raise ValueError("cols and rows must be non-empty and of equal length")
    out, fail = _launch(rpc, cols, rows, cols.numel(), 0, min_alt, max_alt, cs, dev, False, None, 1.0, None)
    if check and int(fail.item()) > 0:
        raise RuntimeError("Max localization iterations (100) exceeded")
    return out


def get_sun_dir(sun_elevation_deg: float, sun_azimuth_deg: float):
    """The one row `get_sun_dirs` tiles (satellite_rgb_dep.py:571-573), rounded to float32 like `.type(FloatTensor)`."""
    el, az = np.radians(sun_elevation_deg), np.radians(sun_azimuth_deg)
    return [float(np.float32(v)) for v in (np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el))]


def image_rays(rpc: RPCModel, height: int, width: int, min_alt, max_alt, cs: str, center: Sequence[float],
               scene_range: float, sun_elevation_deg: Optional[float] = None, sun_azimuth_deg: Optional[float] = None,
               device="cuda", check: bool = True) -> torch.Tensor:
    """All ray records of one image in ONE launch: meshgrid -> get_rays -> normalize_rays [-> hstack sun directions]
    (satellite_rgb_dep.py:353-390): (h*w, 11) float32 CUDA tensor, or (h*w, 8) without a sun position."""
    sun = get_sun_dir(sun_elevation_deg, sun_azimuth_deg) if sun_elevation_deg is not None else None
    out, fail = _launch(rpc, None, None, height * width, width, min_alt, max_alt, cs, device, True, center, scene_range, sun)
    if check and int(fail.item()) > 0:
        raise RuntimeError("Max localization iterations (100) exceeded")
    return out
