"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (numpy float64) of the reference's ray feed geometry ("RPC localisation -> rays"):

    get_rays                      datasets/satellite_rgb_dep.py:23-78     (called per image at :249, :355, :443)
    normalize_rays                datasets/satellite_rgb_dep.py:550-559
    get_sun_dirs                  datasets/satellite_rgb_dep.py:561-576
    sat_utils.rescale_rpc         sat_utils.py:90-108
    sat_utils.latlon_to_ecef_custom   sat_utils.py:110-125

Pinned (tests/test_oracle_vs_reference.py): `latlon_to_ecef_custom`, `normalize_rays`, `get_sun_dirs` bit-exact against the
live reference, and `get_rays` end to end against the live reference's `get_rays` driven by THIS file's RPC model object
(so everything in get_rays except the two third-party calls below is pinned).

PARITY UNPINNED for two third-party dependencies that are neither vendored under /root/reference nor installed here:
  * `rpcm` (requirements.txt:2, no version pin): `RPCModel.localization` = iterative inversion of the RPC00B projection.
    Restated from the package's published source (rpc_model.py: apply_poly / apply_rfm / projection /
    localization_iterative): start at normalised (lon, lat) = (-1, -1), probe steps EPS = 2 then 0.1, project the pixel
    error on the two finite-difference image vectors (assumed orthogonal), iterate until EVERY point is within
    1e-18 squared normalised pixels, at most 100 iterations.
  * `pyproj` / PROJ + `utm` (requirements.txt:19 utm==0.7.0): `sat_utils.utm_from_latlon` (sat_utils.py:148-162) =
    Transformer(+proj=latlong -> +proj=utm +zone=<n><letter>).  Restated as the Krueger series transverse Mercator to
    order n^6 in Karney's formulation (Karney 2011, eqs 7-11, 35), which is what PROJ's tmerc computes to well below a
    nanometre; ellipsoid GRS80 (PROJ's default when a +proj string names none), k0 = 0.9996, false easting 500 km and —
    because "+zone=17R" carries no "+south" — NO false northing in either hemisphere; the zone comes from the FIRST point
    (utm.latlon_to_zone_number, with its Norway / Svalbard exceptions).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np


# ------------------------------------------------------------------------------------------ rpcm restatement (unpinned)
def apply_poly(poly, x, y, z):
    """rpcm.rpc_model.apply_poly: the 20-term RPC00B cubic, x = lat, y = lon, z = alt (all normalised)."""
    out = 0
    out += poly[0]
    out += poly[1] * y + poly[2] * x + poly[3] * z
    out += poly[4] * y * x + poly[5] * y * z + poly[6] * x * z
    out += poly[7] * y * y + poly[8] * x * x + poly[9] * z * z
    out += poly[10] * x * y * z
    out += poly[11] * y * y * y
    out += poly[12] * y * x * x + poly[13] * y * z * z + poly[14] * y * y * x
    out += poly[15] * x * x * x
    out += poly[16] * x * z * z + poly[17] * y * y * z + poly[18] * x * x * z
    out += poly[19] * z * z * z
    return out


def apply_rfm(num, den, x, y, z):
    return apply_poly(num, x, y, z) / apply_poly(den, x, y, z)


@dataclass
class RPCModel:
    """The attributes of rpcm.RPCModel that the path reads (dict_format='rpcm': the dict IS these attributes)."""
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

    def projection(self, lon, lat, alt):
        nlon = (np.asarray(lon) - self.lon_offset) / self.lon_scale
        nlat = (np.asarray(lat) - self.lat_offset) / self.lat_scale
        nalt = (np.asarray(alt) - self.alt_offset) / self.alt_scale
        col = apply_rfm(self.col_num, self.col_den, nlat, nlon, nalt)
        row = apply_rfm(self.row_num, self.row_den, nlat, nlon, nalt)
        return col * self.col_scale + self.col_offset, row * self.row_scale + self.row_offset

    def localization(self, col, row, alt):
        ncol = (np.asarray(col) - self.col_offset) / self.col_scale
        nrow = (np.asarray(row) - self.row_offset) / self.row_scale
        nalt = (np.asarray(alt) - self.alt_offset) / self.alt_scale
        lon, lat = self.localization_iterative(ncol, nrow, nalt)
        return lon * self.lon_scale + self.lon_offset, lat * self.lat_scale + self.lat_offset

    def localization_iterative(self, col, row, alt):
        col, row, alt = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (col, row, alt))
        Xf = np.vstack([col, row]).T
        lon = -col ** 0
        lat = -col ** 0
        EPS = 2
        x0 = apply_rfm(self.col_num, self.col_den, lat, lon, alt)
        y0 = apply_rfm(self.row_num, self.row_den, lat, lon, alt)
        x1 = apply_rfm(self.col_num, self.col_den, lat, lon + EPS, alt)
        y1 = apply_rfm(self.row_num, self.row_den, lat, lon + EPS, alt)
        x2 = apply_rfm(self.col_num, self.col_den, lat + EPS, lon, alt)
        y2 = apply_rfm(self.row_num, self.row_den, lat + EPS, lon, alt)
        n = 0
        while not np.all((x0 - col) ** 2 + (y0 - row) ** 2 < 1e-18):
            if n > 100:
                raise RuntimeError("Max localization iterations (100) exceeded")
            X0 = np.vstack([x0, y0]).T
            e1 = np.vstack([x1, y1]).T - X0
            e2 = np.vstack([x2, y2]).T - X0
            u = Xf - X0
            a1 = np.sum(u * e1, axis=1) / np.sum(e1 * e1, axis=1)
            a2 = np.sum(u * e2, axis=1) / np.sum(e2 * e2, axis=1)
            lon = lon + a1 * EPS
            lat = lat + a2 * EPS
            EPS = .1
            x0 = apply_rfm(self.col_num, self.col_den, lat, lon, alt)
            y0 = apply_rfm(self.row_num, self.row_den, lat, lon, alt)
            x1 = apply_rfm(self.col_num, self.col_den, lat, lon + EPS, alt)
            y1 = apply_rfm(self.row_num, self.row_den, lat, lon + EPS, alt)
            x2 = apply_rfm(self.col_num, self.col_den, lat + EPS, lon, alt)
            y2 = apply_rfm(self.row_num, self.row_den, lat + EPS, lon, alt)
            n += 1
        self.last_iterations = n
        return lon, lat


def rescale_rpc(rpc: RPCModel, alpha: float) -> RPCModel:
    """sat_utils.py:90-108."""
    r = copy.copy(rpc)
    r.row_scale *= float(alpha)
    r.col_scale *= float(alpha)
    r.row_offset *= float(alpha)
    r.col_offset *= float(alpha)
    return r


def synthetic_rpc(view: int = 0) -> RPCModel:
    """The synthetic camera of brdf_nerf_b200.synth.synthetic_rpc_dict as an oracle RPC model."""
    from brdf_nerf_b200.synth import synthetic_rpc_dict
    return RPCModel(**synthetic_rpc_dict(view))


# ------------------------------------------------------------------------------------------ geodesy
WGS84_A = 6378137.0
WGS84_FINV = 298.257223563


def latlon_to_ecef_custom(lat, lon, alt):
    """Geodetic degrees / metres -> geocentric metres, same operation order as sat_utils.py:110-125 (pinned bit-exact):
    prime-vertical radius v = a / sqrt(1 - e2 sin^2 phi), then (v + h) cos phi cos lam, (v + h) cos phi sin lam,
    (v (1 - e2) + h) sin phi with e2 = 1 - (1 - f)^2."""
    phi, lam = lat * (np.pi / 180.0), lon * (np.pi / 180.0)
    flattening = 1 / WGS84_FINV
    ecc2 = 1 - (1 - flattening) * (1 - flattening)
    sin_phi = np.sin(phi)
    v = WGS84_A / np.sqrt(1 - ecc2 * sin_phi * sin_phi)
    ring = (v + alt) * np.cos(phi)
    return ring * np.cos(lam), ring * np.sin(lam), (v * (1 - ecc2) + alt) * sin_phi


def utm_zone_number(latitude: float, longitude: float) -> int:
    """utm.latlon_to_zone_number (utm==0.7.0): 6-degree zones with the Norway / Svalbard exceptions."""
    if 56 <= latitude < 64 and 3 <= longitude < 12:
        return 32
    if 72 <= latitude <= 84 and longitude >= 0:
        if longitude < 9:
            return 31
        elif longitude < 21:
            return 33
        elif longitude < 33:
            return 35
        elif longitude < 42:
            return 37
    return int((longitude + 180) / 6) + 1


GRS80_A = 6378137.0
GRS80_F = 1.0 / 298.257222101
UTM_K0 = 0.9996


def kruger_alpha(n: float):
    """Krueger series coefficients alpha_1..6 to order n^6 (Karney 2011, eq. 35)."""
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    return (n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
            13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
            61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
            49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
            34729 * n5 / 80640 - 3418889 * n6 / 1995840,
            212378941 * n6 / 319334400)


def utm_forward(lats, lons, zone: int):
    """+proj=utm +zone=<zone> (northern formula in both hemispheres, see the module docstring): degrees -> (east, north) m."""
    f = GRS80_F
    n = f / (2 - f)
    e = math.sqrt(f * (2 - f))
    A = GRS80_A / (1 + n) * (1 + n * n / 4 + n ** 4 / 64 + n ** 6 / 256)
    alpha = kruger_alpha(n)
    lon0 = (zone - 1) * 6 - 180 + 3
    phi = np.asarray(lats, dtype=np.float64) * (np.pi / 180.0)
    lam = (np.asarray(lons, dtype=np.float64) - lon0) * (np.pi / 180.0)
    tau = np.tan(phi)
    sigma = np.sinh(e * np.arctanh(e * tau / np.sqrt(1 + tau * tau)))
    taup = tau * np.sqrt(1 + sigma * sigma) - sigma * np.sqrt(1 + tau * tau)
    xip = np.arctan2(taup, np.cos(lam))
    etap = np.arcsinh(np.sin(lam) / np.sqrt(taup * taup + np.cos(lam) ** 2))
    xi, eta = xip.copy(), etap.copy()
    for j, a in enumerate(alpha, start=1):
        xi = xi + a * np.sin(2 * j * xip) * np.cosh(2 * j * etap)
        eta = eta + a * np.cos(2 * j * xip) * np.sinh(2 * j * etap)
    return 500000.0 + UTM_K0 * A * eta, UTM_K0 * A * xi


def utm_from_latlon(lats, lons):
    """sat_utils.py:148-162 (zone from the first point)."""
    lats, lons = np.atleast_1d(lats), np.atleast_1d(lons)
    return utm_forward(lats, lons, utm_zone_number(float(lats[0]), float(lons[0])))


# ------------------------------------------------------------------------------------------ the path
def _ground_points(cols, rows, rpc, alt, cs):
    """All pixels localised on the altitude plane `alt`, in scene coordinates (satellite_rgb_dep.py:46-52 / 55-61)."""
    alts = float(alt) * np.ones(cols.shape)
    lons, lats = rpc.localization(cols, rows, alts)
    if cs == "ecef":
        return np.vstack(latlon_to_ecef_custom(lats, lons, alts)).T
    east, north = utm_from_latlon(lats, lons)
    return np.vstack([east, north, alts]).T


def get_rays(cols, rows, rpc, min_alt, max_alt, cs="ecef"):
    """satellite_rgb_dep.py:23-78: (N, 8) float32 [o(3), d(3), near = 0, far].  The points of maximum altitude are the ones
    nearest to the camera (ray origins, :46-52, :64), those of minimum altitude the far ends (:55-61); d = unit(far - near)
    (:67-68), bounds [0, |far - near|] (:72-73), everything cast to float32 at the end (:76-77)."""
    cols, rows = np.asarray(cols), np.asarray(rows)
    near_pts = _ground_points(cols, rows, rpc, max_alt, cs)
    far_pts = _ground_points(cols, rows, rpc, min_alt, cs)
    delta = far_pts - near_pts
    length = np.linalg.norm(delta, axis=1)
    rays = np.hstack([near_pts, delta / length[:, np.newaxis], np.zeros((len(length), 1)), length[:, np.newaxis]])
    return rays.astype(np.float32)


def normalize_rays(rays, center, scene_range):
    """satellite_rgb_dep.py:550-559: float32 in-place arithmetic with the float32 `center` / `range` tensors."""
    rays = np.array(rays, dtype=np.float32, copy=True)
    c = np.asarray(center, dtype=np.float32)
    r = np.float32(scene_range)
    for k in range(3):
        rays[:, k] -= c[k]
    for k in (0, 1, 2, 6, 7):
        rays[:, k] /= r
    return rays


def get_sun_dirs(sun_elevation_deg, sun_azimuth_deg, n_rays):
    """satellite_rgb_dep.py:561-576."""
    sun_el = np.radians(sun_elevation_deg)
    sun_az = np.radians(sun_azimuth_deg)
    sun_d = np.array([np.sin(sun_az) * np.cos(sun_el), np.cos(sun_az) * np.cos(sun_el), np.sin(sun_el)])
    return np.tile(sun_d, (n_rays, 1)).astype(np.float32)


def image_rays(rpc, h, w, min_alt, max_alt, cs, center, scene_range, sun_elevation_deg, sun_azimuth_deg):
    """One image's training ray records (h*w, 11): satellite_rgb_dep.py:353-390 (meshgrid -> get_rays -> normalize_rays ->
    hstack with the sun directions)."""
    cols, rows = np.meshgrid(np.arange(w), np.arange(h))                 # :353
    rays = get_rays(cols.flatten(), rows.flatten(), rpc, min_alt, max_alt, cs=cs)
    rays = normalize_rays(rays, center, scene_range)
    return np.hstack([rays, get_sun_dirs(sun_elevation_deg, sun_azimuth_deg, rays.shape[0])])
