"""CPU: the ray-feed geometry oracle (oracle/georays_np.py) against the committed golden fixture (live-reference get_rays /
normalize_rays / get_sun_dirs outputs), independent checks of the restated third-party pieces (RPC inversion round trip,
transverse Mercator against meridian-arc quadrature and conformality), the product's host logic (UTM zone, single-pixel
localisation, RPC struct layout) and argument validation through the C ABI (no GPU needed)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from brdf_nerf_b200 import georays as PG
from oracle import georays_np as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "georays.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def test_oracle_get_rays_matches_reference_golden_bit_exact(g):
    rpc = G.synthetic_rpc(0)
    rays = G.get_rays(g["cols"], g["rows"], rpc, float(g["min_alt"]), float(g["max_alt"]), cs="ecef")
    assert rays.dtype == np.float32 and np.array_equal(rays, g["ref_rays_ecef"])
    assert np.array_equal(G.normalize_rays(rays, g["center"], float(g["scene_range"])), g["ref_rays_ecef_norm"])
    assert np.array_equal(G.get_sun_dirs(*g["sun_el_az"], 1), g["ref_sun"])
    utm = G.get_rays(g["cols"], g["rows"], rpc, float(g["min_alt"]), float(g["max_alt"]), cs="utm")
    assert np.array_equal(utm, g["restated_rays_utm"])                    # regression of the restated (unpinned) branch
    assert np.allclose(np.linalg.norm(utm[:, 3:6], axis=1), 1.0, atol=1e-6) and (utm[:, 5] < -0.9).all()
    assert np.array_equal(utm[:, 2], np.full(len(utm), np.float32(g["max_alt"])))        # origins on the max-altitude plane


def test_rpc_inversion_round_trip():
    """The restated localisation really inverts the restated projection (both are the RPC00B model)."""
    for view in range(3):
        rpc = G.synthetic_rpc(view)
        rng = np.random.default_rng(view)
        cols, rows = rng.uniform(0, 2047, 500), rng.uniform(0, 2047, 500)
        alts = rng.uniform(-30, 120, 500)
        lon, lat = rpc.localization(cols, rows, alts)
        c2, r2 = rpc.projection(lon, lat, alts)
        assert np.abs(c2 - cols).max() < 1e-5 and np.abs(r2 - rows).max() < 1e-5
        assert rpc.last_iterations <= 10
    half = G.rescale_rpc(rpc, 0.5)
    lon2, lat2 = half.localization(cols / 2, rows / 2, alts)
    assert np.abs(lon2 - lon).max() < 1e-9 and np.abs(lat2 - lat).max() < 1e-9


def test_transverse_mercator_against_independent_checks():
    """No pyproj here: check the Krueger-series restatement against (1) the meridian arc length by quadrature on the central
    meridian, (2) the UTM scale factor there, (3) conformality (Cauchy-Riemann) off the meridian, (4) east-west symmetry."""
    a, f = G.GRS80_A, G.GRS80_F
    e2 = f * (2 - f)
    for lat in (0.0, 12.5, 30.31, 47.0, 63.2, 80.0):
        e, n = G.utm_forward(np.array([lat]), np.array([-81.0]), 17)          # zone 17: central meridian 81 W
        t = np.linspace(0.0, math.radians(lat), 200001)
        arc = np.trapezoid(a * (1 - e2) / (1 - e2 * np.sin(t) ** 2) ** 1.5, t) if lat else 0.0
        assert abs(e[0] - 500000.0) < 1e-6 and abs(n[0] - G.UTM_K0 * arc) < 2e-5
    h = 1e-6
    for lat, lon in ((30.31, -81.66), (30.31, -78.2), (55.0, -83.9), (-33.9, -80.1)):
        e0, n0 = G.utm_forward(np.array([lat]), np.array([lon]), 17)
        e1, n1 = G.utm_forward(np.array([lat + h]), np.array([lon]), 17)
        e2_, n2 = G.utm_forward(np.array([lat]), np.array([lon + h]), 17)
        phi = math.radians(lat)
        M = a * (1 - e2) / (1 - e2 * math.sin(phi) ** 2) ** 1.5              # metres per radian of latitude
        N = a / math.sqrt(1 - e2 * math.sin(phi) ** 2) * math.cos(phi)       # metres per radian of longitude
        dn_dphi, de_dphi = (n1 - n0)[0] / math.radians(h) / M, (e1 - e0)[0] / math.radians(h) / M
        dn_dlam, de_dlam = (n2 - n0)[0] / math.radians(h) / N, (e2_ - e0)[0] / math.radians(h) / N
        assert abs(dn_dphi - de_dlam) < 2e-6 and abs(de_dphi + dn_dlam) < 2e-6          # conformal
        assert 0.9995 < math.hypot(dn_dphi, de_dphi) < 1.0012                              # UTM scale inside a zone
    ee, nn = G.utm_forward(np.array([40.0, 40.0]), np.array([-81.0 - 2.0, -81.0 + 2.0]), 17)
    assert abs((ee[0] - 500000.0) + (ee[1] - 500000.0)) < 1e-8 and abs(nn[0] - nn[1]) < 1e-8
    _, south = G.utm_forward(np.array([-10.0]), np.array([-81.0]), 17)
    assert south[0] < 0                                                     # "+zone=17R" has no +south: no false northing


def test_utm_published_worked_example():
    """Known-answer anchor for the restated (unpinned) projection: the worked example of the UTM article most readers know —
    the CN Tower, 43 deg 38' 33.24" N, 79 deg 23' 13.7" W, lies in zone 17 at 630 084 m east, 4 833 438 m north (WGS84 / GRS80
    differ by 0.1 mm in the semi-minor axis: invisible at the metre the example is quoted to)."""
    lat = 43 + 38 / 60 + 33.24 / 3600
    lon = -(79 + 23 / 60 + 13.7 / 3600)
    assert G.utm_zone_number(lat, lon) == 17
    e, n = G.utm_forward(np.array([lat]), np.array([lon]), 17)
    assert abs(e[0] - 630084.0) < 1.0 and abs(n[0] - 4833438.0) < 1.0


def test_product_host_logic_matches_oracle():
    rpc_o = G.synthetic_rpc(1)
    d = {k: getattr(rpc_o, k) for k in PG._KEYS + PG._POLYS}
    rpc_p = PG.RPCModel.from_dict(d)
    for col, row, alt in ((0.0, 0.0, 95.0), (2047.0, 13.0, -25.0), (1000.5, 1999.25, 10.0)):
        lon, lat = PG.localize_one(rpc_p, col, row, alt)
        lo, la = rpc_o.localization(np.array([col]), np.array([row]), np.array([alt]))
        assert abs(lon - lo[0]) < 1e-12 and abs(lat - la[0]) < 1e-12
    for lat in np.arange(-80, 84.5, 3.7):
        for lon in np.arange(-180, 180, 2.9):
            assert PG.utm_zone_number(lat, lon) == G.utm_zone_number(lat, lon)
    assert PG.utm_zone_number(60.0, 5.0) == 32 and PG.utm_zone_number(75.0, 10.0) == 33 and PG.utm_zone_number(30.3, -81.66) == 17
    half = PG.rescale_rpc(rpc_p, 0.5)
    assert half.row_scale == rpc_p.row_scale * 0.5 and half.col_offset == rpc_p.col_offset * 0.5 and half.lat_scale == rpc_p.lat_scale
    s = rpc_p.as_struct()
    assert C.sizeof(s) == 90 * 8 and s.col_den[0] == 1.0 and s.alt_scale == rpc_o.alt_scale
    assert PG.get_sun_dir(62.5, 148.0) == [float(v) for v in G.get_sun_dirs(62.5, 148.0, 1)[0]]
    with pytest.raises(ValueError):
        PG.RPCModel.from_dict(dict(d, row_num=[0.0] * 19))
    bad = PG.RPCModel.from_dict(dict(d, col_num=[0.0] * 20, row_num=[0.0] * 20))          # degenerate camera: never converges
    with pytest.raises((RuntimeError, ZeroDivisionError)):
        PG.localize_one(bad, 5.0, 5.0, 0.0)


def test_georays_abi_argument_validation():
    from brdf_nerf_b200 import _lib
    lib = _lib.load()
    s = PG.RPCModel.from_dict({k: getattr(G.synthetic_rpc(0), k) for k in PG._KEYS + PG._POLYS}).as_struct()
    call = lambda **kw: lib.bn_rays_from_rpc(*[kw.get(k, dflt) for k, dflt in (
        ("rpc", C.byref(s)), ("cols", None), ("rows", None), ("n", 16), ("width", 4), ("min_alt", 0.0), ("max_alt", 50.0),
        ("cs", 1), ("zone", 17), ("normalize", 0), ("cx", 0.0), ("cy", 0.0), ("cz", 0.0), ("range", 1.0), ("sun", None),
        ("out", None), ("stride", 8), ("fail", None), ("iters", C.c_void_p(512)), ("stream", None))])
    assert call() == -1 and b"null pointer" in lib.bn_last_error()
    dummy = C.c_void_p(256)                                               # never dereferenced: validation fails first
    assert call(out=dummy, cs=2) == -1 and call(out=dummy, zone=0) == -1 and call(out=dummy, stride=11) == -1
    assert call(out=dummy, width=0) == -1 and call(out=dummy, n=0) == -1 and call(out=dummy, normalize=1, range=0.0) == -1
    with pytest.raises(_lib.BnError):
        PG.get_rays(np.arange(4.0), np.arange(4.0), G.synthetic_rpc(0), 0.0, 50.0, device="cpu")
