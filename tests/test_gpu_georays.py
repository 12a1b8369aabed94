"""GPU parity of the ray-feed kernel (csrc/georays.cu, bn_rays_from_rpc through the C ABI) against the golden fixture
(live-reference get_rays / normalize_rays outputs for cs='ecef') and the oracle (cs='utm').  Bar: the records are float32
casts of float64 results whose only difference to the reference is libm (sin / cos / atanh ... of CUDA vs numpy, <= 2 ulp in
float64): after the cast a value is either identical or one float32 ulp away, and such flips must be rare."""
import os

import numpy as np
import pytest
import torch

from brdf_nerf_b200 import georays as PG
from oracle import georays_np as G

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "georays.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _rpc(view=0):
    o = G.synthetic_rpc(view)
    return PG.RPCModel.from_dict({k: getattr(o, k) for k in PG._KEYS + PG._POLYS}), o


def _assert_f32_close(got, want, what, max_flip_frac=0.02, ulps=1):
    """identical, or <= `ulps` float32 ulp apart on at most `max_flip_frac` of the entries"""
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert got.shape == want.shape, what
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = ulps * np.spacing(np.maximum(np.abs(want), np.abs(got)).astype(np.float32)).astype(np.float64)
    assert (diff <= tol).all(), f"{what}: worst {np.max(diff / np.maximum(tol, 1e-300)):.1f} ulp"
    assert (diff > 0).mean() <= max_flip_frac, f"{what}: {100 * (diff > 0).mean():.2f}% of the entries differ"


def test_get_rays_ecef_vs_reference_golden(cuda, g):
    rpc, _ = _rpc(0)
    rays = PG.get_rays(g["cols"], g["rows"], rpc, float(g["min_alt"]), float(g["max_alt"]), cs="ecef", device=cuda)
    assert rays.shape == (len(g["cols"]), 8) and rays.dtype == torch.float32
    _assert_f32_close(rays.cpu().numpy(), g["ref_rays_ecef"], "get_rays ecef")


def test_get_rays_utm_vs_oracle(cuda, g):
    rpc, _ = _rpc(0)
    rays = PG.get_rays(g["cols"], g["rows"], rpc, float(g["min_alt"]), float(g["max_alt"]), cs="utm", device=cuda)
    _assert_f32_close(rays.cpu().numpy(), g["restated_rays_utm"], "get_rays utm")


@pytest.mark.parametrize("cs", ["ecef", "utm"])
@pytest.mark.parametrize("hw", [(1, 1), (3, 43), (16, 8), (37, 129)])
def test_image_rays_vs_oracle_ragged(cuda, cs, hw):
    """meshgrid + get_rays + normalize_rays + sun columns in one launch; sizes around the 128-ray block."""
    h, w = hw
    rpc, rpc_o = _rpc(1)
    center = (799000.0, -5452800.0, 3200200.0) if cs == "ecef" else (436200.0, 3353400.0, 30.0)
    got = PG.image_rays(rpc, h, w, -25.0, 95.0, cs, center, 400.0, 62.5, 148.0, device=cuda).cpu().numpy()
    want = G.image_rays(rpc_o, h, w, -25.0, 95.0, cs, center, 400.0, 62.5, 148.0)
    assert got.shape == (h * w, 11)
    assert np.array_equal(got[:, 8:], want[:, 8:]) and np.array_equal(got[:, 6], want[:, 6])       # sun columns, near = 0
    _assert_f32_close(got[:, 3:6], want[:, 3:6], "directions", max_flip_frac=0.05)
    _assert_f32_close(got[:, 7], want[:, 7], "far", max_flip_frac=0.05)
    # origins: (float32(o) - c) / r — a one-ulp flip of the float32 cast of a ~5e6 m coordinate is 0.5 m = 1.25e-3 here
    # (the reference's own quantisation); anything else must be identical
    d = np.abs(got[:, 0:3] - want[:, 0:3])
    assert (d <= 0.5 / 400.0 * 1.01).all() and (d > 0).mean() <= 0.05


def test_rescaled_rpc_and_without_sun(cuda):
    rpc, rpc_o = _rpc(2)
    got = PG.image_rays(PG.rescale_rpc(rpc, 0.25), 40, 52, 0.0, 60.0, "utm", (436200.0, 3353400.0, 30.0), 400.0, device=cuda)
    want = G.image_rays(G.rescale_rpc(rpc_o, 0.25), 40, 52, 0.0, 60.0, "utm", (436200.0, 3353400.0, 30.0), 400.0, 0.0, 0.0)[:, :8]
    assert got.shape == (40 * 52, 8)
    _assert_f32_close(got.cpu().numpy()[:, 3:], want[:, 3:], "rescaled rpc", max_flip_frac=0.05)


def test_non_convergence_is_reported(cuda):
    rpc, _ = _rpc(0)
    bad = PG.RPCModel.from_dict({**{k: getattr(rpc, k) for k in PG._KEYS + PG._POLYS}, "col_num": [0.0] * 20, "row_num": [0.0] * 20})
    with pytest.raises(RuntimeError, match="Max localization iterations"):
        PG.get_rays(np.arange(8.0), np.arange(8.0), bad, 0.0, 50.0, cs="ecef", device=cuda)
    rays = PG.get_rays(np.arange(8.0), np.arange(8.0), bad, 0.0, 50.0, cs="ecef", device=cuda, check=False)
    assert rays.shape == (8, 8)                                           # the launch itself survives a degenerate camera


def test_full_image_2048_properties_and_strided_oracle(cuda):
    """BASELINE-size image (4.19 M rays) in one launch: unit directions, zero near, positive far, origins on the max-altitude
    plane, and a strided subset (every 1009th pixel) against the oracle."""
    h = w = 2048
    rpc, rpc_o = _rpc(0)
    rays = PG.get_rays(*np.meshgrid(np.arange(w), np.arange(h)), rpc, -25.0, 95.0, cs="utm", device=cuda)
    assert rays.shape == (h * w, 8)
    assert (rays[:, 3:6].norm(dim=1) - 1).abs().max().item() < 1e-6
    assert bool((rays[:, 6] == 0).all()) and bool((rays[:, 7] > 100).all()) and bool((rays[:, 2] == 95.0).all())
    assert torch.isfinite(rays).all()
    idx = np.arange(0, h * w, 1009)
    want = G.get_rays((idx % w).astype(np.float64), (idx // w).astype(np.float64), rpc_o, -25.0, 95.0, cs="utm")
    _assert_f32_close(rays[torch.from_numpy(idx).to(cuda)].cpu().numpy(), want, "strided subset", max_flip_frac=0.02)
    grid = PG.image_rays(rpc, h, w, -25.0, 95.0, "utm", (0.0, 0.0, 0.0), 1.0, device=cuda)       # the pixel-grid path
    assert torch.equal(grid, rays)                                        # center 0 / range 1: normalisation is the identity
