"""ORACLE (test infrastructure only — never imported by the product path).

Runs the UNMODIFIED reference (read-only tree at $BRDFNERF_REF or /root/reference) in this process
so the restatement in oracle/ can be pinned against it and golden vectors can be generated.
The reference tree does not exist on the GPU box; everything here must therefore only be reached
from tests that skip when the tree is absent, and from oracle/make_golden.py.

Import recipe: the path imports `rasterio` (train_utils.py:9) and, for the losses,
`kornia.losses.ssim` (metrics.py:7); neither is installed nor used on the path, so empty stub
modules are registered before importing.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from typing import List

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # written by oracle/stage_ref.py


def _pick_root() -> str:
    """The live read-only tree when it exists (this container), else the byte-identical staged copy of the path's files
    (oracle/_ref, the GPU box)."""
    live = os.environ.get("BRDFNERF_REF", "/root/reference")
    if os.path.isfile(os.path.join(live, "rendering.py")):
        return live
    return _STAGED


REF_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "rendering.py"))


def kind() -> str:
    """'live' (/root/reference), 'staged' (oracle/_ref) or 'absent'."""
    if not available():
        return "absent"
    return "staged" if os.path.abspath(REF_ROOT) == os.path.abspath(_STAGED) else "live"


_mods = None


def load():
    """Import (once) and return (rendering, load_model, metrics) of the live reference."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in ("rasterio", "kornia", "kornia.losses"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia.losses"].ssim = lambda *a, **k: None
    sys.modules["kornia"].losses = sys.modules["kornia.losses"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import rendering as ref_rendering          # noqa
        from models import load_model as ref_load  # noqa
        import metrics as ref_metrics              # noqa
    _mods = (ref_rendering, ref_load, ref_metrics)
    return _mods


def build_model(args, seed=0):
    """Random-init reference model exactly as the reference does (spsbrdfnerf.py:537-539)."""
    _, ref_load, _ = load()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref_load(args)
    return model


@contextlib.contextmanager
def inject_draws(queue: List[torch.Tensor]):
    """Replace torch.rand / rand_like / randn by a FIFO of pre-drawn tensors (shape-checked)."""
    q = list(queue)
    orig = (torch.rand, torch.rand_like, torch.randn)

    def pop(shape, what):
        if not q:
            raise RuntimeError(f"draw queue exhausted at {what}{tuple(shape)}")
        t = q.pop(0)
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"draw shape mismatch at {what}: want {tuple(shape)}, queued {tuple(t.shape)}")
        return t.clone()

    def _shape(a):
        return tuple(a[0]) if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else tuple(a)

    torch.rand = lambda *a, **k: pop(_shape(a), "rand")
    torch.rand_like = lambda t, **k: pop(t.shape, "rand_like").to(t.dtype)
    torch.randn = lambda *a, **k: pop(_shape(a), "randn")
    try:
        yield q
    finally:
        torch.rand, torch.rand_like, torch.randn = orig


def draw_queue(draws, valid_depth=None, mode="test", dtype=torch.float32):
    """Order of draws inside one reference render_rays call."""
    q = [draws.u_strat, draws.noise1]
    if draws.u_sun is not None:
        q += [draws.u_sun, draws.noise_sun]
    q += [draws.u_pred]
    if mode == "train" and valid_depth is not None:
        q += [draws.u_gt[valid_depth > 0]]
    q += [draws.noise2]
    return [t.to(dtype) for t in q]


def render(model, args, rays, draws, dtype=torch.float32, ts=None, embedding=None, **kw):
    """Call the live reference render_rays with injected draws; stdout is swallowed.  `ts` (N,) int64 + `embedding`
    (nn.Embedding, models['t']) for models built with beta=True (rendering.py:228-229)."""
    ref_rendering, _, _ = load()
    mode = kw.get("mode", "test")
    q = draw_queue(draws, kw.get("valid_depth"), mode, dtype)
    buf = io.StringIO()
    models = {"coarse": model}
    if embedding is not None:
        models["t"] = embedding
    with inject_draws(q) as rest, contextlib.redirect_stdout(buf):
        res, btype = ref_rendering.render_rays(models, args, rays.to(dtype), ts, **kw)
    if rest:
        raise RuntimeError(f"{len(rest)} queued draws were not consumed")
    return res, btype


_ds_mod = None


def load_dataset_module():
    """Import (once) the live reference's `datasets/satellite_rgb_dep.py` WITHOUT running `datasets/__init__.py` (which
    pulls in every dataset and their absent dependencies).  Only the pure-arithmetic methods of `SatelliteRGBDEPDataset`
    are used (`get_latlonalt_from_nerf_prediction`, `calc_normal_from_depth`), called unbound with a stand-in `self`
    that carries `range`, `center`, `cs`.  Stubs: rasterio, rpcm, pytorch3d.transforms (imported at module top, unused by
    those methods) and the sibling `cal_rmse_depth` module (needs matplotlib / plyflatten)."""
    global _ds_mod
    if _ds_mod is not None:
        return _ds_mod
    load()
    import importlib
    for name in ("rpcm", "pytorch3d", "pytorch3d.transforms", "datasets.cal_rmse_depth"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d.transforms"].axis_angle_to_matrix = None
    sys.modules["datasets.cal_rmse_depth"].cal_rmse_depth = None
    if "datasets" not in sys.modules or not hasattr(sys.modules["datasets"], "__path__"):
        pkg = types.ModuleType("datasets")
        pkg.__path__ = [os.path.join(REF_ROOT, "datasets")]
        sys.modules["datasets"] = pkg
    with contextlib.redirect_stdout(io.StringIO()):
        _ds_mod = importlib.import_module("datasets.satellite_rgb_dep")
    return _ds_mod


def ref_latlonalt(rays, depth, scene_range, center, cs="utm"):
    """The live reference's get_latlonalt_from_nerf_prediction (satellite_rgb_dep.py:601-634)."""
    mod = load_dataset_module()
    me = types.SimpleNamespace(range=torch.tensor(float(scene_range)), center=torch.tensor([float(c) for c in center]), cs=cs)
    return mod.SatelliteRGBDEPDataset.get_latlonalt_from_nerf_prediction(me, rays, depth)


def ref_normals_from_pts3d(pts3d):
    """The live reference's sat_utils.calc_normal_from_pts3d(pts3d, valid_depth=None, Flatten=False)[0] (sat_utils.py:16-50)."""
    load_dataset_module()
    import warnings
    import sat_utils as ref_sat_utils
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # torch.cross without dim= is deprecated
        return ref_sat_utils.calc_normal_from_pts3d(pts3d, None, False)[0]


def ref_get_rays(cols, rows, rpc, min_alt, max_alt, cs="ecef"):
    """The live reference's module-level get_rays (satellite_rgb_dep.py:23-78), driven by a duck-typed RPC object (anything
    with `.localization(cols, rows, alts) -> (lons, lats)`; rpcm itself is not installed).  Only cs='ecef' runs here: the
    'utm' branch calls pyproj."""
    mod = load_dataset_module()
    return mod.get_rays(cols, rows, rpc, min_alt, max_alt, cs=cs)


def ref_normalize_rays(rays, scene_range, center):
    mod = load_dataset_module()
    me = types.SimpleNamespace(range=torch.tensor(float(scene_range)), center=torch.tensor([float(c) for c in center]))
    return mod.SatelliteRGBDEPDataset.normalize_rays(me, rays.clone())


def ref_sun_dirs(sun_elevation_deg, sun_azimuth_deg, n_rays):
    mod = load_dataset_module()
    return mod.SatelliteRGBDEPDataset.get_sun_dirs(None, sun_elevation_deg, sun_azimuth_deg, n_rays)


def ref_latlon_to_ecef(lat, lon, alt):
    load_dataset_module()
    import sat_utils as ref_sat_utils
    return ref_sat_utils.latlon_to_ecef_custom(lat, lon, alt)


def ref_ecef_to_latlon(x, y, z):
    load_dataset_module()
    import sat_utils as ref_sat_utils
    return ref_sat_utils.ecef_to_latlon_custom(x, y, z)
