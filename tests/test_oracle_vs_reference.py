"""CPU, only where the reference tree is mounted: pin the oracle against the LIVE reference
(forward bit-exactness of the sampler, result keys, gradients incl. the second-order normal path)."""
import contextlib
import io

import pytest
import torch

from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.synth import make_rays
from oracle import losses_torch as LT
from oracle import ref_harness as RH
from oracle import render_torch as RT

pytestmark = pytest.mark.skipif(not RH.available(), reason="reference tree not mounted (GPU box)")


def _run(cfg, n=40, mode="test", ds=False, **kw):
    args = named_config(cfg)
    ref_model = RH.build_model(args)
    batch = make_rays(n, depth_supervision=ds)
    S1, G = args.n_samples, args.guided_samples
    S = G if kw.get("gsam_only") else S1 + G
    draws = RT.Draws.make(n, S1, G, S, seed=77, with_gt=ds, with_sun=bool(kw.get("bTestSun_v")), s_sun=G if kw.get("gsam_only") else S1)
    extra = dict(valid_depth=batch.valid_depth, target_depths=batch.target_depths, target_std=batch.target_std) if ds else {}
    train = mode == "train"
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        ref, bt = RH.render(ref_model, args, batch.rays, draws, mode=mode, **extra, **kw)
        om = RT.OracleModel(ref_model.state_dict(), args, requires_grad=train)
        ora, bt2, _ = RT.render_rays(om, args, batch.rays, draws, mode=mode, **extra, **kw)
    return args, ref_model, om, batch, ref, ora, bt, bt2


@pytest.mark.parametrize("cfg,kw", [
    ("lambertian", {}), ("lambertian", dict(gsam_only=True)),
    ("rpv111", dict(apply_brdf=True, cos_irra_on=True)), ("rpv111", dict(apply_brdf=False)),
    ("rpv111_multi", dict(apply_brdf=True, cos_irra_on=True)),
    ("hapke_bct", dict(apply_brdf=True, apply_theta=True, cos_irra_on=True)), ("hapke_b", dict(apply_brdf=True)),
    ("microfacet", dict(apply_brdf=True, cos_irra_on=True)),
    ("rpv111", dict(apply_brdf=True, cos_irra_on=True, bTestSun_v=True)),
    ("lambertian_viewdir", {}), ("rpv111_viewdir", dict(apply_brdf=True, cos_irra_on=True)),
])
def test_forward_keys_and_values(cfg, kw):
    _, _, _, _, ref, ora, bt, bt2 = _run(cfg, **kw)
    assert bt == bt2
    assert set(ref) == set(ora), set(ref) ^ set(ora)
    for k in ref:
        assert ref[k].shape == ora[k].shape, k
        if ref[k].dtype == torch.int64:
            assert torch.equal(ref[k], ora[k]), k
        elif k.startswith("z_vals"):
            assert torch.equal(ref[k], ora[k]), f"{k} not bit-exact"
        else:
            assert (ref[k] - ora[k]).abs().max().item() <= 2e-6, k


@pytest.mark.parametrize("cfg,ds,kw", [("lambertian_ds", True, {}), ("rpv111", False, dict(apply_brdf=True, cos_irra_on=True)),
                                       ("hapke_bct", False, dict(apply_brdf=True, apply_theta=True, cos_irra_on=True)),
                                       ("lambertian_viewdir", False, {})])
def test_gradients(cfg, ds, kw):
    args, ref_model, om, batch, ref, ora, _, _ = _run(cfg, n=24, mode="train", ds=ds, **kw)
    _, _, M = RH.load()
    with contextlib.redirect_stdout(io.StringIO()):
        loss_r, _ = M.SNerfLoss(lambda_sc=0., lambda_rgb=args.lambda_rgb)(ref, batch.rgbs)
        if ds:
            dl = M.DepthLoss(lambda_ds=args.ds_lambda, GNLL=False, usealldepth=False, margin=args.margin,
                             stdscale=args.stdscale, subset=True)
            l2, _ = dl(ref, batch.target_depths[:, 0], batch.target_depths[:, 1], target_valid_depth=batch.valid_depth,
                       target_std=batch.target_std)
            loss_r = loss_r + l2
    loss_r.backward()
    loss_o = LT.train_loss(ora, batch, args)
    loss_o.backward()
    assert abs(loss_r.item() - loss_o.item()) <= 1e-6
    for name, p in ref_model.named_parameters():
        go = om.p[name].grad
        if p.grad is None:
            assert go is None or go.abs().max() == 0
            continue
        s = p.grad.abs().max().item()
        assert (p.grad - go).abs().max().item() <= 1e-3 * s + 1e-9, name


def test_regularizer_losses_vs_reference():
    """NormalRegLoss (metrics.py:179-216) and HardSurfaceLoss (metrics.py:263-290) of the live reference on a live
    reference render == the oracle restatements (value, bad-normal percentage, gradients w.r.t. the weights)."""
    args, ref_model, om, batch, ref, ora, _, _ = _run("rpv111", n=24, mode="train", apply_brdf=True, cos_irra_on=True)
    _, _, M = RH.load()
    with contextlib.redirect_stdout(io.StringIO()):
        l_nr, _, perc = M.NormalRegLoss(lambda_nr_reg=0.1, keyword="normal_an")(ref)
        l_hs, _ = M.HardSurfaceLoss(lambda_hs=0.5)(ref)
    o_nr, o_perc = LT.normal_reg_loss(ora, 0.1, "normal_an")
    o_hs = LT.hard_surface_loss(ora, 0.5)
    assert abs(l_nr.item() - o_nr.item()) <= 1e-6 * max(1.0, abs(l_nr.item()))
    assert abs(l_hs.item() - o_hs.item()) <= 1e-7
    assert abs(float(perc) - o_perc) < 1e-4
    (l_nr + l_hs).backward()
    (o_nr + o_hs).backward()
    for name, p in ref_model.named_parameters():
        go = om.p[name].grad
        if p.grad is None:
            assert go is None or go.abs().max() == 0
            continue
        if go is None:                     # autograd of the reference leaves exact zeros where the oracle leaves None
            assert p.grad.abs().max() == 0, name
            continue
        s = p.grad.abs().max().item()
        assert (p.grad - go).abs().max().item() <= 2e-3 * s + 1e-9, name


def test_dsm_cloud_and_normals_vs_reference():
    """the DSM oracle's float64 point cloud and the neighbour-cross-product normals are bit-identical to the
    live reference methods (get_latlonalt_from_nerf_prediction, sat_utils.calc_normal_from_pts3d)."""
    import numpy as np
    from brdf_nerf_b200.synth import SCENE_CENTER, make_tile_rays, tile_surface_depth
    from oracle import dsm_np as D
    h, w = 21, 34
    rays = make_tile_rays(h, w, view=2)
    depth = tile_surface_depth(rays) + 0.01 * torch.rand(h * w, generator=torch.Generator().manual_seed(1))
    rng = 9.75
    e, n, a = RH.ref_latlonalt(rays, depth, rng, SCENE_CENTER)
    e2, n2, a2 = D.latlonalt_from_nerf_prediction(rays.numpy(), depth.numpy(), rng, SCENE_CENTER)
    assert np.array_equal(e, e2) and np.array_equal(n, n2) and np.array_equal(a, a2)
    pts = torch.from_numpy(np.vstack([e, n, a]).T).type(torch.FloatTensor).reshape(h, w, 3)
    assert torch.equal(RH.ref_normals_from_pts3d(pts), D.calc_normal_from_pts3d(pts))
    assert torch.equal(D.normal_from_depth_v2(rays.numpy(), depth.numpy(), h, w, rng, SCENE_CENTER).reshape(h, w, 3),
                       RH.ref_normals_from_pts3d(pts))


def test_georays_pinned_parts_vs_reference():
    """everything of get_rays / normalize_rays / get_sun_dirs / latlon_to_ecef_custom that lives in the
    reference tree is bit-identical in the oracle (the RPC inversion itself is rpcm's: restated, unpinned)."""
    import numpy as np
    from oracle import georays_np as G
    rpc = G.synthetic_rpc(2)
    rng = np.random.default_rng(3)
    cols, rows = rng.uniform(0, 2047, 300), rng.uniform(0, 2047, 300)
    ref = RH.ref_get_rays(cols, rows, rpc, -10.0, 70.0, cs="ecef")
    mine = G.get_rays(cols, rows, rpc, -10.0, 70.0, cs="ecef")
    assert np.array_equal(ref.numpy(), mine)
    c, r = (799000.0, -5452800.0, 3200200.0), 412.3
    assert np.array_equal(RH.ref_normalize_rays(ref, r, c).numpy(), G.normalize_rays(mine, c, r))
    assert np.array_equal(RH.ref_sun_dirs(41.0, 163.5, 7).numpy(), G.get_sun_dirs(41.0, 163.5, 7))
    lat, lon, alt = rng.uniform(-80, 80, 100), rng.uniform(-180, 180, 100), rng.uniform(-100, 9000, 100)
    for a, b in zip(RH.ref_latlon_to_ecef(lat, lon, alt), G.latlon_to_ecef_custom(lat, lon, alt)):
        assert np.array_equal(a, b)


def test_dsm_ecef_to_latlon_vs_reference():
    """cs='ecef' leg of get_latlonalt_from_nerf_prediction: ecef_to_latlon_custom (sat_utils.py:127-146) bit-exact."""
    import numpy as np
    from oracle import dsm_np as D
    from oracle import georays_np as G
    rng = np.random.default_rng(8)
    lat, lon, alt = rng.uniform(-85, 85, 500), rng.uniform(-180, 180, 500), rng.uniform(-100, 5000, 500)
    x, y, z = G.latlon_to_ecef_custom(lat, lon, alt)
    for a, b in zip(RH.ref_ecef_to_latlon(x, y, z), D.ecef_to_latlon_custom(x, y, z)):
        assert np.array_equal(a, b)
    la, lo, al = D.ecef_to_latlon_custom(x, y, z)
    assert np.abs(la - lat).max() < 1e-7 and np.abs(lo - lon).max() < 1e-9 and np.abs(al - alt).max() < 1e-2     # Bowring, one step


def test_reference_checkpoint_resumes_the_trainer():
    """A Lightning-style checkpoint of the LIVE reference module + torch.optim.Adam (main.py:147-150: Adam over
    model.parameters() in registration order) loads into Trainer.load_state_dict: same parameter order, names and shapes."""
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.train import Trainer
    args = named_config("rpv111")
    ref = RH.build_model(args, seed=3)
    opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=5e-4, weight_decay=0)
    g = torch.Generator().manual_seed(2)
    for p in ref.parameters():
        p.grad = torch.randn(p.shape, generator=g) * 1e-3
    opt.step(); opt.step()
    ckpt = {"state_dict": {f"nerf_coarse.{k}": v for k, v in ref.state_dict().items()}, "optimizer_states": [opt.state_dict()],
            "global_step": 2, "epoch": 0}
    torch.manual_seed(0)
    model = load_model(args)
    tr = Trainer(model, args)
    tr.load_state_dict(ckpt)
    assert tr.step_count == 2 and abs(tr.lr - 5e-4) < 1e-12
    assert [k for k, _ in model.named_parameters()] == [k for k, _ in ref.named_parameters()]
    flat = model.flat_params
    for (name, p), q in zip(model.named_parameters(), ref.parameters()):
        off = (p.data_ptr() - flat.data_ptr()) // 4
        assert torch.equal(p.detach(), q.detach()), name
        assert torch.equal(tr.m[off:off + p.numel()].view_as(p), opt.state[q]["exp_avg"]), name
        assert torch.equal(tr.v[off:off + p.numel()].view_as(p), opt.state[q]["exp_avg_sq"]), name


def test_valid_normal_mask_and_zone_letter_vs_reference():
    """Host-side pieces of the tile products against the live reference: the `valid_normal` image of
    sat_utils.calc_normal_from_pts3d (sat_utils.py:19-24) and get_zone's band letter rule."""
    from brdf_nerf_b200 import dsm as PD
    from brdf_nerf_b200 import georays as PG
    RH.load_dataset_module()
    import sat_utils as ref_sat_utils
    g = torch.Generator().manual_seed(6)
    pts = torch.rand(9, 11, 3, generator=g)
    vd = (torch.rand(9, 11, generator=g) > 0.3).float()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, want = ref_sat_utils.calc_normal_from_pts3d(pts, vd, True)
    assert torch.equal(PD.valid_normal_mask(9, 11, "cpu", vd), want)
    assert torch.equal(PD.valid_normal_mask(9, 11, "cpu"), torch.ones(99))
    # utm 0.7.0 letters: C..X in 8-degree bands, X extended to 84 N
    for lat, letter in ((-80.0, "C"), (-72.1, "C"), (-72.0, "D"), (-0.1, "M"), (0.0, "N"), (30.3, "R"), (71.9, "W"), (72.0, "X"), (84.0, "X")):
        assert PG.utm_zone_letter(lat) == letter
    assert PG.utm_zone_letter(-80.1) is None and PG.utm_zone_letter(84.1) is None
    from oracle import georays_np as G
    o = G.synthetic_rpc(0)
    rpc = PG.RPCModel.from_dict({k: getattr(o, k) for k in PG._KEYS + PG._POLYS})
    assert PG.get_zone([0, 5], [0, 7], rpc, -25.0) == (17, "R")


OPTION_VARIANTS = [
    ("hapke_b", dict(b=0, shell_hapke=1), dict(apply_brdf=True)),
    ("hapke_b", dict(b=0, shell_hapke=2), dict(apply_brdf=True)),
    ("hapke_b", dict(b=0, shell_hapke=3), dict(apply_brdf=True, cos_irra_on=True)),
    ("lambertian", dict(mapping=False), {}),
    ("rpv111", dict(mapping=False), dict(apply_brdf=True, cos_irra_on=True)),
    ("rpv111", dict(normal="analystic_learned"), dict(apply_brdf=True, cos_irra_on=True)),
]


@pytest.mark.parametrize("cfg,over,kw", OPTION_VARIANTS)
def test_option_variants_vs_reference(cfg, over, kw):
    """Model options outside the BASELINE configs that the CUDA path claims: shell-Hapke 1-3, no --mapping,
    analytic + learned normals together — oracle == live reference on every result key."""
    args = named_config(cfg, **over)
    ref_model = RH.build_model(args)
    batch = make_rays(24)
    S1, G = args.n_samples, args.guided_samples
    draws = RT.Draws.make(24, S1, G, S1 + G, seed=77)
    with torch.no_grad():
        ref, bt = RH.render(ref_model, args, batch.rays, draws, mode="test", **kw)
        ora, bt2, _ = RT.render_rays(RT.OracleModel(ref_model.state_dict(), args), args, batch.rays, draws, mode="test", **kw)
    assert bt == bt2 and set(ref) == set(ora)
    for k in ref:
        if ref[k].dtype == torch.int64 or k.startswith("z_vals"):
            assert torch.equal(ref[k], ora[k]), k
        else:
            assert (ref[k] - ora[k]).abs().max().item() <= 2e-6, k


@pytest.mark.parametrize("cfg,kw", [("lambertian", dict(mode="test")), ("rpv111", dict(mode="test", apply_brdf=True, cos_irra_on=True)),
                                    ("lambertian_viewdir", dict(mode="test"))])
def test_beta_channel_vs_reference(cfg, kw):
    """beta=True (transient-uncertainty head on [features | t-embedding], spsbrdfnerf.py:571-575, 708-711; rays_t =
    models['t'](ts), rendering.py:228-229): identical state_dict keys and seeded weights, identical result keys (the early
    return of `inference` is skipped: irradiance / rays_d / sun_d appear), every key equal to the live reference."""
    import torch
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.synth import make_rays
    from oracle import ref_harness as RH
    from oracle import render_torch as RT
    if not RH.available():
        pytest.skip("reference tree not available")
    args = named_config(cfg, beta=True)
    n = 12
    batch = make_rays(n)
    draws = RT.Draws.make(n, 64, 64, 128, seed=5)
    ref_model = RH.build_model(args, seed=0)
    torch.manual_seed(1)
    emb = torch.nn.Embedding(args.t_embbeding_vocab, args.t_embbeding_tau)
    ts = torch.arange(n) % 3
    with torch.no_grad():
        ref, bt = RH.render(ref_model, args, batch.rays, draws, ts=ts, embedding=emb, **kw)
    torch.manual_seed(0)
    sd = load_model(args).state_dict()
    rsd = ref_model.state_dict()
    assert list(sd) == list(rsd)
    assert all(torch.equal(sd[k], rsd[k]) for k in sd)
    with torch.no_grad():
        ora, bt2, _ = RT.render_rays(RT.OracleModel(sd, args), args, batch.rays, draws, rays_t=emb(ts), **kw)
    assert bt == bt2 and set(ora) == set(ref)
    assert "beta_coarse" in ref and ref["beta_coarse"].shape == (n, 128, 1) and "irradiance_coarse" in ref
    for k in ref:
        if ref[k].dtype.is_floating_point:
            assert (ora[k].float() - ref[k].float()).abs().max().item() <= 2e-6, k
