"""CPU: the oracle restatement against the golden vectors produced by the live reference
(oracle/make_golden.py).  z-values / indices bit-exact; everything else <= 2e-6 (same torch ops)."""
import numpy as np
import pytest
import torch

from brdf_nerf_b200.models import load_model
from oracle import render_torch as RT

import _golden as G


@pytest.mark.parametrize("name", G.names())
def test_oracle_matches_reference_golden(name):
    g, args, kw, ds = G.load(name)
    torch.manual_seed(0)
    model = load_model(args)
    state = model.state_dict()
    assert G.weights_digest(state) == str(g["weights_sha256"]), "seeded init no longer reproduces the reference weights"
    rays = torch.from_numpy(g["rays"])
    N = rays.shape[0]
    S = g["ref_z_vals"].shape[1]
    zeros = lambda *s: torch.zeros(*s)
    draws = RT.Draws(u_strat=torch.from_numpy(g["u_strat"]), noise1=zeros(N, args.n_samples),
                     u_pred=torch.from_numpy(g["u_pred"]), noise2=zeros(N, S),
                     u_gt=torch.from_numpy(g["u_gt"]) if ds else None,
                     u_sun=torch.from_numpy(g["u_sun"]) if "u_sun" in g else None,
                     noise_sun=zeros(*g["u_sun"].shape) if "u_sun" in g else None)
    om = RT.OracleModel(state, args)
    with torch.no_grad():
        ts, emb = G.time_embedding(g)
        tkw = {} if ts is None else dict(rays_t=emb(ts))
        res, btype, _ = RT.render_rays(om, args, rays, draws, **tkw, **kw, **(G.supervision(g) if ds else {}))
    assert btype == str(g["brdf_type"])
    assert G.bits_equal(res["z_vals_coarse"].numpy(), g["ref_z_vals"]) == 0
    assert G.bits_equal(res["z_vals_unsort_coarse"].numpy(), g["ref_z_vals_unsort"]) == 0
    if "ref_sort_idx" in g:
        # torch.sort leaves the order of equal keys unspecified: compare the permutation modulo ties
        idx, ridx = res["sort_idx_coarse"].numpy(), g["ref_sort_idx"]
        un = g["ref_z_vals_unsort"]
        assert np.array_equal(np.sort(idx, -1), np.sort(ridx, -1))
        assert G.bits_equal(np.take_along_axis(un, idx, 1), np.take_along_axis(un, ridx, 1)) == 0
        ties = (np.diff(g["ref_z_vals"], axis=-1) == 0).any()
        assert ties or np.array_equal(idx, ridx)
    for k in ("depth", "rgb", "weights", "albedo_accu", "sigmas", "nr_vw", "nr_sun", "brdf", "sun", "weights_sc", "beta"):
        if "ref_" + k in g:
            d = np.abs(res[k + "_coarse"].numpy() - g["ref_" + k]).max()
            assert d <= 2e-6, f"{name}: {k} differs from the reference by {d}"


@pytest.mark.parametrize("cfg,kw", [("lambertian", {}), ("rpv111", dict(apply_brdf=True, cos_irra_on=True)),
                                    ("hapke_bct", dict(apply_brdf=True, apply_theta=True, cos_irra_on=True))])
def test_fp64_arbiter_noise_floor(cfg, kw):
    """The float64 run of the oracle is the arbiter for the tolerance statement (SURVEY §8a N-note): the reference
    algorithm's OWN float32 rounding moves rgb by ~1e-5..1e-4 and the accumulated normal by ~5e-5 — two orders below the
    1e-3 bar the CUDA path is held to — while the RAW per-sample analytic normal differs by more than 1e-3 where |grad sigma|
    is tiny, which is why parity on normals is stated on sum_s w n (and on nr_vw / nr_sun), not per sample."""
    import torch
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.synth import make_rays
    from oracle import render_torch as RT
    args = named_config(cfg)
    n = 64
    torch.manual_seed(0)
    state = load_model(args).state_dict()
    batch = make_rays(n)
    S1, G = args.n_samples, args.guided_samples
    draws = RT.Draws.make(n, S1, G, S1 + G, seed=3)
    with torch.no_grad():
        o32, _, _ = RT.render_rays(RT.OracleModel(state, args), args, batch.rays, draws, mode="test", **kw)
        o64, _, _ = RT.render_rays(RT.OracleModel(state, args, dtype=torch.float64), args, batch.rays, draws, mode="test", **kw)
    assert o64["rgb_coarse"].dtype == torch.float64
    for k in ("rgb_coarse", "depth_coarse", "weights_coarse", "albedo_accu_coarse"):
        d = (o32[k].double() - o64[k]).abs().max().item()
        assert d <= 2e-4, f"{k}: float32 vs float64 {d}"
    if "normal_an_coarse" in o32:
        acc32 = (o32["weights_coarse"].unsqueeze(-1) * o32["normal_an_coarse"]).sum(1).double()
        acc64 = (o64["weights_coarse"].unsqueeze(-1) * o64["normal_an_coarse"]).sum(1)
        assert (acc32 - acc64).abs().max().item() <= 2e-4
        raw = (o32["normal_an_coarse"].double() - o64["normal_an_coarse"]).abs().max().item()
        assert raw > 1e-4                                           # the raw normals are NOT reproducible to the accumulated level
