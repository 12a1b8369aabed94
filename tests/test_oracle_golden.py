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
        res, btype, _ = RT.render_rays(om, args, rays, draws, **kw, **(G.supervision(g) if ds else {}))
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
    for k in ("depth", "rgb", "weights", "albedo_accu", "sigmas", "nr_vw", "nr_sun", "brdf", "sun", "weights_sc"):
        if "ref_" + k in g:
            d = np.abs(res[k + "_coarse"].numpy() - g["ref_" + k]).max()
            assert d <= 2e-6, f"{name}: {k} differs from the reference by {d}"
