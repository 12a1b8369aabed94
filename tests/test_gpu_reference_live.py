"""The UNMODIFIED reference, executed on the GPU box itself (files staged in oracle/_ref by oracle/stage_ref.py; the live
tree in the build container), against the CUDA path on the same rays and the same injected random draws — the direct form of
"results identical to the reference's on the same inputs", at sizes beyond the committed 48-ray goldens.
fp32 mode: <= 1e-3 abs on every compared key (north star); bf16 mode: the tolerances of test_gpu_bf16_parity.py."""
import numpy as np
import pytest
import torch

from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.rendering import Draws, render_rays
from brdf_nerf_b200.synth import make_rays
from oracle import ref_harness as RH
from oracle import render_torch as RT

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(RH.kind() == "absent", reason="reference files neither live nor staged")]

CASES = [("lambertian_ds", True, dict(mode="train")),
         ("rpv111", False, dict(mode="test", apply_brdf=True, cos_irra_on=True)),
         ("hapke_bct", False, dict(mode="test", apply_brdf=True, apply_theta=True, cos_irra_on=True)),
         ("microfacet", False, dict(mode="test", apply_brdf=True, cos_irra_on=True))]


@pytest.mark.parametrize("cfg,ds,kw", CASES)
def test_cuda_path_vs_unmodified_reference(cuda, cfg, ds, kw):
    args = named_config(cfg)
    n = 512
    batch = make_rays(n, seed=777, depth_supervision=ds)
    S1, G = args.n_samples, args.guided_samples
    draws = RT.Draws.make(n, S1, G, S1 + G, seed=2468, with_gt=ds)
    sup = dict(valid_depth=batch.valid_depth, target_depths=batch.target_depths, target_std=batch.target_std) if ds else {}
    ref_model = RH.build_model(args, seed=0)
    with torch.no_grad():
        ref, ref_type = RH.render(ref_model, args, batch.rays, draws, **kw, **sup)
    gb = batch.to(cuda)
    supg = dict(valid_depth=gb.valid_depth, target_depths=gb.target_depths, target_std=gb.target_std) if ds else {}
    d = Draws(u_strat=draws.u_strat, u_pred=draws.u_pred, u_gt=draws.u_gt if ds else None)
    # (max abs, mean abs); for the bf16 path the quantities behind the per-sample analytic normal are bounded on the mean
    # and on the 99th percentile (ill-conditioned outliers where |grad sigma| ~ 0, see test_gpu_bf16_parity.py)
    TOL = {"fp32": dict(rgb=(1e-3, 1e-3), depth=(1e-3, 1e-3), weights=(1e-3, 1e-3), albedo_accu=(1e-3, 1e-3), z_vals=(1e-4, 1e-4)),
           "bf16": dict(rgb=(None, 4e-3), depth=(2e-3, 2e-4), weights=(8e-3, 2e-4), albedo_accu=(5e-3, 1e-3), z_vals=(3e-3, 2e-4))}
    for precision in ("fp32", "bf16"):
        torch.manual_seed(0)
        model = load_model(args, precision=precision).to(cuda)
        with torch.no_grad():
            res, btype = render_rays({"coarse": model}, args, gb.rays, None, _draws=d, **kw, **supg)
        assert btype == ref_type
        assert set(res) == set(ref), (sorted(set(res) ^ set(ref)))
        for k, (tmax, tmean) in TOL[precision].items():
            e = (res[k + "_coarse"].cpu() - ref[k + "_coarse"]).abs()
            p99 = torch.quantile(e.flatten()[:4_000_000].float(), 0.99).item()
            print(f"{cfg} [{precision}] {k} vs the unmodified reference: max {e.max().item():.3e}  p99 {p99:.3e}  mean {e.mean().item():.3e}")
            if tmax is not None:
                assert e.max().item() <= tmax, (cfg, precision, k, e.max().item())
            else:
                assert p99 <= 5e-2, (cfg, precision, k, p99)
            assert e.mean().item() <= tmean, (cfg, precision, k, e.mean().item())
        if "normal_an_coarse" in ref:
            acc = (res["weights_coarse"].unsqueeze(-1) * res["normal_an_coarse"]).sum(1).cpu()
            want = (ref["weights_coarse"].unsqueeze(-1) * ref["normal_an_coarse"]).sum(1)
            e = (acc - want).abs()
            print(f"{cfg} [{precision}] accumulated normal: max {e.max().item():.3e} mean {e.mean().item():.3e}")
            if precision == "fp32":
                assert e.max().item() <= 1e-3
            else:
                assert e.mean().item() <= 8e-3
