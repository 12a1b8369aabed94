"""ORACLE (test infrastructure only).  Generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, torch CPU fp32) with injected random draws on small synthetic batches.
Run here (the reference tree is not available on the GPU box):

    python -m oracle.make_golden

Each file holds the inputs (rays, draws, supervision), a checksum of the seeded model weights
(seed 0, `load_model(args)` — brdf_nerf_b200.models.load_model reproduces them bit for bit) and the
reference outputs that the parity tests compare against.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from brdf_nerf_b200.config import named_config          # noqa: E402
from brdf_nerf_b200.synth import make_rays              # noqa: E402
from oracle import ref_harness as RH                    # noqa: E402
from oracle import render_torch as RT                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N = 48

CASES = {
    # name: (config, config overrides, render kwargs, depth supervision, zero_std)
    "lambertian_test": ("lambertian", {}, dict(mode="test"), False, False),
    "lambertian_ds_train": ("lambertian_ds", {}, dict(mode="train"), True, False),
    "lambertian_ds_std0_train": ("lambertian_ds", {}, dict(mode="train"), True, True),
    "lambertian_gsam_only": ("lambertian", {}, dict(mode="test", gsam_only=True), False, False),
    "rpv111_brdf": ("rpv111", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    "rpv111_multi_brdf": ("rpv111_multi", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    "hapke_bct_brdf": ("hapke_bct", {}, dict(mode="test", apply_brdf=True, apply_theta=True, cos_irra_on=True), False, False),
    "hapke_b_brdf": ("hapke_b", {}, dict(mode="test", apply_brdf=True), False, False),
    "microfacet_brdf": ("microfacet", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    "rpv111_learned_normal": ("rpv111", dict(normal="learned"), dict(mode="test", apply_brdf=True), False, False),
    "lambertian_viewdir_test": ("lambertian_viewdir", {}, dict(mode="test"), False, False),
    "rpv111_sunvis_test": ("rpv111", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True, bTestSun_v=True), False, False),
    # model options outside the BASELINE configs
    "hapke_shell1_brdf": ("hapke_b", dict(b=0, shell_hapke=1), dict(mode="test", apply_brdf=True), False, False),
    "hapke_shell2_brdf": ("hapke_b", dict(b=0, shell_hapke=2), dict(mode="test", apply_brdf=True), False, False),
    "hapke_shell3_brdf": ("hapke_b", dict(b=0, shell_hapke=3), dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    "lambertian_nomapping_test": ("lambertian", dict(mapping=False), dict(mode="test"), False, False),
    "rpv111_nomapping_brdf": ("rpv111", dict(mapping=False), dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    "rpv111_an_lr_normals": ("rpv111", dict(normal="analystic_learned"), dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
    # transient-uncertainty channel: beta head on [features | t-embedding], rays_t = models['t'](ts) (rendering.py:228-229)
    "lambertian_beta_test": ("lambertian", dict(beta=True), dict(mode="test"), False, False),
    "rpv111_beta_brdf": ("rpv111", dict(beta=True), dict(mode="test", apply_brdf=True, cos_irra_on=True), False, False),
}
KEEP = ("z_vals", "z_vals_unsort", "sort_idx", "depth", "rgb", "weights", "albedo_accu", "sigmas", "nr_vw", "nr_sun",
        "brdf", "hpk_scl", "sun", "weights_sc", "beta")


def time_inputs(args, n):
    """ts (view index of every ray) and the embedding table models['t'] of a beta case: nn.Embedding seeded with 1,
    created after the model (main.py:113-118)."""
    torch.manual_seed(1)
    emb = torch.nn.Embedding(args.t_embbeding_vocab, args.t_embbeding_tau)
    return torch.arange(n) % 3, emb


def weights_digest(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    only = set(sys.argv[1:])            # optional: regenerate only the named cases
    os.makedirs(OUT, exist_ok=True)
    for name, (cfg, over, kw, ds, zero_std) in CASES.items():
        if only and name not in only:
            continue
        args = named_config(cfg, **over)
        model = RH.build_model(args, seed=0)
        batch = make_rays(N, depth_supervision=ds, zero_std=zero_std)
        S1, G = args.n_samples, args.guided_samples
        gs = bool(kw.get("gsam_only"))
        S = G if gs else S1 + G
        sun = bool(kw.get("bTestSun_v"))
        draws = RT.Draws.make(N, S1, G, S, seed=4321, with_gt=ds, with_sun=sun, s_sun=G if gs else S1)
        extra = dict(valid_depth=batch.valid_depth, target_depths=batch.target_depths, target_std=batch.target_std) if ds else {}
        tkw = {}
        if args.beta:
            ts, emb = time_inputs(args, N)
            tkw = dict(ts=ts, embedding=emb)
        with torch.no_grad():
            res, btype = RH.render(model, args, batch.rays, draws, **tkw, **kw, **extra)
        # normal accumulation used for the tolerance statement on normals (SURVEY §8a N-note)
        out = {"rays": batch.rays.numpy(), "u_strat": draws.u_strat.numpy(), "u_pred": draws.u_pred.numpy(),
               "brdf_type": np.array(btype), "weights_sha256": np.array(weights_digest(model.state_dict()))}
        if ds:
            out.update(valid_depth=batch.valid_depth.numpy(), target_depths=batch.target_depths.numpy(),
                       target_std=batch.target_std.numpy(), u_gt=draws.u_gt.numpy())
        if sun:
            out.update(u_sun=draws.u_sun.numpy())
        if args.beta:
            out.update(ts=tkw["ts"].numpy(), t_weight=tkw["embedding"].weight.detach().numpy())
        for k in KEEP:
            if f"{k}_coarse" in res:
                out["ref_" + k] = res[f"{k}_coarse"].detach().numpy()
        for nk in ("normal_an", "normal_lr"):
            if f"{nk}_coarse" in res:
                nrm = res[f"{nk}_coarse"]
                out[f"ref_{nk}_acc"] = (res["weights_coarse"].unsqueeze(-1) * nrm).sum(1).numpy()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:28s} {btype:10s} {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
