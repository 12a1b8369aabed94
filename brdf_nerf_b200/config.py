"""Configuration for the ray-rendering hot path.

`make_args(**overrides)` returns an ``argparse.Namespace`` carrying the subset of the
reference's training options that the hot path reads (reference ``opt.py:126-354``; the
attributes consumed are listed in SURVEY.md §5 "Config / flags").  Defaults equal the
reference's defaults, except ``model`` which is fixed to ``spsbrdf-nerf`` (the only variant this
package implements) — so a Namespace produced by the reference's own ``Train_parser`` is accepted
unchanged.

`PathConfig.from_args` freezes a Namespace into the plain-old-data view that is handed to the
C ABI (``include/brdfnerf_b200.h``: ``bn_model_cfg`` / ``bn_render_cfg``).
"""
from __future__ import annotations

import argparse
import dataclasses
from typing import Optional

# name -> default, mirroring reference opt.py (line numbers in comments)
_DEFAULTS = dict(
    data="sat",            # opt.py:148
    model="spsbrdf-nerf",  # opt.py:150 (reference default is sps-nerf)
    gpu_id=1,              # opt.py:152
    lr=5e-4,               # opt.py:156
    batch_size=1024,       # opt.py:158
    fc_feat=512,           # opt.py:172
    fc_layers=8,           # opt.py:174
    n_samples=64,          # opt.py:176
    n_importance=0,        # opt.py:178
    noise_std=0.0,         # opt.py:180
    chunk=1024 * 5,        # opt.py:182
    lambda_rgb=1.0,        # opt.py:184
    sc_lambda=0.0,         # opt.py:186
    ds_lambda=0.0,         # opt.py:190
    ds_drop=1.0,           # opt.py:193
    ds_noweights=False,
    t_embbeding_tau=4,     # opt.py:201
    t_embbeding_vocab=30,
    beta=False,            # opt.py:209
    mapping=False,         # opt.py:211
    GNLL=False,
    usealldepth=False,
    guided_samples=64,     # opt.py:217
    margin=1e-4,           # opt.py:219
    stdscale=1.0,          # opt.py:221
    siren=1,               # opt.py:225
    indirect_light=False,
    normal="none",         # opt.py:230
    sun_v="none",          # opt.py:231
    max_train_steps=300000,  # opt.py:162
    brdf_on=1.0,           # opt.py:242
    nrrg_on=0.0,           # opt.py:244
    nr_reg_an_lambda=0.0,  # opt.py:232
    nr_reg_lr_lambda=0.0,  # opt.py:234
    hs_lambda=0.0,         # opt.py:240
    gsam_only_on=1.0,      # opt.py:255
    cos_irra_on=1.0,       # opt.py:257
    std_range=3.0,         # opt.py:259
    MultiBRDF=0,           # opt.py:261
    roughness=False,
    glossy_scale=1.0,
    fresnel_f0=0.04,
    hpk_scl=4.0,           # opt.py:283
    shell_hapke=0,         # opt.py:285
    b=0, c=0, B0=0, h=0, theta=0,   # opt.py:287-296
    funcM=0, funcF=0, funcH=0,      # opt.py:302-307
    dim_RPV=1,             # opt.py:308
    input_viewdir=0,       # opt.py:317
    print_debuginfo=False,
)


def make_args(**overrides) -> argparse.Namespace:
    """Namespace with the reference's defaults for every option the hot path reads."""
    unknown = set(overrides) - set(_DEFAULTS)
    if unknown:
        raise TypeError(f"unknown option(s) for the render path: {sorted(unknown)}")
    d = dict(_DEFAULTS)
    d.update(overrides)
    ns = argparse.Namespace(**d)
    # reference opt.py:340-341: solar-correction weight only survives with a learned sun net
    if ns.sun_v != "learned":
        ns.sc_lambda = 0.0
    return ns


def named_config(name: str, **overrides) -> argparse.Namespace:
    """The five BASELINE.json configurations (SURVEY.md §8d 'configs -> kwargs')."""
    base = dict(mapping=True)
    table = {
        "lambertian": {},
        "lambertian_ds": dict(ds_lambda=10.0),
        "lambertian_viewdir": dict(input_viewdir=1),
        "rpv111_viewdir": dict(funcM=1, funcF=1, funcH=1, dim_RPV=1, normal="analystic", input_viewdir=1),
        "rpv111": dict(funcM=1, funcF=1, funcH=1, dim_RPV=1, normal="analystic"),
        "rpv111_multi": dict(funcM=1, funcF=1, funcH=1, dim_RPV=3, normal="analystic", MultiBRDF=1),
        "hapke_bct": dict(b=1, c=1, theta=1, normal="analystic"),
        "hapke_b": dict(b=1, normal="analystic"),
        "microfacet": dict(roughness=True, normal="analystic"),
    }
    if name not in table:
        raise KeyError(f"unknown config {name!r}; have {sorted(table)}")
    base.update(table[name])
    base.update(overrides)
    return make_args(**base)


@dataclasses.dataclass(frozen=True)
class PathConfig:
    """Frozen POD view of the options; one instance per model handle."""
    feat: int
    layers: int
    n_freq_xyz: int        # 0 => no positional encoding (raw xyz)
    n_freq_dir: int
    input_viewdir: bool
    skip_layer: int
    normal_an: bool
    normal_lr: bool
    brdf: str              # 'none' | 'microfacet' | 'rpv' | 'hapke'
    funcM: int
    funcF: int
    funcH: int
    dim_rpv: int
    hapke_b: bool
    hapke_c: bool
    hapke_theta: bool
    shell_hapke: int
    multi_brdf: bool
    hpk_scl: float
    n_samples: int
    guided_samples: int
    std_range: float
    noise_std: float

    @staticmethod
    def from_args(args, mapping_sizes=(10, 4), skips=(4,)) -> "PathConfig":
        if args.model != "spsbrdf-nerf":
            raise ValueError("only --model spsbrdf-nerf is implemented on this path")
        if not args.siren:
            raise NotImplementedError("siren=0 (ReLU trunk) is outside the hot path scope")
        rpv = bool(args.funcM or args.funcF or args.funcH)
        if args.roughness:
            brdf = "microfacet"      # head priority: reference spsbrdfnerf.py:483-496
        elif rpv:
            brdf = "rpv"
        elif args.b or args.shell_hapke > 0:
            brdf = "hapke"
        else:
            brdf = "none"
        return PathConfig(
            feat=int(args.fc_feat), layers=int(args.fc_layers),
            n_freq_xyz=int(mapping_sizes[0]) if args.mapping else 0,
            n_freq_dir=int(mapping_sizes[1]) if args.mapping else 0,
            input_viewdir=bool(args.input_viewdir), skip_layer=int(skips[0]),
            normal_an=args.normal in ("analystic", "analystic_learned"),
            normal_lr=args.normal in ("learned", "analystic_learned"),
            brdf=brdf, funcM=int(args.funcM), funcF=int(args.funcF), funcH=int(args.funcH),
            dim_rpv=int(args.dim_RPV), hapke_b=bool(args.b), hapke_c=bool(args.c),
            hapke_theta=bool(args.theta), shell_hapke=int(args.shell_hapke),
            multi_brdf=bool(args.MultiBRDF), hpk_scl=float(args.hpk_scl),
            n_samples=int(args.n_samples), guided_samples=int(args.guided_samples),
            std_range=float(args.std_range), noise_std=float(args.noise_std),
        )
