"""Probe (r02): how well-conditioned the bf16-vs-fp32 comparison of the BRDF configurations is, to choose honest parity
statistics.  (1) render at random init: rgb error of the bf16 path against the fp32 mode as a function of the length of the
accumulated normal |sum_s w n| (the BRDF normalises it; a short vector amplifies any perturbation);
(2) PSNR-drift protocols: fp32 vs fp32 (noise floor of the statistic: atomics order), own-precision pretrain + BRDF stage."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.rendering import Draws, render_rays  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402

dev = torch.device("cuda:0")
CASES = {"rpv111": dict(apply_brdf=True, cos_irra_on=True),
         "hapke_bct": dict(apply_brdf=True, apply_theta=True, cos_irra_on=True),
         "microfacet": dict(apply_brdf=True, cos_irra_on=True)}


def mk_draws(n, seed, with_gt=False):
    g = torch.Generator().manual_seed(seed)
    return Draws(u_strat=torch.rand(n, 64, generator=g), u_pred=torch.rand(n, 64, generator=g),
                 u_gt=torch.rand(n, 64, generator=g) if with_gt else None)


def part1():
    n = 1024
    batch = make_rays(n, seed=5).to(dev)
    d = mk_draws(n, 6)
    for cfg, kw in CASES.items():
        args = named_config(cfg)
        res = {}
        for prec in ("fp32", "bf16"):
            torch.manual_seed(0)
            m = load_model(args, precision=prec).to(dev)
            with torch.no_grad():
                res[prec], _ = render_rays({"coarse": m}, args, batch.rays, None, _draws=d, **kw)
        a, b = res["fp32"], res["bf16"]
        acc = (a["weights_coarse"].unsqueeze(-1) * a["normal_an_coarse"]).sum(1)
        accb = (b["weights_coarse"].unsqueeze(-1) * b["normal_an_coarse"]).sum(1)
        ln = acc.norm(dim=-1)
        err = (a["rgb_coarse"] - b["rgb_coarse"]).abs().max(-1)[0]
        print(f"{cfg}: |acc normal| quantiles {[round(float(q), 3) for q in torch.quantile(ln, torch.tensor([0.01, 0.1, 0.5, 0.9], device=dev))]}"
              f"  acc-normal err max {(acc - accb).abs().max().item():.3e}")
        for thr in (0.0, 0.05, 0.1, 0.2, 0.3, 0.5):
            ok = ln >= thr
            if ok.sum() > 0:
                e = err[ok]
                print(f"   |acc n| >= {thr:4.2f}: {int(ok.sum()):5d} rays  rgb err max {e.max().item():.3e}  p99 {torch.quantile(e, 0.99).item():.3e} mean {e.mean().item():.3e}")
        for k in ("nr_vw_coarse", "nr_sun_coarse", "depth_coarse", "albedo_accu_coarse"):
            print(f"   {k}: max err {(a[k] - b[k]).abs().max().item():.3e}")


def psnr(model, args, batch, evs, kw):
    out = 0.0
    for d in evs:
        with torch.no_grad():
            res, _ = render_rays({"coarse": model}, args, batch.rays, None, _draws=d, **kw)
        out += -10.0 * math.log10(((res["rgb_coarse"] - batch.rgbs) ** 2).mean().item())
    return out / len(evs)


def run(cfg, kw, precision, seed, n, pre, steps, batch, evs, pre_precision=None):
    args = named_config(cfg, ds_lambda=10.0)
    torch.manual_seed(seed)
    model = load_model(args, precision=pre_precision or precision).to(dev)
    tr = Trainer(model, args)
    for i in range(pre):
        tr.step(batch, draws=mk_draws(n, 1000 * seed + i, with_gt=True))
    p0 = psnr(model, args, batch, evs, {})
    if pre_precision and pre_precision != precision:
        model.set_precision(precision)
    curve = []
    for i in range(steps):
        tr.step(batch, draws=mk_draws(n, 1000 * seed + 500 + i, with_gt=True), **kw)
        if (i + 1) % 20 == 0:
            curve.append(round(psnr(model, args, batch, evs, kw), 3))
    return p0, curve


def part2():
    n = 256
    for cfg, kw in CASES.items():
        for seed in (0, 1):
            batch = make_rays(n, seed=20240912 + seed, depth_supervision=True).to(dev)
            evs = [mk_draws(n, 9999 + seed), mk_draws(n, 7777 + seed)]
            for pre in (150,):
                r = {}
                for tag, prec in (("fp32", "fp32"), ("fp32'", "fp32"), ("bf16", "bf16")):
                    r[tag] = run(cfg, kw, prec, seed, n, pre, 100, batch, evs)
                    print(f"{cfg} seed {seed} pretrain {pre} [{tag}]: lambertian PSNR {r[tag][0]:.3f}  BRDF-stage curve {r[tag][1]}", flush=True)
                print(f"   noise floor (fp32 vs fp32') {[round(a - b, 3) for a, b in zip(r['fp32'][1], r[chr(102) + 'p32' + chr(39)][1])]}")
                print(f"   drift bf16 - fp32          {[round(b - a, 3) for a, b in zip(r['fp32'][1], r['bf16'][1])]}", flush=True)


def part3():
    """Gentler BRDF stage from a COMMON fp32-pretrained checkpoint: is there a protocol whose fp32-vs-fp32 floor is below
    0.1 dB, so that the north-star criterion can be resolved?  lr of the BRDF stage 5e-4 (reference default), 1e-4, 2e-5."""
    import copy
    n = 256
    for cfg, kw in CASES.items():
        for seed in (0, 1):
            batch = make_rays(n, seed=20240912 + seed, depth_supervision=True).to(dev)
            evs = [mk_draws(n, 9999 + seed), mk_draws(n, 7777 + seed)]
            args = named_config(cfg, ds_lambda=10.0)
            torch.manual_seed(seed)
            model = load_model(args, precision="fp32").to(dev)
            tr = Trainer(model, args)
            for i in range(150):
                tr.step(batch, draws=mk_draws(n, 1000 * seed + i, with_gt=True))
            ckpt = copy.deepcopy(tr.state_dict())
            for lr in (5e-4, 1e-4, 2e-5):
                res = {}
                for tag, prec in (("fp32", "fp32"), ("fp32b", "fp32"), ("bf16", "bf16")):
                    torch.manual_seed(seed)
                    m = load_model(args, precision=prec).to(dev)
                    t2 = Trainer(m, args)
                    t2.load_state_dict(ckpt)
                    t2.lr = lr
                    curve = []
                    for i in range(60):
                        t2.step(batch, draws=mk_draws(n, 1000 * seed + 500 + i, with_gt=True), **kw)
                        if (i + 1) % 20 == 0:
                            curve.append(round(psnr(m, args, batch, evs, kw), 3))
                    res[tag] = curve
                print(f"{cfg} seed {seed} lr {lr:g}: fp32 {res['fp32']}  floor {[round(a - b, 3) for a, b in zip(res['fp32b'], res['fp32'])]}"
                      f"  drift {[round(a - b, 3) for a, b in zip(res['bf16'], res['fp32'])]}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("1", "all"):
        part1()
    if which in ("2", "all"):
        part2()
    if which in ("3", "all"):
        part3()
