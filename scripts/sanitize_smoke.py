"""One small pass over every kernel family for compute-sanitizer: bf16 training step (fused trunk forward,
tcgen05 GEMMs, fused data-gradient chain, heads, compositing, sampler, loss, Adam), a BRDF-stage step with analytic normals
(second-order kernels), an inference render and the fp32 mode.
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.rendering import render_rays  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for cfg, kw, prec in (("lambertian_ds", {}, "bf16"), ("rpv111", dict(apply_brdf=True, cos_irra_on=True), "bf16"),
                          ("lambertian_ds", {}, "fp32")):
        args = named_config(cfg)
        torch.manual_seed(0)
        model = load_model(args, precision=prec).to(dev)
        batch = make_rays(n, depth_supervision=cfg.endswith("_ds")).to(dev)
        tr = Trainer(model, args, use_graph=False)
        for _ in range(2):
            loss = tr.step(batch, **kw)
        with torch.no_grad():
            res, _ = render_rays({"coarse": model}, args, batch.rays, None, **kw)
        torch.cuda.synchronize()
        print(f"{cfg} [{prec}] loss {float(loss):.5f} rgb mean {res['rgb_coarse'].mean().item():.4f}", flush=True)
    print("sanitize_smoke ok")


if __name__ == "__main__":
    main()
