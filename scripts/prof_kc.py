"""K-C kernels (compositing, per-ray shading, per-sample BRDF, sample permutation) at inference-stream size for the ncu
captures of the HBM roofline (north star: achieved GB/s of the compositing / BRDF kernels, dram__bytes vs algorithmic bytes).
    python scripts/prof_kc.py [N]        N rays x 128 samples (default 65536); every kernel is launched 3 times"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import _lib as L  # noqa: E402
from brdf_nerf_b200 import ops  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402


def cfg_rpv(multi):
    c = L.ShadeCfg()
    c.n_channels = 19 if multi else 16
    c.normal_ch, c.param_ch, c.brdf_ch = 4, 7, (16 if multi else -1)
    c.brdf_type, c.funcM, c.funcF, c.funcH = L.BN_BRDF_RPV, 1, 1, 1
    c.multi_brdf, c.irr_mode, c.hpk_scl, c.fresnel_f0 = int(multi), L.BN_IRR_COS, 4.0, 0.04
    return c


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    S = 128
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    rays = make_rays(N).rays.to(dev)
    z = torch.sort(torch.rand(N, S, generator=g) * 0.6, dim=-1)[0].to(dev)
    reps = 3
    for C_ in (4, 16):
        packed = torch.rand(N, S, C_, generator=g).to(dev)
        if C_ == 16:
            packed[..., 4:7] = torch.nn.functional.normalize(packed[..., 4:7] - 0.5, dim=-1)
        for _ in range(reps):
            alpha, trans, w, depth, wsum, acc, acc_irr = ops.composite_forward(z, packed, None, 0.0)
        g_acc = torch.rand(N, C_, device=dev); g_depth = torch.rand(N, device=dev); g_wsum = torch.rand(N, device=dev)
        for _ in range(reps):
            ops.composite_backward(z, packed, None, 0.0, None, alpha, trans, w, g_acc, None, g_depth, g_wsum, None, None)
        if C_ == 16:
            cfg = cfg_rpv(False)
            for _ in range(reps):
                sh = ops.shade_rays_forward(cfg, rays, acc, wsum, None, None, want_normal=True, want_brdf=True)
            g_rgb = torch.rand(N, 3, device=dev)
            for _ in range(reps):
                ops.shade_rays_backward(cfg, rays, acc, wsum, None, None, g_rgb)
        del packed
    # per-sample BRDF (MultiBRDF): 16 channels + 3 BRDF channels per sample
    cfgm = cfg_rpv(True)
    packed = torch.rand(N, S, 19, generator=g).to(dev)
    packed[..., 4:7] = torch.nn.functional.normalize(packed[..., 4:7] - 0.5, dim=-1)
    for _ in range(reps):
        ops.brdf_points_forward(cfgm, rays, packed, want_aux=False)
    gp = torch.rand(N, S, 19, device=dev)
    for _ in range(reps):
        ops.brdf_points_backward(cfgm, rays, packed, gp)
    # sort_idx applied to the packed rows
    idx = torch.argsort(torch.rand(N, S, device=dev), dim=-1)
    rows = torch.rand(N * S, 16, device=dev)
    for _ in range(reps):
        ops.permute_samples(rows, idx, N, 64, 64, 16, scatter=False)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
