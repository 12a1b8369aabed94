"""Density pass micro-benchmark: fused chain kernel vs per-layer GEMMs (BN_NO_CHAIN=1).
    python scripts/bench_chain.py [n_rays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import _lib as L  # noqa: E402
from brdf_nerf_b200 import ops  # noqa: E402
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = 64
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
    F, Lr, E = 512, 8, 60
    flops = 2.0 * n * S * (E * F + (Lr - 2) * F * F + (F + E) * F + F)
    for env in ("", "1"):
        if env:
            os.environ["BN_NO_CHAIN"] = env
        torch.manual_seed(0)
        m = load_model(args, precision="bf16").to(dev)
        m.sync_weights()
        sig = torch.empty((n, S), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, L.MLP_SIGMA_ONLY, tag="ws_sigma")
        fn = lambda: ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, L.MLP_SIGMA_ONLY, sig, 1, ws)
        us = timeit(fn)
        print(f"{'per-layer GEMMs' if env else 'fused chain    '}  P={n * S:8d}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


def train_fwd(n=1024, S=128):
    """training forward (chain kernel + feats + heads GEMMs) vs per-layer GEMMs (BN_NO_CHAIN=1)"""
    import ctypes as C_
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
    lib = L.load()
    for env in ("", "1"):
        if env:
            os.environ["BN_NO_CHAIN"] = env
        torch.manual_seed(0)
        m = load_model(args, precision="bf16").to(dev)
        m.sync_weights()
        flags = m.mlp_flags(train=True)
        C = m.out_channels(flags)
        packed = torch.empty((n, S, C), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, flags, tag="ws_train")
        for _ in range(3):
            ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, packed, C, ws)
        lib.bn_profile_enable(1)
        for _ in range(5):
            ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, packed, C, ws)
        cnt = (C_.c_longlong * 2)(); tms = (C_.c_double * 2)(); work = (C_.c_double * 2)()
        lib.bn_profile_collect(2, cnt, tms, work)
        lib.bn_profile_enable(0)
        print(f"{'per-layer GEMMs' if env else 'fused chain    '}: {cnt[0] // 5} GEMM launches, {tms[0] / 5 * 1e3:8.1f} us, "
              f"{work[0] / max(tms[0], 1e-9) / 1e9:7.1f} TFLOP/s (training forward, P = {n * S})", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_fwd()
        sys.exit(0)
    main()
