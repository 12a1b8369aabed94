"""A/B timing of the fused trunk kernels inside ONE process (box-to-box spread is +-3 %, so variants are only comparable
when they alternate on the same GPU): training forward chain with the one-pass / two-pass second-half epilogue, the density
chain, and the backward with / without the fused data-gradient chain.
    python scripts/ab_chain.py [rays]        (default 1024 rays x 64 / 128 samples, the bench step's shapes)"""
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
from brdf_nerf_b200.train import Trainer  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
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
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    F, Lr, E = 512, 8, 60
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(dev)
    m.sync_weights()
    for S in (64, 128):
        z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
        flops = 2.0 * n * S * (E * F + (Lr - 2) * F * F + (F + E) * F)
        flags = m.mlp_flags(train=True)
        ws = m.workspace(n * S, flags, tag="ws_train")
        sig = torch.empty((n, S), dtype=torch.float32, device=dev)
        fn = lambda: ops.mlp_trunk_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, n * S, 0, None, ws)
        for rep in range(3):
            for tag, envs in (("default", {}), ("two c boxes", {"BN_CHAIN_CBOX2": "1"}), ("no L2 hints", {"BN_NO_L2_HINTS": "1"})):
                for k in ("BN_NO_L2_HINTS",):
                    os.environ.pop(k, None)
                if "BN_CHAIN_CBOX2" in envs:
                    continue          # read once per process (static): timed by its own run of this script
                os.environ.update(envs)
                us = timeit(fn)
                print(f"train chain {tag:16s} P={n * S:7d} rep {rep}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
        os.environ.pop("BN_NO_L2_HINTS", None)
        ws1 = m.workspace(n * S, L.MLP_SIGMA_ONLY, tag="ws_sigma")
        fs = lambda: ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, L.MLP_SIGMA_ONLY, sig, 1, ws1)
        us = timeit(fs)
        print(f"density chain         P={n * S:7d}       : {us:8.1f} us  {(flops + 2.0 * n * S * F) / us / 1e6:7.1f} TFLOP/s", flush=True)
    # whole training step with / without the fused data-gradient chain (the knob is read when the handle is created)
    batch = make_rays(n, depth_supervision=True).to(dev)
    res = {}
    for rep in range(2):
        for tag, env, env2 in (("dgrad chain W4C2", None, None), ("per-layer dgrad", "1", None), ("dgrad chain W3C3", None, "1")):
            for k, v in (("BN_NO_DGRAD_CHAIN", env), ("BN_DCHAIN_W3C3", env2)):
                if v:
                    os.environ[k] = v
                else:
                    os.environ.pop(k, None)
            torch.manual_seed(0)
            mm = load_model(args, precision="bf16").to(dev)
            tr = Trainer(mm, args, use_graph=True)
            us = timeit(lambda: tr.step(batch), iters=30, warm=5)
            lib = L.load()
            lib.bn_profile_enable(1)
            eager = Trainer(mm, args, use_graph=False)
            for _ in range(3):
                eager.step(batch)
            import ctypes as C_
            cnt = (C_.c_longlong * 4)(); tms = (C_.c_double * 4)(); work = (C_.c_double * 4)()
            lib.bn_profile_collect(4, cnt, tms, work)
            lib.bn_profile_enable(0)
            kinds = "  ".join(f"{nm} {tms[i] / 3 * 1e3:7.1f} us ({cnt[i] // 3})" for i, nm in enumerate(("tn", "nt_wgrad", "chain_fwd", "chain_dgrad")))
            print(f"training step, {tag:20s} rep {rep}: {us:8.1f} us/step  {n / us * 1e3:8.1f} k rays/s | per step (serialised): {kinds}", flush=True)
            del tr, mm
            torch.cuda.empty_cache()
    os.environ.pop("BN_NO_DGRAD_CHAIN", None)
    os.environ.pop("BN_DCHAIN_W3C3", None)


if __name__ == "__main__":
    main()
