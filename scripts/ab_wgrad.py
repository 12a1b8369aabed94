"""A/B runs inside ONE process (eager training steps of the benchmark shape, per-launch CUDA-event times from the library's
own profile hooks):
  * the weight-gradient GEMM's L2 prefetch distance (gemm_tc.cuh: nt_prefetch_distance, BN_NT_PREFETCH),
  * (r02n / r02o runs, profiles/: timing experiments through a BN_NT_EXP knob that dropped loads / bias MMAs / epilogue stores;
    the knob produced garbage results by design and was removed again once the answer was in),
  * density out of the fused trunk kernel vs the separate sigma GEMMs (BN_CHAIN_NO_SIG).
    python scripts/ab_wgrad.py [rays]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import _lib as L  # noqa: E402
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    batch = make_rays(n, depth_supervision=True).to(dev)
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(dev)
    tr = Trainer(model, args, use_graph=False)
    lib = L.load()
    lib.bn_profile_enable.restype = C.c_int
    for _ in range(3):
        tr.step(batch)
    torch.cuda.synchronize()
    for rep in range(3):
        for d, env in (("base", {}), ("prefetch 4", {"BN_NT_PREFETCH": "4"}), ("prefetch 8", {"BN_NT_PREFETCH": "8"}),
                       ("prefetch 16", {"BN_NT_PREFETCH": "16"}), ("B multicast (4-CTA)", {"BN_NT_MC": "1"}),
                       ("multicast + prefetch 8", {"BN_NT_MC": "1", "BN_NT_PREFETCH": "8"})):
            for k in ("BN_NT_PREFETCH", "BN_NT_EXP", "BN_CHAIN_NO_SIG", "BN_NT_MC"):
                os.environ.pop(k, None)
            os.environ.update(env)
            tr.step(batch)
            torch.cuda.synchronize()
            lib.bn_profile_enable(1)
            for _ in range(3):
                tr.step(batch)
            torch.cuda.synchronize()
            cnt = (C.c_longlong * 4)(); tms = (C.c_double * 4)(); work = (C.c_double * 4)()
            lib.bn_profile_collect(4, cnt, tms, work)
            lib.bn_profile_enable(0)
            print(f"{d:>18s} rep {rep}: wgrad (nt) {1e3 * tms[1] / 3:7.1f} us/step in {cnt[1] / 3:.0f} launches  "
                  f"{work[1] / max(tms[1], 1e-9) / 1e9:7.1f} TFLOP/s | tn {1e3 * tms[0] / 3:6.1f} us/step in {cnt[0] / 3:.0f} | "
                  f"chain {1e3 * tms[2] / max(cnt[2], 1):6.1f} us | dgrad chain {1e3 * tms[3] / max(cnt[3], 1):6.1f} us", flush=True)
    for k in ("BN_NT_PREFETCH", "BN_NT_EXP", "BN_CHAIN_NO_SIG", "BN_NT_MC"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
