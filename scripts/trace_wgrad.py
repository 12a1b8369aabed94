"""Timeline of one 512 x 512 weight-gradient GEMM (gemm_tc.cuh, NT pair kernel): clock64() stamps of CTA pair 0.
    python scripts/trace_wgrad.py [rays]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BN_NT_TRACE"] = "1"
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
    for _ in range(3):
        tr.step(batch)
    torch.cuda.synchronize()
    for rep in range(2):
        buf = torch.zeros(512 + 16, dtype=torch.int64, device=dev)
        L.check(L.load().bn_debug_chain_trace(model.handle(), C.c_void_p(buf.data_ptr())))
        tr.step(batch)
        torch.cuda.synchronize()
        L.check(L.load().bn_debug_chain_trace(model.handle(), None))
        t = [int(v) for v in buf[512:].cpu()]
        t0 = t[0]
        print(f"rep {rep}: setup done +{t[1] - t0}  first MMA +{t[2] - t0}  last MMA issued +{t[3] - t0}  accumulator complete +{t[6] - t0}  "
              f"epilogue done +{t[7] - t0}")
        print(f"    K blocks {t[5]}  issue phase {t[3] - t[2]} cycles = {(t[3] - t[2]) / max(t[5], 1):.0f} per K block (512 nominal)  "
              f"issuer waited {t[4]} for operands, producer waited {t[8]} for slots")


if __name__ == "__main__":
    main()
