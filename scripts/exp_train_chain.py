"""Times the training forward of the MLP (fused trunk chain kernel + feature layer + heads) at a given size.
    python scripts/exp_train_chain.py [n_rays] [S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import ops  # noqa: E402
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from bench_chain import timeit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(dev)
    m.sync_weights()
    for train in (True, False):
        flags = m.mlp_flags(train=train)
        C = m.out_channels(flags)
        packed = torch.empty((n, S, C), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, flags, tag="ws_train" if train else "ws_full")
        us_all = timeit(lambda: ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, packed, C, ws), iters=20)
        us_trunk = timeit(lambda: ops.mlp_trunk_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, n * S, 0, None, ws), iters=20)
        print(f"{'training' if train else 'inference'} forward P={n * S}: trunk chain {us_trunk:8.1f} us "
              f"({2.0 * n * S * 1896448 / us_trunk / 1e6:6.1f} TFLOP/s), trunk + heads {us_all:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
