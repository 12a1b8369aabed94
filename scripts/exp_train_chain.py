"""Timing experiment on the training-forward chain kernel: which part of its epilogue is exposed?
BN_CHAIN_DBG bits (results are WRONG with any bit set; timing only): 1 no C store, 2 no cosine at all,
4 no H store, 8 no wait for the cosine box.   python scripts/exp_train_chain.py [n_rays] [S]"""
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
from bench_chain import timeit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
    for dbg in (0, 1, 2, 4, 8, 6, 7, 15):
        os.environ["BN_CHAIN_DBG"] = str(dbg)
        torch.manual_seed(0)
        m = load_model(args, precision="bf16").to(dev)
        m.sync_weights()
        flags = m.mlp_flags(train=True)
        C = m.out_channels(flags)
        packed = torch.empty((n, S, C), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, flags, tag="ws_train")
        us = timeit(lambda: ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, packed, C, ws), iters=20)
        print(f"BN_CHAIN_DBG={dbg:2d}: training forward (chain + heads) {us:8.1f} us  P={n * S}", flush=True)
        del m, ws


if __name__ == "__main__":
    main()
