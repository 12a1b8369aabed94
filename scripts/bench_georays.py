"""Ray-feed micro-benchmark on one B200: all ray records of a 2048 x 2048 image from its RPC model in one launch
(bn_rays_from_rpc).  The reference's CPU path is timed beside it by bench.py (`tile_products` leg).
    python scripts/bench_georays.py        prints one JSON line"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import georays as PG  # noqa: E402
from brdf_nerf_b200.synth import synthetic_rpc_dict  # noqa: E402


def run(h=2048, w=2048):
    dev = torch.device("cuda:0")
    rpc = PG.RPCModel.from_dict(synthetic_rpc_dict(0))
    out = {"rays": h * w}
    for cs, center in (("utm", (436200.0, 3353400.0, 30.0)), ("ecef", (799000.0, -5452800.0, 3200200.0))):
        fn = lambda: PG.image_rays(rpc, h, w, -25.0, 95.0, cs, center, 400.0, 62.5, 148.0, device=dev, check=False)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        out[f"image_rays {cs}"] = {"us": us, "Mrays/s": h * w / us, "out_GB/s": h * w * 44 / us / 1e3}
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
