"""Ray-feed micro-benchmark on one B200: all ray records of a 2048 x 2048 image from its RPC model in one launch
(bn_rays_from_rpc), next to the reference's CPU path (numpy restatement of rpcm's localisation + get_rays) on a strided
sample.      python scripts/bench_georays.py        prints one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import georays as PG  # noqa: E402
from oracle import georays_np as G  # noqa: E402


def run(h=2048, w=2048):
    dev = torch.device("cuda:0")
    o = G.synthetic_rpc(0)
    rpc = PG.RPCModel.from_dict({k: getattr(o, k) for k in PG._KEYS + PG._POLYS})
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
    n = 200_000
    idx = np.arange(0, h * w, (h * w) // n)[:n]
    t0 = time.perf_counter()
    G.image_rays(o, 1, 1, -25.0, 95.0, "utm", (0, 0, 0), 1.0, 0.0, 0.0)
    G.get_rays((idx % w).astype(np.float64), (idx // w).astype(np.float64), o, -25.0, 95.0, cs="utm")
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"kind": "port", "cores": 1, "Mrays/s": n / dt / 1e6,
                           "sample": f"{n} strided pixels of the image, numpy float64 (vectorised, as rpcm / the reference)"}
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
