"""K-C micro-benchmark on one B200: achieved HBM GB/s of the compositing kernels at inference-chunk size
(the 1024-ray training batch moves 7 MB and is latency bound; SURVEY 8d asks for N >= 8192).
    python scripts/bench_composite.py [N]      prints one line per kernel: us, algorithmic GB/s, frac of measured HBM peak"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import ops  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, iters=20):
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


def run(N=65536, S=128, report=print):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    peak = hbm_peak()
    out = {}
    for C in (4, 16):
        # three input sets (each >= 130 MB) rotated so that no launch finds its inputs in the 126 MB L2
        sets = []
        for _ in range(3):
            z = torch.sort(torch.rand(N, S, generator=g), dim=-1)[0].to(dev)
            packed = torch.rand(N, S, C, generator=g).to(dev)
            sets.append((z, packed))
        k = [0]

        def fwd():
            k[0] = (k[0] + 1) % 3
            z, packed = sets[k[0]]
            return ops.composite_forward(z, packed, None, 0.0)

        saved = [ops.composite_forward(z, packed, None, 0.0)[:3] for z, packed in sets]     # alpha, T, w per input set
        g_acc = torch.rand(N, C, device=dev); g_depth = torch.rand(N, device=dev); g_wsum = torch.rand(N, device=dev)

        def bwd():
            k[0] = (k[0] + 1) % 3
            z, packed = sets[k[0]]
            alpha, trans, w = saved[k[0]]
            return ops.composite_backward(z, packed, None, 0.0, None, alpha, trans, w, g_acc, None, g_depth, g_wsum, None, None)

        # algorithmic bytes per sample (fp32): fwd reads z + C channels, writes alpha, T, w;
        # bwd re-reads z, C channels, alpha, T, w and writes C channel gradients
        b_fwd = N * S * 4 * (1 + C + 3) + N * 4 * (C + 2)
        b_bwd = N * S * 4 * (1 + C + 3 + C) + N * 4 * (C + 2)
        for name, fn, nbytes in ((f"composite_fwd C={C}", fwd, b_fwd), (f"composite_bwd C={C}", bwd, b_bwd)):
            us = timeit(fn)
            gbs = nbytes / us / 1e3
            out[name] = dict(us=us, gbs=gbs, frac=gbs / peak, bytes=nbytes)
            report(f"{name:22s} N={N} S={S}: {us:8.1f} us  {gbs:7.1f} GB/s  {100 * gbs / peak:5.1f}% of measured HBM peak ({peak:.0f} GB/s)")
        del sets, saved
    return out


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 65536)
