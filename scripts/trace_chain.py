"""Timeline of the fused trunk kernels (mlp_chain.cuh): clock64() stamps of the first 256-point block of CTA pair 0.
    python scripts/trace_chain.py [train]     density pass (default) or the TRAINING forward (h_l / c_l stored);
                                              prints, per layer and column half, cycles relative to the start of layer 0"""
import ctypes as C
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


def main():
    n, S = 4096, 64
    dev = torch.device("cuda:0")
    args = named_config("lambertian_ds")
    rays = make_rays(n).rays.to(dev)
    z = torch.sort(torch.rand(n, S, device=dev) * 0.6, -1)[0].contiguous()
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(dev)
    m.sync_weights()
    train = len(sys.argv) > 1 and sys.argv[1] == "train"
    if train:
        flags = m.mlp_flags(train=True)
        sig = torch.empty((n, S), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, flags, tag="ws_train")
        fn = lambda: ops.mlp_trunk_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, flags, n * S, 0, sig, ws)
    else:
        sig = torch.empty((n, S), dtype=torch.float32, device=dev)
        ws = m.workspace(n * S, L.MLP_SIGMA_ONLY, tag="ws_sigma")
        fn = lambda: ops.mlp_forward(m, rays[:, 0:3], 11, rays[:, 3:6], 11, z, L.MLP_SIGMA_ONLY, sig, 1, ws)
    print("kernel:", "chain::train_chain_kernel (training forward)" if train else "chain::sigma_chain_kernel (density pass)")
    for _ in range(3):
        fn()
    buf = torch.zeros(16 * 2 * 16, dtype=torch.int64, device=dev)
    L.check(L.load().bn_debug_chain_trace(m.handle(), C.c_void_p(buf.data_ptr())))
    fn()
    torch.cuda.synchronize()
    L.check(L.load().bn_debug_chain_trace(m.handle(), None))
    t = buf.cpu().view(16, 2, 16)
    t0 = int(t[0, 0, 0])
    names = ["tmem_free", "kb0", "kb1", "kb2", "kb3", "kb4", "kb5", "kb6", "kb7", "kb8", "issued", "epi_wait", "epi_tfull", "epi_done"]
    durations = {14: "wait_act", 15: "wait_w"}        # cycle counts, not stamps
    print("cycles relative to the first stamp; MMA issuer: tmem_free, kb*, issued; epilogue warp 4: epi_*")
    prev_tfull1 = None
    for l in range(8):
        for h in range(2):
            row = {nm: int(t[l, h, i]) - t0 for i, nm in enumerate(names) if int(t[l, h, i]) != 0}
            dur = "  ".join(f"{nm}={int(t[l, h, i])}" for i, nm in durations.items())
            print(f"layer {l} half {h}: " + "  ".join(f"{k}={v}" for k, v in row.items()) + "  | " + dur)
        tf1 = int(t[l, 1, 12]) - t0
        if prev_tfull1 is not None:
            print(f"   -> layer period (tfull[1] to tfull[1]) = {tf1 - prev_tfull1} cycles (MMA floor 8192 nominal, ~12500 at the measured 195 cycles per pair-MMA)")
        prev_tfull1 = tf1


if __name__ == "__main__":
    main()
