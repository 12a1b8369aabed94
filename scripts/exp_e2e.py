"""Where does the end-to-end overhead of a host-fed training step go?  Times 1024-ray Lambertian+depth steps (CUDA graph)
with (a) inputs resident, (b) H2D + loss D2H issued inline on the compute stream, (c) the copy-stream feed
(Trainer.prefetch / read_loss_async), (d) inline H2D only, (e) inline D2H only.  Prints ms/step of each.
    python scripts/exp_e2e.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
args = named_config("lambertian_ds")
torch.manual_seed(0)
model = load_model(args, precision="bf16").to(dev)
tr = Trainer(model, args, use_graph=True)
host = make_rays(1024, depth_supervision=True).packed(pin=True)
for _ in range(5):
    tr.step(host)
static = tr.static_batch()
pin = torch.empty(4).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(body):
    torch.cuda.synchronize()
    e0.record()
    body()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def resident():
    for _ in range(steps):
        tr.step(static)


def inline(h2d=True, d2h=True, lag=1):
    """lag = how many steps the host runs ahead of the loss it waits for"""
    def body():
        evs = [None] * 8
        pin8 = torch.empty(8).pin_memory()
        for i in range(steps):
            loss = tr.step(host if h2d else static)
            if d2h:
                pin8[i % 8:i % 8 + 1].copy_(loss.reshape(1), non_blocking=True)
            evs[i % 8] = torch.cuda.Event()
            evs[i % 8].record()
            if i >= lag:
                evs[(i - lag) % 8].synchronize()
        evs[(steps - 1) % 8].synchronize()
    return body


def prefetch():
    evs, prev = [None] * 4, None
    staged = tr.prefetch(host)
    for i in range(steps):
        loss = tr.step(staged)
        if i + 1 < steps:
            staged = tr.prefetch(host)
        evs[i % 4] = tr.read_loss_async(loss, pin[i % 4:i % 4 + 1])
        if prev is not None:
            evs[prev].synchronize()
        prev = i % 4
    evs[prev].synchronize()


for rep in range(2):
    print(f"rep {rep}: resident {timed(resident):.4f} ms | inline h2d+d2h {timed(inline()):.4f} | prefetch feed {timed(prefetch):.4f} | "
          f"inline h2d only {timed(inline(True, False)):.4f} | inline d2h only {timed(inline(False, True)):.4f} | "
          f"event-sync only {timed(inline(False, False)):.4f} | inline lag 2 {timed(inline(lag=2)):.4f} | inline lag 4 {timed(inline(lag=4)):.4f} | "
          f"resident again {timed(resident):.4f}", flush=True)
