"""A few eager training steps of a named configuration for an ncu launch list.
    python scripts/prof_cfg.py <config> <n_rays> [apply_theta]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg.csv python scripts/prof_cfg.py hapke_bct 8192 1"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402

cfg, n = sys.argv[1], int(sys.argv[2])
theta = len(sys.argv) > 3 and sys.argv[3] == "1"
dev = torch.device("cuda:0")
args = named_config(cfg)
torch.manual_seed(0)
model = load_model(args, precision="bf16").to(dev)
tr = Trainer(model, args, use_graph=False)
batch = make_rays(n).to(dev)
for _ in range(3):
    tr.step(batch, apply_brdf=True, cos_irra_on=True, apply_theta=theta)
torch.cuda.synchronize()
print("done")
