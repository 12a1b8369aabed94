"""Tiny bf16 training run used under compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.synth import make_rays
from brdf_nerf_b200.train import Trainer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
args = named_config("lambertian_ds")
torch.manual_seed(0)
m = load_model(args, precision=prec).cuda()
tr = Trainer(m, args)
b = make_rays(n, depth_supervision=True).to("cuda")
for i in range(2):
    l = tr.step(b)
    torch.cuda.synchronize()
    print("step", i, float(l))
