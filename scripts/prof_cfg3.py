"""cfg3 (RPV111 + analytic normals, second-order backward): a few eager steps for an ncu launch list."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.synth import make_rays
from brdf_nerf_b200.train import Trainer
dev = torch.device("cuda:0")
args = named_config("rpv111")
torch.manual_seed(0)
model = load_model(args, precision="bf16").to(dev)
tr = Trainer(model, args, use_graph=False)
batch = make_rays(1024).to(dev)
for _ in range(4):
    tr.step(batch, apply_brdf=True, cos_irra_on=True)
torch.cuda.synchronize()
print("done")
