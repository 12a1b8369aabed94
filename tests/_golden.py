"""Helpers shared by the golden-fixture tests."""
import glob
import hashlib
import os

import numpy as np
import torch

from brdf_nerf_b200.config import named_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror oracle/make_golden.py CASES
CASES = {
    "lambertian_test": ("lambertian", {}, dict(mode="test"), False),
    "lambertian_ds_train": ("lambertian_ds", {}, dict(mode="train"), True),
    "lambertian_ds_std0_train": ("lambertian_ds", {}, dict(mode="train"), True),
    "lambertian_gsam_only": ("lambertian", {}, dict(mode="test", gsam_only=True), False),
    "rpv111_brdf": ("rpv111", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "rpv111_multi_brdf": ("rpv111_multi", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "hapke_bct_brdf": ("hapke_bct", {}, dict(mode="test", apply_brdf=True, apply_theta=True, cos_irra_on=True), False),
    "hapke_b_brdf": ("hapke_b", {}, dict(mode="test", apply_brdf=True), False),
    "microfacet_brdf": ("microfacet", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "rpv111_learned_normal": ("rpv111", dict(normal="learned"), dict(mode="test", apply_brdf=True), False),
    "lambertian_viewdir_test": ("lambertian_viewdir", {}, dict(mode="test"), False),
    "rpv111_sunvis_test": ("rpv111", {}, dict(mode="test", apply_brdf=True, cos_irra_on=True, bTestSun_v=True), False),
    "hapke_shell1_brdf": ("hapke_b", dict(b=0, shell_hapke=1), dict(mode="test", apply_brdf=True), False),
    "hapke_shell2_brdf": ("hapke_b", dict(b=0, shell_hapke=2), dict(mode="test", apply_brdf=True), False),
    "hapke_shell3_brdf": ("hapke_b", dict(b=0, shell_hapke=3), dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "lambertian_nomapping_test": ("lambertian", dict(mapping=False), dict(mode="test"), False),
    "rpv111_nomapping_brdf": ("rpv111", dict(mapping=False), dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "rpv111_an_lr_normals": ("rpv111", dict(normal="analystic_learned"), dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
    "lambertian_beta_test": ("lambertian", dict(beta=True), dict(mode="test"), False),
    "rpv111_beta_brdf": ("rpv111", dict(beta=True), dict(mode="test", apply_brdf=True, cos_irra_on=True), False),
}


def names():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))}
    return [n for n in CASES if n in have]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    cfg, over, kw, ds = CASES[name]
    return g, named_config(cfg, **over), dict(kw), ds


def weights_digest(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def supervision(g):
    return dict(valid_depth=torch.from_numpy(g["valid_depth"]), target_depths=torch.from_numpy(g["target_depths"]),
                target_std=torch.from_numpy(g["target_std"]))


def bits_equal(a: np.ndarray, b: np.ndarray) -> int:
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())


def time_embedding(g):
    """(ts, nn.Embedding) of a beta case, rebuilt from the fixture (None, None otherwise)."""
    if "ts" not in g:
        return None, None
    w = torch.from_numpy(g["t_weight"])
    emb = torch.nn.Embedding(w.shape[0], w.shape[1])
    with torch.no_grad():
        emb.weight.copy_(w)
    return torch.from_numpy(g["ts"]), emb
