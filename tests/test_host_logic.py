"""CPU: host-side logic — config POD, synthetic rays, module surface / flat parameter storage,
loss restatement used by the fused step, and the no-CPU-fallback rule."""
import os
import numpy as np
import pytest
import torch

from brdf_nerf_b200.config import PathConfig, make_args, named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.synth import make_rays, make_tile_rays


def test_args_defaults_and_pod():
    a = make_args()
    assert (a.n_samples, a.guided_samples, a.fc_feat, a.fc_layers, a.chunk) == (64, 64, 512, 8, 5120)
    assert a.sc_lambda == 0.0
    with pytest.raises(TypeError):
        make_args(not_an_option=1)
    c = PathConfig.from_args(named_config("rpv111"))
    assert c.brdf == "rpv" and c.normal_an and c.n_freq_xyz == 10 and c.skip_layer == 4
    assert PathConfig.from_args(named_config("microfacet")).brdf == "microfacet"
    assert PathConfig.from_args(named_config("hapke_bct")).hapke_theta


def test_synthetic_rays_are_deterministic_and_well_formed():
    a, b = make_rays(257, depth_supervision=True), make_rays(257, depth_supervision=True)
    assert torch.equal(a.rays, b.rays) and torch.equal(a.target_depths, b.target_depths)
    r = a.rays
    assert r.shape == (257, 11) and r.dtype == torch.float32
    assert torch.allclose(r[:, 3:6].norm(dim=-1), torch.ones(257), atol=1e-6)
    assert torch.allclose(r[:, 8:11].norm(dim=-1), torch.ones(257), atol=1e-6)
    assert (r[:, 5] < 0).all() and (r[:, 10] > 0).all() and (r[:, 6] == 0).all()
    d = a.target_depths[:, 0]
    assert ((d > 0) & (d < r[:, 7])).all()
    s0, s1 = a.shard(0, 2), a.shard(1, 2)
    assert torch.equal(torch.cat([s0.rays, s1.rays]), r[:256])
    assert make_tile_rays(8, 16).shape == (128, 11)


@pytest.mark.parametrize("cfg,n_params", [("lambertian", 2295812), ("rpv111", 2690567), ("rpv111_multi", 2692109),
                                          ("hapke_bct", 2690567), ("microfacet", 2427397)])
def test_module_surface_and_param_counts(cfg, n_params):
    args = named_config(cfg)
    torch.manual_seed(0)
    m = load_model(args)
    assert sum(p.numel() for p in m.parameters()) == n_params          # SURVEY §6 [probe]
    keys = list(m.state_dict())
    assert keys[:2] == ["fc_net.0.weight", "fc_net.0.bias"] and "sigma_from_xyz.0.weight" in keys
    assert m.fc_net[8].weight.shape == (512, 572)                      # skip layer input = [PE(60) | h(512)]
    for attr in ("number_of_outputs", "number_of_outputs_brdf", "normal", "sun_v", "indirect_light", "beta",
                 "roughness", "RPV", "MultiBRDF", "rgb_padding", "glossy_scale", "args"):
        assert hasattr(m, attr)
    # flat storage: parameters are views of one buffer, state_dict round-trips through it
    flat = m.flat_params
    assert flat.numel() >= n_params and flat.numel() % 8 == 0      # tensors start on 32-byte boundaries
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    off = m.offsets()
    assert torch.equal(flat[off["fc_net.2.weight"]:off["fc_net.2.weight"] + 512 * 512].view(512, 512), sd["fc_net.2.weight"])
    m.load_state_dict({k: v + 1 for k, v in sd.items()})
    assert torch.equal(m.flat_params[:10], sd["fc_net.0.weight"].reshape(-1)[:10] + 1)
    m.freeze("fc_net")
    assert not m.fc_net[0].weight.requires_grad and m.sigma_from_xyz[0].weight.requires_grad


def test_unsupported_model_options_fail_loudly():
    with pytest.raises(ValueError):
        load_model(make_args(model="sps-nerf"))
    with pytest.raises(NotImplementedError):
        load_model(named_config("lambertian", indirect_light=True))
    m = load_model(named_config("lambertian", beta=True))                  # beta head reads [features | t-embedding]
    assert tuple(m.beta_from_xyz[0].weight.shape) == (256, 512 + 4) and m.number_of_outputs == 5
    m = load_model(named_config("lambertian", input_viewdir=1))            # colour head reads [features | Mapping(d)]
    assert tuple(m.rgb_from_xyzdir[0].weight.shape) == (256, 512 + 24)
    m = load_model(named_config("lambertian", input_viewdir=1, mapping=False))
    assert tuple(m.rgb_from_xyzdir[0].weight.shape) == (256, 512 + 3)


def test_no_cpu_fallback_on_the_product_path():
    from brdf_nerf_b200 import _lib as L
    from brdf_nerf_b200.rendering import render_rays
    args = named_config("lambertian")
    torch.manual_seed(0)
    m = load_model(args)
    with pytest.raises(L.BnError):
        render_rays({"coarse": m}, args, make_rays(4).rays, None)
    with pytest.raises(L.BnError):
        m(torch.zeros(4, 3))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product package may import it."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brdf_nerf_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle"


def test_schedule_matches_reference_rules():
    """Step-fraction switches of NeRF_pl.training_step (reference main.py:59-68, 196-210, 246-248) and the
    per-epoch StepLR (train_utils.py:117-118, 153-155)."""
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.schedule import Schedule
    args = named_config("rpv111", max_train_steps=1000, brdf_on=0.25, gsam_only_on=0.8, cos_irra_on=0.5, ds_lambda=10.0,
                        ds_drop=0.3, noise_std=1.0, batch_size=100)
    sc = Schedule(args, dataset_len=1000)          # 10 steps per epoch
    seen = [sc.next() for _ in range(1000)]
    f = lambda k: seen[k - 1]                      # flags of the step whose train_steps == k
    assert not f(250).apply_brdf and f(251).apply_brdf
    assert not f(500).apply_theta and f(501).apply_theta
    assert not f(500).cos_irra_on and f(501).cos_irra_on
    assert not f(800).gsam_only and f(801).gsam_only
    assert f(299).use_depth_loss and not f(300).use_depth_loss
    assert abs(f(1).lr - 5e-4) < 1e-12 and abs(f(10).lr - 5e-4 * 0.9) < 1e-12 and abs(f(25).lr - 5e-4 * 0.9 ** 2) < 1e-12
    assert abs(f(3).noise_std - 0.81) < 1e-12
    # data-parallel: train_steps advances by the number of GPUs (main.py:196)
    sc4 = Schedule(args, dataset_len=1000, world_size=4)
    flags = [sc4.next() for _ in range(70)]
    assert not flags[61].apply_brdf and flags[62].apply_brdf        # 63 * 4 = 252 > 250


def test_checkpoint_hand_off_by_prefix():
    """load_ckpt / warm_start (eval.py:26-54, main.py:96-104): stage-2 model takes trunk + sigma + feats (+ rgb unless
    Hapke) from a Lightning-style stage-1 checkpoint; BRDF heads keep their own init; parameters stay flat-buffer views."""
    from brdf_nerf_b200.inference import load_ckpt, warm_start
    torch.manual_seed(1)
    stage1 = load_model(named_config("lambertian"))
    ckpt = {"state_dict": {f"nerf_coarse.{k}": v.clone() for k, v in stage1.state_dict().items()}}
    ckpt["state_dict"]["embedding_t.weight"] = torch.zeros(3, 4)
    for cfg, takes_rgb in (("rpv111", True), ("hapke_b", False)):
        torch.manual_seed(2)
        args = named_config(cfg)
        m = load_model(args)
        before = {k: v.clone() for k, v in m.state_dict().items()}
        loaded = warm_start(m, ckpt, args)
        after = m.state_dict()
        for k in after:
            src = stage1.state_dict().get(k)
            taken = k.startswith(("fc_net", "sigma_from_xyz", "feats_from_xyz")) or (takes_rgb and k.startswith("rgb_from_xyzdir"))
            if taken:
                assert torch.equal(after[k], src) and k in loaded, k
            else:
                assert torch.equal(after[k], before[k]) and k not in loaded, k
        flat = m.flat_params                              # parameters are still views of ONE buffer
        assert all(p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in m.parameters())
    m = load_model(named_config("lambertian"))
    assert set(load_ckpt(m, ckpt, model_name="nerf_coarse")) == set(stage1.state_dict())


def test_tile_shards_cover_the_tile_on_chunk_boundaries():
    from brdf_nerf_b200.inference import tile_shards
    for n, chunk, world in ((2048 * 2048, 5120, 8), (10, 4, 3), (5120, 5120, 8), (12345, 100, 5), (7, 8, 2)):
        sh = tile_shards(n, chunk, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == n
        for (a, b), (c, d) in zip(sh[:-1], sh[1:]):
            assert b == c and a <= b
        assert all(a % chunk == 0 for a, _ in sh if a < n)
        counts = [-(-(b - a) // chunk) for a, b in sh]
        assert max(counts) - min(counts) <= 1


def test_trainer_checkpoint_layout_is_a_lightning_adam_checkpoint():
    """Trainer.state_dict(): "state_dict" under the `nerf_coarse.` prefix + a torch.optim.Adam state dict over the parameters
    in registration order — torch's own Adam loads it, and a checkpoint built from torch's Adam resumes the Trainer (CPU:
    only the host-side bookkeeping runs, no kernel)."""
    import torch
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.train import Trainer
    args = named_config("rpv111")
    torch.manual_seed(0)
    model = load_model(args)
    tr = Trainer(model, args)
    sd0 = tr.state_dict()
    assert sd0["optimizer_states"][0]["state"] == {} and sd0["global_step"] == 0
    assert set(sd0["state_dict"]) == {f"nerf_coarse.{k}" for k in model.state_dict()}
    # pretend 7 steps were taken
    g = torch.Generator().manual_seed(1)
    tr.m.copy_(torch.randn(tr.m.shape, generator=g)); tr.v.copy_(torch.rand(tr.v.shape, generator=g)); tr.step_count = 7; tr.lr = 3e-4
    sd = tr.state_dict()
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1.0)
    opt.load_state_dict(sd["optimizer_states"][0])                     # torch accepts the layout
    assert opt.param_groups[0]["lr"] == 3e-4 and len(opt.state) == len(params)
    flat = model.flat_params
    for p in params:
        off = (p.data_ptr() - flat.data_ptr()) // 4
        assert torch.equal(opt.state[p]["exp_avg"].reshape(-1), tr.m[off:off + p.numel()])
        assert float(opt.state[p]["step"]) == 7.0
    # and back: a checkpoint made of torch's Adam state resumes a fresh Trainer
    torch.manual_seed(5)
    model2 = load_model(args)
    tr2 = Trainer(model2, args)
    ckpt = {"state_dict": sd["state_dict"], "optimizer_states": [opt.state_dict()], "global_step": 7}
    tr2.load_state_dict(ckpt)
    assert tr2.step_count == 7 and tr2.lr == 3e-4
    assert torch.equal(model2.flat_params, model.flat_params)
    flat2 = model2.flat_params
    for p, q in zip(params, model2.parameters()):                      # padding between tensors stays zero
        o1, o2 = (p.data_ptr() - flat.data_ptr()) // 4, (q.data_ptr() - flat2.data_ptr()) // 4
        assert o1 == o2 and torch.equal(tr2.m[o2:o2 + q.numel()], tr.m[o1:o1 + p.numel()])
        assert torch.equal(tr2.v[o2:o2 + q.numel()], tr.v[o1:o1 + p.numel()])
    import pytest
    bad = {"state_dict": sd["state_dict"], "optimizer_states": [{"state": {}, "param_groups": [{"lr": 1e-3, "params": [0, 1]}]}]}
    with pytest.raises(ValueError):
        tr2.load_state_dict(bad)


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """bench.py on a box without a GPU: the reference arm (the unmodified reference files on the host cores when they are
    available — live tree here, oracle/_ref on the GPU box — else the oracle port) prints ONE JSON line with the contract's
    keys, at the SAME 1024-ray step as the product arm; the product arm refuses to run (there is no CPU fallback)."""
    import json
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_harness as RH
    assert d["cpu_baseline"]["kind"] == ("reference" if RH.kind() != "absent" else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["rays_per_gpu"] == 1024 and "1024 rays" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("train rays/s")
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_algorithmic_flops_match_the_survey():
    """bench.py's per-ray work model == SURVEY §8d: 2.00 GFLOP per ray for the Lambertian training step."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from brdf_nerf_b200.config import named_config
    args = named_config("lambertian_ds")
    per_ray = bench.mlp_flops_per_ray(args, 64, 128)
    assert abs(per_ray / 2 - 1.001e9) < 2e6                              # 1.001 G MAC per ray (fwd 414.6 M + bwd 586.5 M)
    shared = bench.mlp_flops_per_ray(args, 64, 128, shared_trunk=True)
    assert abs((per_ray - shared) / 2 - 64 * 1_896_448) < 1                # one trunk evaluation of the 64 stratified points saved


def test_refused_options_cite_their_reason():
    """R19: `beta` is built (tests/test_gpu_beta.py); the two optional channels that are not are refused at
    construction with the reference lines that justify it (not silently ignored)."""
    import pytest
    import torch
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    for over, needle in ((dict(indirect_light=True), "out[..., 5:8]"), (dict(sun_v="learned"), "NameError")):
        args = named_config("lambertian", **over)
        torch.manual_seed(0)
        with pytest.raises(NotImplementedError) as e:
            load_model(args)
        assert needle in str(e.value), str(e.value)


def test_bench_and_package_have_no_undefined_names():
    """bench.py cannot run on the CPU box (no fallback), so a NameError in one of its legs would only surface on the GPU:
    a stdlib-only undefined-name check over bench.py and the package's Python modules."""
    import ast
    import builtins
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + \
        glob.glob(os.path.join(root, "brdf_nerf_b200", "**", "*.py"), recursive=True)
    for path in files:
        tree = ast.parse(open(path).read())
        defined = set(dir(builtins)) | {"__file__"}
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                defined.add(n.name)
            elif isinstance(n, ast.Import):
                defined.update(a.asname or a.name.split(".")[0] for a in n.names)
            elif isinstance(n, ast.ImportFrom):
                defined.update(a.asname or a.name for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                defined.add(n.id)
            elif isinstance(n, ast.arg):
                defined.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                defined.add(n.name)
        missing = sorted({n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)
                          and n.id not in defined})
        assert not missing, f"{os.path.relpath(path, root)}: undefined names {missing}"
