"""Drop-in surface under the usage patterns of the reference's own training / evaluation code: a torch optimizer stepping
the parameters between renders (Lightning + torch.optim.Adam, main.py:147-150), checkpoints loaded after a first render
(eval.py:26-54), several forwards alive before their backwards (NeRF_pl.forward loops over chunks, main.py:127-139; the
module is called once per chunk inside `inference`, spsbrdfnerf.py:117-127), frozen parameters (spsbrdfnerf.py:617-633)."""
import pytest
import torch

from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.rendering import Draws, render_rays
from brdf_nerf_b200.synth import make_rays
from brdf_nerf_b200.train import Trainer
from oracle import render_torch as RT

pytestmark = pytest.mark.gpu


def _draws(n, seed, with_gt=False):
    od = RT.Draws.make(n, 64, 64, 128, seed=seed, with_gt=with_gt)
    return Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt if with_gt else None)


def _render(model, args, rays, draws, **kw):
    res, _ = render_rays({"coarse": model}, args, rays, None, _draws=draws, **kw)
    return res


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_torch_optimizer_step_is_seen_by_the_next_render(cuda, precision):
    """render -> loss.backward() -> torch.optim.Adam.step() -> render: the second render must run on the UPDATED weights
    (the packed bf16 / transposed copies are refreshed), i.e. equal a fresh model that loads the updated state_dict."""
    args = named_config("lambertian")
    n = 128
    batch = make_rays(n).to(cuda)
    d = _draws(n, 5)
    torch.manual_seed(0)
    model = load_model(args, precision=precision).to(cuda)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    r0 = _render(model, args, batch.rays, d)
    ((r0["rgb_coarse"] - batch.rgbs) ** 2).mean().backward()
    opt.step()
    with torch.no_grad():
        r1 = _render(model, args, batch.rays, d)
    assert (r1["rgb_coarse"] - r0["rgb_coarse"].detach()).abs().max().item() > 1e-4, "the optimizer step was not seen"
    torch.manual_seed(1)
    fresh = load_model(args, precision=precision).to(cuda)
    with torch.no_grad():
        before = _render(fresh, args, batch.rays, d)                    # first render on other weights ...
        fresh.load_state_dict(model.state_dict())                       # ... then the checkpoint arrives (eval.py:26-54)
        r2 = _render(fresh, args, batch.rays, d)
    assert (before["rgb_coarse"] - r2["rgb_coarse"]).abs().max().item() > 1e-4, "load_state_dict was not seen"
    tol = 1e-6 if precision == "fp32" else 1e-6                         # same kernels, same packed weights: identical
    for k in ("rgb_coarse", "depth_coarse", "weights_coarse"):
        assert (r1[k] - r2[k]).abs().max().item() <= tol, k


def test_module_forward_after_inplace_update(cuda):
    """SpSBRDFNeRF.forward (PointsFunction): `p.add_()` under no_grad and `mark_weights_dirty()` after a raw `.data` write."""
    args = named_config("lambertian")
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(cuda)
    x = (torch.rand(500, 3, generator=torch.Generator().manual_seed(3)) * 1.6 - 0.8).to(cuda)
    with torch.no_grad():
        a = m(x)
        for p in m.parameters():
            p.add_(0.01 * torch.randn_like(p))
        b = m(x)
        assert (a - b).abs().max().item() > 1e-3
        m.fc_net[2].weight.data.mul_(0.5)           # invisible to every version counter
        m.mark_weights_dirty()
        c = m(x)
        assert (b - c).abs().max().item() > 1e-3


def test_two_forwards_before_their_backwards(cuda):
    """Chunked use of the autograd bridges: two renders (different rays) and two module calls, THEN the backwards.  Every
    call owns its activations, so the gradients equal those of the calls run one at a time."""
    args = named_config("lambertian_ds")
    n = 128
    b1, b2 = make_rays(n, seed=1, depth_supervision=True).to(cuda), make_rays(n, seed=2, depth_supervision=True).to(cuda)
    d1, d2 = _draws(n, 11), _draws(n, 12)
    torch.manual_seed(0)
    model = load_model(args).to(cuda)

    def grads_of(fn):
        model.flat_grads.zero_()
        fn()
        return model.flat_grads.clone()

    def one(batch, d):
        r = _render(model, args, batch.rays, d)
        ((r["rgb_coarse"] - batch.rgbs) ** 2).mean().backward()

    ga, gb = grads_of(lambda: one(b1, d1)), grads_of(lambda: one(b2, d2))

    def both():
        ra = _render(model, args, b1.rays, d1)
        rb = _render(model, args, b2.rays, d2)           # would have overwritten ra's activations in a shared buffer
        (((ra["rgb_coarse"] - b1.rgbs) ** 2).mean() + ((rb["rgb_coarse"] - b2.rgbs) ** 2).mean()).backward()

    gab = grads_of(both)
    ref = ga + gb
    assert (gab - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-9
    # module forward, chunk by chunk, one backward at the end (spsbrdfnerf.py:117-127)
    x = (torch.rand(600, 3, generator=torch.Generator().manual_seed(3)) * 1.6 - 0.8).to(cuda)
    w = torch.randn(600, 4, generator=torch.Generator().manual_seed(4)).to(cuda)
    g_whole = grads_of(lambda: (model(x) * w).sum().backward())
    g_chunks = grads_of(lambda: torch.cat([model(x[i:i + 200]) for i in range(0, 600, 200)]).mul(w).sum().backward())
    assert (g_chunks - g_whole).abs().max().item() <= 1e-4 * g_whole.abs().max().item() + 1e-9


def test_two_models_interleaved(cuda):
    """No state is shared between calls of different models (the result holder is per call)."""
    args = named_config("lambertian")
    n = 64
    batch = make_rays(n).to(cuda)
    d = _draws(n, 7)
    torch.manual_seed(0)
    ma = load_model(args).to(cuda)
    torch.manual_seed(1)
    mb = load_model(args).to(cuda)
    ra = _render(ma, args, batch.rays, d)
    rb = _render(mb, args, batch.rays, d)
    with torch.no_grad():
        ra0 = _render(ma, args, batch.rays, d)
    assert torch.equal(ra["rgb_coarse"].detach(), ra0["rgb_coarse"])
    assert (ra["rgb_coarse"] - rb["rgb_coarse"]).abs().max().item() > 1e-4
    ra["rgb_coarse"].sum().backward()
    rb["rgb_coarse"].sum().backward()
    assert ma.flat_grads.abs().sum().item() > 0 and mb.flat_grads.abs().sum().item() > 0


@pytest.mark.parametrize("graph", [False, True])
def test_frozen_parameters_are_not_updated(cuda, graph):
    """freeze('fc_net') then Trainer steps: only filter(requires_grad) is optimised (main.py:148-150)."""
    args = named_config("lambertian_ds")
    batch = make_rays(256, depth_supervision=True).to(cuda)
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(cuda)
    model.freeze("fc_net")
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tr = Trainer(model, args, use_graph=graph)
    for _ in range(3):
        tr.step(batch)
    torch.cuda.synchronize()
    after = model.state_dict()
    for k in before:
        changed = not torch.equal(before[k], after[k])
        assert changed == (not k.startswith("fc_net")), k


def test_nr_spv_lambda_is_refused(cuda):
    args = named_config("rpv111")
    args.nr_spv_lambda = 0.1
    torch.manual_seed(0)
    with pytest.raises(NotImplementedError, match="nr_spv_lambda"):
        Trainer(load_model(args).to(cuda), args)


def test_reference_rng_draw_order(cuda):
    """`_reference_rng=True` consumes torch's CUDA generator like the reference: rand (N,S1), randn (N,S1)
    [drawn although noise_std == 0], rand (N,G), rand (n_valid,G), randn (N,S).  Replaying that sequence by hand and
    injecting it must give bit-identical samples."""
    args = named_config("lambertian_ds")
    n = 96
    batch = make_rays(n, depth_supervision=True).to(cuda)
    torch.manual_seed(0)
    model = load_model(args).to(cuda)
    with torch.no_grad():
        torch.manual_seed(77)
        res, _ = render_rays({"coarse": model}, args, batch.rays, None, mode="train", valid_depth=batch.valid_depth,
                             target_depths=batch.target_depths, target_std=batch.target_std, _reference_rng=True)
        torch.manual_seed(77)
        u_strat = torch.rand((n, 64), device=cuda)
        torch.randn((n, 64), device=cuda)
        u_pred = torch.rand((n, 64), device=cuda)
        valid = batch.valid_depth.reshape(-1) > 0
        u_gt = torch.zeros((n, 64), device=cuda)
        u_gt[valid] = torch.rand((int(valid.sum()), 64), device=cuda)
        torch.randn((n, 128), device=cuda)
        after_manual = torch.rand(4, device=cuda)
        ref, _ = render_rays({"coarse": model}, args, batch.rays, None, mode="train", valid_depth=batch.valid_depth,
                             target_depths=batch.target_depths, target_std=batch.target_std,
                             _draws=Draws(u_strat=u_strat, u_pred=u_pred, u_gt=u_gt))
        torch.manual_seed(77)
        render_rays({"coarse": model}, args, batch.rays, None, mode="train", valid_depth=batch.valid_depth,
                    target_depths=batch.target_depths, target_std=batch.target_std, _reference_rng=True)
        after_call = torch.rand(4, device=cuda)
    assert torch.equal(res["z_vals_coarse"], ref["z_vals_coarse"])
    assert torch.equal(res["rgb_coarse"], ref["rgb_coarse"])
    assert torch.equal(after_manual, after_call), "the generator is left where the reference leaves it"
