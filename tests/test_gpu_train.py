"""Training-step driver on the GPU: fused step == autograd step; bf16 PSNR-drift criterion."""
import math

import pytest
import torch

from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from brdf_nerf_b200.rendering import Draws, render_rays
from brdf_nerf_b200.synth import make_rays
from brdf_nerf_b200.train import Trainer
from oracle import losses_torch as LT
from oracle import render_torch as RT

pytestmark = pytest.mark.gpu


def test_trainer_step_matches_autograd_path(cuda):
    """Trainer (no tape, grads straight into the flat bucket) == render_rays + loss.backward()."""
    args = named_config("lambertian_ds")
    n = 128
    batch = make_rays(n, depth_supervision=True).to(cuda)
    od = RT.Draws.make(n, 64, 64, 128, seed=3, with_gt=True)
    draws = Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt)
    torch.manual_seed(0)
    m1 = load_model(args).to(cuda)
    torch.manual_seed(0)
    m2 = load_model(args).to(cuda)
    res, _ = render_rays({"coarse": m1}, args, batch.rays, None, mode="train", valid_depth=batch.valid_depth,
                         target_depths=batch.target_depths, target_std=batch.target_std, _draws=draws)
    loss1 = LT.train_loss(res, batch, args)
    m1.flat_grads.zero_()
    loss1.backward()
    tr = Trainer(m2, args, lr=0.0)
    loss2 = tr.step(batch, draws=draws)
    assert abs(loss1.item() - loss2.item()) < 1e-6
    a, b = m1.flat_grads, m2.flat_grads
    assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item() + 1e-9


def _psnr(model, args, batch, draws):
    with torch.no_grad():
        res, _ = render_rays({"coarse": model}, args, batch.rays, None, _draws=draws)
    mse = ((res["rgb_coarse"] - batch.rgbs) ** 2).mean().item()
    return -10.0 * math.log10(mse)


def test_bf16_psnr_drift(cuda):
    """North-star criterion for the bf16 path: <= 0.1 dB PSNR drift vs fp32 after a fixed number of
    synthetic training steps (same rays, same draws, same init).

    A single PSNR reading of a 512-ray model carries +-0.1 dB of run-to-run noise by itself (fp32
    atomics / TMA reduce-adds arrive in a different order on every run and 200 Adam steps amplify the
    last bit), which the sign flips of the printed curve show.  The criterion is therefore asserted
    on the MEAN drift over the eleven checkpoints of the second half of the run (steps 100, 110, .. 200),
    each PSNR evaluated on two independent sets of evaluation draws; every single checkpoint must
    additionally stay within 0.5 dB."""
    args = named_config("lambertian_ds")
    n, checkpoints = 512, (50,) + tuple(range(100, 201, 10))
    batch = make_rays(n, depth_supervision=True).to(cuda)
    ev_draws = []
    for seed in (9999, 7777):
        ev = RT.Draws.make(n, 64, 64, 128, seed=seed)
        ev_draws.append(Draws(u_strat=ev.u_strat, u_pred=ev.u_pred))
    curve = {}
    for precision in ("fp32", "bf16"):
        torch.manual_seed(0)
        model = load_model(args, precision=precision).to(cuda)
        tr = Trainer(model, args)
        curve[precision] = []
        for i in range(max(checkpoints)):
            od = RT.Draws.make(n, 64, 64, 128, seed=100 + i, with_gt=True)
            tr.step(batch, draws=Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt))
            if i + 1 in checkpoints:
                curve[precision].append(sum(_psnr(model, args, batch, d) for d in ev_draws) / len(ev_draws))
    drift = [b - a for a, b in zip(curve["fp32"], curve["bf16"])]
    for k, step in enumerate(checkpoints):
        print(f"PSNR after {step:4d} steps: fp32 {curve['fp32'][k]:.3f} dB   bf16 {curve['bf16'][k]:.3f} dB   "
              f"drift {drift[k]:+.3f} dB")
    tail = [d for d, step in zip(drift, checkpoints) if step >= 100]
    mean_drift = sum(tail) / len(tail)
    print(f"mean drift over steps 100..200: {mean_drift:+.3f} dB   worst single checkpoint {max(abs(d) for d in tail):.3f} dB")
    assert abs(mean_drift) <= 0.1, curve
    assert max(abs(d) for d in tail) <= 0.5, curve


def test_graph_step_equals_eager(cuda):
    args = named_config("lambertian_ds")
    n = 256
    batch = make_rays(n, depth_supervision=True).to(cuda)
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(cuda)
    tr = Trainer(m, args, use_graph=True)
    l0 = tr.step(batch).item()
    for _ in range(5):
        l = tr.step(batch).item()
    assert math.isfinite(l) and l < l0 + 0.05


@pytest.mark.parametrize("n,use_all", [(64, False), (1000, False), (333, True)])
def test_fused_loss_kernel_vs_oracle_losses(cuda, n, use_all):
    """bn_loss_color_depth (SNerfLoss + DepthLoss subset rule, fused with its gradients) == oracle losses + autograd."""
    from brdf_nerf_b200.train import loss_and_grads
    g = torch.Generator().manual_seed(n)
    s = 128
    args = named_config("lambertian_ds", usealldepth=use_all)
    batch = make_rays(n, depth_supervision=True)
    rgb = torch.rand(n, 3, generator=g, requires_grad=True)
    z = torch.sort(torch.rand(n, s, generator=g) * 0.6, -1)[0]
    w = torch.softmax(torch.randn(n, s, generator=g), -1)
    depth = ((w * z).sum(-1) + 0.05 * torch.randn(n, generator=g)).requires_grad_(True)
    res = {"rgb_coarse": rgb, "depth_coarse": depth, "weights_coarse": w, "z_vals_coarse": z}
    ref = LT.train_loss(res, batch, args)
    ref.backward()
    outs = dict(rgb=rgb.detach().to(cuda), depth=depth.detach().to(cuda), weights=w.to(cuda), z=z.to(cuda))
    loss, g_rgb, g_depth = loss_and_grads(args, outs, None, batch.to(cuda), True)
    assert abs(loss.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
    assert torch.allclose(g_rgb.cpu(), rgb.grad, atol=1e-7) and torch.allclose(g_depth.cpu(), depth.grad, atol=1e-7)
    # colour term alone
    loss_c, g_c, g_d = loss_and_grads(args, outs, None, batch.to(cuda), False)
    assert g_d is None and abs(loss_c.item() - LT.color_loss(res, batch.rgbs).item()) < 1e-6


def test_train_loop_pool_and_schedule(cuda):
    """Device-resident ray pool + step-fraction schedule + fused step: the depth loss drops at ds_drop, the learning
    rate follows StepLR per epoch, every ray of the pool is visited once per epoch."""
    from brdf_nerf_b200.schedule import DeviceRayPool
    from brdf_nerf_b200.train import TrainLoop
    args = named_config("lambertian_ds", batch_size=256, max_train_steps=12, ds_drop=0.5)
    pool = make_rays(1024, depth_supervision=True).to(cuda)
    feed = DeviceRayPool(pool, 256, seed=1)
    seen = torch.cat([feed.next_batch().rays for _ in range(4)])
    assert torch.equal(torch.sort(seen[:, 0])[0], torch.sort(pool.rays[:, 0])[0])        # one epoch = a permutation
    torch.manual_seed(0)
    model = load_model(args, precision="bf16").to(cuda)
    loop = TrainLoop(model, args, pool, use_graph=True)
    losses = [loop.step().item() for _ in range(12)]
    assert all(math.isfinite(x) for x in losses)
    assert loop.trainer.use_depth_loss is False and abs(loop.trainer.lr - 5e-4 * 0.9 ** 3) < 1e-12
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("n,s,lam", [(64, 128, (0.1, 0.0, 0.0)), (333, 64, (0.1, 0.05, 0.5)), (100, 128, (0.0, 0.0, 0.5))])
def test_fused_regularizer_kernel_vs_oracle_losses(cuda, n, s, lam):
    """bn_loss_regularizers (NormalRegLoss metrics.py:179-216 for analytic / learned normals + HardSurfaceLoss
    metrics.py:263-290, fused with their gradients) == the oracle restatements + autograd."""
    from brdf_nerf_b200 import ops
    g = torch.Generator().manual_seed(n + s)
    lam_an, lam_lr, lam_hs = lam
    pitch = 13
    rays = make_rays(n).rays
    z = torch.sort(torch.rand(n, s, generator=g) * 0.6, -1)[0]
    w = torch.softmax(torch.randn(n, s, generator=g), -1).requires_grad_(True)
    depth = ((w.detach() * z).sum(-1) + 0.05 * torch.randn(n, generator=g)).requires_grad_(True)
    packed = torch.randn(n, s, pitch, generator=g).requires_grad_(True)
    res = {"weights_coarse": w, "z_vals_coarse": z, "depth_coarse": depth, "rays_d_coarse": (-rays[:, 3:6]).reshape(n, 1, 3),
           "normal_an_coarse": packed[..., 4:7], "normal_lr_coarse": packed[..., 7:10]}
    ref = torch.zeros(())
    percs = [0.0, 0.0]
    if lam_an:
        l, percs[0] = LT.normal_reg_loss(res, lam_an, "normal_an"); ref = ref + l
    if lam_lr:
        l, percs[1] = LT.normal_reg_loss(res, lam_lr, "normal_lr"); ref = ref + l
    if lam_hs:
        ref = ref + LT.hard_surface_loss(res, lam_hs)
    ref.backward()
    loss = torch.full((1,), 0.25, device=cuda)            # accumulated on top of the colour / depth loss
    g_depth0 = torch.full((n,), 0.5, device=cuda)
    gw, gp, gd, bad = ops.loss_regularizers(loss, w.detach().to(cuda), z.to(cuda), depth.detach().to(cuda),
                                            packed.detach().to(cuda), rays.to(cuda), 4, lam_an, 7, lam_lr, lam_hs,
                                            g_depth=g_depth0.clone(), want_bad_count=True)
    assert abs(loss.item() - 0.25 - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert torch.allclose(gw.cpu(), w.grad, atol=1e-6, rtol=1e-5)
    if lam_an or lam_lr:
        assert torch.allclose(gp.cpu(), packed.grad, atol=1e-6, rtol=1e-5)
    else:
        assert gp is None
    if lam_hs:
        assert torch.allclose(gd.cpu() - 0.5, depth.grad, atol=1e-6, rtol=1e-5)
    tot = n * s
    if lam_an:
        assert abs(100.0 * bad[0].item() / tot - percs[0]) < 1e-3
    if lam_lr:
        assert abs(100.0 * bad[1].item() / tot - percs[1]) < 1e-3


def test_trainer_with_regularizers_matches_autograd(cuda):
    """Trainer step with NormalRegLoss + HardSurfaceLoss switched on == render_rays autograd + oracle losses (fp32)."""
    from brdf_nerf_b200.train import Trainer
    args = named_config("rpv111", nr_reg_an_lambda=0.1, hs_lambda=0.5)
    n = 64
    batch = make_rays(n).to(cuda)
    S1, Gs = args.n_samples, args.guided_samples
    od = RT.Draws.make(n, S1, Gs, S1 + Gs, seed=3)
    kw = dict(apply_brdf=True, cos_irra_on=True)
    torch.manual_seed(0)
    m1 = load_model(args, precision="fp32").to(cuda)
    res, _ = render_rays({"coarse": m1}, args, batch.rays, None, mode="train",
                         _draws=Draws(u_strat=od.u_strat, u_pred=od.u_pred), **kw)
    loss = LT.color_loss(res, batch.rgbs, args.lambda_rgb) + LT.normal_reg_loss(res, 0.1)[0] + LT.hard_surface_loss(res, 0.5)
    m1.flat_grads.zero_()
    loss.backward()
    g_ref = m1.flat_grads.clone()
    torch.manual_seed(0)
    m2 = load_model(args, precision="fp32").to(cuda)
    tr = Trainer(m2, args, use_graph=False)
    tr.use_hard_surface = True
    p0 = m2.flat_params.clone()
    loss2 = tr.step(batch, draws=Draws(u_strat=od.u_strat, u_pred=od.u_pred), **kw)
    assert abs(loss2.item() - loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))
    g = m2.flat_grads
    assert (g - g_ref).abs().max().item() <= 1e-4 * g_ref.abs().max().item() + 1e-8
    assert not torch.equal(p0, m2.flat_params)


def test_graph_adam_matches_eager_adam(cuda):
    """bn_adam_step_graph (lr / step counter in device memory, bias corrections derived on the device) == bn_adam_step,
    including after a learning-rate change and when it is replayed from a captured CUDA graph."""
    from brdf_nerf_b200 import ops
    g = torch.Generator().manual_seed(1)
    n = 10007
    p0 = torch.randn(n, generator=g).to(cuda)
    grads = [torch.randn(n, generator=g).to(cuda) * 0.1 for _ in range(6)]
    pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pb, mb, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    state = torch.zeros(4, device=cuda)
    gbuf = torch.zeros_like(p0)
    graph = None
    for t, gr in enumerate(grads, start=1):
        lr = 5e-4 if t < 4 else 4.5e-4
        ops.adam_step(pa, gr, ma, va, lr, t, grad_scale=0.5)
        state[0:1].fill_(lr)
        gbuf.copy_(gr)
        if t < 3:
            ops.adam_step_graph(pb, gbuf, mb, vb, state, grad_scale=0.5)
        else:
            if graph is None:
                torch.cuda.synchronize()
                keep = (pb.clone(), mb.clone(), vb.clone(), state.clone())
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    ops.adam_step_graph(pb, gbuf, mb, vb, state, grad_scale=0.5)
                for dst, src in zip((pb, mb, vb, state), keep):      # capture does not execute; restore in case it did
                    dst.copy_(src)
            graph.replay()
        assert state[1].item() == float(t)
        assert torch.allclose(pa, pb, atol=1e-7, rtol=1e-6), (t, (pa - pb).abs().max().item())
    assert torch.allclose(ma, mb) and torch.allclose(va, vb)


def test_whole_step_graph_advances_optimizer(cuda):
    """The captured step contains the parameter update: step counter, learning-rate changes and eager steps in between
    stay consistent with the device-side optimizer state."""
    args = named_config("lambertian_ds")
    batch = make_rays(256, depth_supervision=True).to(cuda)
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(cuda)
    tr = Trainer(m, args, use_graph=True)
    p0 = m.flat_params.clone()
    losses = [tr.step(batch).item() for _ in range(4)]
    assert tr._graph_updates and tr.step_count == 4 and tr._opt_state[1].item() == 4.0
    assert not torch.equal(p0, m.flat_params)
    od = RT.Draws.make(256, args.n_samples, args.guided_samples, args.n_samples + args.guided_samples, seed=2, with_gt=True)
    tr.step(batch, draws=Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt))     # eager step (injected draws)
    assert tr.step_count == 5
    tr.lr = 4e-4
    losses.append(tr.step(batch).item())
    assert tr.step_count == 6 and tr._opt_state[1].item() == 6.0 and abs(tr._opt_state[0].item() - 4e-4) < 1e-10
    assert all(math.isfinite(x) for x in losses) and losses[-1] < losses[0]


@pytest.mark.parametrize("use_graph", [True, False])
def test_prefetch_feed_delivers_batches_and_losses(cuda, use_graph):
    """Host feed on the copy stream (Trainer.prefetch / read_loss_async): every step consumes exactly the host batch that was
    prefetched for it, and the loss that arrives in pinned memory is that step's loss."""
    args = named_config("lambertian_ds")
    n = 256
    hosts = [make_rays(n, seed=100 + k, depth_supervision=True).packed(pin=True) for k in range(3)]
    torch.manual_seed(0)
    m = load_model(args, precision="bf16").to(cuda)
    tr = Trainer(m, args, use_graph=use_graph)
    pin = torch.full((8,), float("nan")).pin_memory()
    direct, evs, seen = [], [], []
    staged = tr.prefetch(hosts[0])
    for i in range(8):
        if use_graph:                                      # the graph reads its static buffers, refilled inside step()
            loss = tr.step(staged)
            snap = (tr.static_batch().rays.clone(), tr.static_batch().rgbs.clone())
        else:                                              # the eager step reads the staging batch in place
            torch.cuda.current_stream().wait_event(staged._ready)
            snap = (staged.rays.clone(), staged.rgbs.clone())
            loss = tr.step(staged)
        seen.append(snap)
        direct.append(loss.clone())
        staged = tr.prefetch(hosts[(i + 1) % 3])
        evs.append(tr.read_loss_async(loss, pin[i:i + 1]))
    for i, ev in enumerate(evs):
        ev.synchronize()
        assert torch.equal(seen[i][0].cpu(), hosts[i % 3].rays) and torch.equal(seen[i][1].cpu(), hosts[i % 3].rgbs)
        assert pin[i].item() == direct[i].item() and math.isfinite(pin[i].item())
    assert tr.step_count == 8


def test_checkpoint_resume_continues_training(cuda):
    """Trainer.state_dict() / load_state_dict(): 3 steps + checkpoint + 2 steps == restore into a fresh model + the same 2
    steps (injected draws; fp32-atomic ordering in the weight gradients allows 1e-6), in eager and in graph mode the device-
    side optimizer state follows the restored step counter."""
    args = named_config("lambertian_ds")
    n = 128
    batch = make_rays(n, depth_supervision=True).to(cuda)
    S1, G = args.n_samples, args.guided_samples

    def draws(i):
        od = RT.Draws.make(n, S1, G, S1 + G, seed=50 + i, with_gt=True)
        return Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt)

    torch.manual_seed(0)
    m1 = load_model(args, precision="fp32").to(cuda)
    t1 = Trainer(m1, args)
    for i in range(3):
        t1.step(batch, draws=draws(i))
    ckpt = t1.state_dict()
    for i in range(3, 5):
        t1.step(batch, draws=draws(i))
    torch.manual_seed(9)                                               # different init: everything must come from the checkpoint
    m2 = load_model(args, precision="fp32").to(cuda)
    t2 = Trainer(m2, args)
    t2.load_state_dict(ckpt)
    assert t2.step_count == 3
    for i in range(3, 5):
        t2.step(batch, draws=draws(i))
    assert t2.step_count == 5
    assert (m1.flat_params - m2.flat_params).abs().max().item() <= 1e-6
    assert (t1.m - t2.m).abs().max().item() <= 1e-6 and (t1.v - t2.v).abs().max().item() <= 1e-7
    # graph mode: the captured Adam reads lr / step from device memory, refreshed after a restore
    m3 = load_model(args, precision="bf16").to(cuda)
    t3 = Trainer(m3, args, use_graph=True)
    t3.step(batch)
    t3.load_state_dict(ckpt)
    t3.step(batch)
    assert t3.step_count == 4 and t3._opt_state[1].item() == 4.0 and abs(t3._opt_state[0].item() - t3.lr) < 1e-10


def test_train_loop_checkpoint_restores_schedule_and_feed(cuda):
    from brdf_nerf_b200.train import TrainLoop
    args = named_config("lambertian_ds", batch_size=64, max_train_steps=100)
    pool = make_rays(640, depth_supervision=True).to(cuda)
    torch.manual_seed(0)
    loop = TrainLoop(load_model(args, precision="bf16").to(cuda), args, pool, use_graph=False, seed=4)
    for _ in range(12):                                                # crosses an epoch boundary (10 batches per epoch)
        loop.step()
    ckpt = loop.state_dict()
    want = [loop.feed.next_batch().rays.clone() for _ in range(10)]   # the batches the original loop would see next
    torch.manual_seed(1)
    loop2 = TrainLoop(load_model(args, precision="bf16").to(cuda), args, pool, use_graph=False, seed=99)
    loop2.load_state_dict(ckpt)
    assert loop2.schedule.train_steps == 12 and loop2.trainer.step_count == 12 and loop2.feed.epoch == 1 and loop2.feed._pos == 128
    got = [loop2.feed.next_batch().rays.clone() for _ in range(10)]
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    assert ckpt["epoch"] == 1 and abs(loop2.schedule.next().lr - args.lr * 0.9) < 1e-12


def test_lazy_packed_step_equals_materialised_rows(cuda):
    """The Trainer's Lambertian step never materialises the depth-ordered per-sample rows (lazy_packed: compositing gathers the
    MLP's rows through sort_idx, its backward scatters the gradient rows): same loss bit for bit, same gradients up to the
    order of the weight-gradient atomics, as the step with bn_permute_samples in both directions."""
    from brdf_nerf_b200 import rendering as R
    args = named_config("lambertian_ds")
    batch = make_rays(256, depth_supervision=True).to(cuda)
    out = {}
    for lazy in (True, False):
        torch.manual_seed(0)
        model = load_model(args, precision="bf16").to(cuda)
        model.sync_weights()
        od = RT.Draws.make(256, 64, 64, 128, seed=3, with_gt=True)
        from brdf_nerf_b200.rendering import Draws
        d = Draws(u_strat=od.u_strat, u_pred=od.u_pred, u_gt=od.u_gt).to(cuda)
        outs, st = R._forward(model, args, batch.rays, d, train=True, mode="train", valid_depth=batch.valid_depth,
                              target_depths=batch.target_depths, target_std=batch.target_std, apply_brdf=False,
                              bTestNormal=False, bTestSun_v=False, gsam_only=False, apply_theta=False, cos_irra_on=False,
                              lazy_packed=lazy)
        assert (outs["packed"] is None) == lazy
        g_rgb = (outs["rgb"] - batch.rgbs) * (2.0 / outs["rgb"].numel())
        grads = torch.zeros_like(model.flat_params)
        R._backward(model, st, g_rgb, None, None, None, grads)
        out[lazy] = (outs["rgb"].clone(), outs["depth"].clone(), grads)
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    ga, gb = out[True][2], out[False][2]
    assert (ga - gb).abs().max().item() <= 1e-5 * gb.abs().max().item() + 1e-9

