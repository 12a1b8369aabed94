"""K-C parity: compositing forward/backward vs the torch oracle (tolerance 1e-5 abs, fp32)."""
import pytest
import torch

from brdf_nerf_b200 import ops
from oracle import render_torch as RT

pytestmark = pytest.mark.gpu


def _inputs(n, s, c, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(n, s, generator=g) * 0.6, -1)[0]
    packed = torch.rand(n, s, c, generator=g)
    packed[..., 3] = torch.randn(n, s, generator=g) * 4 + 1      # density (relu'd inside)
    noise = torch.randn(n, s, generator=g)
    return z, packed, noise


@pytest.mark.parametrize("n,s", [(1, 64), (130, 64), (37, 128), (5, 100), (3, 1)])
def test_composite_sigma(cuda, n, s):
    z, packed, noise = _inputs(n, s, 4, 3)
    sig = packed[..., 3].contiguous()
    a, T, w, d = RT.volume_weights(z, sig, noise, 0.3)
    ga, gT, gw, gd, gstd = ops.composite_sigma(z.to(cuda), sig.to(cuda), noise.to(cuda), 0.3, want_all=True, want_std=True)
    for x, y in ((a, ga), (T, gT), (w, gw), (d, gd)):
        assert torch.allclose(x, y.cpu(), atol=2e-6, rtol=1e-5)
    std = torch.sqrt((((z - d[:, None]) ** 2) * w).sum(-1))
    assert torch.allclose(std, gstd.cpu(), atol=1e-5)


@pytest.mark.parametrize("n,s,c,noise_std,with_irr", [(64, 128, 4, 0.0, False), (33, 128, 16, 0.2, False),
                                                      (17, 64, 7, 0.0, True), (9, 96, 13, 0.0, False), (4, 128, 19, 0.0, False)])
def test_composite_forward_backward(cuda, n, s, c, noise_std, with_irr):
    z, packed, noise = _inputs(n, s, c, 11)
    g = torch.Generator().manual_seed(5)
    irr = torch.rand(n, s, generator=g) if with_irr else None
    packed_r = packed.clone().requires_grad_(True)
    a, T, w, d = RT.volume_weights(z, packed_r[..., 3], noise, noise_std)
    acc = (w.unsqueeze(-1) * packed_r).sum(1)
    wsum = w.sum(-1)
    outs = ops.composite_forward(z.to(cuda), packed.to(cuda), noise.to(cuda) if noise_std else None, noise_std,
                                 irr.to(cuda) if with_irr else None)
    ga, gT, gw, gd, gws, gacc, gacc_irr = outs
    mask = torch.ones(c, dtype=torch.bool); mask[3] = False
    assert torch.allclose(a, ga.cpu(), atol=2e-6) and torch.allclose(T, gT.cpu(), atol=2e-6)
    assert torch.allclose(w, gw.cpu(), atol=2e-6) and torch.allclose(d, gd.cpu(), atol=2e-6)
    assert torch.allclose(wsum, gws.cpu(), atol=2e-6)
    assert torch.allclose(acc[:, mask], gacc.cpu()[:, mask], atol=5e-6)
    loss_terms = []
    G_acc, G_d, G_ws, G_w = torch.randn(n, c, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, s, generator=g)
    G_acc[:, 3] = 0
    G_direct = torch.randn(n, s, c, generator=g)
    loss = (acc * G_acc).sum() + (d * G_d).sum() + (wsum * G_ws).sum() + (w * G_w).sum() + (packed_r * G_direct).sum()
    G_irr = None
    if with_irr:
        acc_irr = torch.cat([(w.unsqueeze(-1) * irr.unsqueeze(-1) * packed_r[..., :3]).sum(1), (w * irr).sum(-1, keepdim=True)], -1)
        assert torch.allclose(acc_irr, gacc_irr.cpu(), atol=5e-6)
        G_irr = torch.randn(n, 4, generator=g)
        loss = loss + (acc_irr * G_irr).sum()
    loss.backward()
    gp = ops.composite_backward(z.to(cuda), packed.to(cuda), noise.to(cuda) if noise_std else None, noise_std,
                                irr.to(cuda) if with_irr else None, ga, gT, gw, G_acc.to(cuda),
                                G_irr.to(cuda) if with_irr else None, G_d.to(cuda), G_ws.to(cuda), G_w.to(cuda),
                                G_direct.to(cuda))
    ref = packed_r.grad
    err = (gp.cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * max(1.0, scale), f"composite backward max err {err} (scale {scale})"


@pytest.mark.parametrize("n,s1,g,c", [(64, 64, 64, 4), (33, 64, 64, 7), (17, 48, 80, 4), (5, 64, 64, 16)])
def test_composite_through_sort_index(cuda, n, s1, g, c):
    """sort_idx folded into the compositing kernels (the Trainer's lazy_packed path): reading the MLP's generation-order rows
    through the index == compositing the rows bn_permute_samples has put into depth order, bit for bit, forward and backward
    (whose gradient rows come out in generation order)."""
    s = s1 + g
    gen = torch.Generator().manual_seed(31 + n)
    z, packed, noise = _inputs(n, s, c, 5)
    rows = packed.reshape(n * s, c).to(cuda)                     # any content: row order = [N][S1] block, then [N][G] block
    idx = torch.stack([torch.randperm(s, generator=gen) for _ in range(n)]).to(cuda)
    z, noise = z.to(cuda), noise.to(cuda)
    sorted_rows = ops.permute_samples(rows, idx, n, s1, g, c, scatter=False)
    ref = ops.composite_forward(z, sorted_rows, noise, 0.1)
    got = ops.composite_forward(z, rows, noise, 0.1, sort_idx=idx, n_stratified=s1)
    for a, b in zip(ref, got):
        assert (a is None and b is None) or torch.equal(a, b)
    alpha, trans, w = ref[0], ref[1], ref[2]
    g_acc = torch.randn(n, c, generator=gen).to(cuda)
    g_depth, g_wsum = torch.randn(n, generator=gen).to(cuda), torch.randn(n, generator=gen).to(cuda)
    g_w = torch.randn(n, s, generator=gen).to(cuda)
    gs = ops.composite_backward(z, sorted_rows, noise, 0.1, None, alpha, trans, w, g_acc, None, g_depth, g_wsum, g_w, None)
    gr = ops.composite_backward(z, rows, noise, 0.1, None, alpha, trans, w, g_acc, None, g_depth, g_wsum, g_w, None,
                                sort_idx=idx, n_stratified=s1)
    assert torch.equal(gr, ops.permute_samples(gs, idx, n, s1, g, c, scatter=True))


@pytest.mark.parametrize("n,with_depth,use_all", [(256, True, False), (33, True, True), (64, False, False)])
def test_lambertian_render_loss_equals_the_five_launches(cuda, n, with_depth, use_all):
    """bn_lambertian_render_loss (the Trainer's Lambertian K-C: compositing + colour + loss + their backward in one launch, rows
    gathered / scattered through sort_idx) against bn_permute_samples -> bn_composite_forward -> bn_shade_rays_forward ->
    bn_loss_color_depth -> bn_shade_rays_backward -> bn_composite_backward -> bn_permute_samples."""
    from brdf_nerf_b200 import _lib as L
    s1 = g = 64
    s, c = s1 + g, 4
    gen = torch.Generator().manual_seed(77 + n)
    z, packed, _ = _inputs(n, s, c, 9)
    packed[..., 3] = packed[..., 3] * (torch.rand(n, s, generator=gen) > 0.3)           # some empty space
    rows = packed.reshape(n * s, c).to(cuda)
    idx = torch.stack([torch.randperm(s, generator=gen) for _ in range(n)]).to(cuda)
    z = z.to(cuda)
    target = torch.rand(n, 3, generator=gen).to(cuda)
    valid = (torch.rand(n, generator=gen) < 0.7).to(torch.int64).to(cuda)
    tdep = torch.stack([torch.rand(n, generator=gen) * 0.6, torch.rand(n, generator=gen)], -1).to(cuda)
    tstd = (torch.rand(n, generator=gen) * 0.05).to(cuda)
    rays = torch.zeros(n, 11, device=cuda)
    kw = dict(valid_depth=valid if with_depth else None, target_depths=tdep, target_std=tstd,
              lambda_ds=10.0 if with_depth else 0.0, use_all_depth=use_all)
    # five launches (+ two permutations)
    srt = ops.permute_samples(rows, idx, n, s1, g, c, scatter=False)
    alpha, trans, w, depth, wsum, acc, _ = ops.composite_forward(z, srt, None, 0.0)
    cfg = L.ShadeCfg()
    cfg.n_channels, cfg.normal_ch, cfg.param_ch, cfg.brdf_ch = c, -1, -1, -1
    cfg.brdf_type, cfg.irr_mode = L.BN_BRDF_NONE, L.BN_IRR_ONES
    sh = ops.shade_rays_forward(cfg, rays, acc, wsum, None, None, want_normal=False, want_brdf=False)
    loss, g_rgb, g_depth = ops.loss_color_depth(sh["rgb"], target, 1.0, depth=depth, z=z, weights=w, **kw)
    g_acc, g_wsum, _ = ops.shade_rays_backward(cfg, rays, acc, wsum, None, None, g_rgb)
    gs = ops.composite_backward(z, srt, None, 0.0, None, alpha, trans, w, g_acc, None, g_depth, g_wsum, None, None)
    ref_rows = ops.permute_samples(gs, idx, n, s1, g, c, scatter=True)
    # one launch
    loss1, g_rows, rgb1, depth1 = ops.lambertian_render_loss(z, rows, idx, s1, target, 1.0, want_outputs=True, **kw)
    assert torch.equal(rgb1, sh["rgb"]) and torch.equal(depth1, depth)
    assert abs(loss1.item() - loss.item()) <= 1e-6 * abs(loss.item()) + 1e-9          # same terms, atomics in another order
    assert torch.equal(g_rows, ref_rows)


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 * 33 + 5])
def test_device_nan_counter(cuda, n):
    """bn_count_nan == torch.isnan(x).sum() (train_utils.check_nan, train_utils.py:61-78), accumulated without a host sync;
    ragged sizes, unaligned views, infinities do not count."""
    from brdf_nerf_b200 import ops
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n + 1, generator=g)
    x[torch.rand(n + 1, generator=g) < 0.1] = float("nan")
    x[torch.rand(n + 1, generator=g) < 0.05] = float("inf")
    xd = x.to(cuda)
    a, b = xd[:n], xd[1:]                                     # b is 4-byte aligned only
    counters = ops.count_nan([a, b, a])
    assert counters.tolist() == [int(torch.isnan(x[:n]).sum()), int(torch.isnan(x[1:]).sum()), int(torch.isnan(x[:n]).sum())]
    ops.count_nan([a], counters)                              # accumulates into slot 0
    assert counters[0].item() == 2 * int(torch.isnan(x[:n]).sum())
