"""K-A parity: CUDA sample generators vs the numpy oracle, BIT-EXACT (z-values, indices)."""
import numpy as np
import pytest
import torch

from brdf_nerf_b200 import ops
from brdf_nerf_b200.synth import make_rays
from oracle import sampler_np as SN

pytestmark = pytest.mark.gpu


def _pass1_like(n, s, seed):
    """Plausible pass-1 outputs (ascending z, weights summing to <= 1, depth inside the ray)."""
    g = torch.Generator().manual_seed(seed)
    b = make_rays(n, seed=seed, depth_supervision=True)
    t, _ = SN.tables(s, 3.0)
    u = torch.rand(n, s, generator=g)
    z1 = SN.stratified_z(b.rays[:, 6].numpy(), b.rays[:, 7].numpy(), t, u.numpy())
    w = torch.softmax(torch.randn(n, s, generator=g) * 3, -1).numpy().astype(np.float32) * 0.9
    depth = (w * z1).sum(-1).astype(np.float32)
    return b, u, z1, w, depth


@pytest.mark.parametrize("n,s", [(1, 64), (257, 64), (1024, 64), (33, 48), (7, 128)])
def test_stratified_bit_exact(cuda, n, s):
    b = make_rays(n, seed=n * 7 + s)
    t, _ = SN.tables(s, 3.0)
    u = torch.rand(n, s, generator=torch.Generator().manual_seed(5))
    ref = SN.stratified_z(b.rays[:, 6].numpy(), b.rays[:, 7].numpy(), t, u.numpy())
    rays = b.rays.to(cuda)
    tv, _ = ops.sampler_tables(s, 3.0, cuda)
    z = ops.sample_stratified(rays[:, 6], rays[:, 7], 11, tv, u.to(cuda)).cpu().numpy()
    assert z.dtype == np.float32 and np.array_equal(z.view(np.uint32), ref.view(np.uint32))
    assert np.all(np.diff(z, axis=-1) >= 0)


@pytest.mark.parametrize("n,s,with_gt,zero_std", [(257, 64, False, False), (1024, 64, True, False),
                                                  (300, 64, True, True), (19, 48, True, False), (5, 128, False, False)])
def test_guided_and_merge_bit_exact(cuda, n, s, with_gt, zero_std):
    b, u, z1, w, depth = _pass1_like(n, s, seed=100 + n)
    g = torch.Generator().manual_seed(9)
    u_pred, u_gt = torch.rand(n, s, generator=g), torch.rand(n, s, generator=g)
    t, gw = SN.tables(s, 3.0)
    near0, far0 = float(b.rays[0, 6]), float(b.rays[0, 7])
    std_t = torch.zeros(n) if zero_std else b.target_std
    ref_z2 = SN.guided_z(z1, depth, w, near0, far0, 3.0, t, gw, u_pred.numpy(),
                         valid_depth=b.valid_depth.numpy() if with_gt else None,
                         gt_depth=b.target_depths[:, 0].numpy(), gt_std=std_t.numpy(), u_gt=u_gt.numpy())
    rays = b.rays.to(cuda)
    tv, gv = ops.sampler_tables(s, 3.0, cuda)
    td = b.target_depths.to(cuda)
    z2, std = ops.sample_guided(torch.from_numpy(z1).to(cuda), torch.from_numpy(depth).to(cuda),
                                torch.from_numpy(w).to(cuda), tv, gv, u_pred.to(cuda), rays[0:1, 6], rays[0:1, 7], 3.0,
                                valid_depth=b.valid_depth.to(cuda) if with_gt else None, gt_depth=td if with_gt else None,
                                gt_depth_stride=2, gt_std=std_t.to(cuda) if with_gt else None,
                                u_gt=u_gt.to(cuda) if with_gt else None, want_std=True)
    z2n = z2.cpu().numpy()
    mism = int((z2n.view(np.uint32) != ref_z2.view(np.uint32)).sum())
    assert mism == 0, f"{mism} guided z-values differ from the oracle bit pattern"
    assert np.array_equal(std.cpu().numpy().view(np.uint32), SN.depth_std(z1, depth, w).view(np.uint32))
    # merge: sorted values bit-exact; indices exact modulo ties (torch.sort leaves tie order unspecified)
    rz, ridx, run = SN.merge_sorted(z1, ref_z2)
    z, idx, unsort = ops.merge_samples(torch.from_numpy(z1).to(cuda), z2)
    assert np.array_equal(z.cpu().numpy().view(np.uint32), rz.view(np.uint32))
    assert np.array_equal(unsort.cpu().numpy().view(np.uint32), run.view(np.uint32))
    idx = idx.cpu().numpy()
    assert idx.dtype == np.int64 and np.array_equal(np.sort(idx, -1), np.tile(np.arange(2 * s), (n, 1)))
    assert np.array_equal(np.take_along_axis(run, idx, 1).view(np.uint32), rz.view(np.uint32))
    assert np.array_equal(idx, ridx)          # this implementation is stable, like the oracle
    if zero_std and with_gt:
        # std 0 collapses the guided samples onto the GT depth up to the rounding of c(1-t)+ct (App. C.1)
        v = b.valid_depth.numpy() > 0
        gt = b.target_depths[:, 0].numpy()[v]
        assert np.abs(z2n[v] - gt[:, None]).max() <= 1e-6


@pytest.mark.parametrize("n,s,with_gt,noise_std", [(1024, 64, True, 0.0), (257, 64, False, 0.0), (300, 64, True, 1.0),
                                                   (19, 48, True, 0.0), (5, 128, False, 0.0)])
def test_coarse_to_fine_equals_the_three_launches(cuda, n, s, with_gt, noise_std):
    """bn_coarse_to_fine (compositing of the stratified densities + guided samples + merge in one launch, the path render_rays
    takes) == bn_composite_sigma -> bn_sample_guided -> bn_merge_samples, bit for bit, on every output."""
    b, u, z1, _, _ = _pass1_like(n, s, seed=200 + n)
    g = torch.Generator().manual_seed(11)
    sigma = (torch.rand(n, s, generator=g) * 8.0 * (torch.rand(n, s, generator=g) > 0.5)).to(cuda)
    noise = torch.randn(n, s, generator=g).to(cuda)
    u_pred, u_gt = torch.rand(n, s, generator=g).to(cuda), torch.rand(n, s, generator=g).to(cuda)
    rays = b.rays.to(cuda)
    tv, gv = ops.sampler_tables(s, 3.0, cuda)
    z1c = torch.from_numpy(z1).to(cuda)
    kw = dict(valid_depth=b.valid_depth.to(cuda) if with_gt else None, gt_depth=b.target_depths.to(cuda) if with_gt else None,
              gt_depth_stride=2, gt_std=b.target_std.to(cuda) if with_gt else None, u_gt=u_gt if with_gt else None, want_std=True)
    _, _, w1, d1, _ = ops.composite_sigma(z1c, sigma, noise, noise_std)
    z2, std = ops.sample_guided(z1c, d1, w1, tv, gv, u_pred, rays[0:1, 6], rays[0:1, 7], 3.0, **kw)
    z, idx, unsort = ops.merge_samples(z1c, z2)
    f = ops.coarse_to_fine(z1c, sigma, noise, noise_std, tv, gv, u_pred, rays[0:1, 6], rays[0:1, 7], 3.0, **kw)
    for name, a, c in zip(("weights", "depth", "std", "z2", "z", "idx", "unsort"), f, (w1, d1, std, z2, z, idx, unsort)):
        assert torch.equal(a, c), name
    assert bool(torch.isfinite(f[4]).all()) and bool((f[4][:, 1:] >= f[4][:, :-1]).all())


def test_sort_rows(cuda):
    x = torch.rand(77, 64, generator=torch.Generator().manual_seed(1))
    out = ops.sort_rows(x.to(cuda)).cpu()
    assert torch.equal(out, torch.sort(x, -1)[0])


def test_empty_batch_and_bad_args(cuda):
    from brdf_nerf_b200 import _lib as L
    tv, gv = ops.sampler_tables(64, 3.0, cuda)
    u = torch.empty(0, 64, device=cuda)
    rays = torch.zeros(1, 11, device=cuda)
    assert ops.sample_stratified(rays[:, 6], rays[:, 7], 11, tv, u).shape == (0, 64)
    with pytest.raises(L.BnError):
        L.check(L.load().bn_sample_guided(None, None, None, None, None, None, None, None, 3.0, None, None, 1, None,
                                          None, None, None, 4, 64, 64, None))
