"""K-B parity: PE + SIREN MLP forward / backward vs the torch oracle.
fp32 mode: <= 1e-3 abs (north-star tolerance; observed ~1e-5).  bf16 (tcgen05) mode: loose check
against the fp32 mode — its acceptance criterion is the PSNR-drift test in test_gpu_train.py."""
import pytest
import torch

from brdf_nerf_b200.config import named_config
from brdf_nerf_b200.models import load_model
from oracle import render_torch as RT

pytestmark = pytest.mark.gpu


def _models(cfg, cuda, precision="fp32", **over):
    args = named_config(cfg, **over)
    torch.manual_seed(0)
    m = load_model(args, precision=precision)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(cuda)
    return args, m, state


def _pts(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, generator=g) * 1.6 - 0.8


@pytest.mark.parametrize("cfg,kw", [("lambertian", {}), ("rpv111", dict(apply_brdf=True)),
                                    ("hapke_bct", dict(apply_brdf=True, apply_theta=True)),
                                    ("microfacet", dict(apply_brdf=True)), ("rpv111_multi", dict(apply_brdf=True))])
def test_forward_fp32(cuda, cfg, kw):
    args, m, state = _models(cfg, cuda, normal="none")
    x = _pts(777)
    om = RT.OracleModel(state, args)
    with torch.no_grad():
        ref = om.forward(x, **kw)
        sig = m(x.to(cuda), sigma_only=True).cpu()
        out = m(x.to(cuda), **kw).cpu()
    assert torch.allclose(sig, ref["sigma"], atol=1e-4), (sig - ref["sigma"]).abs().max()
    cols = [ref["albedo"], ref["sigma"]]
    for k in ("roughness", "rpv_k", "rpv_theta", "rpv_rhoc", "hpk_b", "hpk_c", "hpk_theta"):
        if k in ref:
            cols.append(ref[k])
    want = torch.cat(cols, -1)
    assert out.shape == want.shape, (out.shape, want.shape)
    err = (out - want).abs().max().item()
    assert err < 1e-4, f"{cfg}: packed output max err {err}"


def test_forward_learned_normal_fp32(cuda):
    args, m, state = _models("rpv111", cuda, normal="learned")
    x = _pts(300, 2)
    om = RT.OracleModel(state, args)
    with torch.no_grad():
        ref = om.forward(x, nr_lr=True, apply_brdf=True)
        out = m(x.to(cuda), nr_lr_on=True, apply_brdf=True).cpu()
    want = torch.cat([ref["albedo"], ref["sigma"], ref["normal_lr"], ref["rpv_k"], ref["rpv_theta"], ref["rpv_rhoc"]], -1)
    assert (out - want).abs().max().item() < 2e-4


@pytest.mark.parametrize("cfg,kw", [("lambertian", {}), ("rpv111", dict(apply_brdf=True)),
                                    ("hapke_bct", dict(apply_brdf=True, apply_theta=True))])
def test_backward_fp32(cuda, cfg, kw):
    args, m, state = _models(cfg, cuda, normal="none")
    x = _pts(513, 4)
    om = RT.OracleModel(state, args, requires_grad=True)
    ref = om.forward(x, **kw)
    cols = [ref["albedo"], ref["sigma"]] + [ref[k] for k in ("rpv_k", "rpv_theta", "rpv_rhoc", "hpk_b", "hpk_c", "hpk_theta") if k in ref]
    want = torch.cat(cols, -1)
    G = torch.randn(want.shape, generator=torch.Generator().manual_seed(8))
    (want * G).sum().backward()
    out = m(x.to(cuda), **kw)
    (out * G.to(cuda)).sum().backward()
    worst = 0.0
    for name, p in m.named_parameters():
        r = om.p[name].grad
        if r is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, name
            continue
        d = (p.grad.cpu() - r).abs().max().item()
        s = r.abs().max().item()
        worst = max(worst, d / (s + 1e-12))
        assert d <= 2e-3 * s + 1e-6, f"{cfg} grad {name}: max diff {d} (scale {s})"
    print(f"{cfg}: worst relative grad error {worst:.2e}")


def test_bf16_tcgen05_close_to_fp32(cuda):
    args, m32, state = _models("rpv111", cuda, normal="none")
    _, m16, _ = _models("rpv111", cuda, precision="bf16", normal="none")
    x = _pts(2000, 6).to(cuda)
    with torch.no_grad():
        a = m32(x, apply_brdf=True)
        b = m16(x, apply_brdf=True)
    err = (a - b).abs().max().item()
    print("bf16 vs fp32 packed max abs diff", err, "mean", (a - b).abs().mean().item())
    assert err < 0.15 and (a - b).abs().mean().item() < 0.02


def test_bf16_backward_direction(cuda):
    """Gradient of the bf16 path points the same way as the fp32 gradient (cosine > 0.98)."""
    args, m32, state = _models("lambertian", cuda, normal="none")
    _, m16, _ = _models("lambertian", cuda, precision="bf16", normal="none")
    x = _pts(4096, 7).to(cuda)
    G = torch.randn(4096, 4, generator=torch.Generator().manual_seed(1)).to(cuda)
    for m in (m32, m16):
        m.flat_grads.zero_()
        (m(x) * G).sum().backward()
    a, b = m32.flat_grads, m16.flat_grads
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    print("bf16 vs fp32 gradient cosine", cos)
    assert cos > 0.98


def _an_cols(ref):
    cols = [ref["albedo"], ref["sigma"], ref["normal_an"]]
    for k in ("roughness", "rpv_k", "rpv_theta", "rpv_rhoc", "hpk_b", "hpk_c", "hpk_theta"):
        if k in ref:
            cols.append(ref[k])
    return torch.cat(cols, -1)


@pytest.mark.parametrize("cfg", ["rpv111", "lambertian"])
def test_analytic_normals_forward_fp32(cuda, cfg):
    """K-B2 forward sweep vs autograd.grad of the oracle (reference calc_normals).  Raw per-sample
    normals are ill-conditioned where |grad sigma| ~ 0 (SURVEY §8a N-note): compare the raw gradient
    direction where |grad| exceeds a floor, tolerance 1e-3."""
    args, m, state = _models(cfg, cuda, normal="analystic")
    x = _pts(1500, 12)
    om = RT.OracleModel(state, args)
    ref = om.forward(x, nr_an=True, apply_brdf=(cfg != "lambertian"))
    ref = {k: v.detach() for k, v in ref.items()}
    with torch.no_grad():
        out = m(x.to(cuda), nr_an_on=True, apply_brdf=(cfg != "lambertian")).cpu()
    want = _an_cols(ref)
    assert out.shape == want.shape
    # recompute the oracle's raw gradient norm to select well-conditioned points
    xx = x.clone().requires_grad_(True)
    sig = om.forward(xx, sigma_only=True)["sigma"]
    (g,) = torch.autograd.grad(sig.sum(), xx)
    ok = g.norm(dim=-1) > 1e-2
    assert ok.float().mean() > 0.5
    err = (out[ok] - want[ok]).abs().max().item()
    assert err < 1e-3, f"analytic normals (well-conditioned points) max err {err}"
    assert (out[:, :4] - want[:, :4]).abs().max().item() < 1e-4


@pytest.mark.parametrize("cfg", ["rpv111", "hapke_bct"])
def test_analytic_normals_second_order_backward_fp32(cuda, cfg):
    """Gradient of a loss on [albedo, sigma, normal_an, brdf params] w.r.t. every weight: the double
    backward through calc_normals (create_graph=True in the reference) vs the hand-derived sweep."""
    args, m, state = _models(cfg, cuda, normal="analystic")
    n = 384
    x = _pts(n, 21)
    kw = dict(apply_brdf=True, apply_theta=True) if cfg == "hapke_bct" else dict(apply_brdf=True)
    om = RT.OracleModel(state, args, requires_grad=True)
    ref = om.forward(x, nr_an=True, **kw)
    want = _an_cols(ref)
    G = torch.randn(want.shape, generator=torch.Generator().manual_seed(8))
    G[:, 4:7] *= 0.05      # keep the ill-conditioned normal term from dominating the comparison
    (want * G).sum().backward()
    out = m(x.to(cuda), nr_an_on=True, **{("apply_brdf" if k == "apply_brdf" else k): v for k, v in kw.items()})
    m.flat_grads.zero_()
    (out * G.to(cuda)).sum().backward()
    worst = 0.0
    for name, p in m.named_parameters():
        r = om.p[name].grad
        if r is None:
            continue
        d = (p.grad.cpu() - r).abs().max().item()
        s = r.abs().max().item()
        worst = max(worst, d / (s + 1e-12))
        assert d <= 5e-3 * s + 1e-6, f"{cfg} second-order grad {name}: max diff {d} (scale {s})"
    print(f"{cfg}: worst relative second-order grad error {worst:.2e}")


@pytest.mark.parametrize("n", [100, 256, 1000, 70001])
def test_sigma_chain_bf16(cuda, n, monkeypatch):
    """Fused density pass (mlp_chain.cuh: all trunk layers + sigma head in one tcgen05 kernel) against
    the fp32 oracle and against the per-layer bf16 GEMM path (BN_NO_CHAIN=1), ragged point counts."""
    args, m, state = _models("lambertian", cuda, precision="bf16")
    x = _pts(n, 11)
    om = RT.OracleModel(state, args)
    with torch.no_grad():
        ref = om.forward(x)["sigma"]
        sig = m(x.to(cuda), sigma_only=True).cpu()
    monkeypatch.setenv("BN_NO_CHAIN", "1")
    args2, m2, _ = _models("lambertian", cuda, precision="bf16")
    with torch.no_grad():
        sig_layered = m2(x.to(cuda), sigma_only=True).cpu()
    e_ref = (sig - ref).abs().max().item()
    e_lay = (sig_layered - ref).abs().max().item()
    print(f"n={n}: chain vs fp32 oracle {e_ref:.3e}   per-layer bf16 vs oracle {e_lay:.3e}   chain vs per-layer {(sig - sig_layered).abs().max().item():.3e}")
    assert torch.isfinite(sig).all()
    # bf16 activations: same error class as the per-layer bf16 path
    assert e_ref <= max(2.0 * e_lay, 2e-2), (e_ref, e_lay)


@pytest.mark.parametrize("cfg,kw,n", [("lambertian", {}, 1000), ("lambertian", {}, 70001),
                                      ("rpv111", dict(apply_brdf=True, nr_an_on=True), 3000)])
def test_inference_chain_bf16(cuda, cfg, kw, n, monkeypatch):
    """Inference forward in bf16: the fused trunk kernel (train_chain_kernel with store_c = 0 / h_from = L-1: only the
    last layer's activations and, for analytic normals, the cosines leave the SM) against the per-layer tcgen05 GEMMs
    (BN_NO_CHAIN=1) and the fp32 mode."""
    over = dict(normal="analystic") if kw.get("nr_an_on") else {}
    _, m32, _ = _models(cfg, cuda, **over)
    _, m16, _ = _models(cfg, cuda, precision="bf16", **over)
    x = _pts(n, 21).to(cuda)
    with torch.no_grad():
        a = m32(x, **kw)
        b = m16(x, **kw)
    monkeypatch.setenv("BN_NO_CHAIN", "1")
    _, m16l, _ = _models(cfg, cuda, precision="bf16", **over)
    with torch.no_grad():
        c = m16l(x, **kw)
    assert torch.isfinite(b).all()
    cols = [0, 1, 2, 3] + list(range(7, b.shape[1])) if kw.get("nr_an_on") else list(range(b.shape[1]))
    e_chain = (a - b)[:, cols].abs().max().item()
    e_layer = (a - c)[:, cols].abs().max().item()
    print(f"{cfg} n={n}: chain vs fp32 {e_chain:.3e}, per-layer bf16 vs fp32 {e_layer:.3e}")
    assert e_chain <= max(2.0 * e_layer, 5e-2), (e_chain, e_layer)
    if kw.get("nr_an_on"):        # unit normals: direction agrees where the per-layer bf16 path agrees with fp32
        na, nb, nc = a[:, 4:7], b[:, 4:7], c[:, 4:7]
        ok = (na * nc).sum(-1) > 0.9
        assert ((na * nb).sum(-1)[ok] > 0.8).float().mean().item() > 0.98


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
@pytest.mark.parametrize("mapping", [True, False])
def test_viewdir_colour_head(cuda, precision, tol, mapping):
    """--input_viewdir: rgb_from_xyzdir reads [features | Mapping(direction)] (spsbrdfnerf.py:689-690); forward against the
    oracle and, in fp32, the gradient of the colour head's first layer (its direction columns included)."""
    args, m, state = _models("lambertian", cuda, precision=precision, input_viewdir=1, mapping=mapping)
    x = _pts(777, 31)
    g = torch.Generator().manual_seed(32)
    d = torch.nn.functional.normalize(torch.randn(777, 3, generator=g), dim=-1)
    om = RT.OracleModel(state, args, requires_grad=True)
    ref = om.forward(x, d=d)
    want = torch.cat([ref["albedo"], ref["sigma"]], -1)
    out = m(x.to(cuda), input_dir=d.to(cuda))
    assert out.shape == want.shape
    err = (out.detach().cpu() - want.detach()).abs().max().item()
    assert err <= tol * max(1.0, want.detach().abs().max().item()), err
    with pytest.raises(ValueError):
        m(x.to(cuda))
    if precision == "fp32":
        w = torch.randn(777, 4, generator=g)
        (want * w).sum().backward()
        m.flat_grads.zero_()
        (out * w.to(cuda)).sum().backward()
        for name, p in m.named_parameters():
            r = om.p[name].grad
            s = r.abs().max().item()
            assert (p.grad.cpu() - r).abs().max().item() <= 2e-3 * s + 1e-7, name
