"""tcgen05 / TMEM / TMA GEMM (bf16) and the CUDA-core GEMM (fp32) in isolation, against torch.matmul."""
import ctypes as C

import pytest
import torch

from brdf_nerf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _gemm(kind, prec, A, B, M, N, K, out=None):
    ldo = N
    if out is None:
        out = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    L.check(L.load().bn_debug_gemm(kind, prec, C.c_void_p(A.data_ptr()), A.stride(0), C.c_void_p(B.data_ptr()), B.stride(0),
                                   C.c_void_p(out.data_ptr()), ldo, M, N, K, L.stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 512, 512), (1000, 512, 576), (4096, 512, 512), (333, 256, 512),
                                   (2048, 1024, 512), (256, 128, 64), (130, 64, 512)])
def test_tn_tcgen05_bf16(cuda, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    out = _gemm(0, L.BN_PREC_BF16, A, B, M, N, K)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 TN gemm max err {err}"


def test_tn_tcgen05_strided_operand(cuda):
    """A operand with a row pitch larger than K (the [enc | h] joint buffer of the skip layer)."""
    M, N, K, ld = 512, 512, 64, 576
    g = torch.Generator().manual_seed(1)
    buf = (torch.randn(M, ld, generator=g)).to(cuda).to(torch.bfloat16)
    A = buf[:, :K]
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    out = _gemm(0, L.BN_PREC_BF16, A, B, M, N, K)
    ref = A.float() @ B.float().t()
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("Mo,No,P", [(512, 512, 4096), (512, 576, 1000), (256, 512, 8192), (512, 64, 3000), (128, 256, 64)])
def test_nt_tcgen05_bf16(cuda, Mo, No, P):
    g = torch.Generator().manual_seed(Mo + No + P)
    A = (torch.randn(P, Mo, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    B = (torch.randn(P, No, generator=g) * 0.5).to(cuda).to(torch.bfloat16)
    ref = A.float().t() @ B.float()
    out = _gemm(1, L.BN_PREC_BF16, A, B, Mo, No, P)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 NT gemm max err {err}"


@pytest.mark.parametrize("Mo,No,P", [(512, 512, 4096), (512, 576, 131072), (512, 512, 1000)])
def test_nt_multicast_variant(cuda, monkeypatch, Mo, No, P):
    """BN_NT_MC=1 (read per call): clusters of two CTA pairs, the B tile multicast between them (gemm_tc.cuh, kMC).  Off by default
    (measured: no gain, the weight gradient is HBM-bound); kept correct."""
    monkeypatch.setenv("BN_NT_MC", "1")
    g = torch.Generator().manual_seed(Mo + No + P)
    A = (torch.randn(P, Mo, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    B = (torch.randn(P, No, generator=g) * 0.5).to(cuda).to(torch.bfloat16)
    ref = A.float().t() @ B.float()
    out = _gemm(1, L.BN_PREC_BF16, A, B, Mo, No, P)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 NT gemm (multicast) max err {err}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 512, 576), (77, 256, 512)])
def test_tn_simt_fp32(cuda, M, N, K):
    g = torch.Generator().manual_seed(3)
    A = torch.randn(M, K, generator=g).to(cuda)
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A @ B.t()
    out = _gemm(0, L.BN_PREC_FP32, A, B, M, N, K)
    assert (out - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("Mo,No,P", [(512, 512, 1000), (512, 576, 333), (16, 512, 4096)])
def test_nt_simt_fp32(cuda, Mo, No, P):
    g = torch.Generator().manual_seed(4)
    A = (torch.randn(P, Mo, generator=g) * 0.1).to(cuda)
    B = torch.randn(P, No, generator=g).to(cuda)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A.t() @ B
    out = _gemm(1, L.BN_PREC_FP32, A, B, Mo, No, P)
    assert (out - ref).abs().max().item() < 1e-3


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("M,N,K,use_add,use_mul", [(4096, 512, 512, False, True), (1000, 512, 512, True, True), (300, 256, 64, False, False),
                                                   (2048, 768, 512, False, True), (130, 64, 512, True, False), (20000, 512, 576, False, True)])
def test_tn_tma_epilogue(cuda, M, N, K, use_add, use_mul):
    """TMA-staged dgrad epilogue: operand boxes in, bf16 boxes out, fused column sums; ragged M."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    add = (torch.randn(M, N, generator=g)).to(cuda).to(torch.bfloat16) if use_add else None
    mul = (torch.randn(M, N, generator=g)).to(cuda).to(torch.bfloat16) if use_mul else None
    out = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=cuda)
    cs = torch.zeros(N, dtype=torch.float32, device=cuda)
    L.check(L.load().bn_debug_gemm_epi(0, _vp(A), A.stride(0), _vp(B), B.stride(0), _vp(out), N, _vp(add), _vp(mul), _vp(cs),
                                       0, 0, M, N, K, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if use_add:
        ref = ref + add.float()
    if use_mul:
        ref = ref * mul.float()
    scale = max(1.0, ref.abs().max().item())
    assert (out.float() - ref).abs().max().item() < 1e-2 * scale            # bf16 rounding of the output
    assert (cs - ref.sum(0)).abs().max().item() < 2e-3 * scale * (M ** 0.5) + 1e-2 * scale


@pytest.mark.parametrize("Mo,No,P,pad_lo,pad_hi", [(512, 512, 4096, 512, 512), (512, 576, 3000, 60, 64), (512, 64, 5000, 60, 64),
                                                   (256, 512, 8192, 512, 512)])
def test_nt_tma_reduce_epilogue(cuda, Mo, No, P, pad_lo, pad_hi):
    """wgrad through cp.reduce.async.bulk (fp32 add) with the [enc|pad|h] -> Linear(572) column remap."""
    g = torch.Generator().manual_seed(Mo + No + P)
    A = (torch.randn(P, Mo, generator=g) * 0.1).to(cuda).to(torch.bfloat16)
    B = (torch.randn(P, No, generator=g) * 0.5).to(cuda).to(torch.bfloat16)
    kreal = No - (pad_hi - pad_lo)
    base = torch.randn(Mo, kreal, generator=g).to(cuda)
    out = base.clone()
    bias = torch.full((Mo,), 3.0, dtype=torch.float32, device=cuda)
    L.check(L.load().bn_debug_gemm_epi(1, _vp(A), A.stride(0), _vp(B), B.stride(0), _vp(out), kreal, None, None, _vp(bias),
                                       pad_lo, pad_hi, Mo, No, P, L.stream_ptr()))
    torch.cuda.synchronize()
    # bias gradient fused into the mainloop (N=16 MMA against ones): column sums of the A operand
    bref = 3.0 + A.float().sum(0)
    assert (bias - bref).abs().max().item() < 2e-3 * max(1.0, bref.abs().max().item())
    full = A.float().t() @ B.float()
    ref = base + torch.cat([full[:, :pad_lo], full[:, pad_hi:]], dim=1)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"TMA reduce wgrad max err {err}"
