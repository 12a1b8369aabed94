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
