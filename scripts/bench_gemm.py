"""GEMM micro-benchmark on one B200: isolates mainloop / epilogue costs of the tcgen05 kernel.
    python scripts/bench_gemm.py            (prints one line per variant: us, TFLOP/s)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brdf_nerf_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    P, N, K = 131072, 512, 512
    g = torch.Generator().manual_seed(0)
    # several operand sets so that consecutive launches do not hit in L2 (each set ~0.5 GB)
    sets = []
    for i in range(3):
        A = (torch.randn(P, K, generator=g) * 0.5).to(dev).to(torch.bfloat16)
        mul = torch.randn(P, N, generator=g).to(dev).to(torch.bfloat16)
        out = torch.empty(P, N, dtype=torch.bfloat16, device=dev)
        sets.append((A, mul, out))
    B = (torch.randn(N, K, generator=g) * 0.1).to(dev).to(torch.bfloat16)
    sink = torch.zeros(16, dtype=torch.float32, device=dev)
    cs = torch.zeros(N, dtype=torch.float32, device=dev)
    dW = torch.zeros(N, K, dtype=torch.float32, device=dev)
    flops = 2.0 * P * N * K
    k = [0]

    def nxt():
        k[0] = (k[0] + 1) % len(sets)
        return sets[k[0]]

    def null():
        A, _, _ = nxt()
        L.check(lib.bn_debug_gemm(2, L.BN_PREC_BF16, vp(A), K, vp(B), K, vp(sink), N, P, N, K, L.stream_ptr()))

    def plain():
        A, _, out = nxt()
        L.check(lib.bn_debug_gemm_epi(0, vp(A), K, vp(B), K, vp(out), N, None, None, None, 0, 0, P, N, K, L.stream_ptr()))

    def mul():
        A, m, out = nxt()
        L.check(lib.bn_debug_gemm_epi(0, vp(A), K, vp(B), K, vp(out), N, None, vp(m), None, 0, 0, P, N, K, L.stream_ptr()))

    def mulcs():
        A, m, out = nxt()
        L.check(lib.bn_debug_gemm_epi(0, vp(A), K, vp(B), K, vp(out), N, None, vp(m), vp(cs), 0, 0, P, N, K, L.stream_ptr()))

    def wgrad():
        A, m, _ = nxt()
        L.check(lib.bn_debug_gemm_epi(1, vp(A), K, vp(m), N, vp(dW), K, None, None, None, K, K, K, N, P, L.stream_ptr()))

    bias = torch.zeros(K, dtype=torch.float32, device=dev)

    def wgrad_bias():
        A, m, _ = nxt()
        L.check(lib.bn_debug_gemm_epi(1, vp(A), K, vp(m), N, vp(dW), K, None, None, vp(bias), K, K, K, N, P, L.stream_ptr()))

    def cublas():
        A, _, out = nxt()
        torch.matmul(A, B.t(), out=out)

    for name, fn in (("mainloop only (null epilogue)", null), ("bf16 store", plain), ("mul + store (dgrad)", mul),
                     ("mul + store + colsum", mulcs), ("wgrad NT + TMA reduce", wgrad), ("wgrad NT + TMA reduce + bias MMA", wgrad_bias), ("torch.matmul (cuBLAS) bf16 out", cublas)):
        us = timeit(fn)
        print(f"{name:34s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
