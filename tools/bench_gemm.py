"""Micro-benchmark of tt_gemm_bf16_tn vs torch.matmul (cuBLAS) on the path's GEMM shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


shapes = [(8192, 1024, 1024), (8192, 3072, 1024), (8192, 4096, 1024), (8192, 1024, 4096),
          (800, 1024, 1024), (800, 2048, 1024), (800, 4096, 1024), (800, 1024, 4096),
          (800, 5002, 1024), (8192, 8192, 8192), (16384, 4096, 4096)]
for M, N, K in shapes:
    a = torch.randn(M, K, device='cuda').bfloat16()
    b = torch.randn(N, K, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda')
    out16 = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    t32 = timeit(lambda: ops.gemm_tn(a, b, out=out))
    t16 = timeit(lambda: ops.gemm_tn(a, b, out16=out16, want32=False))
    tc = timeit(lambda: torch.matmul(a, b.t()))
    fl = 2.0 * M * N * K
    print('M=%5d N=%5d K=%5d  tt(fp32 out) %7.1f us %7.1f TF | tt(bf16 out) %7.1f us %7.1f TF | cublas %7.1f us %7.1f TF'
          % (M, N, K, t32 * 1e6, fl / t32 / 1e12, t16 * 1e6, fl / t16 / 1e12, tc * 1e6, fl / tc / 1e12))
