"""Epilogue cost study: RoBERTa-layer GEMM shapes with/without bias, GELU, residual, rotating
buffers (cold weights) vs hot L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402


def run(M, N, K, o16, bias, act, res16, rotate):
    nb = 8 if rotate else 1
    A = [torch.randn(M, K, device='cuda').bfloat16() for _ in range(nb)]
    W = [torch.randn(N, K, device='cuda').bfloat16() for _ in range(nb)]
    O = [torch.empty(M, N, device='cuda', dtype=torch.bfloat16 if o16 else torch.float32) for _ in range(nb)]
    b = torch.randn(N, device='cuda') if bias else None
    R = [torch.randn(M, N, device='cuda').bfloat16() for _ in range(nb)] if res16 else None

    def one(i):
        kw = dict(out16=O[i % nb], want32=False) if o16 else dict(out=O[i % nb])
        ops.gemm_tn(A[i % nb], W[i % nb], bias=b, act=act, residual16=R[i % nb] if R else None, **kw)
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            one(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 80 * 1e3
    print('M=%5d N=%5d K=%5d o16=%d bias=%d act=%d res16=%d rotate=%d  %7.1f us %7.1f TF'
          % (M, N, K, o16, bias, act, res16, rotate, us, 2.0 * M * N * K / us / 1e6))


for rot in (0, 1):
    run(8192, 3072, 1024, 1, 0, 0, 0, rot)
    run(8192, 3072, 1024, 1, 1, 0, 0, rot)
    run(8192, 4096, 1024, 1, 0, 0, 0, rot)
    run(8192, 4096, 1024, 1, 1, 0, 0, rot)
    run(8192, 4096, 1024, 1, 1, 2, 0, rot)
    run(8192, 1024, 1024, 0, 0, 0, 0, rot)
    run(8192, 1024, 1024, 0, 1, 0, 1, rot)
    run(8192, 1024, 4096, 0, 1, 0, 1, rot)
