"""Decoder-sized GEMMs (M = 800 rows, fp32 output, the signatures of one train step) timed from a
CUDA graph.  TT_B200_LIB=<path> times another build of the library (A/B of two versions).
    python tools/gemm_small.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402

# (M, N, K, trans_a, trans_b, bias)
SIGS = [(800, 1024, 1024, 0, 0, 1), (800, 2048, 1024, 0, 0, 1), (800, 4096, 1024, 0, 0, 1),
        (800, 1024, 4096, 0, 0, 1), (800, 1024, 1024, 0, 1, 0), (1024, 1024, 800, 1, 1, 0),
        (800, 4096, 1024, 0, 1, 0), (1024, 4096, 800, 1, 1, 0), (800, 48, 1024, 0, 0, 0),
        (800, 496, 1024, 0, 0, 0), (800, 5002, 1024, 0, 0, 0), (256, 1024, 1024, 0, 0, 1)]
for M, N, K, ta, tb, bias in SIGS:
    def mk(r, c):
        return [torch.randn(r, (c + 7) // 8 * 8, device='cuda').bfloat16()[:, :c] for _ in range(2)]
    A = mk(K, M) if ta else mk(M, K)
    B = mk(K, N) if tb else mk(N, K)
    out = torch.zeros(M, N, device='cuda')
    b = torch.zeros(N, device='cuda') if bias else None

    def one(i):
        ops.gemm_tn(A[i % 2], B[i % 2], out=out, bias=b, trans_a=bool(ta), trans_b=bool(tb))
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            one(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 80 * 1e3)
    print('M=%5d N=%5d K=%5d ta=%d tb=%d  %6.2f us  %6.1f TF' % (M, N, K, ta, tb, best, 2.0 * M * N * K / best / 1e6),
          flush=True)
