"""The three GEMMs of a ResNet-152 stage-3 bottleneck (batch 16: M = 16*14*14 = 3136) with their real
epilogues (folded-BN bias, ReLU, bf16 residual), timed from a CUDA graph; the frozen image encoder
launches each of them 36 times per step (tell_b200/models/resnet.py features_nhwc).
   python tools/resnet_gemm.py            # table
   python tools/resnet_gemm.py one N,K,r  # a few eager launches of one signature (for ncu)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402

M = 3136
SIGS = [(256, 1024, 0), (256, 2304, 0), (1024, 256, 1)]      # (N, K, residual)


def make(N, K, res):
    a = [torch.randn(M, K, device='cuda').bfloat16() for _ in range(2)]
    w = torch.randn(N, K, device='cuda').bfloat16()
    b = torch.randn(N, device='cuda')
    r = torch.randn(M, N, device='cuda').bfloat16() if res else None
    o = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)

    def run(i):
        ops.gemm_tn(a[i % 2], w, bias=b, residual16=r, act=ops.ACT_RELU, out16=o, want32=False, want16=True)
    return run


if len(sys.argv) > 1 and sys.argv[1] == 'one':
    N, K, res = [int(x) for x in sys.argv[2].split(',')]
    run = make(N, K, res)
    for i in range(6):
        run(i)
    torch.cuda.synchronize()
    sys.exit(0)

for N, K, res in SIGS:
    run = make(N, K, res)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            run(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / 80
    byts = 2.0 * (M * K + N * K + M * N * (2 if res else 1))
    print('M=%d N=%-5d K=%-5d res=%d  %6.2f us  %7.1f TFLOP/s  %6.2f TB/s operand+output bytes'
          % (M, N, K, res, us, 2.0 * M * N * K / us / 1e6, byts / us / 1e6), flush=True)
