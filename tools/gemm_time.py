"""Time GEMM shapes from a CUDA graph (device time per launch):
   python tools/gemm_time.py M,N,K[,rows] ...     rows = device-side row limit (m_limit)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402

for spec in sys.argv[1:]:
    v = [int(x) for x in spec.split(',')]
    M, N, K = v[:3]
    rows = v[3] if len(v) > 3 else M
    a = [torch.randn(M, K, device='cuda').bfloat16() for _ in range(2)]
    w = [torch.randn(N, K, device='cuda').bfloat16() for _ in range(2)]
    o = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    lim = torch.tensor([rows], dtype=torch.int32, device='cuda') if rows != M else None
    for i in range(3):
        ops.gemm_tn(a[i % 2], w[i % 2], out16=o, want32=False, m_limit=lim)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            ops.gemm_tn(a[i % 2], w[i % 2], out16=o, want32=False, m_limit=lim)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / 80
    print('%-28s BN=%s  %7.1f us  %7.1f TFLOP/s' % (spec, os.environ.get('TT_GEMM2_BN', 'auto'), us,
                                                    2.0 * rows * N * K / us / 1e6), flush=True)
