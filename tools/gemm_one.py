"""Run one GEMM shape a few times (for ncu --set full):  python tools/gemm_one.py M N K [bf16out]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
o16 = len(sys.argv) > 4 and sys.argv[4] == '1'
a = torch.randn(M, K, device='cuda').bfloat16()
w = torch.randn(N, K, device='cuda').bfloat16()
o = torch.empty(M, N, device='cuda', dtype=torch.bfloat16 if o16 else torch.float32)
kw = dict(out16=o, want32=False) if o16 else dict(out=o)
for _ in range(5):
    ops.gemm_tn(a, w, **kw)
torch.cuda.synchronize()
