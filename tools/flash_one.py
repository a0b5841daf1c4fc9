"""Run the packed RoBERTa self-attention kernels a few times (for ncu / timing):
   python tools/flash_one.py [tc5=1]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops  # noqa: E402

tc5 = (sys.argv[1] if len(sys.argv) > 1 else '1') == '1'
rs = np.random.RandomState(1234)
B, S, H, D = 16, 512, 16, 64
lens = rs.randint(256, 513, size=B)
if os.environ.get('TT_LEN'):
    lens = np.full(B, int(os.environ['TT_LEN']))
cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device='cuda')
qkv = (torch.randn(B * S, 3 * H * D, device='cuda') * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.flash_self_attn_varlen(qkv, cu, B, S, H, D, tc5=tc5)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        ops.flash_self_attn_varlen(qkv, cu, B, S, H, D, tc5=tc5)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    g.replay()
e.record()
torch.cuda.synchronize()
print('tc5=%d  tokens %d  %.1f us per launch' % (tc5, int(lens.sum()), s.elapsed_time(e) * 1e3 / 50))
