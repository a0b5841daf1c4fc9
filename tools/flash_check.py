"""Max error of the tcgen05 flash kernel against fp32 attention per sample + which rows differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import ops
torch.manual_seed(3)
H, D = 16, 64
E = H * D
lens = [int(a) for a in sys.argv[1:]] or [512, 300, 129, 7]
B, S = len(lens), 512
n = sum(lens)
qkv = torch.randn(B * S, 3 * E) * 0.7
qkv[:, :E] *= D ** -0.5
q16 = qkv.to(torch.bfloat16)
cu = torch.cat([torch.zeros(1, dtype=torch.long), torch.tensor(lens).cumsum(0)]).int()
x = q16.float()
want = torch.zeros(n, E)
for b in range(B):
    r0, r1 = int(cu[b]), int(cu[b + 1])
    q = x[r0:r1, :E].view(-1, H, D).transpose(0, 1)
    k = x[r0:r1, E:2 * E].view(-1, H, D).transpose(0, 1)
    v = x[r0:r1, 2 * E:].view(-1, H, D).transpose(0, 1)
    p = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    want[r0:r1] = (p @ v).transpose(0, 1).reshape(-1, E)
got = ops.flash_self_attn_varlen(q16.cuda(), cu.cuda(), B, S, H, D, tc5=True)[:n].float().cpu()
err = (got - want).abs()
print('max err %.4g of %.4g' % (err.max().item(), want.abs().max().item()))
for b in range(B):
    r0, r1 = int(cu[b]), int(cu[b + 1])
    e = err[r0:r1].view(r1 - r0, H, D)
    rows = (e.amax(dim=(1, 2)) > 2e-2).nonzero().flatten()
    print('sample %d len %d: bad rows %d' % (b, r1 - r0, rows.numel()), rows[:10].tolist(), rows[-5:].tolist(),
          'per-64-dim-half err', e[:, :, :32].max().item(), e[:, :, 32:].max().item())
