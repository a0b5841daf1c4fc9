"""Collect every GEMM shape of one train step and time each unique shape in isolation (CUDA graph
replay of 20 launches, so CPU launch overhead is excluded)."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import config, ops  # noqa: E402
from tell_b200.parallel import FlatGradients  # noqa: E402

dev = torch.device('cuda', 0)
config.set_precision('bf16')
model = bench.build_model(dev)
fg = FlatGradients(model.parameters())
host = bench.make_batch(16)
b = {k: v.to(dev) for k, v in host.items()}
shapes = collections.Counter()
orig = ops.gemm_tn


def hook(a, bb, out=None, out16=None, **kw):
    key = (a.shape[0], bb.shape[0], a.shape[1], out16 is not None or kw.get('want16', False),
           kw.get('m_limit') is not None)
    shapes[key] += 1
    return orig(a, bb, out=out, out16=out16, **kw)


ops.gemm_tn = hook
import tell_b200.functional as Fn  # noqa: E402
out = model(context={'roberta': b['article']}, image=b['image'], caption={'roberta': b['caption']},
            face_embeds=b['faces'], obj_embeds=b['objs'], metadata=None)
out['loss'].backward()
torch.cuda.synchronize()
ops.gemm_tn = orig
del model, fg
torch.cuda.empty_cache()

rows = []
for (M, N, K, o16, lim), cnt in shapes.items():
    Kp = (K + 7) // 8 * 8
    a = torch.randn(M, Kp, device=dev).bfloat16()[:, :K]
    w = torch.randn(N, Kp, device=dev).bfloat16()[:, :K]
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if o16 else torch.float32)
    kw = dict(out16=o, want32=False) if o16 else dict(out=o)
    for _ in range(3):
        ops.gemm_tn(a, w, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.gemm_tn(a, w, **kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 100 * 1e3
    # cuBLAS on the same shape (target, not product)
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ac, wt = a.contiguous(), w.contiguous().t()
    for _ in range(3):
        torch.matmul(ac, wt, out=ob)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(20):
            torch.matmul(ac, wt, out=ob)
    torch.cuda.synchronize()
    s.record()
    for _ in range(5):
        g2.replay()
    e.record()
    torch.cuda.synchronize()
    us_cb = s.elapsed_time(e) / 100 * 1e3
    rows.append((us * cnt, cnt, M, N, K, o16, lim, us, 2.0 * M * N * K / us / 1e6, us_cb))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print('unique shapes %d, launches %d, sum of isolated times %.2f ms' % (len(rows), sum(r[1] for r in rows), tot / 1e3))
for r in rows[:60]:
    print('tot %8.1f us  x%3d  M=%6d N=%6d K=%6d o16=%d lim=%d  %7.1f us  %7.1f TF   cublas %7.1f us' % r)
