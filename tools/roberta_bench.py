"""RoBERTa-large forward (B=16, S=512) under CUDA-graph replay: sustained time per pass + clocks."""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200.models import RobertaEncoder  # noqa: E402

dev = torch.device('cuda', 0)
with torch.device(dev):
    enc = RobertaEncoder()
enc = enc.to(dev).eval()
ids = bench.make_batch(16)['article'].to(dev)
for _ in range(2):
    enc.all_hiddens(ids)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    enc.all_hiddens(ids)
torch.cuda.synchronize()
flop = 24 * 2 * 8192 * 1024 * (3072 + 1024 + 4096 + 4096) + 24 * 4 * 256 * 512 * 512 * 64
samp = bench.ClockSampler(0)
samp.start()
for reps in (1, 5, 50, 200):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print('reps %3d: %.3f ms per forward, %.1f TF' % (reps, ms, flop / ms / 1e9))
samp.stop_flag = True
samp.join(timeout=2)
print(samp.summary())
print(sorted(set(r[0] for r in samp.rows)))
