"""Device time of each frozen encoder alone (CUDA graph replay, batch 16): ResNet-152 in both
BatchNorm modes, RoBERTa-large on the packed NYTimes-shaped batch.
    python tools/encoder_times.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import _lib, config  # noqa: E402

dev = torch.device('cuda', 0)
config.set_precision('bf16')
model = bench.build_model(dev)
host = bench.make_batch(16)
b = {k: v.to(dev) for k, v in host.items()}
n_real = int((host['article'] != 1).sum())


def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    fn()
    torch.cuda.synchronize()
    launches = _lib.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3), launches


out = {}
for mode in ('running', 'batch'):
    model.resnet.bn_mode = mode
    out['resnet152_' + mode] = graph_time(lambda: model.resnet.features_nhwc(b['image']))
out['roberta_large_packed'] = graph_time(lambda: model.roberta.all_hiddens(b['article'], n_real))
print(json.dumps({k: {'ms': v[0], 'launches': v[1]} for k, v in out.items()}))
