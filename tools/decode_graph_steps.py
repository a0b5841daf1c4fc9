"""Graphed greedy decode (B = 256) under ncu: steps 0-1 eager, the captured step replayed 4 times.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... \
       python tools/decode_graph_steps.py
The last replay = the last `launches_per_replay` rows of the launch list (printed at the end)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import _lib, config  # noqa: E402

B = int(os.environ.get('TT_B', '256'))
dev = torch.device('cuda', 0)
config.set_precision('bf16')
model = bench.build_model(dev).eval()
model.decode_graph = True
host = bench.make_batch(B)
with torch.no_grad():
    b = {k: v.to(dev) for k, v in host.items()}
    cap = {'roberta': b['article'].new_zeros(B, 2)}
    cap_ids, _, contexts = model._forward({'roberta': b['article']}, b['image'], cap, b['faces'], b['objs'])
    model.gen_len = 6
    model._generate(cap_ids, contexts, early_exit=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model._generate(cap_ids, contexts, early_exit=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('done')
