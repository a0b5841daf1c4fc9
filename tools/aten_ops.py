"""Which torch (ATen) kernels still run inside one train step, with shapes:
   python tools/aten_ops.py"""
import os
import sys
from collections import Counter

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import config  # noqa: E402

dev = torch.device('cuda', 0)
config.set_precision('bf16')
config.manual_seed(1234)
config.enable_zero_arena(dev)
_orig_seed = config.next_seed
config.next_seed = lambda: _orig_seed() & (2 ** 62 - 1)   # the profiler records int args as int64
model = bench.build_model(dev)
params = [p for p in model.parameters() if p.requires_grad]
host = bench.make_batch(16)
pristine = {k: v.to(dev) for k, v in host.items()}


def step():
    b = {k: v.clone() for k, v in pristine.items()}
    for p in params:
        p.grad = None
    out = model(context={'roberta': b['article']}, image=b['image'], caption={'roberta': b['caption']},
                face_embeds=b['faces'], obj_embeds=b['objs'], metadata=None)
    out['loss'].backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
c = Counter()
t = Counter()
for e in prof.events():
    if e.name.startswith('aten::') and e.device_time > 0 and e.name not in ('aten::empty', 'aten::view'):
        key = (e.name, str(e.input_shapes)[:90])
        c[key] += 1
        t[key] += e.device_time
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:40]:
    print('%8.1f us %4d  %s %s' % (v, c[k], k[0], k[1]))
