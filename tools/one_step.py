"""One eager train step of the bench workload between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import config  # noqa: E402
from tell_b200.parallel import FlatGradients  # noqa: E402

dev = torch.device('cuda', 0)
config.set_precision(os.environ.get('TT_PRECISION', 'bf16'))
config.manual_seed(1234)
config.enable_zero_arena(dev)
config.enable_wgrad_stream(int(os.environ.get('TT_WGRAD', '0')))
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
    return out['loss']


for _ in range(int(os.environ.get('TT_WARM', '2'))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()      # ncu --profile-from-start off (covers the autograd thread too)
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('loss', float(loss))
