"""Per-kernel device time of the captured greedy-decode step (batch 256), from the torch profiler
(CUPTI) over one generate() call whose steps are graph replays.
    python tools/decode_kernel_times.py > profiles/r2_decode_kernel_times.txt"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import config  # noqa: E402

B = int(os.environ.get('TT_B', '256'))
dev = torch.device('cuda', 0)
config.set_precision('bf16')
model = bench.build_model(dev).eval()
model.gen_len = 50
host = bench.make_batch(B)
with torch.no_grad():
    b = {k: v.to(dev) for k, v in host.items()}
    cap = {'roberta': b['article'].new_zeros(B, 2)}
    cap_ids, _, contexts = model._forward({'roberta': b['article']}, b['image'], cap, b['faces'], b['objs'])
    for _ in range(2):
        model._generate(cap_ids, contexts, early_exit=False)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        model._generate(cap_ids, contexts, early_exit=False)       # graph reused: step 0 eager + 49 replays
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    nm = re.sub(r'\(.*', '', e.name)[:70]
    agg[nm][0] += e.device_time
    agg[nm][1] += 1
tot = sum(v[0] for v in agg.values())
print('one generate() call, batch %d, 50 steps (step 0 eager incl. the K|V projections and cache build, 49 graph '
      'replays): sum of kernel durations %.1f ms over %d launches' % (B, tot / 1e3, sum(v[1] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print('%-72s %9.1f us %5.1f%% %5d  avg %7.1f' % (k, v[0], 100 * v[0] / tot, v[1], v[0] / v[1]))
