"""Per-kernel device time of the two step graphs (frozen encoders / decoder fwd+loss+bwd), each REPLAYED
alone under the torch profiler (CUPTI): warm-L2, back-to-back timings as they occur in the captured
step -- unlike the ncu launch list, whose per-kernel times are cold-cache and serialised.
    python tools/graph_kernel_times.py > profiles/r2_graph_kernel_times.txt"""
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

dev = torch.device('cuda', 0)
config.set_precision('bf16')
config.manual_seed(1234)
config.enable_device_step(dev)
config.enable_zero_arena(dev)
config.encoder_sm_cap = int(os.environ.get('TT_ENC_SMS', '88'))
config.set_gemm_occupancy_weight(float(os.environ.get('TT_OCC_W', '1.0')))
model = bench.build_model(dev, 'batch')
params = [p for p in model.parameters() if p.requires_grad]
host = bench.make_batch(16)
st = {k: v.to(dev) for k, v in host.items()}
pristine = {k: v.clone() for k, v in st.items()}
n_real = int((host['article'] != 1).sum())
enc = [None]


def encode():
    enc[0] = model.encode({'roberta': st['article']}, st['image'], n_real_tokens=n_real)


def train():
    for p in params:
        p.grad = None
    config.advance_device_step()
    out = model(context={'roberta': st['article']}, image=st['image'], caption={'roberta': st['caption']},
                face_embeds=st['faces'], obj_embeds=st['objs'], metadata=None, encoded=enc[0])
    out['loss'].backward()


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        for k in ('faces', 'objs'):
            st[k].copy_(pristine[k])
        encode()
        train()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
with torch.cuda.graph(g1):
    encode()
with torch.cuda.graph(g2, pool=g1.pool()):
    train()
torch.cuda.synchronize()
REPS = 5
for name, g in (('frozen encoders (ResNet-152 || RoBERTa-large)', g1), ('decoder forward + loss + backward', g2)):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(REPS):
            g.replay()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    t0, t1 = 1e30, 0.0
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        nm = re.sub(r'\(.*', '', e.name)[:70]
        agg[nm][0] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
        agg[nm][1] += 1
    tot = sum(v[0] for v in agg.values()) / REPS
    print('=== %s: sum of kernel durations %.1f us per replay over %d launches' %
          (name, tot, sum(v[1] for v in agg.values()) // REPS))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print('%-72s %9.1f us %5.1f%% %5d  avg %7.1f' % (k, v[0] / REPS, 100 * v[0] / REPS / tot, v[1] // REPS,
                                                       v[0] / v[1]))
