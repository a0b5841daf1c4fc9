"""A few eager decode steps between cudaProfilerStart/Stop for an ncu launch list:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/decode_steps.py"""
import os
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
model.decode_graph = False
host = bench.make_batch(B)
with torch.no_grad():
    b = {k: v.to(dev) for k, v in host.items()}
    cap = {'roberta': b['article'].new_zeros(B, 2)}
    cap_ids, _, contexts = model._forward({'roberta': b['article']}, b['image'], cap, b['faces'], b['objs'])
    model.gen_len = 4
    model._generate(cap_ids, contexts, early_exit=False)
    torch.cuda.synchronize()
    # profile steps 2..3 of a fresh 4-step decode: patch forward_tbc to toggle the profiler
    dec = model.decoder
    orig = dec.forward_tbc
    n = [0]

    def hooked(*a, **k):
        if n[0] == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        n[0] += 1
        return orig(*a, **k)
    dec.forward_tbc = hooked
    model._generate(cap_ids, contexts, early_exit=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('done')
