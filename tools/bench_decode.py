"""Greedy-decode latency, BASELINE.json configs[3] / SURVEY.md 8d cfg 4: batch 256, 512-token article,
4 faces, 16 objects, seed [B,2] zeros, EXACTLY 50 steps (EOS early exit disabled for timing).

    python tools/bench_decode.py [--batch 256] [--steps 50] [--reps 3]

Prints one JSON line: one-off context time (ResNet + RoBERTa + K|V projections happen inside the
first decoder step) and the 50-step decode time measured with CUDA events."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
import bench  # noqa: E402
from tell_b200 import config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--steps', type=int, default=50)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--eager', action='store_true', help='disable the captured decode step')
args = ap.parse_args()
dev = torch.device('cuda', 0)
config.set_precision('bf16')
config.manual_seed(1234)
model = bench.build_model(dev).eval()
model.gen_len = args.steps
model.decode_graph = not args.eager
phases = {}
if os.environ.get('TT_DECODE_PHASES') == '1':
    object.__setattr__(model, 'decode_timing', phases)
B = args.batch
host = bench.make_batch(B)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
res = []
with torch.no_grad():
    for rep in range(args.reps + 2):
        b = {k: v.to(dev) for k, v in host.items()}
        cap = {'roberta': b['article'].new_zeros(B, 2)}
        e0, e1, e2 = ev(), ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        cap_ids, _, contexts = model._forward({'roberta': b['article']}, b['image'], cap, b['faces'], b['objs'])
        e1.record()
        lp, ids, _ = model._generate(cap_ids, contexts, early_exit=False)
        e2.record()
        torch.cuda.synchronize()
        if rep > 1:          # two warm-up repetitions (lazy paths, allocator pools, clocks)
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        assert ids.shape == (B, 1 + args.steps), ids.shape
ctx_ms = sorted(r[0] for r in res)[len(res) // 2]
dec_ms = sorted(r[1] for r in res)[len(res) // 2]
print(json.dumps({'metric': 'greedy decode latency', 'batch': B, 'steps': args.steps,
                  'encoders_ms': round(ctx_ms, 2), 'decode_ms': round(dec_ms, 2),
                  'ms_per_step': round(dec_ms / args.steps, 3),
                  'captions_per_s_decode_only': round(B / (dec_ms * 1e-3), 1),
                  'captions_per_s_incl_encoders': round(B / ((dec_ms + ctx_ms) * 1e-3), 1),
                  'decode_ms_min': round(min(r[1] for r in res), 2),
                  'phases': {k: round(v, 2) for k, v in phases.items()}, 'decode_graph': bool(model.decode_graph), 'decode_ms_reps': [round(r[1], 1) for r in res],
                  'note': 'steps 0-1 eager (one-off K|V projections of the four contexts, graph capture), '
                          'steps 2.. replay one captured decode step'}))
