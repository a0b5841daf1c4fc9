import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import config, synth  # noqa: E402
from tell_b200.models import DynamicConvFacesObjectsDecoder  # noqa: E402
from tell_b200.testing import build_decoder  # noqa: E402

t0 = time.time()
cfg = synth.CFG_TINY
config.set_precision(sys.argv[1] if len(sys.argv) > 1 else 'bf16x3')
sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=4.0)
dec = build_decoder(cfg, DynamicConvFacesObjectsDecoder, sd).cuda().eval()
cap, ctx = synth.decoder_inputs(cfg, 3, 9, 11, 3, 4, 5, seed=1234)
inp, tgt = cap[:, :-1].contiguous().cuda(), cap[:, 1:].contiguous().cuda()
cctx = {k: v.cuda() for k, v in ctx.items()}
torch.cuda.synchronize()
print('setup %.2fs' % (time.time() - t0))


def step():
    out, _ = dec({'roberta': inp}, cctx)
    loss, _ = dec.adaptive_softmax.fused_loss(out, tgt)
    loss.backward()
    torch.cuda.synchronize()


for i in range(3):
    t = time.time()
    step()
    print('step %d %.3fs' % (i, time.time() - t))
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
