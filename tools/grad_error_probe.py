"""Per-parameter gradient error of the CUDA decoder (bf16x3 and bf16) against the CPU oracle's
fp32 autograd at full width on a small batch -- finds which op family limits backward parity.
    python tools/grad_error_probe.py [B T S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'transform-and-tell_b200')]
import restate  # noqa: E402
from tell_b200 import config, synth  # noqa: E402
from tell_b200.models import DynamicConvFacesObjectsDecoder  # noqa: E402
from tell_b200.testing import build_decoder  # noqa: E402

B, T, S = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (2, 16, 64)
cfg = synth.CFG_FULL
sd = synth.decoder_state_dict(cfg, seed=1, logit_gain=2.0)
cap, ctx = synth.decoder_inputs(cfg, B=B, T=T, S=S, F=4, O=8, P=9, seed=77)
inp, tgt = cap[:, :-1].contiguous(), cap[:, 1:].contiguous()
ocfg = synth.oracle_cfg(cfg)
# oracle gradients (fp64 for a clean reference)
clones = {}
sd_r = {}
for k, v in sd.items():
    if v.is_floating_point() and 'version' not in k and '_float_tensor' not in k and 'position' not in k:
        if id(v) not in clones:
            clones[id(v)] = v.clone().requires_grad_(True)
        sd_r[k] = clones[id(v)]
    else:
        sd_r[k] = v
ctx_r = {k: (v.clone().requires_grad_(k == 'article') if v.is_floating_point() else v) for k, v in ctx.items()}
ro, _ = restate.decoder_forward(inp, ctx_r, sd_r, ocfg)
_, n, rl = restate.adaptive_loss(ro, tgt, sd_r, ocfg['cutoffs'])
rl.backward()
for prec in ('bf16x3', 'bf16'):
    config.set_precision(prec)
    dec = build_decoder(cfg, DynamicConvFacesObjectsDecoder, sd).cuda().eval()
    cctx = {k: v.cuda() for k, v in ctx.items()}
    cctx['article'].requires_grad_(True)
    out, _ = dec({'roberta': inp.cuda()}, cctx)
    loss, ntok = dec.adaptive_softmax.fused_loss(out, tgt.cuda())
    loss.backward()
    rows = []
    for name, p in dec.named_parameters():
        ref = sd_r[name].grad
        if ref is None or p.grad is None:
            continue
        e = (p.grad.double().cpu() - ref).abs().max().item() / max(1e-30, ref.abs().max().item())
        rows.append((e, name))
    e = (cctx['article'].grad.double().cpu() - ctx_r['article'].grad).abs().max().item() / ctx_r['article'].grad.abs().max().item()
    rows.append((e, 'd_article'))
    rows.sort(reverse=True)
    print('==', prec, 'out err', (out.detach().double().cpu() - ro.detach()).abs().max().item(),
          'loss', loss.item(), rl.item())
    for e, name in rows[:25]:
        print('  %.3e  %s' % (e, name))
    print('  ... median %.3e' % rows[len(rows) // 2][0])
    del dec
