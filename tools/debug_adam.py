import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, restate
from tell_b200.optim import BertAdam
import test_optim_gpu as T
ref = T._make(); state = [dict() for _ in ref]
params = [torch.nn.Parameter(p.clone().cuda()) for p in T._make()]
opt = BertAdam(params, **T.HYPER)
for step in range(14):
    gs = T._grads(step)
    restate.bert_adam_step(ref, gs, state, **T.HYPER)
    for p, g in zip(params, gs):
        p.grad = g.cuda()
    opt.step()
    st = opt._groups[0]
    off = 0
    row = []
    for p, r, s in zip(params, ref, state):
        n = r.numel()
        m = st.m[off:off+n].cpu().view_as(r); v = st.v[off:off+n].cpu().view_as(r)
        row.append('%.1e/%.1e/%.1e' % ((p.detach().cpu()-r).abs().max().item(), ((m-s['next_m']).abs().max()/s['next_m'].abs().max()).item(), ((v-s['next_v']).abs().max()/s['next_v'].abs().max()).item()))
        off += n
    print(step, 'lr %.3e' % opt.get_lr()[0], 'clip', [round(x, 6) for x in st.scratch[2:].cpu().tolist()], row)
