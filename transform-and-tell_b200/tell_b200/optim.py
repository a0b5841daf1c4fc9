"""BertAdam on the B200 path.

The reference trains with `trainer.optimizer.type: bert_adam`
(expt/nytimes/9_transformer_objects/config.yaml:126-149), i.e. pytorch-pretrained-bert's BertAdam
reached through allennlp 0.9 and stepped at tell/training/callback_apex_trainer.py:238; under apex
O2 the model runs in fp16 with fp32 master weights.  Here the parameters ARE the fp32 master copy
(the bf16 tensor-core operands are rebuilt from them by the weight bank's one `tt_weight_prep`
launch per step), and the whole step of a parameter group is `tt_bertadam_step`: three launches
for all tensors instead of ~10 torch kernels per tensor.

Same constructor keywords, `step()` / `zero_grad()` / `state_dict()` surface and parameter-group
semantics (`parameter_groups` regexes of the config become ordinary torch-style groups).
Differences, all deliberate and documented:
  * the step counter lives on the device (one per group, shared by its tensors -- in the reference
    every tensor's state['step'] holds the same value), so a step captured in a CUDA graph keeps
    advancing the warm-up schedule on replay;
  * `step(loss=...)` takes the device loss: a NaN loss skips the update on the device, the
    NaN-batch skip of callback_apex_trainer.py:225-227 without a host sync;
  * the per-tensor clip scales the gradient inside the update instead of rewriting p.grad.
"""
import ctypes
import re

import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_void_p

SCHEDULES = {'none': 0, None: 0, 'warmup_linear': 1, 'warmup_constant': 2}


class TtAdamSeg(ctypes.Structure):
    _fields_ = [('p', c_void_p), ('g', c_void_p), ('m', c_void_p), ('v', c_void_p), ('n', c_ll),
                ('chunk0', c_int), ('pad_', c_int)]


class TtAdamHyper(ctypes.Structure):
    _fields_ = [('lr', c_float), ('b1', c_float), ('b2', c_float), ('e', c_float),
                ('weight_decay', c_float), ('max_grad_norm', c_float), ('warmup', c_float),
                ('schedule', c_int), ('t_total', c_ll)]


class _Group:
    """Device state of one parameter group."""

    def __init__(self, params, device):
        self.params = params
        n = sum(p.numel() for p in params)
        self.m = torch.zeros(n, dtype=torch.float32, device=device)      # next_m, all tensors
        self.v = torch.zeros(n, dtype=torch.float32, device=device)      # next_v
        self.step = torch.zeros(1, dtype=torch.int64, device=device)
        chunk = int(_lib.lib().tt_bertadam_chunk())
        self.chunk0, c = [], 0
        for p in params:
            self.chunk0.append(c)
            c += (p.numel() + chunk - 1) // chunk
        self.total_chunks = c
        self.partial = torch.empty(max(c, 1), dtype=torch.float32, device=device)
        self.scratch = torch.zeros(2 + len(params), dtype=torch.float32, device=device)
        self.table = None
        self.table_key = None

    def segments(self):
        """Device segment table; re-uploaded only when a parameter or gradient pointer changed
        (gradients are freshly allocated tensors unless they live in a flat buffer / a graph)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in self.params)
        if key != self.table_key:
            segs, off = [], 0
            for p, c0 in zip(self.params, self.chunk0):
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise _lib.TtError('BertAdam needs contiguous fp32 parameters and gradients')
                n = p.numel()
                segs.append(TtAdamSeg(p.data_ptr(), g.data_ptr(), self.m.data_ptr() + 4 * off,
                                      self.v.data_ptr() + 4 * off, n, c0, 0))
                off += n
            raw = bytes((TtAdamSeg * len(segs))(*segs))
            host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            if torch.cuda.is_current_stream_capturing():
                raise _lib.TtError('BertAdam: gradient pointers changed inside a CUDA-graph capture; '
                                   'run one eager step with the same gradient buffers first')
            self.table = host.to(self.m.device)
            self.table_key = key
        return self.table


class BertAdam:
    """BertAdam(params, lr, warmup=-1, t_total=-1, schedule='warmup_linear', b1=0.9, b2=0.999,
    e=1e-6, weight_decay=0.01, max_grad_norm=1.0) -- pytorch-pretrained-bert's signature."""

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule='warmup_linear', b1=0.9,
                 b2=0.999, e=1e-6, weight_decay=0.01, max_grad_norm=1.0):
        if lr < 0.0:
            raise ValueError('Invalid learning rate: {} - should be >= 0.0'.format(lr))
        if schedule not in SCHEDULES:
            raise ValueError('Invalid schedule parameter: {}'.format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError('Invalid warmup: {} - should be in [0.0, 1.0[ or -1'.format(warmup))
        if not 0.0 <= b1 < 1.0:
            raise ValueError('Invalid b1 parameter: {} - should be in [0.0, 1.0['.format(b1))
        if not 0.0 <= b2 < 1.0:
            raise ValueError('Invalid b2 parameter: {} - should be in [0.0, 1.0['.format(b2))
        if not e >= 0.0:
            raise ValueError('Invalid epsilon value: {} - should be >= 0.0'.format(e))
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{'params': params}]
        self.param_groups, self._groups = [], []
        seen = set()
        for g in params:
            group = dict(defaults)
            group.update({k: v for k, v in g.items() if k != 'params'})
            ps = []
            for p in g['params']:
                if p.requires_grad and id(p) not in seen:        # tied weights step once
                    seen.add(id(p))
                    ps.append(p)
            group['params'] = ps
            self.param_groups.append(group)
            if ps:
                if not ps[0].is_cuda:
                    raise _lib.TtError('BertAdam (B200 path) needs CUDA parameters; no CPU fallback')
                self._groups.append(_Group(ps, ps[0].device))
            else:
                self._groups.append(None)

    @classmethod
    def from_config(cls, named_parameters, lr, parameter_groups=None, **kw):
        """The `optimizer:` block of a reference config: `parameter_groups` is allennlp's list of
        [[regex, ...], {overrides}] pairs (config.yaml:137-149); parameters matching no regex form
        the last group, as allennlp.training.optimizers.Optimizer.from_params does."""
        named = [(n, p) for n, p in named_parameters if p.requires_grad]
        if not parameter_groups:
            return cls([p for _, p in named], lr=lr, **kw)
        groups = [{'params': []} for _ in range(len(parameter_groups) + 1)]
        for k, (_, overrides) in enumerate(parameter_groups):
            groups[k].update(overrides)
        for name, p in named:
            for k, (regexes, _) in enumerate(parameter_groups):
                if any(re.search(r, name) for r in regexes):
                    groups[k]['params'].append(p)
                    break
            else:
                groups[-1]['params'].append(p)
        return cls(groups, lr=lr, **kw)

    def zero_grad(self, set_to_none=True):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    def get_lr(self):
        """lr each group's LAST step used (device value; this call synchronises)."""
        return [float(g.scratch[0].item()) if g is not None else 0.0 for g in self._groups]

    def step(self, closure=None, loss=None):
        """One update of every group.  `loss`: optional device scalar -- NaN skips the step."""
        out = closure() if closure is not None else None
        stream = c_void_p(torch.cuda.current_stream().cuda_stream)
        for group, st in zip(self.param_groups, self._groups):
            if st is None:
                continue
            if any(p.grad is None for p in st.params):
                raise _lib.TtError('BertAdam.step: a parameter has no gradient (the fused step '
                                   'updates whole groups; freeze it with requires_grad=False)')
            table = st.segments()
            h = TtAdamHyper(group['lr'], group['b1'], group['b2'], group['e'], group['weight_decay'],
                            group['max_grad_norm'], max(group['warmup'], 0.0),   # 0.6.2: -1 -> 0
                            SCHEDULES[group['schedule']], int(group['t_total']))
            _lib.call('tt_bertadam_step', c_void_p(table.data_ptr()), c_int(len(st.params)),
                      c_int(st.total_chunks), ctypes.byref(h), c_void_p(st.step.data_ptr()),
                      c_void_p(loss.data_ptr()) if loss is not None else c_void_p(0),
                      c_void_p(st.partial.data_ptr()), c_void_p(st.scratch.data_ptr()), stream)
        return out

    # -- checkpointing (torch.optim-like, enough for resume)
    def state_dict(self):
        return {'groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups],
                'state': [None if s is None else {'next_m': s.m.clone(), 'next_v': s.v.clone(),
                                                  'step': s.step.clone()} for s in self._groups]}

    def load_state_dict(self, sd):
        for g, src in zip(self.param_groups, sd['groups']):
            g.update(src)
        for s, src in zip(self._groups, sd['state']):
            if s is not None and src is not None:
                s.m.copy_(src['next_m'])
                s.v.copy_(src['next_v'])
                s.step.copy_(src['step'])
