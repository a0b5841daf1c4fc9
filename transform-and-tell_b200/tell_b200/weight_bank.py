"""Weight bank: every bf16 GEMM weight operand of a decoder prepared by ONE launch per step.

The reference recomputes w = g * v / ||v|| inside every GehringLinear.forward (tell/modules/
linear.py:30-34) and feeds fp32 weights to addmm.  On the B200 path every trainable matrix is a
bf16 tcgen05 operand that has to be rebuilt after each optimizer step; done per module that is
~250 tiny HBM-bound launches on the decoder's critical path.  The bank does it with
`tt_weight_prep` (one launch, device-resident segment table) and runs all weight-norm backwards
with `tt_wnorm_bwd_multi` (one launch at the end of the backward).

* weight-normalised linears are enumerated statically (GehringLinear modules);
* every other weight operand form (plain matrices, row-wise concatenations such as [Wk; Wv] or
  the fused Q projection, K-axis concatenations such as the embedding band projections) is
  RECORDED during the first forward that runs with the bank and served from the table afterwards.

Only the throughput mode ('bf16') uses the bank; the parity mode ('bf16x3') keeps per-call casts.
"""
import ctypes

import weakref

import torch
from torch.autograd import Function

from . import _lib, config, ops
from ._lib import c_int, c_ll, c_void_p


class TtPrepSeg(ctypes.Structure):
    _fields_ = [('src', c_void_p), ('g', c_void_p), ('w32', c_void_p), ('norm', c_void_p),
                ('dst16', c_void_p), ('ld_src', c_ll), ('ld_dst', c_ll), ('rows', c_int),
                ('cols', c_int), ('row0', c_int), ('pad_', c_int)]


class TtWnormBwdSeg(ctypes.Structure):
    _fields_ = [('dw', c_void_p), ('v', c_void_p), ('g', c_void_p), ('norm', c_void_p),
                ('dv', c_void_p), ('dg', c_void_p), ('rows', c_int), ('cols', c_int),
                ('row0', c_int), ('pad_', c_int)]


def _upload(structs, device):
    """ctypes struct array -> device byte tensor (the table the kernels read)."""
    arr = (type(structs[0]) * len(structs))(*structs)
    raw = bytes(arr)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def _pad8(n):
    return (n + 7) // 8 * 8


class _WN:
    """One weight-normalised matrix: parameters v [O,I], g [O,1] and the bank's buffers."""
    __slots__ = ('v', 'g', 'w32', 'norm', 'b16', 'dw', 'dv', 'dg')


LIVE = weakref.WeakSet()      # every bank alive in this process (ops.cast_bf16 guards their handles)


class WeightBank:
    def __init__(self, module):
        LIVE.add(self)
        from .modules.linear import GehringLinear
        self.module = module
        self.device = next(module.parameters()).device
        self.params = [(p, p.data_ptr()) for p in module.parameters()]
        self.storages = {p.untyped_storage().data_ptr() for p, _ in self.params}
        self.wn = []
        mods = [m for m in module.modules() if isinstance(m, GehringLinear) and m.weight_norm]
        n_el = sum(m.weight_v.numel() for m in mods)
        n_rows = sum(m.weight_v.shape[0] for m in mods)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self._w32 = torch.empty(n_el, **f32)
        self._dw = torch.zeros(n_el, **f32)
        self._dv = torch.empty(n_el, **f32)
        self._norm = torch.empty(n_rows, **f32)
        self._dg = torch.empty(n_rows, **f32)
        self._wn16 = torch.empty(n_el, dtype=torch.bfloat16, device=dev)
        off = roff = 0
        self.by_v = {}
        self.dw_dest = {}
        for m in mods:
            e = _WN()
            O, I = m.weight_v.shape
            assert I % 8 == 0, 'weight-normalised matrices need in_features % 8 == 0'
            e.v, e.g = m.weight_v, m.weight_g
            e.w32 = self._w32[off:off + O * I].view(O, I)
            e.dw = self._dw[off:off + O * I].view(O, I)
            e.dv = self._dv[off:off + O * I].view(O, I)
            e.b16 = self._wn16[off:off + O * I].view(O, I)
            e.norm = self._norm[roff:roff + O]
            e.dg = self._dg[roff:roff + O].view(O, 1)
            off += O * I
            roff += O
            self.wn.append(e)
            self.by_v[id(m.weight_v)] = e
            self.dw_dest[e.w32.data_ptr()] = e.dw
        self.map = {}          # lookup key -> bf16 operand served from the table
        self.log = {}          # recorded during the first banked forward: key -> (kind, tensors)
        self.extra = []        # finalised recorded entries: (kind, tensors, bf16 buffer)
        self.table = None
        self.bwd_table = None
        self.total_rows = 0
        self.current = None    # {id(weight_v): effective weight} of the running forward
        self.claimed = set()
        self.prepared_once = False
        self.write_w32 = False
        self._build_tables()

    # ------------------------------------------------------------------ tables
    def _build_tables(self):
        segs, row0 = [], 0

        def add(src, g, w32, norm, dst, rows, cols, ld_src, ld_dst):
            nonlocal row0
            segs.append(TtPrepSeg(src.data_ptr(), g.data_ptr() if g is not None else None,
                                  w32.data_ptr() if w32 is not None else None,
                                  norm.data_ptr() if norm is not None else None, dst.data_ptr(),
                                  ld_src, ld_dst, rows, cols, row0, 0))
            row0 += rows
        self.map = {}
        for e in self.wn:
            O, I = e.v.shape
            # the fp32 effective weight is only a HANDLE here (autograd node output, lookup key): every
            # consumer takes the bf16 operand from the table, so its 4 bytes per element are not
            # written (write_w32 = True restores them; functional.operand refuses to cast the handle)
            add(e.v, e.g, e.w32 if self.write_w32 else None, e.norm, e.b16, O, I, I, I)
            self.map[('s', e.w32.data_ptr(), (O, I))] = e.b16
        for kind, tensors, buf, key in self.extra:
            if kind == 's':
                x = tensors[0]
                add(x, None, None, None, buf, x.shape[0], x.shape[1], x.stride(0), buf.stride(0))
            elif kind == 'rows':
                r0 = 0
                for m in tensors:
                    add(m, None, None, None, buf[r0:], m.shape[0], m.shape[1], m.stride(0),
                        buf.stride(0))
                    r0 += m.shape[0]
            else:   # 'kcat': concatenation along K
                c0 = 0
                for m in tensors:
                    add(m, None, None, None, buf[:, c0:], m.shape[0], m.shape[1], m.stride(0),
                        buf.stride(0))
                    c0 += m.shape[1]
            self.map[key] = buf
        self.total_rows = row0
        self.table = _upload(segs, self.device) if segs else None
        self.n_segs = len(segs)
        bsegs, row0 = [], 0
        for e in self.wn:
            O, I = e.v.shape
            bsegs.append(TtWnormBwdSeg(e.dw.data_ptr(), e.v.data_ptr(), e.g.data_ptr(),
                                       e.norm.data_ptr(), e.dv.data_ptr(), e.dg.data_ptr(), O, I,
                                       row0, 0))
            row0 += O
        self.bwd_rows = row0
        self.bwd_table = _upload(bsegs, self.device) if bsegs else None
        self.n_bsegs = len(bsegs)

    def _finalize_recording(self):
        """Turn the operand forms recorded during the previous forward into table segments."""
        if not self.log or torch.cuda.is_current_stream_capturing():
            return
        for key, (kind, tensors) in self.log.items():
            if kind == 's':
                rows, cols = tensors[0].shape
            elif kind == 'rows':
                rows, cols = sum(m.shape[0] for m in tensors), tensors[0].shape[1]
            else:
                rows, cols = tensors[0].shape[0], sum(m.shape[1] for m in tensors)
            buf = ops.bf16_buffer(rows, cols, self.device)
            self.extra.append((kind, tensors, buf, key))
        self.log = {}
        self._build_tables()

    def _moved(self):
        return any(p.data_ptr() != ptr for p, ptr in self.params)

    # ------------------------------------------------------------------ per-step entry points
    def usable(self):
        return config.precision == 'bf16' and not self._moved()

    def prepare(self):
        """One launch: every weight-norm + every recorded operand cast."""
        self._finalize_recording()
        if self.table is None:
            return
        _lib.call('tt_weight_prep', c_void_p(self.table.data_ptr()), c_int(self.n_segs),
                  c_int(self.total_rows), ops._stream())

    def wnorm_bwd(self):
        if self.bwd_table is None:
            return
        _lib.call('tt_wnorm_bwd_multi', c_void_p(self.bwd_table.data_ptr()), c_int(self.n_bsegs),
                  c_int(self.bwd_rows), ops._stream())

    def effective_weights(self):
        """Runs the step's preparation (through autograd when gradients are recorded) and returns
        {id(weight_v): fp32 effective weight}."""
        if torch.is_grad_enabled() and any(e.v.requires_grad for e in self.wn):
            flat = []
            for e in self.wn:
                flat += [e.v, e.g]
            ws = _BankWNormFn.apply(self, *flat)
        else:
            self.prepare()
            ws = [e.w32 for e in self.wn]
        return {id(e.v): w for e, w in zip(self.wn, ws)}

    def is_unwritten_handle(self, t):
        """True for a view of the fp32 effective-weight buffer whose contents are not maintained."""
        return (not self.write_w32) and t.untyped_storage().data_ptr() == self._w32.untyped_storage().data_ptr()

    # ------------------------------------------------------------------ lookups (forward only)
    def _backed(self, t):
        return t.untyped_storage().data_ptr() in self.storages

    def single(self, x):
        key = ('s', x.data_ptr(), tuple(x.shape))
        hit = self.map.get(key)
        if hit is None and key not in self.log and self._backed(x) and x.dim() == 2 \
                and x.stride(1) == 1:
            self.log[key] = ('s', (x.detach(),))      # detached: must not pin an autograd graph
        return hit

    def group(self, kind, mats):
        key = (kind,) + tuple((m.data_ptr(), tuple(m.shape)) for m in mats)
        hit = self.map.get(key)
        if hit is None and key not in self.log and all(
                self._backed(m) and m.dim() == 2 and m.stride(1) == 1 for m in mats):
            self.log[key] = (kind, tuple(m.detach() for m in mats))
        return hit

    def claim_dw(self, wptr):
        """Destination for dL/dw of a weight-normalised matrix (first use in this backward)."""
        d = self.dw_dest.get(wptr)
        if d is None or wptr in self.claimed:
            return None
        self.claimed.add(wptr)
        return d


class _BankWNormFn(Function):
    """All weight norms of the decoder as ONE autograd node: forward = tt_weight_prep, backward =
    tt_wnorm_bwd_multi (runs once every dL/dw has been produced)."""

    @staticmethod
    def forward(ctx, bank, *vg):
        bank.prepare()
        ctx.bank = bank
        return tuple(e.w32.detach() for e in bank.wn)

    @staticmethod
    def backward(ctx, *dws):
        bank = ctx.bank
        from . import functional
        functional.wgrad_join()          # dL/dw GEMMs forked onto the weight-gradient stream
        for e, dw in zip(bank.wn, dws):
            if dw is None:
                e.dw.zero_()
            elif dw.data_ptr() != e.dw.data_ptr():
                e.dw.copy_(dw)
        bank.claimed.clear()
        # The gradients returned below are views of the bank's persistent dv / dg buffers, which
        # AccumulateGrad adopts as p.grad without a copy.  If a previous backward's gradient is still
        # there (gradient accumulation, zero_grad(set_to_none=False), two forwards before one step),
        # the launch below would overwrite it and autograd would then add the new gradient to
        # itself: move the old value out of the bank first so that `p.grad += new` is what it says.
        for e in bank.wn:
            for p_, buf in ((e.v, e.dv), (e.g, e.dg)):
                g_ = p_.grad
                if g_ is not None and g_.untyped_storage().data_ptr() == buf.untyped_storage().data_ptr():
                    p_.grad = g_.clone()
        bank.wnorm_bwd()
        grads = []
        for e in bank.wn:
            # fresh view objects: AccumulateGrad can adopt them without a copy
            grads += [e.dv.view_as(e.dv), e.dg.view_as(e.dg)]
        return (None,) + tuple(grads)


# The bank serving the decoder forward that is currently running (None outside of one).
ACTIVE = None


class activate:
    def __init__(self, bank):
        self.bank = bank

    def __enter__(self):
        global ACTIVE
        self.prev = ACTIVE
        ACTIVE = self.bank
        return self.bank

    def __exit__(self, *exc):
        global ACTIVE
        ACTIVE = self.prev
        if self.bank is not None:
            self.bank.current = None
        return False
