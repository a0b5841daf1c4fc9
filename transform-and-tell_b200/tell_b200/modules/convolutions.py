"""tell/modules/convolutions/{dynamic,lightweight}.py on the B200 kernels."""
import torch
import torch.nn as nn

from .. import config, ops, twin
from .. import functional as Fn
from ..utils import get_incremental_state, set_incremental_state
from .linear import linear


class _WeightLinear(nn.Module):
    def __init__(self, in_features, out_features, bias):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None


_OWNED = '__tt_owned_windows__'


def _step_window(module, incremental_state, X):
    """Fixed [K-1,B,C] time-ordered input buffer of an incremental decode (T = 1, no autograd).
    The reference's buffer grows from 0 to K-1 rows (dynamic.py:95-99); zero rows are exactly the
    causal zero padding it would apply, so a zero-initialised full-size buffer gives the same
    numbers with step-independent shapes (which is what lets a decode step be captured in a CUDA
    graph).  The buffer is owned by the state and updated in place by tt_dynconv_step."""
    K = module.kernel_size
    _, B, C = X.shape
    buf = get_incremental_state(module, incremental_state, 'input_buffer')
    owned_set = incremental_state.setdefault(_OWNED, set())   # buffers this module may overwrite
    owned = buf is not None and buf.data_ptr() in owned_set
    if K == 1:
        return None
    if buf is None:
        buf = torch.zeros((K - 1, B, C), dtype=X.dtype, device=X.device)
    elif buf.shape[0] < K - 1:
        pad = torch.zeros((K - 1 - buf.shape[0], B, C), dtype=X.dtype, device=X.device)
        buf = torch.cat([pad, buf], dim=0)
    elif not owned or not buf.is_contiguous():
        buf = buf.clone()            # may alias a caller's tensor (prefix fed with T > 1)
    set_incremental_state(module, incremental_state, 'input_buffer', buf)
    owned_set.add(buf.data_ptr())
    return buf


def _is_decode_step(X, incremental_state, query=None):
    return (incremental_state is not None and X.shape[0] == 1 and query is None
            and not torch.is_grad_enabled())


class DynamicConv1dTBC(nn.Module):
    """dynamic.py:25-361.  Supports the configuration every shipped model uses
    (weight_softmax, padding_l = K-1, no renorm_padding, no in_proj, query = X)."""

    def __init__(self, input_size, kernel_size=1, padding_l=None, num_heads=1, weight_dropout=0.,
                 weight_softmax=False, renorm_padding=False, bias=False, conv_bias=False,
                 query_size=None, in_proj=False):
        super().__init__()
        if renorm_padding or in_proj:
            raise NotImplementedError('renorm_padding / in_proj are not used by any shipped model')
        if padding_l is not None and padding_l != kernel_size - 1:
            raise NotImplementedError('only causal padding_l = kernel_size - 1 is supported')
        self.input_size = input_size
        self.query_size = input_size if query_size is None else query_size
        self.kernel_size = kernel_size
        self.padding_l = padding_l
        self.num_heads = num_heads
        self.weight_dropout = weight_dropout
        self.weight_softmax = weight_softmax
        self.renorm_padding = renorm_padding
        self.weight_linear = _WeightLinear(self.query_size, num_heads * kernel_size, bias)
        self.conv_bias = nn.Parameter(torch.zeros(input_size)) if conv_bias else None

    def forward(self, X, incremental_state=None, query=None, unfold=None):
        """X [T,B,C] -> [T,B,C]; with incremental_state the last K-1 inputs are buffered
        (dynamic.py:95-99) and only the new rows are returned (:115-116)."""
        assert X.dim() == 3 and X.shape[2] == self.input_size
        if _is_decode_step(X, incremental_state, query) and not self.training:
            # one new time step: only its own tap logits and its own output row are computed
            _, B, C = X.shape
            x_new = X[0].contiguous()
            window = _step_window(self, incremental_state, X)
            z = linear(x_new, self.weight_linear.weight, self.weight_linear.bias)
            if Fn._tw() and C % 8 == 0:
                out, out16 = ops.dynconv_step(window, x_new, z.contiguous(), self.num_heads,
                                              self.kernel_size, self.weight_softmax, twin=True)
                twin.put(out, out16)               # operand of linear2
            else:
                out = ops.dynconv_step(window, x_new, z.contiguous(), self.num_heads, self.kernel_size,
                                       self.weight_softmax)
            return out.view(1, B, C)
        prev = None
        if incremental_state is not None:
            prev = get_incremental_state(self, incremental_state, 'input_buffer')
            if prev is not None:
                X = torch.cat([prev, X], dim=0)
            set_incremental_state(self, incremental_state, 'input_buffer',
                                  X[-self.kernel_size + 1:] if self.kernel_size > 1 else X[:0])
        q = X if query is None else query
        T, B, C = X.shape
        X = X.contiguous()
        z = linear(q, self.weight_linear.weight, self.weight_linear.bias)
        p = self.weight_dropout if self.training else 0.0
        out = Fn.DynConvFn.apply(X, z.contiguous(), self.num_heads, self.kernel_size,
                                 self.weight_softmax, p, config.next_seed() if p > 0 else 0)
        if prev is not None:
            out = out[prev.shape[0]:]
        if self.conv_bias is not None:
            raise NotImplementedError('conv_bias is not used by any shipped model')
        return out

    def reorder_incremental_state(self, incremental_state, new_order):
        buf = get_incremental_state(self, incremental_state, 'input_buffer')
        if buf is not None:
            set_incremental_state(self, incremental_state, 'input_buffer',
                                  buf.index_select(1, new_order))


class LightweightConv1dTBC(nn.Module):
    """lightweight.py:88-240: static depthwise taps weight [H,1,K], softmax-normalised."""

    def __init__(self, input_size, kernel_size=1, padding_l=None, num_heads=1, weight_dropout=0.,
                 weight_softmax=False, bias=False):
        super().__init__()
        if padding_l is not None and padding_l != kernel_size - 1:
            raise NotImplementedError('only causal padding_l = kernel_size - 1 is supported')
        self.input_size = input_size
        self.kernel_size = kernel_size
        self.padding_l = padding_l
        self.num_heads = num_heads
        self.weight_dropout = weight_dropout
        self.weight_softmax = weight_softmax
        self.weight = nn.Parameter(torch.empty(num_heads, 1, kernel_size))
        nn.init.xavier_uniform_(self.weight)
        self.bias = nn.Parameter(torch.zeros(input_size)) if bias else None

    def forward(self, X, incremental_state=None, unfold=False):
        if _is_decode_step(X, incremental_state) and not self.training:
            _, B, C = X.shape
            x_new = X[0].contiguous()
            window = _step_window(self, incremental_state, X)
            out = ops.dynconv_step(window, x_new, self.weight.view(self.num_heads, -1).contiguous(),
                                   self.num_heads, self.kernel_size, self.weight_softmax,
                                   broadcast=True)
            return out.view(1, B, C)
        prev = None
        if incremental_state is not None:
            prev = get_incremental_state(self, incremental_state, 'input_buffer')
            if prev is not None:
                X = torch.cat([prev, X], dim=0)
            set_incremental_state(self, incremental_state, 'input_buffer',
                                  X[-self.kernel_size + 1:] if self.kernel_size > 1 else X[:0])
        p = self.weight_dropout if self.training else 0.0
        out = Fn.LightConvFn.apply(X.contiguous(), self.weight.view(self.num_heads, -1),
                                   self.num_heads, self.kernel_size, self.weight_softmax, p,
                                   config.next_seed() if p > 0 else 0)
        if prev is not None:
            out = out[prev.shape[0]:]
        if self.bias is not None:
            raise NotImplementedError('bias is not used by any shipped model')
        return out
