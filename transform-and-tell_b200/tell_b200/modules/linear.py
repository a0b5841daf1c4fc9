"""tell/modules/linear.py:8-49 on the B200 kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops, weight_bank


def linear(x, weight, bias=None, alpha=1.0, act=ops.ACT_NONE, twin_out=False):
    """F.linear over the last dim through the tcgen05 GEMM.  twin_out: the result feeds another GEMM,
    so the epilogue also writes its bf16 operand (tell_b200/twin.py)."""
    shape = x.shape
    y = Fn.LinearFn.apply(x.reshape(-1, shape[-1]), weight, bias, alpha, act, twin_out)
    return y.view(*shape[:-1], weight.shape[0])


class Linear(nn.Module):
    """Plain nn.Linear parameter holder (state-dict keys weight / bias)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return linear(x, self.weight, self.bias)


class GehringLinear(nn.Module):
    """Weight-normalised linear with Gehring init (linear.py:8-34).
    Parameters: weight_g [out,1], weight_v [out,in], bias [out]."""

    def __init__(self, in_features, out_features, dropout=0, bias=True, weight_norm=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.dropout = dropout
        self.weight_norm = weight_norm
        std = math.sqrt((1 - dropout) / in_features)
        w = torch.empty(out_features, in_features).normal_(mean=0, std=std)
        if weight_norm:
            self.weight_g = nn.Parameter(w.norm(dim=1, keepdim=True))
            self.weight_v = nn.Parameter(w)
        else:
            self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None

    def effective_weight(self):
        if self.weight_norm:
            bank = weight_bank.ACTIVE
            if bank is not None and bank.current is not None:
                w = bank.current.get(id(self.weight_v))     # from the step's one weight_prep launch
                if w is not None:
                    return w
            return Fn.WeightNormFn.apply(self.weight_v, self.weight_g)
        return self.weight

    def forward(self, x, act=ops.ACT_NONE, twin_out=False):
        return linear(x, self.effective_weight(), self.bias, 1.0, act, twin_out)


class TiedLinear(nn.Module):
    """linear.py:37-49: a linear layer whose weight is another module's parameter."""

    def __init__(self, weight, transpose):
        super().__init__()
        self.weight = weight
        self.transpose = transpose

    def forward(self, x):
        w = self.weight.t().contiguous() if self.transpose else self.weight
        return linear(x, w)


class LayerNorm(nn.Module):
    """nn.LayerNorm parameter holder; the fused residual+dropout+LN kernel does the work."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = eps

    def forward(self, h, residual=None, p=0.0, seed=0):
        shape = h.shape
        h2 = h.reshape(-1, shape[-1])
        r2 = residual.reshape(-1, shape[-1]) if residual is not None else None
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        elif h2.requires_grad is False or h2.is_leaf:
            h2 = h2.clone()   # never clobber a caller-owned tensor
        y = Fn.ResidualLayerNormFn.apply(h2, r2, self.weight, self.bias, p, seed, self.eps)
        return y.view(shape)
