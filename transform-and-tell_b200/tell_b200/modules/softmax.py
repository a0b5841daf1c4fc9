"""tell/modules/softmax.py (AdaptiveSoftmax, TiedHeadModule) and
tell/modules/criteria/adaptive_loss.py (AdaptiveLoss) on the B200 kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops
from ..registry import Registrable
from .linear import TiedLinear, linear


class _Weight(nn.Module):
    def __init__(self, out_features, in_features):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return linear(x, self.weight)


class TiedHeadModule(nn.Module):
    """softmax.py:11-40: [tied band-0 embedding ; class_proj] logits."""

    def __init__(self, weights, input_dim, n_classes):
        super().__init__()
        tied_emb, _ = weights
        self.num_words, emb_dim = tied_emb.size()
        if input_dim != emb_dim:
            raise NotImplementedError('input_dim != embedding dim is not used by the shipped configs')
        self.word_proj = TiedLinear(tied_emb, transpose=False)
        self.n_classes = n_classes
        if n_classes > 0:
            self.class_proj = _Weight(n_classes, input_dim)
        self.out_dim = self.num_words + n_classes
        self.register_buffer('_float_tensor', torch.FloatTensor(1).zero_())

    def forward(self, X):
        X = X.reshape(-1, X.shape[-1])
        return torch.cat([linear(X, self.word_proj.weight), linear(X, self.class_proj.weight)], 1)


class AdaptiveSoftmax(nn.Module):
    """softmax.py:43-222 with tied adaptive input embeddings (the only mode the shipped configs
    use: adaptive_inputs given, dropout 0)."""

    def __init__(self, vocab_size, input_dim, cutoff, dropout, factor=4., adaptive_inputs=None,
                 tie_proj=False):
        super().__init__()
        cutoff = list(cutoff)
        if not cutoff or vocab_size > cutoff[-1]:
            cutoff.append(vocab_size)
        assert vocab_size == cutoff[-1]
        if adaptive_inputs is None:
            raise NotImplementedError('untied adaptive softmax is not used by the shipped configs')
        if dropout:
            raise NotImplementedError('adaptive_softmax_dropout is 0 in the shipped configs')
        self.vocab_size = vocab_size
        self.cutoff = cutoff
        self.dropout = dropout
        self.input_dim = input_dim
        self.factor = factor
        n_tails = len(cutoff) - 1
        self.head = TiedHeadModule(adaptive_inputs.weights_for_band(0), input_dim, n_tails)
        self.tail = nn.ModuleList()
        for i in range(n_tails):
            tied_emb, tied_proj = adaptive_inputs.weights_for_band(i + 1)
            proj = TiedLinear(tied_proj, transpose=True) if tie_proj \
                else _Weight(tied_proj.shape[1], input_dim)
            self.tail.append(nn.ModuleList([proj, nn.Identity(), TiedLinear(tied_emb, False)]))
        self.register_buffer('version', torch.LongTensor([1]))

    def _tail_weights(self):
        out = []
        for t in self.tail:
            proj = t[0].weight.t().contiguous() if isinstance(t[0], TiedLinear) else t[0].weight
            out += [proj, t[2].weight]
        return out

    def fused_loss(self, X, target, padding_idx=1):
        """(loss / ln2 / ntokens [1], ntokens int32 [1]) without materialising anything on the
        host (adaptive_loss.py:27-73 + transformer_faces_objects.py:82-90)."""
        X2 = X.reshape(-1, X.shape[-1])
        t = target.reshape(-1).contiguous()
        return Fn.AdaptiveLossFn.apply(X2.contiguous(), t, tuple(self.cutoff), padding_idx,
                                       self.head.word_proj.weight, self.head.class_proj.weight,
                                       *self._tail_weights())

    def adapt_target(self, target):
        """softmax.py:144-167 (API-compatible; synchronises with the host like the reference)."""
        target = target.reshape(-1)
        new_target = [target.clone()]
        target_idxs = []
        for i in range(len(self.cutoff) - 1):
            mask = target.ge(self.cutoff[i]) & target.lt(self.cutoff[i + 1])
            new_target[0][mask] = self.cutoff[0] + i
            if mask.any():
                target_idxs.append(mask.nonzero().squeeze(1))
                new_target.append(target[mask] - self.cutoff[i])
            else:
                target_idxs.append(None)
                new_target.append(None)
        return new_target, target_idxs

    def forward(self, X, target):
        """softmax.py:169-191: (list of per-cluster logits, list of per-cluster targets)."""
        X = X.reshape(-1, X.shape[-1])
        new_target, target_idxs = self.adapt_target(target)
        output = [self.head(X)]
        for i, idx in enumerate(target_idxs):
            if idx is not None:
                h = linear(X.index_select(0, idx), self._tail_weights()[2 * i])
                output.append(linear(h, self.tail[i][2].weight))
            else:
                output.append(None)
        return output, new_target

    @torch.no_grad()
    def _cluster_logits(self, X2):
        a16 = Fn.operand(X2, 'a')
        c0, nt = self.cutoff[0], len(self.cutoff) - 1
        hw16 = Fn.concat_rows_operand([self.head.word_proj.weight, self.head.class_proj.weight],
                                      'b', X2.device)
        M = X2.shape[0]
        head = ops.gemm_tn(a16, hw16, out=ops.f32_padded(M, hw16.shape[0], X2))
        tails = []
        tw = self._tail_weights()
        for i in range(nt):
            h = ops.gemm_tn(a16, Fn.operand(tw[2 * i], 'b'))
            tails.append(ops.gemm_tn(Fn.operand(h, 'a'), Fn.operand(tw[2 * i + 1], 'b'),
                                     out=ops.f32_padded(M, tw[2 * i + 1].shape[0], X2)))
        return head, tails

    @torch.no_grad()
    def get_log_prob(self, X, target=None):
        """softmax.py:193-222: full-vocabulary log-probabilities [B,T,V]."""
        assert target is None
        B, T, E = X.shape
        head, tails = self._cluster_logits(X.reshape(B * T, E).contiguous())
        lp, _, _ = ops.adaptive_logprob(head, tails, self.cutoff, True, False)
        return lp.view(B, T, self.vocab_size)

    @torch.no_grad()
    def greedy(self, X):
        """argmax token id and its log-prob per row, never materialising [*, V] log-probs
        (topk(1) + multinomial over one candidate == argmax, transformer_faces_objects.py:443-464)."""
        E = X.shape[-1]
        head, tails = self._cluster_logits(X.reshape(-1, E).contiguous())
        _, ids, lps = ops.adaptive_logprob(head, tails, self.cutoff, False, True)
        return ids, lps


class Criterion(nn.Module, Registrable):
    pass


@Criterion.register('adaptive_loss')
class AdaptiveLoss(Criterion):
    """criteria/adaptive_loss.py:11-73."""

    def __init__(self, padding_idx=1):
        super().__init__()
        self.padding_idx = padding_idx
        self.sentence_avg = False

    def forward(self, adaptive_softmax, net_output, decoder_target, reduction='sum'):
        """Returns (summed loss [1], sample_size int) like the reference (one host sync for the
        int).  The model's hot path uses fused() instead."""
        if reduction != 'sum':
            raise NotImplementedError("only reduction='sum' is used on this path")
        loss, ntok = adaptive_softmax.fused_loss(net_output[0], decoder_target, self.padding_idx)
        n = int(ntok.item())
        return loss * (math.log(2) * n), n

    def fused(self, adaptive_softmax, net_output, decoder_target):
        """(loss / ln2 / ntokens, ntokens) as device tensors: no host synchronisation."""
        return adaptive_softmax.fused_loss(net_output[0], decoder_target, self.padding_idx)
