"""RoBERTa article encoder on B200: the `roberta.extract_features(ids, return_all_hiddens=True)`
call of transformer_faces_objects.py:352-353 (fairseq hub `roberta.large`, an un-vendored
third-party dependency of the reference -- see DESIGN.md).  Frozen (`no_grad: ^roberta`), so it is
inference only: bf16 activations, tcgen05 GEMMs with fused bias/GELU/residual epilogues, flash
self-attention on tensor cores, LayerNorm writing each layer's hidden state straight into the
[L+1, B*S, E] buffer the layer mix reads.  State-dict keys follow fairseq's
`decoder.sentence_encoder.*` layout."""
import torch
import torch.nn as nn

from .. import ops


class _LN(nn.Module):
    def __init__(self, e):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(e))
        self.bias = nn.Parameter(torch.zeros(e))


class _Lin(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(o))


class _SelfAttn(nn.Module):
    def __init__(self, e):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * e, e).normal_(std=0.02))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * e))
        self.out_proj = _Lin(e, e)


class _Layer(nn.Module):
    def __init__(self, e, ffn):
        super().__init__()
        self.self_attn = _SelfAttn(e)
        self.self_attn_layer_norm = _LN(e)
        self.fc1, self.fc2 = _Lin(e, ffn), _Lin(ffn, e)
        self.final_layer_norm = _LN(e)


class _SentenceEncoder(nn.Module):
    def __init__(self, vocab, e, n_layers, ffn, max_pos, init_device=None):
        super().__init__()
        self.embed_tokens = nn.Embedding(vocab, e, padding_idx=1, device=init_device)
        self.embed_positions = nn.Embedding(max_pos, e, padding_idx=1, device=init_device)
        self.emb_layer_norm = _LN(e)
        self.layers = nn.ModuleList([_Layer(e, ffn) for _ in range(n_layers)])


class _Decoder(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.sentence_encoder = _SentenceEncoder(*a, **k)


class RobertaEncoder(nn.Module):
    """roberta.large = 24 layers, E 1024, 16 heads, FFN 4096, vocab 50265, 514 positions."""

    def __init__(self, n_layers=24, embed_dim=1024, heads=16, ffn=4096, vocab=50265, max_pos=514,
                 padding_idx=1):
        super().__init__()
        self.decoder = _Decoder(vocab, embed_dim, n_layers, ffn, max_pos)
        self.n_layers, self.embed_dim, self.heads = n_layers, embed_dim, heads
        self.padding_idx = padding_idx
        self._prep = None
        # Skip the padding: pack the real tokens of the batch and run every GEMM / attention /
        # LayerNorm of the encoder on the packed rows only (all consumers mask the padding).
        self.varlen = True

    def _apply(self, fn, *a, **k):
        self._prep = None
        object.__setattr__(self, '_qkv_key', None)
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)      # in-place copies: the derived bf16 operands are stale
        self._prep = None
        return r

    def prepare(self):
        """bf16 GEMM operands, with the query scaling d^-0.5 folded into Wq / bq (frozen weights)."""
        E, d = self.embed_dim, self.embed_dim // self.heads
        se = self.decoder.sentence_encoder
        prep = []
        for l in se.layers:
            w = l.self_attn.in_proj_weight.detach().float().clone()
            b = l.self_attn.in_proj_bias.detach().float().clone()
            w[:E] *= d ** -0.5
            b[:E] *= d ** -0.5
            prep.append(dict(
                wqkv=w.to(torch.bfloat16), bqkv=b.contiguous(),
                wo=l.self_attn.out_proj.weight.detach().to(torch.bfloat16),
                bo=l.self_attn.out_proj.bias.detach().float(),
                w1=l.fc1.weight.detach().to(torch.bfloat16), b1=l.fc1.bias.detach().float(),
                w2=l.fc2.weight.detach().to(torch.bfloat16), b2=l.fc2.bias.detach().float()))
        self._prep = prep
        return self

    @torch.no_grad()
    def all_hiddens(self, ids, n_real_tokens=0):
        """ids [B,S] int64 -> (bf16 [L+1, B*S, E], key padding mask uint8 [B*S]).
        With `varlen` the rows of padding tokens are zero instead of the encoder's values there.
        n_real_tokens: optional host-side count of non-padding tokens (the data loader knows it);
        a GEMM tile-scheduling hint only."""
        if self._prep is None:
            self.prepare()
        if self.varlen:
            return self._all_hiddens_packed(ids, int(n_real_tokens))
        se = self.decoder.sentence_encoder
        B, S = ids.shape
        E, H = self.embed_dim, self.heads
        L = self.n_layers
        hid = torch.empty((L + 1, B * S, E), dtype=torch.bfloat16, device=ids.device)
        x, is_pad = ops.roberta_embed(ids.contiguous(), se.embed_tokens.weight,
                                      se.embed_positions.weight, self.padding_idx)
        ops.ln_fwd16(x, se.emb_layer_norm.weight, se.emb_layer_norm.bias, hid[0], row_zero=is_pad)
        tmp = torch.empty((B * S, E), dtype=torch.float32, device=ids.device)
        x16 = torch.empty((B * S, E), dtype=torch.bfloat16, device=ids.device)
        for i, (l, p) in enumerate(zip(se.layers, self._prep)):
            h_in = hid[i]
            qkv = ops.gemm_tn(h_in, p['wqkv'], bias=p['bqkv'], want32=False, want16=True)
            a = ops.flash_self_attn(qkv, is_pad, B, S, H, E // H)
            ops.gemm_tn(a, p['wo'], out=tmp, bias=p['bo'], residual16=h_in)
            ops.ln_fwd16(tmp, l.self_attn_layer_norm.weight, l.self_attn_layer_norm.bias, x16)
            f = ops.gemm_tn(x16, p['w1'], bias=p['b1'], act=ops.ACT_GELU, want32=False, want16=True)
            ops.gemm_tn(f, p['w2'], out=tmp, bias=p['b2'], residual16=x16)
            ops.ln_fwd16(tmp, l.final_layer_norm.weight, l.final_layer_norm.bias, hid[i + 1])
        return hid, is_pad

    def _all_hiddens_packed(self, ids, hint=0):
        """Variable-length forward: row counts are device values (cu_seqlens[B] is the m_limit of
        every GEMM), so there is no host sync and the whole thing captures into a CUDA graph; the
        buffers are sized for the padded batch."""
        se = self.decoder.sentence_encoder
        B, S = ids.shape
        E, H = self.embed_dim, self.heads
        L = self.n_layers
        dev = ids.device
        R = B * S
        ids = ids.contiguous()
        hid = torch.empty((L + 1, R, E), dtype=torch.bfloat16, device=dev)
        inv_map, cu = ops.varlen_prepare(ids, self.padding_idx)
        ntok = cu[B:B + 1]
        x, is_pad = ops.roberta_embed(ids, se.embed_tokens.weight, se.embed_positions.weight,
                                      self.padding_idx)
        h = [torch.empty((R, E), dtype=torch.bfloat16, device=dev) for _ in range(2)]   # packed
        ops.ln_fwd16_varlen(x, se.emb_layer_norm.weight, se.emb_layer_norm.bias, y_packed=h[0],
                            y_padded=hid[0], inv_map=inv_map, x_packed=False)
        tmp = torch.empty((R, E), dtype=torch.bfloat16, device=dev)      # pre-norm residual sums
        x16 = torch.empty((R, E), dtype=torch.bfloat16, device=dev)
        # zeros: rows beyond the packed token count are masked keys of the last sample, and a masked
        # key still enters P.V with weight 0 (0 * garbage NaN would poison the row)
        # persistent: zero-filled once, afterwards every row only ever holds finite GEMM outputs, so
        # the step does not pay a 50 MB memset (callers on two streams must not share one encoder)
        key = (R, str(dev))
        if getattr(self, '_qkv_key', None) == key:
            qkv = self._qkv_buf
        elif torch.cuda.is_current_stream_capturing():
            qkv = torch.zeros((R, 3 * E), dtype=torch.bfloat16, device=dev)   # graph-private, not kept
        else:
            qkv = torch.zeros((R, 3 * E), dtype=torch.bfloat16, device=dev)
            object.__setattr__(self, '_qkv_buf', qkv)
            object.__setattr__(self, '_qkv_key', key)
        f = torch.empty((R, se.layers[0].fc1.weight.shape[0]), dtype=torch.bfloat16, device=dev)
        for i, (l, p) in enumerate(zip(se.layers, self._prep)):
            h_in, h_out = h[i & 1], h[(i + 1) & 1]
            ops.gemm_tn(h_in, p['wqkv'], out16=qkv, bias=p['bqkv'], want32=False, m_limit=ntok, m_hint=hint)
            a = ops.flash_self_attn_varlen(qkv, cu, B, S, H, E // H)
            ops.gemm_tn(a, p['wo'], out16=tmp, bias=p['bo'], residual16=h_in, want32=False,
                        m_limit=ntok, m_hint=hint)
            ops.ln_fwd16_varlen(tmp, l.self_attn_layer_norm.weight, l.self_attn_layer_norm.bias,
                                y_packed=x16, count=ntok)
            ops.gemm_tn(x16, p['w1'], out16=f, bias=p['b1'], act=ops.ACT_GELU, want32=False,
                        m_limit=ntok, m_hint=hint)
            ops.gemm_tn(f, p['w2'], out16=tmp, bias=p['b2'], residual16=x16, want32=False,
                        m_limit=ntok, m_hint=hint)
            ops.ln_fwd16_varlen(tmp, l.final_layer_norm.weight, l.final_layer_norm.bias,
                                y_packed=h_out, y_padded=hid[i + 1], inv_map=inv_map)
        return hid, is_pad

    @torch.no_grad()
    def extract_features(self, tokens, return_all_hiddens=False):
        """fairseq hub-interface API: list of L+1 fp32 [B,S,E] tensors (or the last one)."""
        hid, _ = self.all_hiddens(tokens)
        B, S = tokens.shape
        outs = [ops.bf16_to_f32(h).view(B, S, -1) for h in (hid if return_all_hiddens else hid[-1:])]
        return outs if return_all_hiddens else outs[0]
