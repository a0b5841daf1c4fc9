"""InceptionResnetV1 ("FaceNet") face embedder -- tell/facenet/inception_resnet_v1.py:184-299 -- on
the B200 kernels.  It produces the 512-d `face_embeds` the caption model attends over
(scripts/detect_facenet_nytimes.py, tell/tasks/captioner.py:291-303).

Inference only (the reference uses it in eval mode under no_grad).  Activations are NHWC bf16;
every BasicConv2d (conv, BatchNorm eps=1e-3, ReLU; :10-34) is im2col + the tcgen05 GEMM with the
folded BatchNorm bias and ReLU in the epilogue; the residual blocks' `conv2d(cat) * scale + x`
(:57-60, :85-88, :114-118) is ONE GEMM whose epilogue applies scale, adds the bf16 identity and the
ReLU; branch outputs are written directly into channel slices of the concatenation buffer, and the
1x1 convolutions that open the branches of a block share one GEMM.  Module tree and state-dict keys
are the reference's (conv2d_1a.conv.weight, repeat_1.0.branch1.1.bn.running_var, ...), so a
`vggface2` checkpoint loads with strict=True."""
import torch
import torch.nn as nn

from .. import ops


class BasicConv2d(nn.Module):
    def __init__(self, in_planes, out_planes, kernel_size, stride, padding=0):
        super().__init__()
        kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
        ph, pw = (padding, padding) if isinstance(padding, int) else padding
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=(kh, kw), stride=stride,
                              padding=(ph, pw), bias=False)
        self.bn = nn.BatchNorm2d(out_planes, eps=0.001, momentum=0.1, affine=True)
        self.geom = (kh, kw, stride, ph, pw)


def _pad8(n):
    return (n + 7) // 8 * 8


def _fold(conv, bn=None):
    """conv (+BatchNorm running statistics) -> (bf16 [Cout, Kp] in (kh,kw,cin) order, fp32 bias)."""
    w = conv.weight.detach().float()
    cout = w.shape[0]
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        bias = bn.bias.detach().float() - bn.running_mean.float() * scale
        w = w * scale.view(-1, 1, 1, 1)
    else:
        bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=w.device)
    w = w.permute(0, 2, 3, 1).reshape(cout, -1)
    out = torch.zeros((cout, _pad8(w.shape[1])), dtype=torch.bfloat16, device=w.device)
    out[:, :w.shape[1]] = w.to(torch.bfloat16)
    return out, bias.contiguous()


def _cat_rows(folded):
    """Several folded convolutions with the same input -> one GEMM operand (rows concatenated)."""
    return torch.cat([w for w, _ in folded], 0).contiguous(), torch.cat([b for _, b in folded]).contiguous()


def conv_apply(x, folded, geom, act=ops.ACT_RELU, out=None, alpha=1.0, residual16=None):
    """x [B,H,W,C] bf16 view -> [B,Ho,Wo,Cout] bf16 (written into `out`, possibly a channel slice)."""
    w, b = folded
    kh, kw, stride, ph, pw = geom
    B, H, W, C = x.shape
    if kh == 1 and kw == 1 and stride == 1 and x.is_contiguous():
        cols, Ho, Wo = x.view(B * H * W, C), H, W
    else:
        cols, Ho, Wo = ops.im2col_nhwc_hw(x, kh, kw, stride, ph, pw)
    cout = w.shape[0]
    if out is None:
        out = torch.empty((B, Ho, Wo, cout), dtype=torch.bfloat16, device=x.device)
    o2 = out.as_strided((B * Ho * Wo, cout), (out.stride(2), 1), out.storage_offset())
    ops.gemm_tn(cols, w, out16=o2, bias=b, act=act, alpha=alpha, residual16=residual16, want32=False)
    return out


class _ResidualBlock(nn.Module):
    """Block35 / Block17 / Block8 (:37-119): parallel conv branches, concat, 1x1 conv with bias,
    out * scale + x, optional ReLU."""
    IN = 0
    MID = 0          # channels of every branch output
    TAILS = ()       # per extra branch: geometries of the convolutions after its opening 1x1

    def __init__(self, scale=1.0, noReLU=False):
        super().__init__()
        self.scale, self.noReLU = scale, noReLU
        self.branch0 = BasicConv2d(self.IN, self.MID, kernel_size=1, stride=1)
        for i, tail in enumerate(self.TAILS, start=1):
            mods = [BasicConv2d(self.IN, self.MID, kernel_size=1, stride=1)]
            mods += [BasicConv2d(self.MID, self.MID, kernel_size=k, stride=1, padding=p) for k, p in tail]
            setattr(self, 'branch%d' % i, nn.Sequential(*mods))
        self.conv2d = nn.Conv2d(self.MID * (1 + len(self.TAILS)), self.IN, kernel_size=1, stride=1)
        self._f = None

    def prepare(self):
        nb = 1 + len(self.TAILS)
        heads = [_fold(self.branch0.conv, self.branch0.bn)]
        tails = []
        for i in range(1, nb):
            seq = getattr(self, 'branch%d' % i)
            heads.append(_fold(seq[0].conv, seq[0].bn))
            tails.append([(_fold(m.conv, m.bn), m.geom) for m in list(seq)[1:]])
        # GEMM column order of the fused opening 1x1s: [branch1 .. branch_{nb-1} | branch0], so that
        # branch0 lands in its final concat slot when the buffer is laid out the same way; instead
        # keep it simple: heads go to a scratch buffer, branch0 is copied by its own GEMM slot.
        self._f = dict(heads=_cat_rows(heads), tails=tails, out=_fold(self.conv2d))
        return self

    def forward_nhwc(self, x):
        if self._f is None:
            self.prepare()
        f = self._f
        B, H, W, C = x.shape
        nb, mid = 1 + len(self.TAILS), self.MID
        # all opening 1x1 convolutions as ONE GEMM -> [B,H,W, nb*mid]; branch0's slice is final
        cat = conv_apply(x, f['heads'], (1, 1, 1, 0, 0))
        for i, tail in enumerate(f['tails'], start=1):
            sl = cat[..., i * mid:(i + 1) * mid]
            cur = sl
            for j, (fw, geom) in enumerate(tail):
                last = j == len(tail) - 1
                cur = conv_apply(cur, fw, geom, out=sl if last else None)
        act = ops.ACT_NONE if self.noReLU else ops.ACT_RELU
        x2 = x.view(B * H * W, C)
        return conv_apply(cat, f['out'], (1, 1, 1, 0, 0), act=act, alpha=self.scale, residual16=x2)


class Block35(_ResidualBlock):
    IN, MID = 256, 32
    TAILS = ([(3, 1)], [(3, 1), (3, 1)])


class Block17(_ResidualBlock):
    IN, MID = 896, 128
    TAILS = ([((1, 7), (0, 3)), ((7, 1), (3, 0))],)


class Block8(_ResidualBlock):
    IN, MID = 1792, 192
    TAILS = ([((1, 3), (0, 1)), ((3, 1), (1, 0))],)


class _Reduction(nn.Module):
    """Mixed_6a / Mixed_7a (:122-181): strided conv branches + a 3x3/2 max-pool, concatenated."""
    SPECS = ()      # per branch: list of (cin, cout, k, stride, pad)

    def __init__(self):
        super().__init__()
        for i, spec in enumerate(self.SPECS):
            mods = [BasicConv2d(ci, co, kernel_size=k, stride=s, padding=p) for ci, co, k, s, p in spec]
            setattr(self, 'branch%d' % i, mods[0] if len(mods) == 1 and self.SINGLE0 and i == 0
                    else nn.Sequential(*mods))
        setattr(self, 'branch%d' % len(self.SPECS), nn.MaxPool2d(3, stride=2))
        self._f = None

    SINGLE0 = False

    def _mods(self, i):
        b = getattr(self, 'branch%d' % i)
        return [b] if isinstance(b, BasicConv2d) else list(b)

    def prepare(self):
        self._f = [[(_fold(m.conv, m.bn), m.geom) for m in self._mods(i)] for i in range(len(self.SPECS))]
        return self

    def forward_nhwc(self, x):
        if self._f is None:
            self.prepare()
        B, H, W, C = x.shape
        Ho, Wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        couts = [spec[-1][1] for spec in self.SPECS]
        out = torch.empty((B, Ho, Wo, sum(couts) + C), dtype=torch.bfloat16, device=x.device)
        c0 = 0
        for convs, co in zip(self._f, couts):
            cur = x
            for j, (fw, geom) in enumerate(convs):
                last = j == len(convs) - 1
                cur = conv_apply(cur, fw, geom, out=out[..., c0:c0 + co] if last else None)
            c0 += co
        ops.maxpool_nhwc(x, 3, 2, out=out[..., c0:c0 + C])
        return out


class Mixed_6a(_Reduction):
    SINGLE0 = True
    SPECS = ([(256, 384, 3, 2, 0)],
             [(256, 192, 1, 1, 0), (192, 192, 3, 1, 1), (192, 256, 3, 2, 0)])


class Mixed_7a(_Reduction):
    SPECS = ([(896, 256, 1, 1, 0), (256, 384, 3, 2, 0)],
             [(896, 256, 1, 1, 0), (256, 256, 3, 2, 0)],
             [(896, 256, 1, 1, 0), (256, 256, 3, 1, 1), (256, 256, 3, 2, 0)])


class InceptionResnetV1(nn.Module):
    """Same constructor as the reference (:201-262).  Pretrained weights are a network download in
    the reference (:312-338); here `pretrained` must be None and a checkpoint is loaded with
    load_state_dict()."""

    def __init__(self, pretrained=None, classify=False, num_classes=None, dropout_prob=0.6):
        super().__init__()
        if pretrained is not None:
            raise RuntimeError('no network access: build with pretrained=None and load_state_dict() '
                               'a vggface2 / casia-webface checkpoint')
        if num_classes is None:
            raise Exception('At least one of "pretrained" or "num_classes" must be specified')
        self.pretrained, self.classify, self.num_classes = pretrained, classify, num_classes
        self.conv2d_1a = BasicConv2d(3, 32, kernel_size=3, stride=2)
        self.conv2d_2a = BasicConv2d(32, 32, kernel_size=3, stride=1)
        self.conv2d_2b = BasicConv2d(32, 64, kernel_size=3, stride=1, padding=1)
        self.maxpool_3a = nn.MaxPool2d(3, stride=2)
        self.conv2d_3b = BasicConv2d(64, 80, kernel_size=1, stride=1)
        self.conv2d_4a = BasicConv2d(80, 192, kernel_size=3, stride=1)
        self.conv2d_4b = BasicConv2d(192, 256, kernel_size=3, stride=2)
        self.repeat_1 = nn.Sequential(*[Block35(scale=0.17) for _ in range(5)])
        self.mixed_6a = Mixed_6a()
        self.repeat_2 = nn.Sequential(*[Block17(scale=0.10) for _ in range(10)])
        self.mixed_7a = Mixed_7a()
        self.repeat_3 = nn.Sequential(*[Block8(scale=0.20) for _ in range(5)])
        self.block8 = Block8(noReLU=True)
        self.avgpool_1a = nn.AdaptiveAvgPool2d(1)
        self.dropout = nn.Dropout(dropout_prob)
        self.last_linear = nn.Linear(1792, 512, bias=False)
        self.last_bn = nn.BatchNorm1d(512, eps=0.001, momentum=0.1, affine=True)
        self.logits = nn.Linear(512, num_classes)
        self._f = None

    def _invalidate(self):
        self._f = None
        for m in self.modules():
            if isinstance(m, (_ResidualBlock, _Reduction)):
                m._f = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)      # in-place copies: folded operands are stale
        self._invalidate()
        return r

    def prepare(self):
        stem = [self.conv2d_1a, self.conv2d_2a, self.conv2d_2b, self.conv2d_3b, self.conv2d_4a,
                self.conv2d_4b]
        f = {'stem': [(_fold(m.conv, m.bn), m.geom) for m in stem]}
        bn = self.last_bn
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        w = self.last_linear.weight.detach().float() * scale.view(-1, 1)
        f['last'] = (w.to(torch.bfloat16).contiguous(),
                     (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous())
        f['logits'] = (self.logits.weight.detach().to(torch.bfloat16).contiguous(),
                       self.logits.bias.detach().float().contiguous())
        self._f = f
        return self

    @torch.no_grad()
    def forward(self, x):
        """x [B,3,H,W] fp32 (fixed_image_standardization applied by the caller, as in the
        reference) -> (l2-normalised embeddings [B,512] fp32, logits [B,num_classes] fp32)."""
        if self.training:
            raise RuntimeError('InceptionResnetV1 on the B200 path is inference-only (call .eval())')
        if self._f is None:
            self.prepare()
        f = self._f
        B = x.shape[0]
        (w, b), (kh, kw, s, ph, pw) = f['stem'][0]
        cols, Ho, Wo = ops.im2col_nchw_f32(x.contiguous(), kh, kw, s, ph, w.shape[1])
        h = ops.gemm_tn(cols, w, bias=b, act=ops.ACT_RELU, want32=False, want16=True).view(B, Ho, Wo, -1)
        h = conv_apply(h, *f['stem'][1])
        h = conv_apply(h, *f['stem'][2])
        h = ops.maxpool_nhwc(h, 3, 2)
        h = conv_apply(h, *f['stem'][3])
        h = conv_apply(h, *f['stem'][4])
        h = conv_apply(h, *f['stem'][5])
        for blk in self.repeat_1:
            h = blk.forward_nhwc(h)
        h = self.mixed_6a.forward_nhwc(h)
        for blk in self.repeat_2:
            h = blk.forward_nhwc(h)
        h = self.mixed_7a.forward_nhwc(h)
        for blk in self.repeat_3:
            h = blk.forward_nhwc(h)
        h = self.block8.forward_nhwc(h)
        pooled = ops.avgpool_nhwc(h)                                   # [B,1792] fp32 (dropout: eval)
        emb = ops.gemm_tn(ops.cast_bf16(pooled), f['last'][0], bias=f['last'][1])
        emb = ops.l2norm_rows(emb)
        logits = ops.gemm_tn(ops.cast_bf16(emb), f['logits'][0], bias=f['logits'][1])
        return emb, logits
