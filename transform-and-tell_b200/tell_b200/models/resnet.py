"""ResNet-152 feature extractor (tell/models/resnet.py:12-117, torchvision Bottleneck) on B200.

No gradients (the reference freezes it: `no_grad: ^resnet`, config.yaml:150-152).  Activations
are NHWC bf16; every convolution is im2col + the tcgen05 GEMM.

BatchNorm follows nn.BatchNorm2d: `bn_mode`
  'auto'     (default) batch statistics when the module is in train(), running statistics in eval()
             -- exactly what the reference does: its training step calls model.train()
             (callback_apex_trainer.py:259), which leaves the FROZEN backbone normalising with batch
             statistics and moving its running statistics every step; evaluate / demo run eval();
  'running'  always the running statistics, folded into the GEMM weights with bias, identity branch
             and ReLU in the GEMM epilogue;
  'batch'    always batch statistics: the GEMM writes the raw convolution and accumulates the
             per-channel sums in its epilogue (`col_stats`); `tt_bn_apply_bf16` normalises (+ identity,
             ReLU) in place -- except between conv1 and conv2 of a Bottleneck, where the 3x3 im2col
             applies bn1 + ReLU while it gathers (`tt_im2col_nhwc_bn`) -- and the running statistics /
             num_batches_tracked move as in nn.BatchNorm2d.  6 launches per Bottleneck instead of 10.
State-dict keys are torchvision's (conv1.weight, bn1.*, layerL.i.convJ.weight, ...)."""
import torch
import torch.nn as nn

from .. import ops


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


class _Conv(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode='fan_out', nonlinearity='relu')
        self.k, self.stride, self.padding = k, stride, padding


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1, self.bn1 = _Conv(inplanes, planes, 1), _BN(planes)
        self.conv2, self.bn2 = _Conv(planes, planes, 3, stride, 1), _BN(planes)
        self.conv3, self.bn3 = _Conv(planes, planes * 4, 1), _BN(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.ModuleList([_Conv(inplanes, planes * 4, 1, stride), _BN(planes * 4)])


def _fold(conv, bn, fold=True):
    """(bf16 GEMM weight [Cout, Kp] in (kh,kw,cin) order, fp32 bias [Cout]); fold=False keeps the
    plain convolution weight (batch-statistics mode normalises after the GEMM)."""
    w = conv.weight.detach().float()
    if fold:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        bias = bn.bias.detach().float() - bn.running_mean.float() * scale
        w = w * scale.view(-1, 1, 1, 1)
    else:
        bias = None
    w = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    K = w.shape[1]
    Kp = (K + 7) // 8 * 8
    out = torch.zeros((w.shape[0], Kp), dtype=torch.bfloat16, device=w.device)
    out[:, :K] = w.to(torch.bfloat16)
    return out, (bias.contiguous() if bias is not None else None)


class ResNetFeatureExtractor(nn.Module):
    def __init__(self, layers=(3, 8, 36, 3), num_classes=1000):
        super().__init__()
        self.conv1, self.bn1 = _Conv(3, 64, 7, 2, 3), _BN(64)
        inplanes = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if li == 0 else 2
            blocks = []
            for bi in range(n):
                s = stride if bi == 0 else 1
                blocks.append(_Bottleneck(inplanes, planes, s, bi == 0 and (s != 1 or inplanes != planes * 4)))
                inplanes = planes * 4
            setattr(self, 'layer%d' % (li + 1), nn.ModuleList(blocks))
        self.fc = nn.Linear(512 * 4, num_classes)     # kept for state-dict compatibility, unused
        self._folded = None
        self._plain = None
        self.bn_mode = 'auto'
        self.momentum = 0.1

    def prepare(self, fold=True):
        """bf16 GEMM weights, made once (weights are frozen): BN folded in (running statistics) or
        the plain convolutions (batch statistics)."""
        f = {'stem': _fold(self.conv1, self.bn1, fold)}
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, 'layer%d' % li)):
                key = (li, bi)
                f[key] = [_fold(blk.conv1, blk.bn1, fold), _fold(blk.conv2, blk.bn2, fold),
                          _fold(blk.conv3, blk.bn3, fold),
                          _fold(blk.downsample[0], blk.downsample[1], fold) if blk.downsample else None]
        if fold:
            self._folded = f
        else:
            self._plain = f
            self._n_stats = 2 * sum(m.weight.numel() for m in self.modules() if isinstance(m, _BN))
        return self

    def _apply(self, fn, *a, **k):
        self._folded = None
        self._plain = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)      # in-place copies: derived operands are stale
        self._folded = None
        self._plain = None
        return r

    def uses_batch_stats(self):
        return self.bn_mode == 'batch' or (self.bn_mode == 'auto' and self.training)

    # ------------------------------------------------------------------ one convolution + BN
    def _operands(self, batch_stats):
        if batch_stats:
            if self._plain is None:
                self.prepare(fold=False)
            self._folded = None            # running statistics are about to move
            return self._plain
        if self._folded is None:
            self.prepare()
        return self._folded

    class _Pass:
        """Per-forward state: which BN mode, and (batch mode) the zeroed arena of per-channel sums
        -- ONE memset for all 155 BatchNorm layers of a forward."""

        def __init__(self, net, device, batch_stats):
            self.net, self.batch = net, batch_stats
            self.f = net._operands(batch_stats)
            self.off = 0
            self.arena = (torch.zeros(net._n_stats, dtype=torch.float32, device=device)
                          if batch_stats else None)

        def conv_bn(self, cols, wb, bn, relu=True, residual=None):
            """cols [M,K] bf16 (im2col rows or the NHWC activation itself for 1x1) -> [M,Cout] bf16."""
            w, b = wb
            if not self.batch:
                return ops.gemm_tn(cols, w, bias=b, residual16=residual,
                                   act=ops.ACT_RELU if relu else ops.ACT_NONE, want32=False, want16=True)
            y, st = self.conv_raw(cols, wb)
            return ops.bn_apply_(y, st, bn.weight, bn.bias, bn.eps, residual=residual, relu=relu,
                                 running_mean=bn.running_mean, running_var=bn.running_var,
                                 momentum=self.net.momentum, num_batches_tracked=bn.num_batches_tracked)

        def conv_raw(self, cols, wb):
            """Batch-statistics mode: the raw convolution [M,Cout] bf16 and its per-channel sums
            [2*Cout], accumulated in the GEMM epilogue (col_stats) -- no pass over the output."""
            w, _ = wb
            C = w.shape[0]
            st = self.arena[self.off:self.off + 2 * C]
            self.off = (self.off + 2 * C) % self.arena.numel()
            if cols.shape[0] >= 32:
                y = ops.gemm_tn(cols, w, want32=False, want16=True, col_stats=st)
            else:       # fewer rows than one epilogue warp owns: separate reduction
                y = ops.gemm_tn(cols, w, want32=False, want16=True)
                ops.bn_stats(y, st)
            return y, st

    def _stem(self, ps, image):
        B = image.shape[0]
        w = ps.f['stem'][0]
        cols, Ho, Wo = ops.im2col_nchw_f32(image.contiguous(), 7, 7, 2, 3, w.shape[1])
        x = ps.conv_bn(cols, ps.f['stem'], self.bn1)
        return ops.maxpool3x3s2_nhwc(x.view(B, Ho, Wo, 64))

    def _block(self, ps, x, li, bi):
        blk = getattr(self, 'layer%d' % li)[bi]
        c1, c2, c3, ds = ps.f[(li, bi)]
        Bx, H, W, C = x.shape
        x2 = x.view(Bx * H * W, C)
        planes = c1[0].shape[0]
        if ps.batch:
            # bn1 + ReLU are applied by the 3x3 convolution's im2col while it gathers: the normalised
            # activation between conv1 and conv2 is never written
            o, st = ps.conv_raw(x2, c1)
            bn = blk.bn1
            cols, Ho, Wo = ops.im2col_nhwc_bn(o.view(Bx, H, W, planes), 3, 3, blk.conv2.stride, 1, st,
                                              o.shape[0], bn.weight, bn.bias, bn.eps, bn.running_mean,
                                              bn.running_var, self.momentum, bn.num_batches_tracked)
        else:
            o = ps.conv_bn(x2, c1, blk.bn1)
            cols, Ho, Wo = ops.im2col_nhwc(o.view(Bx, H, W, planes), 3, 3, blk.conv2.stride, 1)
        o = ps.conv_bn(cols, c2, blk.bn2)
        if ds is not None:
            if blk.conv2.stride != 1:
                idc, _, _ = ops.im2col_nhwc(x, 1, 1, blk.conv2.stride, 0)
            else:
                idc = x2
            idn = ps.conv_bn(idc, ds, blk.downsample[1], relu=False)
        else:
            idn = x2
        return ps.conv_bn(o, c3, blk.bn3, residual=idn).view(Bx, Ho, Wo, planes * 4)

    @torch.no_grad()
    def features_nhwc(self, image):
        """image [B,3,H,W] fp32 -> [B, H/32, W/32, 2048] bf16 (NHWC)."""
        ps = self._Pass(self, image.device, self.uses_batch_stats())
        x = self._stem(ps, image)
        for li in range(1, 5):
            for bi in range(len(getattr(self, 'layer%d' % li))):
                x = self._block(ps, x, li, bi)
        return x

    @torch.no_grad()
    def stem_nhwc(self, image):
        """conv1 + bn1 + relu + maxpool alone (tests: per-stage parity)."""
        return self._stem(self._Pass(self, image.device, self.uses_batch_stats()), image)

    @torch.no_grad()
    def block_nhwc(self, x, li, bi):
        """One Bottleneck (layer `li` in 1..4, block `bi`) on an NHWC bf16 activation."""
        return self._block(self._Pass(self, x.device, self.uses_batch_stats()), x.contiguous(), li, bi)

    def forward(self, x, pool=False):
        """resnet.py:92-117 API: [B,3,224,224] -> [B,2048,7,7] fp32 (or pooled [B,2048])."""
        feats = ops.bf16_to_f32(self.features_nhwc(x))
        out = feats.permute(0, 3, 1, 2).contiguous()
        return out.mean(dim=(2, 3)) if pool else out


def resnet152(pretrained=False, **kwargs):
    """resnet.py:184-192.  Pretrained ImageNet weights need the network (unavailable here); load a
    torchvision state dict with load_state_dict() when one is available."""
    if pretrained:
        raise RuntimeError('no network access: load a resnet152 state_dict explicitly')
    return ResNetFeatureExtractor((3, 8, 36, 3), **kwargs)
