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
  'batch'    always batch statistics: GEMM writes the raw convolution, `tt_bn_stats_bf16` reduces
             per-channel sums, `tt_bn_apply_bf16` normalises (+ identity, ReLU) in place and updates
             running_mean / running_var / num_batches_tracked.
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

    @torch.no_grad()
    def features_nhwc(self, image):
        """image [B,3,H,W] fp32 -> [B, H/32, W/32, 2048] bf16 (NHWC)."""
        if self.uses_batch_stats():
            return self._features_batch_stats(image)
        if self._folded is None:
            self.prepare()
        f = self._folded
        B = image.shape[0]
        w, b = f['stem']
        cols, Ho, Wo = ops.im2col_nchw_f32(image.contiguous(), 7, 7, 2, 3, w.shape[1])
        x = ops.gemm_tn(cols, w, bias=b, act=ops.ACT_RELU, want32=False, want16=True)
        x = ops.maxpool3x3s2_nhwc(x.view(B, Ho, Wo, 64))
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, 'layer%d' % li)):
                (w1, b1), (w2, b2), (w3, b3), ds = f[(li, bi)]
                Bx, H, W, C = x.shape
                x2 = x.view(Bx * H * W, C)
                o = ops.gemm_tn(x2, w1, bias=b1, act=ops.ACT_RELU, want32=False, want16=True)
                planes = w1.shape[0]
                cols, Ho, Wo = ops.im2col_nhwc(o.view(Bx, H, W, planes), 3, 3, blk.conv2.stride, 1)
                o = ops.gemm_tn(cols, w2, bias=b2, act=ops.ACT_RELU, want32=False, want16=True)
                if ds is not None:
                    if blk.conv2.stride != 1:
                        idc, _, _ = ops.im2col_nhwc(x, 1, 1, blk.conv2.stride, 0)
                    else:
                        idc = x2
                    idn = ops.gemm_tn(idc, ds[0], bias=ds[1], want32=False, want16=True)
                else:
                    idn = x2
                x = ops.gemm_tn(o, w3, bias=b3, residual16=idn, act=ops.ACT_RELU, want32=False,
                                want16=True).view(Bx, Ho, Wo, planes * 4)
        return x

    def _features_batch_stats(self, image):
        """Train-mode forward (batch statistics; running statistics move, as in the reference's
        training step).  The folded running-statistics operands become stale and are dropped."""
        if self._plain is None:
            self.prepare(fold=False)
        self._folded = None
        f = self._plain
        B = image.shape[0]
        arena = torch.zeros(self._n_stats, dtype=torch.float32, device=image.device)
        off = [0]

        def bn_(y2d, bn, residual=None, relu=True):
            C = y2d.shape[1]
            st = arena[off[0]:off[0] + 2 * C]
            off[0] += 2 * C
            ops.bn_stats(y2d, st)
            return ops.bn_apply_(y2d, st, bn.weight, bn.bias, bn.eps, residual=residual, relu=relu,
                                 running_mean=bn.running_mean, running_var=bn.running_var,
                                 momentum=self.momentum, num_batches_tracked=bn.num_batches_tracked)

        w, _ = f['stem']
        cols, Ho, Wo = ops.im2col_nchw_f32(image.contiguous(), 7, 7, 2, 3, w.shape[1])
        x = bn_(ops.gemm_tn(cols, w, want32=False, want16=True), self.bn1)
        x = ops.maxpool3x3s2_nhwc(x.view(B, Ho, Wo, 64))
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, 'layer%d' % li)):
                (w1, _), (w2, _), (w3, _), ds = f[(li, bi)]
                Bx, H, W, C = x.shape
                x2 = x.view(Bx * H * W, C)
                o = bn_(ops.gemm_tn(x2, w1, want32=False, want16=True), blk.bn1)
                planes = w1.shape[0]
                cols, Ho, Wo = ops.im2col_nhwc(o.view(Bx, H, W, planes), 3, 3, blk.conv2.stride, 1)
                o = bn_(ops.gemm_tn(cols, w2, want32=False, want16=True), blk.bn2)
                if ds is not None:
                    if blk.conv2.stride != 1:
                        idc, _, _ = ops.im2col_nhwc(x, 1, 1, blk.conv2.stride, 0)
                    else:
                        idc = x2
                    idn = bn_(ops.gemm_tn(idc, ds[0], want32=False, want16=True), blk.downsample[1],
                              relu=False)
                else:
                    idn = x2
                x = bn_(ops.gemm_tn(o, w3, want32=False, want16=True), blk.bn3,
                        residual=idn).view(Bx, Ho, Wo, planes * 4)
        return x

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
