"""MTCNN proposal / refine / output networks -- tell/facenet/mtcnn.py:11-159 -- on the B200 kernels.

Only the three convolutional networks are here (they are the device work of face detection); the
image pyramid, box regression and NMS around them (tell/facenet/utils/detect_face.py) are host-side
numpy in the reference and stay out of scope.  Same module tree and state-dict keys
as the reference (conv1.weight, prelu1.weight, dense4.weight, ...), so the vendored
data/{pnet,rnet,onet}.pt checkpoints load with strict=True.

Layout: NHWC bf16 with channel counts padded to multiples of 8 (10->16, 28->32: the padded weights
and PReLU slopes are zero, so the padded channels stay exactly 0).  conv = im2col + tcgen05 GEMM
(+bias in the epilogue), then an in-place PReLU, ceil-mode max-pools; the dense heads are one GEMM
whose fp32 output holds [prob logits | box | landmarks] side by side, softmax on the first two
columns.  The reference flattens NCHW as (w, h, c) before the dense layer (:95, :148
`x.permute(0, 3, 2, 1)`); the dense weight's columns are permuted once at load time to the NHWC
(h, w, c) order instead of moving activations."""
import torch
import torch.nn as nn

from .. import ops


def _pad8(n):
    return (n + 7) // 8 * 8


def _fold_conv(conv, cin_pad):
    """nn.Conv2d -> (bf16 [Cout_pad, Kp] in (kh,kw,cin_pad) order, fp32 bias [Cout_pad])."""
    w = conv.weight.detach().float()
    cout, cin, kh, kw = w.shape
    wp = torch.zeros((_pad8(cout), kh, kw, cin_pad), device=w.device)
    wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
    wp = wp.reshape(_pad8(cout), -1)
    out = torch.zeros((wp.shape[0], _pad8(wp.shape[1])), dtype=torch.bfloat16, device=w.device)
    out[:, :wp.shape[1]] = wp.to(torch.bfloat16)
    b = torch.zeros(_pad8(cout), device=w.device)
    b[:cout] = conv.bias.detach().float()
    return out, b


def _slope(prelu):
    s = prelu.weight.detach().float()
    out = torch.zeros(_pad8(s.numel()), device=s.device)
    out[:s.numel()] = s
    return out


class _Net(nn.Module):
    def _apply(self, fn, *a, **k):
        self._f = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)      # in-place copies: folded operands are stale
        self._f = None
        return r

    def _conv(self, x, folded, slope, k):
        """x [B,H,W,Cp] bf16 -> PReLU(conv_kxk(x)) [B,H-k+1,W-k+1,Coutp] bf16."""
        w, b = folded
        B = x.shape[0]
        cols, Ho, Wo = ops.im2col_nhwc_hw(x, k, k, 1, 0, 0)
        y = ops.gemm_tn(cols, w, bias=b, want32=False, want16=True)
        ops.prelu_bf16_(y, slope)
        return y.view(B, Ho, Wo, w.shape[0])

    def _stem(self, x, folded, slope, k):
        """NCHW fp32 image batch -> first conv + PReLU, NHWC bf16."""
        w, b = folded
        B = x.shape[0]
        cols, Ho, Wo = ops.im2col_nchw_f32(x.contiguous(), k, k, 1, 0, w.shape[1])
        y = ops.gemm_tn(cols, w, bias=b, want32=False, want16=True)
        ops.prelu_bf16_(y, slope)
        return y.view(B, Ho, Wo, w.shape[0])

    @staticmethod
    def _dense_weight(lin, C, Hh, Ww, Cp):
        """dense weight over the reference's (w, h, c) flatten -> bf16 [out, Hh*Ww*Cp] over (h, w, c)."""
        w = lin.weight.detach().float().view(-1, Ww, Hh, C).permute(0, 2, 1, 3)       # [out,h,w,c]
        wp = torch.zeros((w.shape[0], Hh, Ww, Cp), device=w.device)
        wp[..., :C] = w
        return wp.reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()

    @staticmethod
    def _heads(lins, device):
        """Several nn.Linear heads on the same input -> one bf16 operand [sum_out (padded to 8), in]."""
        w = torch.cat([l.weight.detach().float() for l in lins], 0)
        b = torch.cat([l.bias.detach().float() for l in lins])
        n = _pad8(w.shape[0])
        wp = torch.zeros((n, w.shape[1]), device=device)
        wp[:w.shape[0]] = w
        bp = torch.zeros(n, device=device)
        bp[:b.numel()] = b
        return wp.to(torch.bfloat16).contiguous(), bp


class PNet(_Net):
    """mtcnn.py:11-51.  Fully convolutional: x [B,3,H,W] -> (box regression [B,4,H',W'],
    face probability [B,2,H',W']), fp32, NCHW like the reference."""

    def __init__(self, pretrained=False):
        super().__init__()
        if pretrained:
            raise RuntimeError('load the vendored pnet.pt with load_state_dict(); no implicit file access')
        self.conv1 = nn.Conv2d(3, 10, kernel_size=3)
        self.prelu1 = nn.PReLU(10)
        self.pool1 = nn.MaxPool2d(2, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(10, 16, kernel_size=3)
        self.prelu2 = nn.PReLU(16)
        self.conv3 = nn.Conv2d(16, 32, kernel_size=3)
        self.prelu3 = nn.PReLU(32)
        self.conv4_1 = nn.Conv2d(32, 2, kernel_size=1)
        self.softmax4_1 = nn.Softmax(dim=1)
        self.conv4_2 = nn.Conv2d(32, 4, kernel_size=1)
        self.training = False
        self._f = None

    def prepare(self):
        dev = self.conv1.weight.device
        head_w = torch.cat([self.conv4_1.weight, self.conv4_2.weight], 0).detach().float().view(6, 32)
        hw = torch.zeros((8, 32), device=dev)
        hw[:6] = head_w
        hb = torch.zeros(8, device=dev)
        hb[:6] = torch.cat([self.conv4_1.bias, self.conv4_2.bias]).detach().float()
        self._f = dict(c1=_fold_conv(self.conv1, 3), s1=_slope(self.prelu1),
                       c2=_fold_conv(self.conv2, 16), s2=_slope(self.prelu2),
                       c3=_fold_conv(self.conv3, 16), s3=_slope(self.prelu3),
                       head=(hw.to(torch.bfloat16).contiguous(), hb))
        return self

    @torch.no_grad()
    def forward(self, x):
        if self._f is None:
            self.prepare()
        f = self._f
        h = self._stem(x, f['c1'], f['s1'], 3)
        h = ops.maxpool_nhwc(h, 2, 2, ceil_mode=True)
        h = self._conv(h, f['c2'], f['s2'], 3)
        h = self._conv(h, f['c3'], f['s3'], 3)
        B, Ho, Wo, C = h.shape
        out = ops.gemm_tn(h.view(B * Ho * Wo, C), f['head'][0], bias=f['head'][1])     # [rows, 8] fp32
        ops.softmax2_(out, 0)
        out = out.view(B, Ho, Wo, 8)
        a = out[..., 0:2].permute(0, 3, 1, 2).contiguous()
        b = out[..., 2:6].permute(0, 3, 1, 2).contiguous()
        return b, a


class RNet(_Net):
    """mtcnn.py:54-101: x [B,3,24,24] -> (box [B,4], prob [B,2])."""

    def __init__(self, pretrained=False):
        super().__init__()
        if pretrained:
            raise RuntimeError('load the vendored rnet.pt with load_state_dict(); no implicit file access')
        self.conv1 = nn.Conv2d(3, 28, kernel_size=3)
        self.prelu1 = nn.PReLU(28)
        self.pool1 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(28, 48, kernel_size=3)
        self.prelu2 = nn.PReLU(48)
        self.pool2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv3 = nn.Conv2d(48, 64, kernel_size=2)
        self.prelu3 = nn.PReLU(64)
        self.dense4 = nn.Linear(576, 128)
        self.prelu4 = nn.PReLU(128)
        self.dense5_1 = nn.Linear(128, 2)
        self.softmax5_1 = nn.Softmax(dim=1)
        self.dense5_2 = nn.Linear(128, 4)
        self.training = False
        self._f = None

    def prepare(self):
        dev = self.conv1.weight.device
        self._f = dict(c1=_fold_conv(self.conv1, 3), s1=_slope(self.prelu1),
                       c2=_fold_conv(self.conv2, 32), s2=_slope(self.prelu2),
                       c3=_fold_conv(self.conv3, 48), s3=_slope(self.prelu3),
                       d4=(self._dense_weight(self.dense4, 64, 3, 3, 64),
                           self.dense4.bias.detach().float().contiguous()),
                       s4=_slope(self.prelu4),
                       head=self._heads([self.dense5_1, self.dense5_2], dev))
        return self

    @torch.no_grad()
    def forward(self, x):
        if self._f is None:
            self.prepare()
        f = self._f
        B = x.shape[0]
        h = self._stem(x, f['c1'], f['s1'], 3)
        h = ops.maxpool_nhwc(h, 3, 2, ceil_mode=True)
        h = self._conv(h, f['c2'], f['s2'], 3)
        h = ops.maxpool_nhwc(h, 3, 2, ceil_mode=True)
        h = self._conv(h, f['c3'], f['s3'], 2)
        if h.shape[1:] != (3, 3, 64):
            raise ValueError('RNet expects 24x24 crops (got feature map %s)' % (tuple(h.shape),))
        d = ops.gemm_tn(h.view(B, -1), f['d4'][0], bias=f['d4'][1], want32=False, want16=True)
        ops.prelu_bf16_(d, f['s4'])
        out = ops.gemm_tn(d, f['head'][0], bias=f['head'][1])                          # [B, 8] fp32
        ops.softmax2_(out, 0)
        return out[:, 2:6].contiguous(), out[:, 0:2].contiguous()


class ONet(_Net):
    """mtcnn.py:104-159: x [B,3,48,48] -> (box [B,4], landmarks [B,10], prob [B,2])."""

    def __init__(self, pretrained=False):
        super().__init__()
        if pretrained:
            raise RuntimeError('load the vendored onet.pt with load_state_dict(); no implicit file access')
        self.conv1 = nn.Conv2d(3, 32, kernel_size=3)
        self.prelu1 = nn.PReLU(32)
        self.pool1 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3)
        self.prelu2 = nn.PReLU(64)
        self.pool2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=3)
        self.prelu3 = nn.PReLU(64)
        self.pool3 = nn.MaxPool2d(2, 2, ceil_mode=True)
        self.conv4 = nn.Conv2d(64, 128, kernel_size=2)
        self.prelu4 = nn.PReLU(128)
        self.dense5 = nn.Linear(1152, 256)
        self.prelu5 = nn.PReLU(256)
        self.dense6_1 = nn.Linear(256, 2)
        self.softmax6_1 = nn.Softmax(dim=1)
        self.dense6_2 = nn.Linear(256, 4)
        self.dense6_3 = nn.Linear(256, 10)
        self.training = False
        self._f = None

    def prepare(self):
        dev = self.conv1.weight.device
        self._f = dict(c1=_fold_conv(self.conv1, 3), s1=_slope(self.prelu1),
                       c2=_fold_conv(self.conv2, 32), s2=_slope(self.prelu2),
                       c3=_fold_conv(self.conv3, 64), s3=_slope(self.prelu3),
                       c4=_fold_conv(self.conv4, 64), s4=_slope(self.prelu4),
                       d5=(self._dense_weight(self.dense5, 128, 3, 3, 128),
                           self.dense5.bias.detach().float().contiguous()),
                       s5=_slope(self.prelu5),
                       head=self._heads([self.dense6_1, self.dense6_2, self.dense6_3], dev))
        return self

    @torch.no_grad()
    def forward(self, x):
        if self._f is None:
            self.prepare()
        f = self._f
        B = x.shape[0]
        h = self._stem(x, f['c1'], f['s1'], 3)
        h = ops.maxpool_nhwc(h, 3, 2, ceil_mode=True)
        h = self._conv(h, f['c2'], f['s2'], 3)
        h = ops.maxpool_nhwc(h, 3, 2, ceil_mode=True)
        h = self._conv(h, f['c3'], f['s3'], 3)
        h = ops.maxpool_nhwc(h, 2, 2, ceil_mode=True)
        h = self._conv(h, f['c4'], f['s4'], 2)
        if h.shape[1:] != (3, 3, 128):
            raise ValueError('ONet expects 48x48 crops (got feature map %s)' % (tuple(h.shape),))
        d = ops.gemm_tn(h.view(B, -1), f['d5'][0], bias=f['d5'][1], want32=False, want16=True)
        ops.prelu_bf16_(d, f['s5'])
        out = ops.gemm_tn(d, f['head'][0], bias=f['head'][1])                          # [B, 16] fp32
        ops.softmax2_(out, 0)
        return out[:, 2:6].contiguous(), out[:, 6:16].contiguous(), out[:, 0:2].contiguous()
