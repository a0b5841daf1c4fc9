"""Face encoders on the B200 path against golden vectors produced by the
reference's own modules (oracle/gen_golden.py: tell/facenet/inception_resnet_v1.py, mtcnn.py) and
the building-block kernels of convnets.cu against torch.

Tolerances: activations are bf16 (8 mantissa bits) through up to ~70 convolution layers; the
reference is fp32.  Embeddings are unit vectors: 4e-2 of the largest component and cosine > 0.999;
P/R/O-Net probabilities 2e-2 absolute, regressions 4e-2 of the largest magnitude."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'facenet.npz')


def T_(a):
    return torch.from_numpy(np.asarray(a))


def _nhwc16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def test_inception_resnet_v1_matches_reference_golden():
    from tell_b200 import synth
    from tell_b200.facenet import InceptionResnetV1
    g = np.load(GOLD)
    net = InceptionResnetV1(num_classes=10)
    sd = synth.shaped_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed=3)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    emb, logits = net(T_(g['irv1_x']).cuda())
    emb, logits = emb.cpu(), logits.cpu()
    want, want_l = T_(g['irv1_emb']), T_(g['irv1_logits'])
    assert emb.shape == (3, 512) and logits.shape == (3, 10)
    assert (emb.norm(dim=1) - 1).abs().max().item() < 1e-5
    assert (emb - want).abs().max().item() < 4e-2 * want.abs().max().item()
    assert ((emb * want).sum(1)).min().item() > 0.999
    assert (logits - want_l).abs().max().item() < 4e-2 * max(1.0, want_l.abs().max().item())
    with pytest.raises(RuntimeError):
        net.train()(T_(g['irv1_x'][:1]).cuda())


@pytest.mark.parametrize('kind', ['synth', 'real'])
def test_mtcnn_nets_match_reference_golden(kind):
    """Seeded synthetic weights, and the reference's vendored checkpoints (carried in the fixture)."""
    from tell_b200 import synth
    from tell_b200.facenet import ONet, PNet, RNet
    g = np.load(GOLD)
    for name, cls in (('pnet', PNet), ('rnet', RNet), ('onet', ONet)):
        net = cls()
        if kind == 'synth':
            sd = synth.shaped_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed=5)
        else:
            sd = {k: T_(g['%s_real_w.%s' % (name, k)]) for k in net.state_dict()}
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        outs = net(T_(g[name + '_x']).cuda())
        n_out = 2 if name != 'onet' else 3
        assert len(outs) == n_out
        for i, o in enumerate(outs):
            want = T_(g['%s_%s_out%d' % (name, kind, i)])
            o = o.cpu()
            assert o.shape == want.shape, (name, i, o.shape, want.shape)
            is_prob = i == n_out - 1
            tol = 2e-2 if is_prob else 4e-2 * max(1.0, want.abs().max().item())
            assert (o - want).abs().max().item() < tol, (name, kind, i, (o - want).abs().max().item())
            if is_prob:
                assert (o.sum(1) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize('H,W,k,s,p,ceil', [(24, 24, 3, 2, 0, True), (11, 7, 2, 2, 0, True),
                                             (37, 53, 3, 2, 0, False), (10, 10, 3, 2, 1, False),
                                             (9, 9, 2, 2, 0, True)])
def test_maxpool_matches_torch(H, W, k, s, p, ceil):
    from tell_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 16, H, W).to(torch.bfloat16).float()
    want = F.max_pool2d(x, k, s, p, ceil_mode=ceil)
    assert ops.conv_out_size(H, k, s, p, ceil) == want.shape[2]
    assert ops.conv_out_size(W, k, s, p, ceil) == want.shape[3]
    # into a channel slice of a wider buffer, from a channel slice of a wider buffer
    wide_in = torch.zeros(2, H, W, 40, dtype=torch.bfloat16, device='cuda')
    wide_in[..., 8:24] = _nhwc16(x)
    wide_out = torch.full((2, want.shape[2], want.shape[3], 32), 7.0, dtype=torch.bfloat16, device='cuda')
    ops.maxpool_nhwc(wide_in[..., 8:24], k, s, p, ceil, out=wide_out[..., 16:32])
    got = wide_out[..., 16:32].float().cpu().permute(0, 3, 1, 2)
    assert torch.equal(got, want)
    assert (wide_out[..., :16] == 7).all()


@pytest.mark.parametrize('kh,kw,s,ph,pw', [(1, 7, 1, 0, 3), (7, 1, 1, 3, 0), (3, 3, 2, 0, 0),
                                            (1, 3, 1, 0, 1), (2, 2, 1, 0, 0), (1, 1, 1, 0, 0)])
def test_im2col_hw_conv_matches_torch(kh, kw, s, ph, pw):
    """im2col (rectangular kernel, per-axis padding, strided input view) + GEMM == F.conv2d."""
    from tell_b200 import ops
    torch.manual_seed(1)
    B, C, H, W, Co = 2, 16, 13, 9, 24
    x = torch.randn(B, C, H, W).to(torch.bfloat16).float()
    w = (torch.randn(Co, C, kh, kw) * 0.1).to(torch.bfloat16).float()
    want = F.conv2d(x, w, None, s, (ph, pw))
    wide = torch.zeros(B, H, W, 48, dtype=torch.bfloat16, device='cuda')
    wide[..., 16:32] = _nhwc16(x)
    cols, Ho, Wo = ops.im2col_nhwc_hw(wide[..., 16:32], kh, kw, s, ph, pw)
    w2 = w.permute(0, 2, 3, 1).reshape(Co, -1).to(torch.bfloat16).cuda()
    got = ops.gemm_tn(cols, w2).view(B, Ho, Wo, Co).permute(0, 3, 1, 2).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())


def test_prelu_l2norm_softmax2_avgpool():
    from tell_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(37, 24).to(torch.bfloat16)
    slope = torch.rand(24)
    want = F.prelu(x.float(), slope).to(torch.bfloat16)
    wide = torch.zeros(37, 40, dtype=torch.bfloat16, device='cuda')
    wide[:, 8:32] = x.cuda()
    ops.prelu_bf16_(wide[:, 8:32], slope.cuda())
    assert torch.equal(wide[:, 8:32].cpu(), want) and (wide[:, :8] == 0).all()
    v = torch.randn(9, 512)
    assert (ops.l2norm_rows(v.cuda()).cpu() - F.normalize(v, p=2, dim=1)).abs().max().item() < 1e-6
    z = torch.randn(50, 8)
    got = ops.softmax2_(z.clone().cuda(), 0).cpu()
    assert (got[:, :2] - F.softmax(z[:, :2], dim=1)).abs().max().item() < 1e-6
    assert torch.equal(got[:, 2:], z[:, 2:])
    a = torch.randn(3, 5, 5, 32).to(torch.bfloat16)
    assert (ops.avgpool_nhwc(a.cuda()).cpu() - a.float().mean(dim=(1, 2))).abs().max().item() < 1e-5
