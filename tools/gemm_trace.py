"""Per-CTA timeline of the 1-CTA tcgen05 GEMM (tt_gemm_set_trace): where do the microseconds of a
small GEMM go?   python tools/gemm_trace.py [M N K]..."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import _lib, ops  # noqa: E402

NAMES = ['entry', 'setup', 'tma0', 'land0', 'mmaN', 'accrdy', 'epi', 'exit']
RESNET = '--resnet' in sys.argv      # bf16 output, bias, ReLU, bf16 residual (ResNet conv3 epilogue)
args = [int(x) for x in sys.argv[1:] if x != '--resnet']
shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [
    (800, 112, 112), (800, 1024, 1024), (800, 1024, 4096), (800, 4096, 1024), (3136, 256, 2304)]
lib = _lib.lib()
lib.tt_gemm_set_trace.argtypes = [ctypes.c_void_p]
lib.tt_gemm_set_trace.restype = None
for (M, N, K) in shapes:
    a = torch.randn(M, K, device='cuda').bfloat16()
    w = torch.randn(N, K, device='cuda').bfloat16()
    o = torch.empty(M, N, device='cuda')
    if RESNET:
        o16 = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
        bias = torch.randn(N, device='cuda')
        res = torch.randn(M, N, device='cuda').bfloat16()
        _plain = ops.gemm_tn

        def _resnet(a_, w_, out=None):
            return _plain(a_, w_, bias=bias, residual16=res, act=ops.ACT_RELU, out16=o16, want32=False,
                          want16=True)
        gemm = _resnet
    else:
        gemm = ops.gemm_tn
    tr = torch.zeros(148 * 8, dtype=torch.int64, device='cuda')
    for _ in range(3):
        gemm(a, w, out=o)
    torch.cuda.synchronize()
    lib.tt_gemm_set_trace(ctypes.c_void_p(tr.data_ptr()))
    gemm(a, w, out=o)
    gemm(a, w, out=o)   # second launch: warm, back-to-back
    torch.cuda.synchronize()
    lib.tt_gemm_set_trace(ctypes.c_void_p(0))
    t = tr.view(148, 8).cpu()
    t = t[t[:, 0] > 0]
    if t.shape[0] == 0:
        print('M=%d N=%d K=%d went to the 2-CTA kernel (no trace)' % (M, N, K))
        continue
    t0 = t[:, 0].min()
    rel = (t - t0).float() / 1e3
    print('M=%d N=%d K=%d  ctas=%d' % (M, N, K, t.shape[0]))
    for i, n in enumerate(NAMES):
        c = rel[:, i]
        print('   %-7s min %6.2f  med %6.2f  max %6.2f us' % (n, c.min(), c.median(), c.max()))
