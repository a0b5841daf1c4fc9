"""Staged (TMA store / TMA residual) epilogue vs the register epilogue on the bf16-output GEMM
signatures of the frozen encoders, timed from a CUDA graph (16 launches, rotating operand sets).
    python tools/gemm_staged.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
from tell_b200 import _lib, ops  # noqa: E402

# (M, N, K, bias, act, residual16, label)
SIGS = [(8192, 3072, 1024, 1, 0, 0, 'roberta qkv'), (8192, 1024, 1024, 1, 0, 1, 'roberta out_proj'),
        (8192, 4096, 1024, 1, 2, 0, 'roberta fc1 gelu'), (8192, 1024, 4096, 1, 0, 1, 'roberta fc2'),
        (5957, 3072, 1024, 1, 0, 0, 'packed qkv'), (5957, 4096, 1024, 1, 2, 0, 'packed fc1'),
        (8192, 8192, 1024, 1, 0, 0, 'article k|v x4 layers'),
        (3136, 256, 1024, 1, 1, 0, 'resnet l3 conv1'), (3136, 256, 2304, 1, 1, 0, 'resnet l3 conv2'),
        (3136, 1024, 256, 1, 1, 1, 'resnet l3 conv3+id'), (12544, 128, 512, 1, 1, 0, 'resnet l2 conv1'),
        (12544, 512, 128, 1, 1, 1, 'resnet l2 conv3+id'), (50176, 64, 256, 1, 1, 0, 'resnet l1 conv1'),
        (50176, 256, 64, 1, 1, 1, 'resnet l1 conv3+id'), (784, 2048, 512, 1, 1, 1, 'resnet l4 conv3+id'),
        (200704, 64, 152, 1, 1, 0, 'resnet stem')]


def timeit(M, N, K, bias, act, res):
    nb = 4
    A = [(torch.randn(M, K, device='cuda') / 8).bfloat16() for _ in range(nb)]
    W = [(torch.randn(N, K, device='cuda') / 8).bfloat16() for _ in range(nb)]
    O = [torch.empty(M, N, device='cuda', dtype=torch.bfloat16) for _ in range(nb)]
    b = torch.randn(N, device='cuda') if bias else None
    R = [torch.randn(M, N, device='cuda').bfloat16() for _ in range(nb)] if res else None

    def one(i):
        ops.gemm_tn(A[i % nb], W[i % nb], bias=b, act=act, residual16=R[i % nb] if R else None,
                    out16=O[i % nb], want32=False)
    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            one(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 80 * 1e3)
    return best, O[0].clone()


for M, N, K, bias, act, res, label in SIGS:
    row = []
    outs = []
    for on in (0, 1):
        _lib.lib().tt_gemm_set_staged_epilogue(on)
        torch.manual_seed(1)
        us, o = timeit(M, N, K, bias, act, res)
        row.append(us)
        outs.append(o)
    same = torch.equal(outs[0], outs[1])
    fl = 2.0 * M * N * K
    print('%-24s M=%6d N=%5d K=%5d  register %7.2f us (%6.1f TF)  staged %7.2f us (%6.1f TF)  x%.2f  %s'
          % (label, M, N, K, row[0], fl / row[0] / 1e6, row[1], fl / row[1] / 1e6, row[0] / row[1],
             'bit-identical' if same else 'DIFFERENT'), flush=True)
