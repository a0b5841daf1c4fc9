"""Which tile configuration is fastest for each GEMM signature of the step?  Every signature is
timed (CUDA graph, 16 launches, rotating operands) under each forced configuration; the tile
choice is read from the environment once per process, so each configuration is a subprocess.
    python tools/gemm_tile_sweep.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.path.join(%r, 'transform-and-tell_b200'))
from tell_b200 import ops
SIGS = %r
tag = os.environ.get('CFG_TAG')
for (M, N, K, o16, bias, act, res, stats) in SIGS:
    nb = 3
    A = [(torch.randn(M, K, device='cuda') / 8).bfloat16() for _ in range(nb)]
    W = [(torch.randn(N, K, device='cuda') / 8).bfloat16() for _ in range(nb)]
    O = [torch.empty(M, N, device='cuda', dtype=torch.bfloat16 if o16 else torch.float32) for _ in range(nb)]
    b = torch.randn(N, device='cuda') if bias else None
    R = [torch.randn(M, N, device='cuda').bfloat16() for _ in range(nb)] if res else None
    st = torch.zeros(2 * N, device='cuda') if stats else None
    def one(i):
        kw = dict(out16=O[i %% nb], want32=False) if o16 else dict(out=O[i %% nb])
        ops.gemm_tn(A[i %% nb], W[i %% nb], bias=b, act=act, residual16=R[i %% nb] if R else None, col_stats=st, **kw)
    try:
        for i in range(3): one(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(16): one(i)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5): g.replay()
            e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) / 80 * 1e3)
        print('%%s|%%d,%%d,%%d,%%d,%%d,%%d,%%d,%%d|%%.2f' %% (tag, M, N, K, o16, bias, act, res, stats, best), flush=True)
    except Exception as ex:
        print('%%s|%%d,%%d,%%d,%%d,%%d,%%d,%%d,%%d|nan' %% (tag, M, N, K, o16, bias, act, res, stats), flush=True)
'''
# (M, N, K, bf16 out, bias, act, residual16, col_stats)
SIGS = [
    # ResNet-152, batch 16, running-statistics mode (bias + ReLU [+ identity])
    (3136, 256, 1024, 1, 1, 1, 0, 0), (3136, 256, 2304, 1, 1, 1, 0, 0), (3136, 1024, 256, 1, 1, 1, 1, 0),
    (12544, 128, 512, 1, 1, 1, 0, 0), (12544, 128, 1152, 1, 1, 1, 0, 0), (12544, 512, 128, 1, 1, 1, 1, 0),
    (50176, 64, 256, 1, 1, 1, 0, 0), (50176, 64, 576, 1, 1, 1, 0, 0), (50176, 256, 64, 1, 1, 1, 1, 0),
    (784, 512, 2048, 1, 1, 1, 0, 0), (784, 512, 4608, 1, 1, 1, 0, 0), (784, 2048, 512, 1, 1, 1, 1, 0),
    (200704, 64, 152, 1, 1, 1, 0, 0),
    # batch-statistics mode (raw output + column sums)
    (3136, 256, 1024, 1, 0, 0, 0, 1), (3136, 256, 2304, 1, 0, 0, 0, 1), (3136, 1024, 256, 1, 0, 0, 0, 1),
    (12544, 512, 128, 1, 0, 0, 0, 1), (50176, 256, 64, 1, 0, 0, 0, 1),
    # RoBERTa (packed 5957 rows approximated by M) and the decoder's big K|V projection
    (5957, 1024, 1024, 1, 1, 0, 1, 0), (5957, 3072, 1024, 1, 1, 0, 0, 0), (5957, 4096, 1024, 1, 1, 2, 0, 0),
    (5957, 1024, 4096, 1, 1, 0, 1, 0),
    # decoder (fp32 out)
    (800, 1024, 1024, 0, 1, 0, 0, 0), (800, 2048, 1024, 0, 1, 0, 0, 0), (800, 4096, 1024, 0, 1, 0, 0, 0),
    (800, 1024, 4096, 0, 1, 0, 0, 0), (800, 5002, 1024, 0, 0, 0, 0, 0), (256, 1024, 1024, 0, 1, 0, 0, 0),
    (256, 4096, 1024, 0, 1, 0, 0, 0), (256, 1024, 4096, 0, 1, 0, 0, 0),
]
CFGS = [('auto', {}), ('pair256', {'TT_GEMM2_BN': '256'}), ('pair128', {'TT_GEMM2_BN': '128'}),
        ('one256', {'TT_GEMM_2CTA': '0', 'TT_GEMM_BN': '256'}), ('one128', {'TT_GEMM_2CTA': '0', 'TT_GEMM_BN': '128'}),
        ('one64', {'TT_GEMM_2CTA': '0', 'TT_GEMM_BN': '64'}), ('one32', {'TT_GEMM_2CTA': '0', 'TT_GEMM_BN': '32'})]
res = {}
for tag, envx in CFGS:
    env = dict(os.environ, CFG_TAG=tag, **envx)
    r = subprocess.run([sys.executable, '-c', CHILD % (ROOT, SIGS)], env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if '|' in line:
            t, sig, us = line.split('|')
            res.setdefault(sig, {})[t] = float(us)
print('%-44s' % 'M,N,K,o16,bias,act,res,stats' + ''.join('%9s' % c[0] for c in CFGS) + '   best')
for sig in res:
    row = res[sig]
    best = min((v, k) for k, v in row.items() if v == v)
    print('%-44s' % sig + ''.join('%9.2f' % row.get(c[0], float('nan')) for c in CFGS) + '   %s (%.0f%% of auto)'
          % (best[1], 100 * best[0] / row['auto']))
