"""Time a few GEMM shapes under every tile width (TT_GEMM_BN is read once per process, so each
configuration runs in a subprocess)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.path.join(%r, 'transform-and-tell_b200'))
from tell_b200 import ops
for (M, N, K) in %r:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
    o = torch.empty(M, N, device='cuda')
    for _ in range(3): ops.gemm_tn(a, w, out=o)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): ops.gemm_tn(a, w, out=o)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 100 * 1e3
    print('BN=%%s M=%%5d N=%%5d K=%%5d  %%7.2f us  %%7.1f TF' %% (os.environ.get('TT_GEMM_BN','auto'), M, N, K, us, 2.0*M*N*K/us/1e6))
'''
shapes = [(800, 1024, 1024), (800, 1024, 4096), (800, 4096, 1024), (1024, 1024, 800),
          (1024, 1024, 8192), (3136, 256, 2304), (8192, 1024, 1024), (8192, 4096, 1024), (128, 128, 4096), (128, 256, 64)]
for bn in ['auto', '32', '64', '128', '256']:
    env = dict(os.environ)
    if bn != 'auto':
        env['TT_GEMM_BN'] = bn
    subprocess.run([sys.executable, '-c', CHILD % (ROOT, shapes)], env=env)
