#!/bin/bash
# ncu --set full captures of the round-2 kernels (one GPU):   gpurun -- bash tools/ncu_r2.sh
mkdir -p gpurun_out
cat > /tmp/_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.environ['GRAFT_REPO_ROOT'] if 'GRAFT_REPO_ROOT' in os.environ else '.', 'transform-and-tell_b200'))
from tell_b200 import ops
what = sys.argv[1]
if what == 'fc1':       # RoBERTa fc1: 8192 x 4096 x 1024, bias + GELU, bf16 out (CTA-pair kernel, staged epilogue)
    a = (torch.randn(8192, 1024, device='cuda') / 8).bfloat16(); w = (torch.randn(4096, 1024, device='cuda') / 8).bfloat16()
    b = torch.randn(4096, device='cuda'); o = torch.empty(8192, 4096, device='cuda', dtype=torch.bfloat16)
    f = lambda: ops.gemm_tn(a, w, bias=b, act=ops.ACT_GELU, out16=o, want32=False)
elif what == 'dec':     # decoder GEMM: 800 x 1024 x 1024, bias, fp32 out (single-CTA kernel, register epilogue)
    a = (torch.randn(800, 1024, device='cuda') / 8).bfloat16(); w = (torch.randn(1024, 1024, device='cuda') / 8).bfloat16()
    b = torch.randn(1024, device='cuda'); o = torch.empty(800, 1024, device='cuda')
    f = lambda: ops.gemm_tn(a, w, bias=b, out=o)
elif what == 'dec4096': # decoder fc1: 800 x 4096 x 1024
    a = (torch.randn(800, 1024, device='cuda') / 8).bfloat16(); w = (torch.randn(4096, 1024, device='cuda') / 8).bfloat16()
    b = torch.randn(4096, device='cuda'); o = torch.empty(800, 4096, device='cuda')
    f = lambda: ops.gemm_tn(a, w, bias=b, act=ops.ACT_RELU, out=o)
elif what == 'conv3':   # ResNet stage-3 expanding 1x1 conv + identity + ReLU
    a = (torch.randn(3136, 256, device='cuda') / 8).bfloat16(); w = (torch.randn(1024, 256, device='cuda') / 8).bfloat16()
    b = torch.randn(1024, device='cuda'); r = torch.randn(3136, 1024, device='cuda').bfloat16(); o = torch.empty(3136, 1024, device='cuda', dtype=torch.bfloat16)
    f = lambda: ops.gemm_tn(a, w, bias=b, residual16=r, act=ops.ACT_RELU, out16=o, want32=False)
elif what == 'bn':      # train-mode BatchNorm pieces at the stage-2 size
    x = torch.randn(12544, 512, device='cuda').bfloat16(); res = torch.randn(12544, 512, device='cuda').bfloat16()
    st = torch.zeros(1024, device='cuda'); g = torch.ones(512, device='cuda'); be = torch.zeros(512, device='cuda')
    ops.bn_stats(x, st)
    x4 = torch.randn(16, 28, 28, 128, device='cuda').bfloat16(); st4 = torch.zeros(256, device='cuda'); ops.bn_stats(x4.view(-1, 128), st4)
    g4 = torch.ones(128, device='cuda'); b4 = torch.zeros(128, device='cuda')
    def f():
        ops.bn_apply_(x, st, g, be, 1e-5, residual=res, relu=True)
        ops.im2col_nhwc_bn(x4, 3, 3, 1, 1, st4, 12544, g4, b4, 1e-5)
for _ in range(4):
    f()
torch.cuda.synchronize()
PY
run() {   # name  kernel-regex  arg
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/r2_$1 python /tmp/_one.py $3 > gpurun_out/ncu_$1.log 2>&1
}
run gemm2_fc1 gemm2_bf16_kernel fc1
run gemm_dec_800x1024x1024 gemm_bf16_tn_kernel dec
run gemm_dec_800x4096x1024 gemm_bf16_tn_kernel dec4096
run gemm_conv3 'gemm' conv3
run bn_apply bn_apply_kernel bn
run im2col_bn im2col_nhwc_bn_rows_kernel bn
ls -la gpurun_out/*.ncu-rep
