#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small-shape tests of the hand-written
# mbarrier / TMEM / TMA pipelines (GEMM incl. staged epilogue, tcgen05 flash attention, cross
# attention, DynamicConv, train-mode BatchNorm).  Logs -> gpurun_out/sanitizer_*.log
#   gpurun -- bash tools/sanitize.sh
mkdir -p gpurun_out
SEL_GEMM='test_gemm_plain and (33-40-8 or 128-128-64 or 128-256-128 or 100-48-1024) or staged_epilogue_bit and (777 or 40-32) or column_statistics and (45-64 or 98-2048)'
SEL_OPS='test_flash_attention_tcgen05 or test_attention_tensor_core_path or test_dynconv_fwd_bwd or test_im2col_with_fused or test_layernorm_fwd_bwd'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_gemm_gpu.py -q -x -k "$SEL_GEMM" > gpurun_out/sanitizer_${tool}_gemm.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_ops_gpu.py -q -x -k "$SEL_OPS" > gpurun_out/sanitizer_${tool}_ops.log 2>&1
done
# round-2 additions: operand-twin kernels (twin.cu), multi-tile dK|dV, the warp-specialised flash kernel is in SEL_OPS
SEL_TWIN='test_elementwise_twins or test_layernorm_multi or test_dynconv_twins'
SEL_DKV='test_attention_dkv_tiles_per_cta'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_twin_gpu.py -q -x -k "$SEL_TWIN" > gpurun_out/sanitizer_${tool}_twin.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_ops_gpu.py -q -x -k "$SEL_DKV" > gpurun_out/sanitizer_${tool}_dkv.log 2>&1
done
for f in gpurun_out/sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $f | tail -4; done
