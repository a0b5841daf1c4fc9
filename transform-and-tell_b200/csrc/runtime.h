#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tt {
void count_launch(int n);
// Device scalar mixed into every dropout seed (see tt_set_rng_step_ptr); may be null.
const unsigned long long* rng_step_ptr();
// 2-D bf16 tensor map, 128-byte swizzle, zero fill out of bounds.
// inner = contiguous extent (elements), ld_elems = row stride (elements, multiple of 8).
int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows);
// Same with the swizzle width chosen (64 or 128 bytes = the box's inner extent in bytes).
int make_tmap_bf16_2d_sw(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                         uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes);

// Kernel launch helper; TT_PDL=1 turns on programmatic dependent launch (measured neutral under
// CUDA-graph replay, so off by default).
bool pdl_enabled();
void set_pdl(int on);
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace tt
