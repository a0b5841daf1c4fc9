#pragma once
#include <cuda.h>
#include <stdint.h>

namespace tt {
void count_launch(int n);
// Device scalar mixed into every dropout seed (see tt_set_rng_step_ptr); may be null.
const unsigned long long* rng_step_ptr();
// 2-D bf16 tensor map, 128-byte swizzle, zero fill out of bounds.
// inner = contiguous extent (elements), ld_elems = row stride (elements, multiple of 8).
int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows);
}  // namespace tt
