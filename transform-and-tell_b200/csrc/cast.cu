// fp32 -> bf16 GEMM-operand preparation (optionally transposed, optionally split hi/lo for the
// error-compensated "bf16x3" parity mode).  HBM-bound: 4 B read + 2 B (or 6 B) written per element.
#include "common.cuh"
#include "runtime.h"

namespace tt {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// seg_a/b/c: which of {hi, lo} goes into K-segment 0/1/2:  A side = hi,lo,hi ; B side = hi,hi,lo
template <int SPLIT>
__device__ __forceinline__ void seg_values(float x, __nv_bfloat16 (&o)[3]) {
  __nv_bfloat16 hi, lo;
  split_bf16(x, hi, lo);
  if (SPLIT == 1) { o[0] = hi; o[1] = lo; o[2] = hi; }
  else            { o[0] = hi; o[1] = hi; o[2] = lo; }
}

template <int SPLIT>
__device__ __forceinline__ void cast_store1(float x, __nv_bfloat16* __restrict__ d, long long seg) {
  if (SPLIT == 0) {
    d[0] = __float2bfloat16_rn(x);
  } else {
    __nv_bfloat16 o[3];
    seg_values<SPLIT>(x, o);
#pragma unroll
    for (int s = 0; s < 3; ++s) d[s * seg] = o[s];
  }
}

// One thread per 4 consecutive columns (16 B load, 8 B stores); the cols % 4 leftover columns of
// each row (the 5002-wide adaptive-softmax head) are done one element per thread afterwards.
// idx_t = unsigned whenever rows * cols fits: the row/column split is the only arithmetic here.
template <int SPLIT, typename idx_t>
__global__ void cast_rows_kernel(const float* __restrict__ src, long long ld_src,
                                 __nv_bfloat16* __restrict__ dst, long long ld_dst, int rows,
                                 int cols, long long seg) {
  pdl_prologue();
  const idx_t c4 = static_cast<idx_t>(cols >> 2);
  const idx_t total = static_cast<idx_t>(rows) * c4;
  const idx_t step = static_cast<idx_t>(gridDim.x) * blockDim.x;
  const idx_t first = blockIdx.x * static_cast<idx_t>(blockDim.x) + threadIdx.x;
  for (idx_t i = first; i < total; i += step) {
    const idx_t r = i / c4;
    const int c = static_cast<int>(i - r * c4) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
    const float x[4] = {v.x, v.y, v.z, v.w};
    if (SPLIT == 0) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(x[0], x[1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(x[2], x[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = u;
    } else {
      __nv_bfloat16 o[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j) seg_values<SPLIT>(x[j], o[j]);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        __nv_bfloat162 p0 = __halves2bfloat162(o[0][s], o[1][s]);
        __nv_bfloat162 p1 = __halves2bfloat162(o[2][s], o[3][s]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(dst + r * ld_dst + s * seg + c) = u;
      }
    }
  }
  const int rem = cols & 3;
  if (rem) {
    const int cbase = cols - rem;
    const idx_t tail = static_cast<idx_t>(rows) * rem;
    for (idx_t i = first; i < tail; i += step) {
      const idx_t r = i / static_cast<idx_t>(rem);
      const int c = cbase + static_cast<int>(i - r * rem);
      cast_store1<SPLIT>(src[r * ld_src + c], dst + r * ld_dst + c, seg);
    }
  }
}

template <int SPLIT>
__global__ void cast_rows_scalar_kernel(const float* __restrict__ src, long long ld_src,
                                        __nv_bfloat16* __restrict__ dst, long long ld_dst,
                                        int rows, int cols, long long seg) {
  pdl_prologue();
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols);
    const int c = static_cast<int>(i - static_cast<long long>(r) * cols);
    const float x = src[r * ld_src + c];
    if (SPLIT == 0) {
      dst[r * ld_dst + c] = __float2bfloat16_rn(x);
    } else {
      __nv_bfloat16 o[3];
      seg_values<SPLIT>(x, o);
      for (int s = 0; s < 3; ++s) dst[r * ld_dst + s * seg + c] = o[s];
    }
  }
}

// dst[c, s*rows + r] = seg_s(src[r, c]); 32x32 tiles through shared memory, both sides coalesced.
template <int SPLIT>
__global__ void cast_transpose_kernel(const float* __restrict__ src, long long ld_src,
                                      __nv_bfloat16* __restrict__ dst, long long ld_dst, int rows,
                                      int cols, long long seg) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      const float x = tile[threadIdx.x][j];
      if (SPLIT == 0) {
        dst[c * ld_dst + r] = __float2bfloat16_rn(x);
      } else {
        __nv_bfloat16 o[3];
        seg_values<SPLIT>(x, o);
#pragma unroll
        for (int s = 0; s < 3; ++s) dst[c * ld_dst + s * seg + r] = o[s];
      }
    }
  }
}

template <int SPLIT>
static int cast_dispatch(const float* src, long long ld_src, __nv_bfloat16* dst, long long ld_dst,
                         int rows, int cols, int transpose, long long seg, cudaStream_t s) {
  if (seg <= 0) seg = transpose ? rows : cols;
  if (transpose) {
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
    launch_k(cast_transpose_kernel<SPLIT>, dim3(grid), dim3(block), 0, s, src, ld_src, dst, ld_dst, rows, cols, seg);
    return check_launch("cast_transpose_kernel");
  }
  const bool vec = (cols >= 4) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) && (SPLIT == 0 || seg % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  const long long work = vec ? static_cast<long long>(rows) * (cols / 4)
                             : static_cast<long long>(rows) * cols;
  long long blocks = ceil_div_ll(work, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (vec && static_cast<long long>(rows) * cols < (1ll << 31) - (1ll << 24))
    launch_k(cast_rows_kernel<SPLIT, unsigned>, dim3((int)blocks), dim3(256), 0, s, src, ld_src, dst, ld_dst, rows, cols, seg);
  else if (vec)
    launch_k(cast_rows_kernel<SPLIT, long long>, dim3((int)blocks), dim3(256), 0, s, src, ld_src, dst, ld_dst, rows, cols, seg);
  else
    launch_k(cast_rows_scalar_kernel<SPLIT>, dim3((int)blocks), dim3(256), 0, s, src, ld_src, dst, ld_dst, rows, cols,
                                                               seg);
  return check_launch("cast_rows_kernel");
}

}  // namespace tt

extern "C" int tt_cast_bf16(const float* src, long long ld_src, void* dst, long long ld_dst,
                            int rows, int cols, int transpose, int split, long long seg_stride,
                            void* stream) {
  using namespace tt;
  TT_REQUIRE(src && dst, "tt_cast_bf16: null pointer");
  TT_REQUIRE(rows >= 0 && cols >= 0, "tt_cast_bf16: bad shape");
  TT_REQUIRE(split >= 0 && split <= 2, "tt_cast_bf16: split must be 0, 1 or 2");
  if (rows == 0 || cols == 0) return TT_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  switch (split) {
    case 0: return cast_dispatch<0>(src, ld_src, d, ld_dst, rows, cols, transpose, seg_stride, s);
    case 1: return cast_dispatch<1>(src, ld_src, d, ld_dst, rows, cols, transpose, seg_stride, s);
    default: return cast_dispatch<2>(src, ld_src, d, ld_dst, rows, cols, transpose, seg_stride, s);
  }
}
