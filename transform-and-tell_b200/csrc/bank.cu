// Weight bank: ONE launch prepares every GEMM weight operand of the decoder for a step, and ONE
// launch runs every weight-norm backward.
//
// The reference rebuilds w = g * v / ||v|| inside each GehringLinear.forward (linear.py:30-34,
// nn.utils.weight_norm) and hands fp32 weights to addmm.  Here every trainable matrix is consumed
// as a bf16 tcgen05 operand, which must be rebuilt after every optimizer step.  Doing that per
// layer costs ~250 launches of 3-17 us on the decoder's critical path (HBM-bound kernels too small
// to fill the machine); batched, the same 1.4 GB of traffic is one bandwidth-bound launch.
//
// Work decomposition: one warp per matrix row, rows of all segments numbered consecutively
// (row0 = exclusive prefix sum, ascending), segment found by binary search.  HBM-bound:
// 4 B read + 2 B written per element (+4 B when the fp32 effective weight is kept).
#include "common.cuh"
#include "runtime.h"

namespace tt {

template <typename Seg>
__device__ __forceinline__ int find_seg(const Seg* __restrict__ segs, int nsegs, int row) {
  int lo = 0, hi = nsegs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].row0 <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  uint2 u;
  u.x = pack_bf16(a, b);
  u.y = pack_bf16(c, d);
  return u;
}

__global__ void __launch_bounds__(256) weight_prep_kernel(const TtPrepSeg* __restrict__ segs, int nsegs,
                                                          int total_rows) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int row = warp_global; row < total_rows; row += nwarps) {
    const TtPrepSeg sg = segs[find_seg(segs, nsegs, row)];
    const int r = row - sg.row0;
    const float* src = sg.src + static_cast<long long>(r) * sg.ld_src;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sg.dst16) + static_cast<long long>(r) * sg.ld_dst;
    float* w32 = sg.w32 ? sg.w32 + static_cast<long long>(r) * sg.cols : nullptr;
    const int cols = sg.cols;
    const bool vec = (cols & 3) == 0 && (sg.ld_src & 3) == 0 && (sg.ld_dst & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(sg.src) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(sg.dst16) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(sg.w32) & 15) == 0;
    float sc = 1.f;
    if (sg.g != nullptr) {
      float s = 0.f;
      if (vec) {
        for (int i = lane * 4; i < cols; i += 128) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
          s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      } else {
        for (int i = lane; i < cols; i += 32) s += src[i] * src[i];
      }
      s = warp_sum(s);
      const float nrm = sqrtf(s);
      sc = sg.g[r] / nrm;
      if (lane == 0 && sg.norm) sg.norm[r] = nrm;
    }
    if (vec) {
      for (int i = lane * 4; i < cols; i += 128) {
        float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        if (w32) *reinterpret_cast<float4*>(w32 + i) = v;
        *reinterpret_cast<uint2*>(dst + i) = pack4_bf16(v.x, v.y, v.z, v.w);
      }
    } else {
      for (int i = lane; i < cols; i += 32) {
        const float v = src[i] * sc;
        if (w32) w32[i] = v;
        dst[i] = __float2bfloat16_rn(v);
      }
    }
  }
}

// dg[o] = <dw,v>/||v|| ; dv = g/||v|| * (dw - v <dw,v>/||v||^2)      (backward of linear.py:30-34)
__global__ void __launch_bounds__(256) wnorm_bwd_multi_kernel(const TtWnormBwdSeg* __restrict__ segs,
                                                              int nsegs, int total_rows) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int row = warp_global; row < total_rows; row += nwarps) {
    const TtWnormBwdSeg sg = segs[find_seg(segs, nsegs, row)];
    const int r = row - sg.row0;
    const int cols = sg.cols;
    const float* vr = sg.v + static_cast<long long>(r) * cols;
    const float* dr = sg.dw + static_cast<long long>(r) * cols;
    float* out = sg.dv + static_cast<long long>(r) * cols;
    const bool vec = (cols & 3) == 0 && (reinterpret_cast<uintptr_t>(sg.v) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(sg.dw) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(sg.dv) & 15) == 0;
    float s = 0.f;
    if (vec) {
      for (int i = lane * 4; i < cols; i += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dr + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(vr + i));
        s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
    } else {
      for (int i = lane; i < cols; i += 32) s += dr[i] * vr[i];
    }
    s = warp_sum(s);
    const float nrm = sg.norm[r];
    const float gg = sg.g[r];
    if (lane == 0) sg.dg[r] = s / nrm;
    const float a_ = gg / nrm, b_ = s / (nrm * nrm);
    if (vec) {
      for (int i = lane * 4; i < cols; i += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(dr + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(vr + i));
        *reinterpret_cast<float4*>(out + i) = make_float4(a_ * (a.x - b.x * b_), a_ * (a.y - b.y * b_),
                                                          a_ * (a.z - b.z * b_), a_ * (a.w - b.w * b_));
      }
    } else {
      for (int i = lane; i < cols; i += 32) out[i] = a_ * (dr[i] - vr[i] * b_);
    }
  }
}

static inline int bank_grid(int total_rows) {
  const int want = ceil_div(total_rows, 8);
  const int cap = num_sms() * 8;     // 8 CTAs x 8 warps per SM: full occupancy for a streaming kernel
  return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace tt

extern "C" int tt_weight_prep(const TtPrepSeg* segs_dev, int nsegs, int total_rows, void* stream) {
  using namespace tt;
  TT_REQUIRE(segs_dev != nullptr && nsegs > 0 && total_rows > 0, "tt_weight_prep: empty table");
  launch_k(weight_prep_kernel, dim3(bank_grid(total_rows)), dim3(256), 0, (cudaStream_t)stream, segs_dev,
           nsegs, total_rows);
  return check_launch("weight_prep_kernel");
}

extern "C" int tt_wnorm_bwd_multi(const TtWnormBwdSeg* segs_dev, int nsegs, int total_rows,
                                  void* stream) {
  using namespace tt;
  TT_REQUIRE(segs_dev != nullptr && nsegs > 0 && total_rows > 0, "tt_wnorm_bwd_multi: empty table");
  launch_k(wnorm_bwd_multi_kernel, dim3(bank_grid(total_rows)), dim3(256), 0, (cudaStream_t)stream,
           segs_dev, nsegs, total_rows);
  return check_launch("wnorm_bwd_multi_kernel");
}
