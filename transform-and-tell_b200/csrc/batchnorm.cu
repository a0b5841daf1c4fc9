// Train-mode BatchNorm2d of the frozen ResNet-152: the reference's training step
// runs model.train() (tell/training/callback_apex_trainer.py:259), so every nn.BatchNorm2d of
// tell/models/resnet.py:12-117 normalises with the statistics of the CURRENT batch and moves its
// running statistics -- only the parameters are frozen.  The convolution itself is the tcgen05 GEMM
// writing the raw bf16 NHWC output [M = B*H*W, C]; the two kernels here are the HBM-bound remainder:
//   bn_stats_kernel : per-channel sum / sum of squares over the M pixels (fp32, one atomic per
//                     channel per CTA) -- 1 read of the activation
//   bn_apply_kernel : y = (x - mean) * rsqrt(var_biased + eps) * gamma + beta (+ identity) -> ReLU,
//                     in place, plus the running-statistics update (momentum, UNBIASED variance,
//                     num_batches_tracked) -- 1 read + 1 write (+ 1 read of the identity)
// Thread layout of both: a thread owns 8 consecutive channels (one 16-byte load) for a strided set
// of rows, so a warp reads 512 contiguous bytes and the per-channel scale/shift are computed once.
#include "common.cuh"
#include "runtime.h"

namespace tt {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// stats[0..C) += sum_r x[r,c];  stats[C..2C) += sum_r x[r,c]^2.   C/8 <= 256 column groups.
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long pitch,
                                                       long long M, int C, float* __restrict__ stats) {
  pdl_prologue();
  extern __shared__ float sh[];                 // [2*C]
  const int ncg = C >> 3;
  const int rows_per_pass = blockDim.x / ncg;   // >= 1
  const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (rl < rows_per_pass) {
    const long long step = static_cast<long long>(gridDim.x) * rows_per_pass;
    long long r = static_cast<long long>(blockIdx.x) * rows_per_pass + rl;
    // two rows in flight per thread
    for (; r + step < M; r += 2 * step) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + r * pitch) + cg);
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(x + (r + step) * pitch) + cg);
      float fa[8], fb[8];
      unpack8(a, fa);
      unpack8(b, fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += fa[i] + fb[i];
        q[i] += fa[i] * fa[i] + fb[i] * fb[i];
      }
    }
    if (r < M) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + r * pitch) + cg);
      float fa[8];
      unpack8(a, fa);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += fa[i];
        q[i] += fa[i] * fa[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&sh[cg * 8 + i], s[i]);
      atomicAdd(&sh[C + cg * 8 + i], q[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], sh[i]);
}

// In place on x.  residual (bf16, may be null) is added after the affine transform, before the ReLU.
// The per-channel scale / shift are computed once per CTA into shared memory (a thread then reads the
// 8 it owns), two rows are in flight per thread.
__global__ void __launch_bounds__(256) bn_apply_kernel(__nv_bfloat16* __restrict__ x, long long pitch, long long M,
                                                       int C, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       const __nv_bfloat16* __restrict__ residual, long long rpitch,
                                                       int relu, float* __restrict__ running_mean,
                                                       float* __restrict__ running_var, float momentum,
                                                       long long* __restrict__ num_batches_tracked) {
  pdl_prologue();
  extern __shared__ float sh[];            // [C] scale, [C] shift
  const float inv_m = 1.f / static_cast<float>(M);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[c] * inv_m;
    const float var = fmaxf(stats[C + c] * inv_m - mean * mean, 0.f);
    const float a = gamma[c] * rsqrtf(var + eps);
    sh[c] = a;
    sh[C + c] = beta[c] - mean * a;
    if (blockIdx.x == 0 && running_mean != nullptr) {
      const float unbiased = var * (M > 1 ? static_cast<float>(M) / static_cast<float>(M - 1) : 1.f);
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  __syncthreads();
  const int ncg = C >> 3;
  const int rows_per_pass = blockDim.x / ncg;
  const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
  if (rl >= rows_per_pass) return;
  float sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = sh[cg * 8 + i];
    sf[i] = sh[C + cg * 8 + i];
  }
  auto apply = [&](uint4 v, uint4 rv, bool has_res) -> uint4 {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], sc[i], sf[i]);
    if (has_res) {
      float g[8];
      unpack8(rv, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return o;
  };
  const bool has_res = residual != nullptr;
  const long long step = static_cast<long long>(gridDim.x) * rows_per_pass;
  long long r = static_cast<long long>(blockIdx.x) * rows_per_pass + rl;
  for (; r + step < M; r += 2 * step) {
    uint4* p0 = reinterpret_cast<uint4*>(x + r * pitch) + cg;
    uint4* p1 = reinterpret_cast<uint4*>(x + (r + step) * pitch) + cg;
    const uint4 v0 = *p0, v1 = *p1;
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (has_res) {
      r0 = __ldg(reinterpret_cast<const uint4*>(residual + r * rpitch) + cg);
      r1 = __ldg(reinterpret_cast<const uint4*>(residual + (r + step) * rpitch) + cg);
    }
    *p0 = apply(v0, r0, has_res);
    *p1 = apply(v1, r1, has_res);
  }
  if (r < M) {
    uint4* p0 = reinterpret_cast<uint4*>(x + r * pitch) + cg;
    const uint4 v0 = *p0;
    uint4 r0 = make_uint4(0, 0, 0, 0);
    if (has_res) r0 = __ldg(reinterpret_cast<const uint4*>(residual + r * rpitch) + cg);
    *p0 = apply(v0, r0, has_res);
  }
}

static inline int bn_grid(long long M, int C) {
  const int rows_per_pass = 256 / (C / 8);
  long long g = ceil_div_ll(M, static_cast<long long>(rows_per_pass) * 8);   // >= 8 rows per thread
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (g > cap) g = cap;
  return static_cast<int>(g > 0 ? g : 1);
}

}  // namespace tt

using namespace tt;

extern "C" int tt_bn_stats_bf16(const void* x, long long pitch, long long M, int C, float* stats, void* stream) {
  TT_REQUIRE(x && stats, "tt_bn_stats_bf16: null pointer");
  TT_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048 && pitch % 8 == 0 && pitch >= C,
             "tt_bn_stats_bf16: C must be a multiple of 8 in [8, 2048], pitch a multiple of 8 >= C");
  if (M <= 0) return TT_OK;
  launch_k(bn_stats_kernel, dim3(bn_grid(M, C)), dim3(256), static_cast<size_t>(2 * C) * sizeof(float),
           (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x), pitch, M, C, stats);
  return check_launch("bn_stats_kernel");
}

extern "C" int tt_bn_apply_bf16(void* x, long long pitch, long long M, int C, const float* stats,
                                const float* gamma, const float* beta, float eps, const void* residual,
                                long long rpitch, int relu, float* running_mean, float* running_var,
                                float momentum, long long* num_batches_tracked, void* stream) {
  TT_REQUIRE(x && stats && gamma && beta, "tt_bn_apply_bf16: null pointer");
  TT_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048 && pitch % 8 == 0 && pitch >= C,
             "tt_bn_apply_bf16: C must be a multiple of 8 in [8, 2048], pitch a multiple of 8 >= C");
  TT_REQUIRE(residual == nullptr || (rpitch % 8 == 0 && rpitch >= C), "tt_bn_apply_bf16: bad residual pitch");
  TT_REQUIRE((running_mean == nullptr) == (running_var == nullptr),
             "tt_bn_apply_bf16: running_mean and running_var go together");
  if (M <= 0) return TT_OK;
  launch_k(bn_apply_kernel, dim3(bn_grid(M, C)), dim3(256), static_cast<size_t>(2 * C) * sizeof(float),
           (cudaStream_t)stream,
           reinterpret_cast<__nv_bfloat16*>(x), pitch, M, C, stats, gamma, beta, eps,
           reinterpret_cast<const __nv_bfloat16*>(residual), rpitch, relu, running_mean, running_var, momentum,
           num_batches_tracked);
  return check_launch("bn_apply_kernel");
}
