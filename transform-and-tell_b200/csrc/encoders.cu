// Kernels of the two frozen encoders on the hot path (inference-only, bf16 activations):
//   ResNet-152 image encoder (tell/models/resnet.py:92-117): NHWC im2col feeding the tcgen05 GEMM
//     (bias = folded BatchNorm, residual + ReLU in the GEMM epilogue), 3x3/2 max-pool;
//   RoBERTa-large article encoder (fairseq, called at transformer_faces_objects.py:352-353):
//     embedding + LayerNorm, fp32->bf16 LayerNorm, flash self-attention on mma.sync bf16 tensor cores.
#include <mma.h>

#include <cstdlib>

#include "common.cuh"
#include "runtime.h"

namespace tt {

// ------------------------------------------------------------------------------------------ im2col
// out[(b,ho,wo), (kh,kw,c)] = in[b, ho*s - p + kh, wo*s - p + kw, c]  (zero outside), K padded to Kp.
// NHWC bf16 input, 8 channels (16 B) per thread.  The index arithmetic is the cost of this kernel
// (one load + one store per element): idx_t = unsigned whenever the element count fits (64-bit
// div/mod made it issue-bound at 1.3 TB/s), and two elements per iteration are in flight.
template <typename idx_t>
__device__ __forceinline__ uint4 im2col_gather(const __nv_bfloat16* __restrict__ in, idx_t i, int H, int W,
                                               int C, int C8, int Ho, int Wo, int KH, int KW, int stride,
                                               int pad, idx_t chunks_per_row) {
  const int ch = static_cast<int>(i % chunks_per_row);
  const idx_t row = i / chunks_per_row;
  uint4 v = make_uint4(0, 0, 0, 0);
  const int tap = ch / C8;
  if (tap < KH * KW) {
    const int c8 = ch - tap * C8;
    const int kh = tap / KW, kw = tap - kh * KW;
    const int wo = static_cast<int>(row % static_cast<idx_t>(Wo));
    const idx_t t = row / static_cast<idx_t>(Wo);
    const int ho = static_cast<int>(t % static_cast<idx_t>(Ho));
    const long long b = static_cast<long long>(t / static_cast<idx_t>(Ho));
    const int hi = ho * stride - pad + kh, wi = wo * stride - pad + kw;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      v = __ldg(reinterpret_cast<const uint4*>(in + ((b * H + hi) * W + wi) * C) + c8);
  }
  return v;
}

template <typename idx_t>
__global__ void im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                   int B, int H, int W, int C, int Ho, int Wo, int KH, int KW,
                                   int stride, int pad, int Kp) {
  pdl_prologue();
  const int C8 = C >> 3;
  const idx_t chunks_per_row = static_cast<idx_t>(Kp >> 3);
  const idx_t total = static_cast<idx_t>(B) * Ho * Wo * chunks_per_row;
  const idx_t step = static_cast<idx_t>(gridDim.x) * blockDim.x;
  idx_t i = blockIdx.x * static_cast<idx_t>(blockDim.x) + threadIdx.x;
  uint4* o = reinterpret_cast<uint4*>(out);
  for (; i < total && total - i > step; i += 2 * step) {       // i and i + step both in range
    const uint4 v0 = im2col_gather<idx_t>(in, i, H, W, C, C8, Ho, Wo, KH, KW, stride, pad, chunks_per_row);
    const uint4 v1 = im2col_gather<idx_t>(in, i + step, H, W, C, C8, Ho, Wo, KH, KW, stride, pad, chunks_per_row);
    o[i] = v0;
    o[i + step] = v1;
  }
  if (i < total) o[i] = im2col_gather<idx_t>(in, i, H, W, C, C8, Ho, Wo, KH, KW, stride, pad, chunks_per_row);
}
// Warp-per-output-row form (every ResNet layer): the (b, ho, wo) decomposition is paid once per
// row, the lanes walk the row's 16-byte chunks -- contiguous 512 B loads and stores per warp
// instruction -- and tap -> (kh, kw) is a multiply-shift (kw_inv = 65536 / KW + 1, exact for
// tap < 65536 / KW), so no division is left in the element loop.
template <bool POW2>
__global__ void im2col_nhwc_rows_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                        int B, int H, int W, int C, int Ho, int Wo, int KH, int KW,
                                        int stride, int pad, int Kp, int c8_shift, int kw_inv) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const int cpr = Kp >> 3, C8 = C >> 3, taps = KH * KW;
  const int rows = B * Ho * Wo;
  for (int row = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < rows; row += warps) {
    const int wo = row % Wo, t = row / Wo;
    const int ho = t % Ho, b = t / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const __nv_bfloat16* src = in + static_cast<long long>(b) * H * W * C;
    uint4* dst = reinterpret_cast<uint4*>(out) + static_cast<long long>(row) * cpr;
#pragma unroll 4
    for (int ch = lane; ch < cpr; ch += 32) {
      const int tap = POW2 ? (ch >> c8_shift) : (ch / C8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tap < taps) {
        const int c8 = ch - tap * C8;
        const int kh = (tap * kw_inv) >> 16, kw = tap - kh * KW;
        const int hi = h0 + kh, wi = w0 + kw;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W)
          v = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<long long>(hi) * W + wi) * C) + c8);
      }
      dst[ch] = v;
    }
  }
}
// Same gather with the producer's train-mode BatchNorm + ReLU applied on the fly: `in` is the RAW
// output of the previous (1x1) convolution, sc/sh = per-channel scale / shift derived from its column
// sums (the GEMM epilogue's col_stats) once per CTA into shared memory.  Padding taps stay zero (the
// reference pads the NORMALISED activation).  Block 0 moves the running statistics.
template <bool POW2>
__global__ void im2col_nhwc_bn_rows_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                           int B, int H, int W, int C, int Ho, int Wo, int KH, int KW,
                                           int stride, int pad, int Kp, int c8_shift, int kw_inv,
                                           const float* __restrict__ stats, long long n_stat,
                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float eps, float* __restrict__ running_mean,
                                           float* __restrict__ running_var, float momentum,
                                           long long* __restrict__ nbt) {
  pdl_prologue();
  extern __shared__ float sc_sh[];          // [C] scale, [C] shift
  float* sc = sc_sh;
  float* sh = sc_sh + C;
  const float inv_n = 1.f / static_cast<float>(n_stat);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[c] * inv_n;
    const float var = fmaxf(stats[C + c] * inv_n - mean * mean, 0.f);
    const float a = gamma[c] * rsqrtf(var + eps);
    sc[c] = a;
    sh[c] = beta[c] - mean * a;
    if (blockIdx.x == 0 && running_mean != nullptr) {
      const float unbiased = var * (n_stat > 1 ? static_cast<float>(n_stat) / static_cast<float>(n_stat - 1) : 1.f);
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const int cpr = Kp >> 3, C8 = C >> 3, taps = KH * KW;
  const int rows = B * Ho * Wo;
  for (int row = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < rows; row += warps) {
    const int wo = row % Wo, t = row / Wo;
    const int ho = t % Ho, b = t / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const __nv_bfloat16* src = in + static_cast<long long>(b) * H * W * C;
    uint4* dst = reinterpret_cast<uint4*>(out) + static_cast<long long>(row) * cpr;
#pragma unroll 4
    for (int ch = lane; ch < cpr; ch += 32) {
      const int tap = POW2 ? (ch >> c8_shift) : (ch / C8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tap < taps) {
        const int c8 = ch - tap * C8;
        const int kh = (tap * kw_inv) >> 16, kw = tap - kh * KW;
        const int hi = h0 + kh, wi = w0 + kw;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<long long>(hi) * W + wi) * C) + c8);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&v);
          const float4 a0 = *reinterpret_cast<const float4*>(sc + c8 * 8), a1 = *reinterpret_cast<const float4*>(sc + c8 * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sh + c8 * 8), b1 = *reinterpret_cast<const float4*>(sh + c8 * 8 + 4);
          float2 f;
          f = __bfloat1622float2(h2[0]);
          o2[0] = __floats2bfloat162_rn(fmaxf(fmaf(f.x, a0.x, b0.x), 0.f), fmaxf(fmaf(f.y, a0.y, b0.y), 0.f));
          f = __bfloat1622float2(h2[1]);
          o2[1] = __floats2bfloat162_rn(fmaxf(fmaf(f.x, a0.z, b0.z), 0.f), fmaxf(fmaf(f.y, a0.w, b0.w), 0.f));
          f = __bfloat1622float2(h2[2]);
          o2[2] = __floats2bfloat162_rn(fmaxf(fmaf(f.x, a1.x, b1.x), 0.f), fmaxf(fmaf(f.y, a1.y, b1.y), 0.f));
          f = __bfloat1622float2(h2[3]);
          o2[3] = __floats2bfloat162_rn(fmaxf(fmaf(f.x, a1.z, b1.z), 0.f), fmaxf(fmaf(f.y, a1.w, b1.w), 0.f));
        }
      }
      dst[ch] = v;
    }
  }
}
// First layer: NCHW fp32 image (C = 3).  K index = (kh*KW + kw)*C + c.  One thread builds 8
// consecutive k of one output pixel (a 16-byte store); the gathers hit the L1/L2-resident image.
// CC / CKW > 0: compile-time channel count / kernel width (the 7x7, 3-channel ResNet stem), so the
// per-element k -> (tap, c) -> (kh, kw) decode is multiply-shift instead of runtime division.
template <typename idx_t, int CC, int CKW>
__global__ void im2col_nchw_f32_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                       int B, int H, int W, int C_rt, int Ho, int Wo, int KH, int KW_rt,
                                       int stride, int pad, int Kp) {
  const int C = CC > 0 ? CC : C_rt;
  const int KW = CKW > 0 ? CKW : KW_rt;
  pdl_prologue();
  const idx_t chunks = static_cast<idx_t>(Kp >> 3);
  const idx_t total = static_cast<idx_t>(B) * Ho * Wo * chunks;
  const int K = KH * KW * C;
  for (idx_t i = blockIdx.x * static_cast<idx_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<idx_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const idx_t row = i / chunks;
    const int wo = static_cast<int>(row % static_cast<idx_t>(Wo));
    const idx_t t = row / static_cast<idx_t>(Wo);
    const int ho = static_cast<int>(t % static_cast<idx_t>(Ho));
    const int b = static_cast<int>(t / static_cast<idx_t>(Ho));
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      v[e] = 0.f;
      if (k < K) {
        const int c = k % C, tap = k / C;
        const int kh = tap / KW, kw = tap - kh * KW;
        const int hi = h0 + kh, wi = w0 + kw;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W)
          v[e] = __ldg(in + ((static_cast<long long>(b) * C + c) * H + hi) * W + wi);
      }
    }
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}

// 3x3 stride-2 pad-1 max pool, NHWC bf16, 8 channels per thread (resnet.py:98 maxpool).
__global__ void maxpool3x3s2_nhwc_kernel(const __nv_bfloat16* __restrict__ in,
                                         __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                         int Ho, int Wo) {
  pdl_prologue();
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    long long t = i / C8;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * 2 - 1 + kh;
      if (hi < 0 || hi >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * 2 - 1 + kw;
        if (wi < 0 || wi >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + hi) * W + wi) * C) + c8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[2 * j] = fmaxf(m[2 * j], __low2float(h2[j]));
          m[2 * j + 1] = fmaxf(m[2 * j + 1], __high2float(h2[j]));
        }
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(m[2 * j], m[2 * j + 1]);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                   long long n) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------------------------------ LayerNorm -> bf16
// y16 = LN(x) * gamma + beta, optionally zeroing rows flagged in row_zero (padding positions of the
// RoBERTa embedding output).  One warp per row.
__global__ void ln_fwd16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                const uint8_t* __restrict__ row_zero, int N, int E, float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int E4 = E >> 2;
  for (int r = warp_global; r < N; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * E);
    float s = 0.f;
    for (int c = lane; c < E4; c += 32) {
      const float4 v = __ldg(xr + c);
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mu = warp_sum(s) / E;
    float ss = 0.f;
    for (int c = lane; c < E4; c += 32) {
      const float4 v = __ldg(xr + c);
      const float a = v.x - mu, b = v.y - mu, d = v.z - mu, e = v.w - mu;
      ss += (a * a + b * b) + (d * d + e * e);
    }
    const float rs = rsqrtf(warp_sum(ss) / E + eps);
    const float keep = (row_zero && row_zero[r]) ? 0.f : 1.f;
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(r) * E);
    for (int c = lane; c < E4; c += 32) {
      const float4 v = __ldg(xr + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(keep * ((v.x - mu) * rs * g.x + bt.x),
                                                keep * ((v.y - mu) * rs * g.y + bt.y));
      __nv_bfloat162 p1 = __floats2bfloat162_rn(keep * ((v.z - mu) * rs * g.z + bt.z),
                                                keep * ((v.w - mu) * rs * g.w + bt.w));
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      yr[c] = u;
    }
  }
}

// Variable-length packing of the article batch: inv_map[r] = index of padded row r = b*S + t among
// the non-pad rows (sample-major order) or -1 for padding, cu[b] = first packed row of sample b,
// cu[B] = total number of real tokens (used as the device-side row limit of every RoBERTa GEMM).
// One CTA per sample: its offset is the number of real tokens in all earlier samples.
__global__ void __launch_bounds__(256) varlen_prepare_kernel(const long long* __restrict__ ids, int B, int S,
                                                             int pad, int* __restrict__ inv_map,
                                                             int* __restrict__ cu) {
  pdl_prologue();
  __shared__ int red[32];
  __shared__ int wsum[8];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  auto block_total = [&](int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int t = lane < 8 ? red[lane] : 0;
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return __shfl_sync(0xffffffffu, t, 0);
  };
  int before = 0;
  for (long long j = threadIdx.x; j < static_cast<long long>(b) * S; j += 256) before += ids[j] != pad;
  before = block_total(before);
  // own sample: thread t owns the contiguous tokens [t*per, (t+1)*per)
  const int per = (S + 255) / 256;
  const int t0 = threadIdx.x * per;
  const long long* row = ids + static_cast<long long>(b) * S;
  int mine = 0;
  for (int t = t0; t < t0 + per && t < S; ++t) mine += row[t] != pad;
  int incl = mine;                                   // inclusive scan within the warp
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < w; ++i) woff += wsum[i];
  int pos = before + woff + incl - mine;
  for (int t = t0; t < t0 + per && t < S; ++t) {
    const bool real = row[t] != pad;
    inv_map[static_cast<long long>(b) * S + t] = real ? pos : -1;
    pos += real;
  }
  if (threadIdx.x == 255) {
    cu[b] = before;
    if (b == B - 1) cu[B] = pos;                     // thread 255 ends the scan: pos = total
  }
}

// LayerNorm (fp32 in, bf16 out) between the packed and padded layouts.  One WARP per row, the row
// lives in registers (MAXC float4 per lane, all loads issued back to back, shuffle reductions, no
// block barrier): HBM-bound, 4 B read + 2 B (4 B when both layouts are written) per element.
//   inv_map == null : rows i < *count:  y_packed[i] = LN(x[i])                  (mid-layer LN)
//   inv_map != null : padded rows r < R: c = inv_map[r];
//                       c < 0 : y_padded[r] = 0 (padding stays finite for the masked consumers)
//                       else  : y = LN(x[x_packed ? c : r]) -> y_packed[c] and y_padded[r]
template <int MAXC, bool X16>   // E <= 128 * MAXC; X16: the input rows are bf16
__global__ void __launch_bounds__(256) ln_fwd16_varlen_kernel(
    const void* __restrict__ xv, int x_packed, const float* __restrict__ gamma,
    const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_packed,
    __nv_bfloat16* __restrict__ y_padded, const int* __restrict__ inv_map,
    const int* __restrict__ count, int R, int E, float eps) {
  pdl_prologue();
  const int E4 = E >> 2;
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int limit = (inv_map == nullptr && count != nullptr) ? min(R, __ldg(count)) : R;
  const float inv_e = 1.f / E;
  for (int r = warp_global; r < limit; r += nwarps) {
    int c = r, src = r;
    if (inv_map != nullptr) {
      c = inv_map[r];
      if (c < 0) {
        if (y_padded)
          for (int i = lane; i < E4; i += 32)
            reinterpret_cast<uint2*>(y_padded + static_cast<long long>(r) * E)[i] = make_uint2(0u, 0u);
        continue;
      }
      src = x_packed ? c : r;
    }
    float4 v[MAXC];
    float s = 0.f;
    if (X16) {
      const uint2* xr = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(xv) +
                                                       static_cast<long long>(src) * E);
      uint2 u[MAXC];
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        const int i = lane + k * 32;
        u[k] = i < E4 ? __ldg(xr + i) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
        v[k] = make_float4(__low2float(h2[0]), __high2float(h2[0]), __low2float(h2[1]), __high2float(h2[1]));
      }
    } else {
      const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv) +
                                                         static_cast<long long>(src) * E);
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        const int i = lane + k * 32;
        v[k] = i < E4 ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < MAXC; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float mu = warp_sum(s) * inv_e;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      if (lane + k * 32 < E4) {
        const float a = v[k].x - mu, b = v[k].y - mu, d = v[k].z - mu, e = v[k].w - mu;
        ss += (a * a + b * b) + (d * d + e * e);
      }
    }
    const float rs = rsqrtf(warp_sum(ss) * inv_e + eps);
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int i = lane + k * 32;
      if (i < E4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
        uint2 u;
        u.x = pack_bf16((v[k].x - mu) * rs * g.x + bt.x, (v[k].y - mu) * rs * g.y + bt.y);
        u.y = pack_bf16((v[k].z - mu) * rs * g.z + bt.z, (v[k].w - mu) * rs * g.w + bt.w);
        if (y_packed) reinterpret_cast<uint2*>(y_packed + static_cast<long long>(c) * E)[i] = u;
        if (y_padded && inv_map) reinterpret_cast<uint2*>(y_padded + static_cast<long long>(r) * E)[i] = u;
      }
    }
  }
}

// RoBERTa embedding sum: x[r,:] = tok[ids[r]] + pos[position(r)], positions = pad+1+index for non-pad
// tokens (learned positional embedding, fairseq utils.make_positions), pad rows flagged for zeroing.
__global__ void roberta_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                     const float* __restrict__ pos, float* __restrict__ x,
                                     uint8_t* __restrict__ is_pad, int B, int S, int E4, int pad) {
  pdl_prologue();
  const int r = blockIdx.x;
  const int b = r / S, t = r - b * S;
  const long long id = ids[r];
  // position = pad + (number of non-pad tokens in ids[b, 0..t]) for non-pad tokens (cumsum form)
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (int j = threadIdx.x; j <= t; j += blockDim.x) c += ids[static_cast<long long>(b) * S + j] != pad;
  atomicAdd(&cnt, c);
  __syncthreads();
  const bool real = id != pad;
  const int p = real ? pad + cnt : pad;
  if (threadIdx.x == 0) is_pad[r] = real ? 0 : 1;
  const float4* tr = reinterpret_cast<const float4*>(tok) + id * E4;
  const float4* pr = reinterpret_cast<const float4*>(pos) + static_cast<long long>(p) * E4;
  float4* xr = reinterpret_cast<float4*>(x) + static_cast<long long>(r) * E4;
  for (int j = threadIdx.x; j < E4; j += blockDim.x) {
    const float4 a = __ldg(tr + j), q = __ldg(pr + j);
    xr[j] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

// ------------------------------------------------------------------------------------------ flash self-attention
// qkv [B*S, 3E] bf16 (q pre-scaled), key padding mask [B,S]; out [B*S, E] bf16.  D = 64.
// CTA = 64 queries (4 warps x 16 rows) of one (b,h); loop over 64-key tiles; S = QK^T and O += PV on
// mma.sync.m16n8k16 bf16 with fp32 accumulation, online softmax in registers.
constexpr int FA_BM = 64, FA_BN = 64, FA_D = 64, FA_LD = FA_D + 8;  // +8 bf16 pad: conflict-free ldmatrix

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// CTA = 128 queries (8 warps x 16 rows) of one (b,h); K/V tiles of 64 keys are double-buffered with
// cp.async so the loads of tile j+1 overlap the MMAs/softmax of tile j.
constexpr int FA_THREADS = 256;
constexpr int FA_QROWS = 128;

__global__ void __launch_bounds__(FA_THREADS)
flash_self_attn_kernel(const __nv_bfloat16* __restrict__ qkv, const uint8_t* __restrict__ mask,
                       __nv_bfloat16* __restrict__ out, int B, int S, int H,
                       const int* __restrict__ cu) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t fa_smem[];
  typedef __nv_bfloat16 (*Tile)[FA_LD];
  Tile sQ = reinterpret_cast<Tile>(fa_smem);                                   // [128][72]
  Tile sK0 = reinterpret_cast<Tile>(fa_smem + FA_QROWS * FA_LD * 2);           // 2 x [64][72]
  Tile sV0 = reinterpret_cast<Tile>(fa_smem + (FA_QROWS + 2 * FA_BN) * FA_LD * 2);
  float* sMask = reinterpret_cast<float*>(fa_smem + (FA_QROWS + 4 * FA_BN) * FA_LD * 2);  // [2][64]
  const int E = H * FA_D;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int q0 = blockIdx.x * FA_QROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const long long ld = 3LL * E;
  // padded layout: sample b owns rows [b*S, (b+1)*S) and `mask` flags its padding keys;
  // packed (variable-length) layout: rows [cu[b], cu[b+1]) hold its real tokens only.
  long long row0 = static_cast<long long>(b) * S;
  int len = S;
  if (cu != nullptr) {
    row0 = cu[b];
    len = cu[b + 1] - cu[b];
    if (q0 >= len) return;
  }
  const __nv_bfloat16* qbase = qkv + row0 * ld + h * FA_D;
  const __nv_bfloat16* kbase = qbase + E;
  const __nv_bfloat16* vbase = qbase + 2 * E;
  constexpr float LOG2E = 1.4426950408889634f;

  auto load_kv = [&](int buf, int j0) {
    Tile sK = sK0 + buf * FA_BN;
    Tile sV = sV0 + buf * FA_BN;
    for (int i = threadIdx.x; i < FA_BN * (FA_D / 8); i += FA_THREADS) {
      const int r = i >> 3, c8 = i & 7;
      const bool ok = j0 + r < len;
      const long long row = ok ? (j0 + r) : 0;
      cp_async16(&sK[r][c8 * 8], kbase + row * ld + c8 * 8, ok);
      cp_async16(&sV[r][c8 * 8], vbase + row * ld + c8 * 8, ok);
    }
    if (threadIdx.x < FA_BN) {
      const int j = j0 + threadIdx.x;
      sMask[buf * FA_BN + threadIdx.x] =
          (j < len && !(mask && mask[static_cast<long long>(b) * S + j])) ? 0.f : -INFINITY;
    }
  };

  for (int i = threadIdx.x; i < FA_QROWS * (FA_D / 8); i += FA_THREADS) {
    const int r = i >> 3, c8 = i & 7;
    const bool ok = q0 + r < len;
    cp_async16(&sQ[r][c8 * 8], qbase + (ok ? (q0 + r) : 0) * ld + c8 * 8, ok);
  }
  load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3],
                &sQ[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max in log2 units

  const int ntiles = (len + FA_BN - 1) / FA_BN;
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_kv(buf ^ 1, (t + 1) * FA_BN);   // prefetch next tile
    cp_async_commit();
    Tile sK = sK0 + buf * FA_BN;
    Tile sV = sV0 + buf * FA_BN;
    const float* mk = sMask + buf * FA_BN;
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3,
                    &sK[np * 16 + (lane & 7) + ((lane >> 4) << 3)][ks * 16 + ((lane >> 3) & 1) * 8]);
        mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float k0 = mk[n * 8 + 2 * tg], k1 = mk[n * 8 + 2 * tg + 1];
      s[n][0] = s[n][0] * LOG2E + k0; s[n][1] = s[n][1] * LOG2E + k1;
      s[n][2] = s[n][2] * LOG2E + k0; s[n][3] = s[n][3] * LOG2E + k1;
      tm0 = fmaxf(tm0, fmaxf(s[n][0], s[n][1]));
      tm1 = fmaxf(tm1, fmaxf(s[n][2], s[n][3]));
    }
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
    tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
    tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
    const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
    const float c0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - mn0);
    const float c1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - mn1);
    const float base0 = (mn0 == -INFINITY) ? 0.f : mn0, base1 = (mn1 == -INFINITY) ? 0.f : mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p0 = exp2f(s[n][0] - base0), p1 = exp2f(s[n][1] - base0);
      const float p2 = exp2f(s[n][2] - base1), p3 = exp2f(s[n][3] - base1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3,
                          &sV[kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]);
        mma_bf16_16816(o[2 * np], pf[kk], b0, b1);
        mma_bf16_16816(o[2 * np + 1], pf[kk], b2, b3);
      }
    }
    cp_async_wait<0>();   // next tile landed
    __syncthreads();      // everyone done with this tile's buffers before they are overwritten
  }
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = h * FA_D + n * 8 + 2 * tg;
    if (r0 < len)
      *reinterpret_cast<uint32_t*>(out + (row0 + r0) * E + col) =
          pack_bf16(o[n][0] * i0, o[n][1] * i0);
    if (r1 < len)
      *reinterpret_cast<uint32_t*>(out + (row0 + r1) * E + col) =
          pack_bf16(o[n][2] * i1, o[n][3] * i1);
  }
}
constexpr int FA_SMEM = (FA_QROWS + 4 * FA_BN) * FA_LD * 2 + 2 * FA_BN * 4;

static inline int flat_grid3(long long n) {
  long long g = ceil_div_ll(n, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace tt

using namespace tt;

extern "C" int tt_im2col_nhwc(const void* in, void* out, int B, int H, int W, int C, int KH, int KW,
                              int stride, int pad, int Kp, void* stream) {
  TT_REQUIRE(in && out, "tt_im2col_nhwc: null pointer");
  TT_REQUIRE(C % 8 == 0 && Kp % 8 == 0 && Kp >= KH * KW * C, "tt_im2col_nhwc: C, Kp must be multiples of 8");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (Kp / 8);
  if (total <= 0) return TT_OK;
  const int C8 = C / 8;
  static int rows_form = -1;
  if (rows_form < 0) {
    const char* e = getenv("TT_IM2COL_ROWS");
    rows_form = (e && e[0] == '0') ? 0 : 1;       // TT_IM2COL_ROWS=0: flat form (experiments)
  }
  if (rows_form && total < (1ll << 31) && KH * KW < 65536 / KW) {
    const long long rows = static_cast<long long>(B) * Ho * Wo;
    long long ctas = ceil_div_ll(rows, 8);
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (ctas > cap) ctas = cap;
    const bool pow2 = (C8 & (C8 - 1)) == 0;
    int sh = 0;
    while ((1 << sh) < C8) ++sh;
    const int kw_inv = 65536 / KW + 1;
    if (pow2) {
      launch_k(im2col_nhwc_rows_kernel<true>, dim3((int)ctas), dim3(256), 0, (cudaStream_t)stream,
          reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
          Ho, Wo, KH, KW, stride, pad, Kp, sh, kw_inv);
    } else {
      launch_k(im2col_nhwc_rows_kernel<false>, dim3((int)ctas), dim3(256), 0, (cudaStream_t)stream,
          reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
          Ho, Wo, KH, KW, stride, pad, Kp, sh, kw_inv);
    }
  } else if (total < (1ll << 31) - (1ll << 24)) {        // headroom: i + 2 * step must not wrap
    launch_k(im2col_nhwc_kernel<unsigned>, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream,
        reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
        Ho, Wo, KH, KW, stride, pad, Kp);
  } else {
    launch_k(im2col_nhwc_kernel<long long>, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream,
        reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
        Ho, Wo, KH, KW, stride, pad, Kp);
  }
  return check_launch("im2col_nhwc_kernel");
}

extern "C" int tt_im2col_nhwc_bn(const void* in, void* out, int B, int H, int W, int C, int KH, int KW, int stride,
                                 int pad, int Kp, const float* stats, long long n_stat, const float* gamma,
                                 const float* beta, float eps, float* running_mean, float* running_var,
                                 float momentum, long long* num_batches_tracked, void* stream) {
  TT_REQUIRE(in && out && stats && gamma && beta, "tt_im2col_nhwc_bn: null pointer");
  TT_REQUIRE(C % 8 == 0 && C <= 4096 && Kp % 8 == 0 && Kp >= KH * KW * C,
             "tt_im2col_nhwc_bn: C (<= 4096), Kp must be multiples of 8");
  TT_REQUIRE(n_stat > 0, "tt_im2col_nhwc_bn: n_stat must be positive");
  TT_REQUIRE((running_mean == nullptr) == (running_var == nullptr),
             "tt_im2col_nhwc_bn: running_mean and running_var go together");
  TT_REQUIRE(KH * KW < 65536 / KW, "tt_im2col_nhwc_bn: kernel too large");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const long long rows = static_cast<long long>(B) * Ho * Wo;
  if (rows <= 0) return TT_OK;
  TT_REQUIRE(rows * (Kp / 8) < (1ll << 31), "tt_im2col_nhwc_bn: problem too large");
  long long ctas = ceil_div_ll(rows, 8);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (ctas > cap) ctas = cap;
  const int C8 = C / 8;
  const bool pow2 = (C8 & (C8 - 1)) == 0;
  int sh = 0;
  while ((1 << sh) < C8) ++sh;
  if (pow2)
    launch_k(im2col_nhwc_bn_rows_kernel<true>, dim3((int)ctas), dim3(256), static_cast<size_t>(2 * C) * sizeof(float),
             (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B,
             H, W, C, Ho, Wo, KH, KW, stride, pad, Kp, sh, 65536 / KW + 1, stats, n_stat, gamma, beta, eps,
             running_mean, running_var, momentum, num_batches_tracked);
  else
    launch_k(im2col_nhwc_bn_rows_kernel<false>, dim3((int)ctas), dim3(256), static_cast<size_t>(2 * C) * sizeof(float),
             (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B,
             H, W, C, Ho, Wo, KH, KW, stride, pad, Kp, sh, 65536 / KW + 1, stats, n_stat, gamma, beta, eps,
             running_mean, running_var, momentum, num_batches_tracked);
  return check_launch("im2col_nhwc_bn_rows_kernel");
}

extern "C" int tt_im2col_nchw_f32(const float* in, void* out, int B, int H, int W, int C, int KH,
                                  int KW, int stride, int pad, int Kp, void* stream) {
  TT_REQUIRE(in && out, "tt_im2col_nchw_f32: null pointer");
  TT_REQUIRE(Kp >= KH * KW * C && Kp % 8 == 0, "tt_im2col_nchw_f32: Kp must cover KH*KW*C and be a multiple of 8");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (Kp / 8);
  if (total <= 0) return TT_OK;
  if (total < (1ll << 31) - (1ll << 24) && C == 3 && KW == 7) {
    launch_k(im2col_nchw_f32_kernel<unsigned, 3, 7>, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream,
        in, reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, Ho, Wo, KH, KW, stride, pad, Kp);
  } else if (total < (1ll << 31) - (1ll << 24)) {
    launch_k(im2col_nchw_f32_kernel<unsigned, 0, 0>, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream,
        in, reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, Ho, Wo, KH, KW, stride, pad, Kp);
  } else {
    launch_k(im2col_nchw_f32_kernel<long long, 0, 0>, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream,
        in, reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, Ho, Wo, KH, KW, stride, pad, Kp);
  }
  return check_launch("im2col_nchw_f32_kernel");
}

extern "C" int tt_maxpool3x3s2_nhwc(const void* in, void* out, int B, int H, int W, int C,
                                    void* stream) {
  TT_REQUIRE(in && out, "tt_maxpool3x3s2_nhwc: null pointer");
  TT_REQUIRE(C % 8 == 0, "tt_maxpool3x3s2_nhwc: C must be a multiple of 8");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  if (total <= 0) return TT_OK;
  launch_k(maxpool3x3s2_nhwc_kernel, dim3(flat_grid3(total)), dim3(256), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
      Ho, Wo);
  return check_launch("maxpool3x3s2_nhwc_kernel");
}

extern "C" int tt_bf16_to_f32(const void* in, float* out, long long n, void* stream) {
  TT_REQUIRE(in && out, "tt_bf16_to_f32: null pointer");
  if (n <= 0) return TT_OK;
  launch_k(bf16_to_f32_kernel, dim3(flat_grid3(n)), dim3(256), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(in), out, n);
  return check_launch("bf16_to_f32_kernel");
}

extern "C" int tt_ln_fwd16(const float* x, const float* gamma, const float* beta, void* y16,
                           const uint8_t* row_zero, int N, int E, float eps, void* stream) {
  TT_REQUIRE(x && gamma && beta && y16, "tt_ln_fwd16: null pointer");
  TT_REQUIRE(E % 4 == 0, "tt_ln_fwd16: E must be a multiple of 4");
  if (N <= 0) return TT_OK;
  long long g = ceil_div_ll(N, 8);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (g > cap) g = cap;
  launch_k(ln_fwd16_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, 
      x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y16), row_zero, N, E, eps);
  return check_launch("ln_fwd16_kernel");
}

extern "C" int tt_roberta_embed(const long long* ids, const float* tok, const float* pos, float* x,
                                uint8_t* is_pad, int B, int S, int E, int pad, void* stream) {
  TT_REQUIRE(ids && tok && pos && x && is_pad, "tt_roberta_embed: null pointer");
  TT_REQUIRE(E % 4 == 0, "tt_roberta_embed: E must be a multiple of 4");
  if (B * S <= 0) return TT_OK;
  launch_k(roberta_embed_kernel, dim3(B * S), dim3(128), 0, (cudaStream_t)stream, ids, tok, pos, x, is_pad, B, S, E / 4,
                                                               pad);
  return check_launch("roberta_embed_kernel");
}

extern "C" int tt_flash_self_attn(const void* qkv, const uint8_t* key_padding_mask, void* out, int B,
                                  int S, int H, int D, void* stream) {
  TT_REQUIRE(qkv && out, "tt_flash_self_attn: null pointer");
  TT_REQUIRE(D == FA_D, "tt_flash_self_attn: head_dim must be %d (got %d)", FA_D, D);
  if (B <= 0 || S <= 0) return TT_OK;
  dim3 grid(ceil_div(S, FA_QROWS), B * H);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(flash_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    attr_set = true;
  }
  launch_k(flash_self_attn_kernel, dim3(grid), dim3(FA_THREADS), FA_SMEM, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(qkv), key_padding_mask,
      reinterpret_cast<__nv_bfloat16*>(out), B, S, H, (const int*)nullptr);
  return check_launch("flash_self_attn_kernel");
}

extern "C" int tt_flash_self_attn_varlen(const void* qkv, const int* cu_seqlens, void* out, int B,
                                         int S_max, int H, int D, void* stream) {
  TT_REQUIRE(qkv && out && cu_seqlens, "tt_flash_self_attn_varlen: null pointer");
  TT_REQUIRE(D == FA_D, "tt_flash_self_attn_varlen: head_dim must be %d (got %d)", FA_D, D);
  if (B <= 0 || S_max <= 0) return TT_OK;
  dim3 grid(ceil_div(S_max, FA_QROWS), B * H);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(flash_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    attr_set = true;
  }
  launch_k(flash_self_attn_kernel, dim3(grid), dim3(FA_THREADS), FA_SMEM, (cudaStream_t)stream,
      reinterpret_cast<const __nv_bfloat16*>(qkv), (const uint8_t*)nullptr,
      reinterpret_cast<__nv_bfloat16*>(out), B, S_max, H, cu_seqlens);
  return check_launch("flash_self_attn_kernel");
}

extern "C" int tt_varlen_prepare(const long long* ids, int B, int S, int pad, int* inv_map,
                                 int* cu_seqlens, void* stream) {
  TT_REQUIRE(ids && inv_map && cu_seqlens, "tt_varlen_prepare: null pointer");
  if (B <= 0 || S <= 0) return TT_OK;
  launch_k(varlen_prepare_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, ids, B, S, pad, inv_map,
           cu_seqlens);
  return check_launch("varlen_prepare_kernel");
}

extern "C" int tt_ln_fwd16_varlen(const void* x, int x_bf16, int x_packed, const float* gamma, const float* beta,
                                  void* y_packed, void* y_padded, const int* inv_map,
                                  const int* count_ptr, int R, int E, float eps, void* stream) {
  TT_REQUIRE(x && gamma && beta && (y_packed || y_padded), "tt_ln_fwd16_varlen: null pointer");
  TT_REQUIRE(E % 4 == 0 && E <= 1024, "tt_ln_fwd16_varlen: E must be a multiple of 4 and <= 1024 (got %d)", E);
  if (R <= 0) return TT_OK;
  const int cap = num_sms() * 8;
  const int want = ceil_div(R, 8);
  const dim3 grid(want < cap ? want : cap);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y_packed);
  __nv_bfloat16* yd = reinterpret_cast<__nv_bfloat16*>(y_padded);
  cudaStream_t st = (cudaStream_t)stream;
#define TT_LN_VARLEN(MAXC, X16) \
  launch_k(ln_fwd16_varlen_kernel<MAXC, X16>, grid, dim3(256), 0, st, x, x_packed, gamma, beta, yp, yd, inv_map, \
           count_ptr, R, E, eps)
  if (E <= 256) { if (x_bf16) TT_LN_VARLEN(2, true); else TT_LN_VARLEN(2, false); }
  else { if (x_bf16) TT_LN_VARLEN(8, true); else TT_LN_VARLEN(8, false); }
#undef TT_LN_VARLEN
  return check_launch("ln_fwd16_varlen_kernel");
}
