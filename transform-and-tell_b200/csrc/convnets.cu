// Kernels of the face encoders: InceptionResnetV1 ("FaceNet",
// tell/facenet/inception_resnet_v1.py:184-299) and the MTCNN P/R/O networks
// (tell/facenet/mtcnn.py:11-159).  Inference only, NHWC bf16 activations; every convolution is an
// im2col (rectangular kernels, per-axis padding, strided input views) feeding the tcgen05 GEMM with
// the folded BatchNorm / bias / residual-scale / ReLU in its epilogue.  The kernels here are the
// HBM-bound glue: im2col, max/avg pooling, PReLU, row l2-normalisation, 2-way softmax.
#include "common.cuh"
#include "runtime.h"

namespace tt {

static inline int flat_grid_cn(long long n) {
  long long g = ceil_div_ll(n, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// out[(b,ho,wo), (kh,kw,c)] = in[b, ho*s - ph + kh, wo*s - pw + kw, c] (zero outside), K padded to
// Kp.  `in` is an NHWC view whose pixel pitch may exceed C (a channel slice of a wider buffer).
// 8 channels (16 B) per thread.
__global__ void im2col_nhwc_hw_kernel(const __nv_bfloat16* __restrict__ in, long long in_pitch,
                                      __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                      int Ho, int Wo, int KH, int KW, int stride, int pad_h,
                                      int pad_w, int Kp) {
  pdl_prologue();
  const int C8 = C >> 3;
  const int chunks_per_row = Kp >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * chunks_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks_per_row);
    const long long row = i / chunks_per_row;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int tap = ch / C8;
    if (tap < KH * KW) {
      const int c8 = ch - tap * C8;
      const int kh = tap / KW, kw = tap - kh * KW;
      const int wo = static_cast<int>(row % Wo);
      const long long t = row / Wo;
      const int ho = static_cast<int>(t % Ho);
      const int b = static_cast<int>(t / Ho);
      const int hi = ho * stride - pad_h + kh, wi = wo * stride - pad_w + kw;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W)
        v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + hi) * W + wi) * in_pitch) + c8);
    }
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// k x k max pool, stride s, padding p, NHWC bf16 with pixel pitches (so the result can land in a
// channel slice of a concatenation buffer).  Windows are clipped to the input (ceil_mode outputs
// of nn.MaxPool2d are windows that start inside the input).  8 channels per thread.
__global__ void maxpool_nhwc_kernel(const __nv_bfloat16* __restrict__ in, long long in_pitch,
                                    __nv_bfloat16* __restrict__ out, long long out_pitch, int B, int H,
                                    int W, int C, int Ho, int Wo, int k, int stride, int pad) {
  pdl_prologue();
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    long long t = i / C8;
    const long long pix = t;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kh = 0; kh < k; ++kh) {
      const int hi = ho * stride - pad + kh;
      if (hi < 0 || hi >= H) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int wi = wo * stride - pad + kw;
        if (wi < 0 || wi >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + hi) * W + wi) * in_pitch) + c8);
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
    reinterpret_cast<uint4*>(out + pix * out_pitch)[c8] = o;
  }
}

// Global average pool: in [B, HW, C] bf16 (pixel pitch C) -> out [B, C] fp32.  One thread per (b, c8).
__global__ void avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B,
                                    int HW, int C) {
  pdl_prologue();
  const int C8 = C >> 3;
  const int total = B * C8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c8 = i % C8, b = i / C8;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (static_cast<long long>(b) * HW + p) * C) + c8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[2 * j] += __low2float(h2[j]);
        a[2 * j + 1] += __high2float(h2[j]);
      }
    }
    const float inv = 1.f / HW;
    float* o = out + static_cast<long long>(b) * C + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * inv;
  }
}

// x[r, c] = x > 0 ? x : slope[c] * x, in place on bf16 rows of pitch `pitch` (nn.PReLU(C)).
__global__ void prelu_kernel(__nv_bfloat16* __restrict__ x, long long pitch, const float* __restrict__ slope,
                             long long rows, int C) {
  pdl_prologue();
  const int C8 = C >> 3;
  const long long total = rows * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const long long r = i / C8;
    uint4* p = reinterpret_cast<uint4*>(x + r * pitch) + c8;
    uint4 v = *p;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(slope + c8 * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(slope + c8 * 8) + 1);
    const float sl[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = __low2float(h2[j]), b = __high2float(h2[j]);
      a = a > 0.f ? a : a * sl[2 * j];
      b = b > 0.f ? b : b * sl[2 * j + 1];
      h2[j] = __floats2bfloat162_rn(a, b);
    }
    *p = v;
  }
}

// y[r, :] = x[r, :] / max(||x[r, :]||_2, eps)   (F.normalize(p=2, dim=1)); one warp per row.
__global__ void l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D,
                                   float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int r = warp_global; r < N; r += nwarps) {
    const float* xr = x + static_cast<long long>(r) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i] * xr[i];
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), eps);
    for (int i = lane; i < D; i += 32) y[static_cast<long long>(r) * D + i] = xr[i] * inv;
  }
}

// In-place softmax over columns [c0, c0+2) of fp32 rows (nn.Softmax(dim=1) on the 2-way face/non-face
// logits of P/R/O-Net).
__global__ void softmax2_kernel(float* __restrict__ x, long long ld, long long rows, int c0) {
  pdl_prologue();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    float* p = x + r * ld + c0;
    const float a = p[0], b = p[1];
    const float m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    const float inv = 1.f / (ea + eb);
    p[0] = ea * inv;
    p[1] = eb * inv;
  }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_conv_out_size(int in, int k, int stride, int pad, int ceil_mode) {
  const int num = in + 2 * pad - k;
  if (num < 0) return 0;
  int o = (ceil_mode ? (num + stride - 1) / stride : num / stride) + 1;
  if (ceil_mode && (o - 1) * stride >= in + pad) --o;   // last window must start inside the input
  return o;
}

extern "C" int tt_im2col_nhwc_hw(const void* in, long long in_pitch, void* out, int B, int H, int W, int C,
                                 int KH, int KW, int stride, int pad_h, int pad_w, int Kp, void* stream) {
  TT_REQUIRE(in && out, "tt_im2col_nhwc_hw: null pointer");
  TT_REQUIRE(C % 8 == 0 && Kp % 8 == 0 && Kp >= KH * KW * C && in_pitch % 8 == 0 && in_pitch >= C,
             "tt_im2col_nhwc_hw: C, Kp and the pixel pitch must be multiples of 8 (C=%d Kp=%d pitch=%lld)",
             C, Kp, in_pitch);
  TT_REQUIRE(stride >= 1 && KH >= 1 && KW >= 1 && pad_h >= 0 && pad_w >= 0, "tt_im2col_nhwc_hw: bad geometry");
  const int Ho = (H + 2 * pad_h - KH) / stride + 1, Wo = (W + 2 * pad_w - KW) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (Kp / 8);
  if (total <= 0) return TT_OK;
  launch_k(im2col_nhwc_hw_kernel, dim3(flat_grid_cn(total)), dim3(256), 0, (cudaStream_t)stream,
           reinterpret_cast<const __nv_bfloat16*>(in), in_pitch, reinterpret_cast<__nv_bfloat16*>(out), B, H, W,
           C, Ho, Wo, KH, KW, stride, pad_h, pad_w, Kp);
  return check_launch("im2col_nhwc_hw_kernel");
}

extern "C" int tt_maxpool_nhwc(const void* in, long long in_pitch, void* out, long long out_pitch, int B,
                               int H, int W, int C, int k, int stride, int pad, int ceil_mode, void* stream) {
  TT_REQUIRE(in && out, "tt_maxpool_nhwc: null pointer");
  TT_REQUIRE(C % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0 && in_pitch >= C && out_pitch >= C,
             "tt_maxpool_nhwc: C and the pixel pitches must be multiples of 8");
  TT_REQUIRE(k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k, "tt_maxpool_nhwc: bad window");
  const int Ho = tt_conv_out_size(H, k, stride, pad, ceil_mode);
  const int Wo = tt_conv_out_size(W, k, stride, pad, ceil_mode);
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  if (total <= 0) return TT_OK;
  launch_k(maxpool_nhwc_kernel, dim3(flat_grid_cn(total)), dim3(256), 0, (cudaStream_t)stream,
           reinterpret_cast<const __nv_bfloat16*>(in), in_pitch, reinterpret_cast<__nv_bfloat16*>(out), out_pitch,
           B, H, W, C, Ho, Wo, k, stride, pad);
  return check_launch("maxpool_nhwc_kernel");
}

extern "C" int tt_avgpool_nhwc(const void* in, float* out, int B, int HW, int C, void* stream) {
  TT_REQUIRE(in && out, "tt_avgpool_nhwc: null pointer");
  TT_REQUIRE(C % 8 == 0 && HW > 0, "tt_avgpool_nhwc: C must be a multiple of 8, HW > 0");
  if (B <= 0) return TT_OK;
  launch_k(avgpool_nhwc_kernel, dim3(flat_grid_cn(static_cast<long long>(B) * (C / 8))), dim3(256), 0,
           (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(in), out, B, HW, C);
  return check_launch("avgpool_nhwc_kernel");
}

extern "C" int tt_prelu_bf16(void* x, long long pitch, const float* slope, long long rows, int C, void* stream) {
  TT_REQUIRE(x && slope, "tt_prelu_bf16: null pointer");
  TT_REQUIRE(C % 8 == 0 && pitch % 8 == 0 && pitch >= C, "tt_prelu_bf16: C and pitch must be multiples of 8");
  TT_REQUIRE((reinterpret_cast<uintptr_t>(slope) & 15) == 0, "tt_prelu_bf16: slope must be 16-byte aligned");
  if (rows <= 0) return TT_OK;
  launch_k(prelu_kernel, dim3(flat_grid_cn(rows * (C / 8))), dim3(256), 0, (cudaStream_t)stream,
           reinterpret_cast<__nv_bfloat16*>(x), pitch, slope, rows, C);
  return check_launch("prelu_kernel");
}

extern "C" int tt_l2norm_rows(const float* x, float* y, int N, int D, float eps, void* stream) {
  TT_REQUIRE(x && y, "tt_l2norm_rows: null pointer");
  if (N <= 0 || D <= 0) return TT_OK;
  const int g = ceil_div(N, 8);
  launch_k(l2norm_rows_kernel, dim3(g < num_sms() * 8 ? g : num_sms() * 8), dim3(256), 0, (cudaStream_t)stream,
           x, y, N, D, eps);
  return check_launch("l2norm_rows_kernel");
}

extern "C" int tt_softmax2(float* x, long long ld, long long rows, int c0, void* stream) {
  TT_REQUIRE(x, "tt_softmax2: null pointer");
  TT_REQUIRE(c0 >= 0 && ld >= c0 + 2, "tt_softmax2: columns [c0, c0+2) must lie inside a row");
  if (rows <= 0) return TT_OK;
  launch_k(softmax2_kernel, dim3(flat_grid_cn(rows)), dim3(256), 0, (cudaStream_t)stream, x, ld, rows, c0);
  return check_launch("softmax2_kernel");
}
