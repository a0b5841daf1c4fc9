// "Twin" producers of the decoder layer: row / elementwise kernels that write their result BOTH as the
// fp32 tensor autograd carries and as the bf16 tensor the next tcgen05 GEMM reads as its operand, so
// the ~18 standalone fp32 -> bf16 operand casts per decoder layer (and 9 per decode step) disappear;
// and the multi-problem LayerNorms of a layer's parallel context branches
// (decoder_faces_objects.py:272-352: x_c = LN_c(X + dropout(attn_c(X))) for image / article / faces /
// objects, then torch.cat) as ONE launch forward and ONE backward instead of one per context plus the
// adds that summed the residual gradients.
//
// Every kernel computes exactly what its single-output predecessor in rowops.cu computes (same
// arithmetic order, same dropout element index) -- the fp32 outputs are bit-identical, the bf16
// outputs are their round-to-nearest-even (tests/test_twin_gpu.py).
#include "common.cuh"
#include "runtime.h"

namespace tt {

static inline int tw_flat_grid(long long n, int per_thread) {
  long long g = ceil_div_ll(n, 256LL * per_thread);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, const float4& v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// y = x * dropmask/(1-p) (F.dropout; rowops.cu dropout_kernel), 4 elements per thread.
__global__ void __launch_bounds__(256)
dropout_tw_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                  long long n4, float p, unsigned long long seed, const unsigned long long* step_ptr) {
  pdl_prologue();
  seed = mix_seed(seed, step_ptr);
  const float inv_keep = 1.f / (1.f - p);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const unsigned long long base = 4ull * static_cast<unsigned long long>(i);
    v.x *= dropout_scale(seed, base, p, inv_keep);
    v.y *= dropout_scale(seed, base + 1, p, inv_keep);
    v.z *= dropout_scale(seed, base + 2, p, inv_keep);
    v.w *= dropout_scale(seed, base + 3, p, inv_keep);
    if (y) reinterpret_cast<float4*>(y)[i] = v;
    if (y16) store_bf16x4(y16 + 4 * i, v);
  }
}

// dx = dy * (y > 0)  (backward of the ReLU fused into fc1's GEMM epilogue, decoder_faces_objects.py:357)
__global__ void __launch_bounds__(256)
relu_bwd_tw_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                   __nv_bfloat16* __restrict__ dx16, long long n4) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy) + i);
    const float4 a = __ldg(reinterpret_cast<const float4*>(y) + i);
    float4 v;
    v.x = a.x > 0.f ? d.x : 0.f;
    v.y = a.y > 0.f ? d.y : 0.f;
    v.z = a.z > 0.f ? d.z : 0.f;
    v.w = a.w > 0.f ? d.w : 0.f;
    if (dx) reinterpret_cast<float4*>(dx)[i] = v;
    if (dx16) store_bf16x4(dx16 + 4 * i, v);
  }
}

// GLU (nn.GLU, decoder_faces_objects.py:193-195,259-260): rowops.cu glu_fwd / glu_bwd + bf16 twins.
__global__ void __launch_bounds__(256)
glu_fwd_tw_kernel(const float* __restrict__ h, float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
                  long long n4, int C4) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C4;
    const int c = static_cast<int>(i - r * C4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(h) + r * 2 * C4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(h) + r * 2 * C4 + C4 + c);
    float4 o;
    o.x = a.x * sigmoidf_(b.x); o.y = a.y * sigmoidf_(b.y);
    o.z = a.z * sigmoidf_(b.z); o.w = a.w * sigmoidf_(b.w);
    reinterpret_cast<float4*>(out)[i] = o;
    if (out16) store_bf16x4(out16 + 4 * i, o);
  }
}
__global__ void __launch_bounds__(256)
glu_bwd_tw_kernel(const float* __restrict__ dout, const float* __restrict__ h, float* __restrict__ dh,
                  __nv_bfloat16* __restrict__ dh16, long long n4, int C4) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C4;
    const int c = static_cast<int>(i - r * C4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(h) + r * 2 * C4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(h) + r * 2 * C4 + C4 + c);
    const float4 d = __ldg(reinterpret_cast<const float4*>(dout) + i);
    float4 da, db;
    float s;
    s = sigmoidf_(b.x); da.x = d.x * s; db.x = d.x * a.x * s * (1.f - s);
    s = sigmoidf_(b.y); da.y = d.y * s; db.y = d.y * a.y * s * (1.f - s);
    s = sigmoidf_(b.z); da.z = d.z * s; db.z = d.z * a.z * s * (1.f - s);
    s = sigmoidf_(b.w); da.w = d.w * s; db.w = d.w * a.w * s * (1.f - s);
    reinterpret_cast<float4*>(dh)[r * 2 * C4 + c] = da;
    reinterpret_cast<float4*>(dh)[r * 2 * C4 + C4 + c] = db;
    if (dh16) {
      store_bf16x4(dh16 + 4 * (r * 2 * C4 + c), da);
      store_bf16x4(dh16 + 4 * (r * 2 * C4 + C4 + c), db);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// n <= 4 LayerNorms that share their residual: for context c (blockIdx.y)
//   x_c = res + dropout_c(h_c)   (written back into h_c: what the backward needs)
//   Y[:, c*E:(c+1)*E] = LN(x_c) * gamma_c + beta_c      fp32 (ldy) and bf16 (ldy16)
// One CTA per row, the row lives in registers (one float4 per thread, E <= 1024): the arithmetic of
// rowops.cu ln_fwd_row_kernel.  n = 1 is the plain residual + dropout + post-LayerNorm of
// decoder_faces_objects.py:263-266 / :361-364 with its operand twin.
constexpr int LN_MAX = 4;
struct LnFwdMultiArgs {
  float* h[LN_MAX];
  const float* gamma[LN_MAX];
  const float* beta[LN_MAX];
  float* mean[LN_MAX];
  float* rstd[LN_MAX];
  unsigned long long seed[LN_MAX];
  const float* res;
  float* y;
  long long ldy;
  __nv_bfloat16* y16;
  long long ldy16;
  int N, E;
  float eps, p;
  const unsigned long long* step_ptr;
};

__device__ __forceinline__ float2 tw_block_sum2(float a, float b, float2* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) red[w] = make_float2(a, b);
  __syncthreads();
  float2 t = lane < 8 ? red[lane] : make_float2(0.f, 0.f);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
    t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
  }
  t.x = __shfl_sync(0xffffffffu, t.x, 0);
  t.y = __shfl_sync(0xffffffffu, t.y, 0);
  return t;
}

__global__ void __launch_bounds__(256)
ln_fwd_multi_kernel(const __grid_constant__ LnFwdMultiArgs a) {
  pdl_prologue();
  __shared__ float2 red[8];
  const int ctx = blockIdx.y;
  const int E = a.E, E4 = E >> 2;
  const unsigned long long seed = mix_seed(a.seed[ctx], a.step_ptr);
  const float p = a.p;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const float inv_e = 1.f / static_cast<float>(E);
  const int c = threadIdx.x;
  const bool on = c < E4;
  float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), bt = gm;
  if (on) {
    gm = __ldg(reinterpret_cast<const float4*>(a.gamma[ctx]) + c);
    bt = __ldg(reinterpret_cast<const float4*>(a.beta[ctx]) + c);
  }
  float* h = a.h[ctx];
  for (int r = blockIdx.x; r < a.N; r += gridDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      float4* hr = reinterpret_cast<float4*>(h + static_cast<long long>(r) * E);
      v = hr[c];
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.res) q = __ldg(reinterpret_cast<const float4*>(a.res + static_cast<long long>(r) * E) + c);
      if (p > 0.f) {
        const unsigned long long base = static_cast<unsigned long long>(r) * E + 4ull * c;
        v.x *= dropout_scale(seed, base, p, inv_keep);
        v.y *= dropout_scale(seed, base + 1, p, inv_keep);
        v.z *= dropout_scale(seed, base + 2, p, inv_keep);
        v.w *= dropout_scale(seed, base + 3, p, inv_keep);
      }
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      hr[c] = v;
    }
    const float mu = tw_block_sum2((v.x + v.y) + (v.z + v.w), 0.f, red).x * inv_e;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (on) { d0 = v.x - mu; d1 = v.y - mu; d2 = v.z - mu; d3 = v.w - mu; }
    const float var = tw_block_sum2((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3), 0.f, red).x * inv_e;
    const float rs = rsqrtf(var + a.eps);
    if (c == 0) {
      if (a.mean[ctx]) a.mean[ctx][r] = mu;
      if (a.rstd[ctx]) a.rstd[ctx][r] = rs;
    }
    if (on) {
      float4 o;
      o.x = d0 * rs * gm.x + bt.x;
      o.y = d1 * rs * gm.y + bt.y;
      o.z = d2 * rs * gm.z + bt.z;
      o.w = d3 * rs * gm.w + bt.w;
      const long long col = static_cast<long long>(ctx) * E + 4 * c;
      if (a.y) *reinterpret_cast<float4*>(a.y + static_cast<long long>(r) * a.ldy + col) = o;
      if (a.y16) store_bf16x4(a.y16 + static_cast<long long>(r) * a.ldy16 + col, o);
    }
  }
}

// Backward of the above for all n contexts in one launch (rowops.cu ln_bwd_kernel per context):
//   dx_c = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dY[:, cE:(c+1)E] * gamma_c
//   dh_c = dx_c * dropmask_c/(1-p)   -> fp32 dh[c] (optional) and bf16 dh16[:, cE:(c+1)E] (optional)
//   dX   = sum_c dx_c                (the residual gradient; replaces n-1 axpby launches)
// The 8 warps of a CTA work on 8/n rows x n contexts at a time (warp = (row slot, context)): the n
// contexts of a row run side by side instead of one after the other, the n dx rows meet in shared
// memory and are summed in context order (bit-identical to adding separately stored dx_c one by one).
// dgamma_c / dbeta_c: register accumulators per warp (fixed context), CTA reduction, one atomic per
// column per CTA.
struct LnBwdMultiArgs {
  const float* x[LN_MAX];
  const float* mean[LN_MAX];
  const float* rstd[LN_MAX];
  const float* gamma[LN_MAX];
  float* dgamma[LN_MAX];
  float* dbeta[LN_MAX];
  float* dh[LN_MAX];
  unsigned long long seed[LN_MAX];
  const float* dy;
  long long lddy;
  float* dx;
  __nv_bfloat16* dh16;
  long long lddh16;
  int N, E, n;
  float p;
  const unsigned long long* step_ptr;
};

constexpr int LNB_WARPS = 8;

template <int MAXC>
__global__ void __launch_bounds__(LNB_WARPS * 32)
ln_bwd_multi_kernel(const __grid_constant__ LnBwdMultiArgs a) {
  pdl_prologue();
  extern __shared__ float lnm_red[];   // [LNB_WARPS][E]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = a.n;
  const int rpi = LNB_WARPS / n;             // rows per CTA iteration (n = 3: two warps idle)
  const int rs = warp / n, ctx = warp - rs * n;
  const bool warp_on = warp < rpi * n;
  const int E = a.E, E4 = E >> 2;
  const float p = a.p;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const unsigned long long seed = mix_seed(a.seed[ctx], a.step_ptr);
  const float* __restrict__ xc = a.x[ctx];
  const float* __restrict__ gamma = a.gamma[ctx];
  float* dh = a.dh[ctx];
  float4 acc_g[MAXC], acc_b[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int base = blockIdx.x * rpi; base < a.N; base += gridDim.x * rpi) {      // CTA-uniform trip count
    const int r = base + rs;
    const bool on = warp_on && r < a.N;
    if (on) {
      const float4* dyr = reinterpret_cast<const float4*>(a.dy + static_cast<long long>(r) * a.lddy +
                                                          static_cast<long long>(ctx) * E);
      const float4* xr = reinterpret_cast<const float4*>(xc + static_cast<long long>(r) * E);
      const float mu = a.mean[ctx][r], rstd = a.rstd[ctx][r];
      float s1 = 0.f, s2 = 0.f;
      float4 g4[MAXC], xh4[MAXC];
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < E4) {
          const float4 d = __ldg(dyr + c);
          const float4 xv = __ldg(xr + c);
          const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
          float4 xh, g;
          xh.x = (xv.x - mu) * rstd; xh.y = (xv.y - mu) * rstd;
          xh.z = (xv.z - mu) * rstd; xh.w = (xv.w - mu) * rstd;
          g.x = d.x * gm.x; g.y = d.y * gm.y; g.z = d.z * gm.z; g.w = d.w * gm.w;
          s1 += (g.x + g.y) + (g.z + g.w);
          s2 += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
          acc_g[i].x += d.x * xh.x; acc_g[i].y += d.y * xh.y;
          acc_g[i].z += d.z * xh.z; acc_g[i].w += d.w * xh.w;
          acc_b[i].x += d.x; acc_b[i].y += d.y; acc_b[i].z += d.z; acc_b[i].w += d.w;
          g4[i] = g; xh4[i] = xh;
        }
      }
      s1 = warp_sum(s1) / E;
      s2 = warp_sum(s2) / E;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < E4) {
          float4 o;
          o.x = rstd * (g4[i].x - s1 - xh4[i].x * s2);
          o.y = rstd * (g4[i].y - s1 - xh4[i].y * s2);
          o.z = rstd * (g4[i].z - s1 - xh4[i].z * s2);
          o.w = rstd * (g4[i].w - s1 - xh4[i].w * s2);
          if (a.dx) {
            if (n == 1) reinterpret_cast<float4*>(a.dx + static_cast<long long>(r) * E)[c] = o;
            else reinterpret_cast<float4*>(lnm_red + static_cast<long long>(warp) * E)[c] = o;
          }
          if (p > 0.f) {
            const unsigned long long bidx = static_cast<unsigned long long>(r) * E + 4ull * c;
            o.x *= dropout_scale(seed, bidx, p, inv_keep);
            o.y *= dropout_scale(seed, bidx + 1, p, inv_keep);
            o.z *= dropout_scale(seed, bidx + 2, p, inv_keep);
            o.w *= dropout_scale(seed, bidx + 3, p, inv_keep);
          }
          if (dh) reinterpret_cast<float4*>(dh + static_cast<long long>(r) * E)[c] = o;
          if (a.dh16)
            store_bf16x4(a.dh16 + static_cast<long long>(r) * a.lddh16 + static_cast<long long>(ctx) * E + 4 * c, o);
        }
      }
    }
    if (n > 1 && a.dx) {
      // dX[row] = ((dx_0 + dx_1) + dx_2) + dx_3: plain adds in context order
      __syncthreads();
      for (int i = threadIdx.x; i < rpi * E4; i += blockDim.x) {
        const int row = i / E4, c = i - row * E4;
        if (base + row < a.N) {
          const float4* src = reinterpret_cast<const float4*>(lnm_red + static_cast<long long>(row * n) * E) + c;
          float4 t = src[0];
          for (int k = 1; k < n; ++k) {
            const float4 v = src[static_cast<long long>(k) * E4];
            t.x = __fadd_rn(v.x, t.x); t.y = __fadd_rn(v.y, t.y);
            t.z = __fadd_rn(v.z, t.z); t.w = __fadd_rn(v.w, t.w);
          }
          reinterpret_cast<float4*>(a.dx + static_cast<long long>(base + row) * E)[c] = t;
        }
      }
      __syncthreads();
    }
  }
  // ---- dgamma / dbeta: the warps of one context add their register accumulators through shared
  //      memory, one global atomic per column per CTA and context
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    __syncthreads();
    if (warp_on) {
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < E4) reinterpret_cast<float4*>(lnm_red + static_cast<long long>(warp) * E)[c] = q == 0 ? acc_g[i] : acc_b[i];
      }
    }
    __syncthreads();
    for (int c = 0; c < n; ++c) {
      float* dst = q == 0 ? a.dgamma[c] : a.dbeta[c];
      if (dst == nullptr) continue;        // CTA-uniform
      for (int i = threadIdx.x; i < E; i += blockDim.x) {
        float t = 0.f;
        for (int k = 0; k < rpi; ++k) t += lnm_red[static_cast<long long>(k * n + c) * E + i];
        atomicAdd(dst + i, t);
      }
    }
  }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_dropout_tw(const float* x, float* y, void* y16, long long n, float p,
                             unsigned long long seed, void* stream) {
  TT_REQUIRE(x && (y || y16), "tt_dropout_tw: null pointer");
  TT_REQUIRE(p >= 0.f && p < 1.f, "tt_dropout_tw: bad p");
  TT_REQUIRE(n % 4 == 0, "tt_dropout_tw: n must be a multiple of 4 (got %lld)", n);
  TT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(y16) & 7) == 0, "tt_dropout_tw: misaligned pointer");
  if (n <= 0) return TT_OK;
  launch_k(dropout_tw_kernel, dim3(tw_flat_grid(n / 4, 1)), dim3(256), 0, (cudaStream_t)stream, x, y,
           reinterpret_cast<__nv_bfloat16*>(y16), n / 4, p, seed, rng_step_ptr());
  return check_launch("dropout_tw_kernel");
}

extern "C" int tt_relu_bwd_tw(const float* dy, const float* y, float* dx, void* dx16, long long n,
                              void* stream) {
  TT_REQUIRE(dy && y && (dx || dx16), "tt_relu_bwd_tw: null pointer");
  TT_REQUIRE(n % 4 == 0, "tt_relu_bwd_tw: n must be a multiple of 4 (got %lld)", n);
  TT_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx16) & 7) == 0,
             "tt_relu_bwd_tw: misaligned pointer");
  if (n <= 0) return TT_OK;
  launch_k(relu_bwd_tw_kernel, dim3(tw_flat_grid(n / 4, 1)), dim3(256), 0, (cudaStream_t)stream, dy, y, dx,
           reinterpret_cast<__nv_bfloat16*>(dx16), n / 4);
  return check_launch("relu_bwd_tw_kernel");
}

extern "C" int tt_glu_fwd_tw(const float* h, float* out, void* out16, long long N, int C, void* stream) {
  TT_REQUIRE(h && out, "tt_glu_fwd_tw: null pointer");
  TT_REQUIRE(C > 0 && C % 4 == 0, "tt_glu_fwd_tw: C must be a multiple of 4");
  if (N <= 0) return TT_OK;
  const long long n4 = N * (C / 4);
  launch_k(glu_fwd_tw_kernel, dim3(tw_flat_grid(n4, 1)), dim3(256), 0, (cudaStream_t)stream, h, out,
           reinterpret_cast<__nv_bfloat16*>(out16), n4, C / 4);
  return check_launch("glu_fwd_tw_kernel");
}

extern "C" int tt_glu_bwd_tw(const float* dout, const float* h, float* dh, void* dh16, long long N, int C,
                             void* stream) {
  TT_REQUIRE(dout && h && dh, "tt_glu_bwd_tw: null pointer");
  TT_REQUIRE(C > 0 && C % 4 == 0, "tt_glu_bwd_tw: C must be a multiple of 4");
  if (N <= 0) return TT_OK;
  const long long n4 = N * (C / 4);
  launch_k(glu_bwd_tw_kernel, dim3(tw_flat_grid(n4, 1)), dim3(256), 0, (cudaStream_t)stream, dout, h, dh,
           reinterpret_cast<__nv_bfloat16*>(dh16), n4, C / 4);
  return check_launch("glu_bwd_tw_kernel");
}

extern "C" int tt_ln_fwd_multi(const TtLnFwdMulti* p, void* stream) {
  TT_REQUIRE(p, "tt_ln_fwd_multi: null argument block");
  TT_REQUIRE(p->n >= 1 && p->n <= LN_MAX, "tt_ln_fwd_multi: n must be 1..%d (got %d)", LN_MAX, p->n);
  TT_REQUIRE(p->E > 0 && p->E % 4 == 0 && p->E <= 1024, "tt_ln_fwd_multi: E must be a multiple of 4, <= 1024 (got %d)", p->E);
  TT_REQUIRE(p->y || p->y16, "tt_ln_fwd_multi: no output");
  TT_REQUIRE((p->y == nullptr || p->ldy % 4 == 0) && (p->y16 == nullptr || p->ldy16 % 4 == 0),
             "tt_ln_fwd_multi: ldy / ldy16 must be multiples of 4");
  TT_REQUIRE(p->p_drop >= 0.f && p->p_drop < 1.f, "tt_ln_fwd_multi: bad dropout p");
  if (p->N <= 0) return TT_OK;
  LnFwdMultiArgs a{};
  for (int c = 0; c < p->n; ++c) {
    TT_REQUIRE(p->h[c] && p->gamma[c] && p->beta[c], "tt_ln_fwd_multi: null pointer (context %d)", c);
    a.h[c] = p->h[c]; a.gamma[c] = p->gamma[c]; a.beta[c] = p->beta[c];
    a.mean[c] = p->mean[c]; a.rstd[c] = p->rstd[c]; a.seed[c] = p->seed[c];
  }
  a.res = p->res; a.y = p->y; a.ldy = p->ldy;
  a.y16 = reinterpret_cast<__nv_bfloat16*>(p->y16); a.ldy16 = p->ldy16;
  a.N = p->N; a.E = p->E; a.eps = p->eps; a.p = p->p_drop; a.step_ptr = rng_step_ptr();
  const int rows = p->N < num_sms() * 8 ? p->N : num_sms() * 8;
  launch_k(ln_fwd_multi_kernel, dim3(rows, p->n), dim3(256), 0, (cudaStream_t)stream, a);
  return check_launch("ln_fwd_multi_kernel");
}

extern "C" int tt_ln_bwd_multi(const TtLnBwdMulti* p, void* stream) {
  TT_REQUIRE(p, "tt_ln_bwd_multi: null argument block");
  TT_REQUIRE(p->n >= 1 && p->n <= LN_MAX, "tt_ln_bwd_multi: n must be 1..%d (got %d)", LN_MAX, p->n);
  TT_REQUIRE(p->E > 0 && p->E % 4 == 0 && p->E <= 1024 && p->lddy % 4 == 0,
             "tt_ln_bwd_multi: E must be a multiple of 4, <= 1024 (got %d); lddy a multiple of 4", p->E);
  TT_REQUIRE(p->dy, "tt_ln_bwd_multi: null dy");
  TT_REQUIRE(p->dh16 == nullptr || p->lddh16 % 4 == 0, "tt_ln_bwd_multi: lddh16 must be a multiple of 4");
  if (p->N <= 0) return TT_OK;
  LnBwdMultiArgs a{};
  for (int c = 0; c < p->n; ++c) {
    TT_REQUIRE(p->x[c] && p->mean[c] && p->rstd[c] && p->gamma[c], "tt_ln_bwd_multi: null pointer (context %d)", c);
    a.x[c] = p->x[c]; a.mean[c] = p->mean[c]; a.rstd[c] = p->rstd[c]; a.gamma[c] = p->gamma[c];
    a.dgamma[c] = p->dgamma[c]; a.dbeta[c] = p->dbeta[c]; a.dh[c] = p->dh[c]; a.seed[c] = p->seed[c];
  }
  a.dy = p->dy; a.lddy = p->lddy; a.dx = p->dx;
  a.dh16 = reinterpret_cast<__nv_bfloat16*>(p->dh16); a.lddh16 = p->lddh16;
  a.N = p->N; a.E = p->E; a.n = p->n; a.p = p->p_drop; a.step_ptr = rng_step_ptr();
  long long g = ceil_div_ll(p->N, LNB_WARPS / p->n);
  if (g > num_sms()) g = num_sms();
  const int grid = static_cast<int>(g > 0 ? g : 1);
  const size_t smem = LNB_WARPS * static_cast<size_t>(p->E) * sizeof(float);
  if (p->E <= 256)
    launch_k(ln_bwd_multi_kernel<2>, dim3(grid), dim3(LNB_WARPS * 32), smem, (cudaStream_t)stream, a);
  else
    launch_k(ln_bwd_multi_kernel<8>, dim3(grid), dim3(LNB_WARPS * 32), smem, (cudaStream_t)stream, a);
  return check_launch("ln_bwd_multi_kernel");
}
