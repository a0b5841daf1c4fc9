// DynamicConv1dTBC core (tell/modules/convolutions/dynamic.py:285-336, _forward_expanded):
//   p[t,b,h,:]   = softmax_K(z[t,b,h,:])                 (all K taps, even those that hit t' < 0)
//   w            = DropConnect(p)                         (dynamic.py:305-306)
//   out[t,b,hR+r] = sum_k w[t,b,h,k] * x[t-(K-1)+k, b, hR+r]   (x = 0 for negative time)
// The reference builds a [B*H, T, T+K-1] band matrix and bmm's it (O(T^2) memory); here one CTA owns
// a (batch, head, 32-step) tile: the x window is staged once in shared memory (HBM-bound, ~2x reuse
// factor (TT+K-1)/TT instead of K), one warp per time step, the K<=32 taps live one per lane and the
// softmax is a warp-shuffle reduction.
// LightweightConv1dTBC (lightweight.py:88-240) is the same kernel with z broadcast over (t,b).
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int DC_TT = 32;      // time steps per CTA
constexpr int DC_MAXK = 32;    // taps (one per lane)
constexpr int DC_MAXR = 64;    // channels per head
constexpr int DC_ROWS = DC_TT + DC_MAXK - 1;
constexpr int DC_LD = DC_MAXR + 1;  // padded row pitch (bank-conflict free column walks)

struct DynConvArgs {
  const float* x;     // [T,B,C]
  const float* z;     // [T,B,H,K] logits (or [H,K] when z_tb_stride == 0)
  long long z_tb_stride;  // elements between consecutive (t,b) rows of z (H*K, or 0 = broadcast)
  float* out;         // [T,B,C]
  float* probs;       // [T,B,H,K] softmax output saved for backward (may be null)
  int T, B, C, H, K;
  int softmax;
  float p_drop;
  unsigned long long seed;
  const unsigned long long* step_ptr;
  __nv_bfloat16* out16;   // optional bf16 twin of out (the GEMM operand of linear2), same layout
};

__global__ void __launch_bounds__(256)
dynconv_fwd_kernel(DynConvArgs a) {
  pdl_prologue();
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ float xs[DC_ROWS][DC_LD];
  const int R = a.C / a.H;
  const int K = a.K;
  const int bh = blockIdx.x;
  const int b = bh / a.H, h = bh - b * a.H;
  const int t0 = blockIdx.y * DC_TT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = DC_TT + K - 1;
  // stage x[t0-K+1 .. t0+TT-1, b, hR .. hR+R)
  for (int i = threadIdx.x; i < nrows * R; i += blockDim.x) {
    const int row = i / R, c = i - row * R;
    const int t = t0 - (K - 1) + row;
    float v = 0.f;
    if (t >= 0 && t < a.T) v = __ldg(a.x + (static_cast<long long>(t) * a.B + b) * a.C + h * R + c);
    xs[row][c] = v;
  }
  __syncthreads();
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int tt = warp; tt < DC_TT; tt += (blockDim.x >> 5)) {
    const int t = t0 + tt;
    if (t >= a.T) break;  // warp-uniform
    const long long tb = static_cast<long long>(t) * a.B + b;
    float zk = -INFINITY;
    if (lane < K) zk = __ldg(a.z + tb * a.z_tb_stride + h * K + lane);
    float pk;
    if (a.softmax) {
      const float m = warp_max(zk);
      const float e = lane < K ? expf(zk - m) : 0.f;
      const float s = warp_sum(e);
      pk = e / s;
    } else {
      pk = lane < K ? zk : 0.f;
    }
    const long long widx = (tb * a.H + h) * K + lane;
    if (a.probs && lane < K) a.probs[widx] = pk;
    float wk = pk;
    if (a.p_drop > 0.f && lane < K)
      wk *= dropout_scale(a.seed, static_cast<unsigned long long>(widx), a.p_drop, inv_keep);
    for (int c0 = 0; c0 < R; c0 += 32) {  // warp-uniform trip count: the shuffles need all lanes
      const int c = c0 + lane;
      const int cc = c < R ? c : 0;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += __shfl_sync(0xffffffffu, wk, k) * xs[tt + k][cc];
      if (c < R) {
        a.out[tb * a.C + h * R + c] = acc;
        if (a.out16) a.out16[tb * a.C + h * R + c] = __float2bfloat16_rn(acc);
      }
    }
  }
}

// One incremental decoding step (dynamic.py:95-116 with T = 1): the reference concatenates the
// <= K-1 buffered inputs with the new one, recomputes the convolution over the whole window and keeps
// the last row.  Here `window` is a fixed [K-1, B, C] time-ordered buffer (zero rows = the causal zero
// padding of the first steps): one thread per (b, channel) column computes the single output that is
// needed, then shifts its column by one step and appends the new input -- the state update and the
// convolution are one pass over the window (HBM-bound: (K-1)*B*C*8 bytes per step).
__global__ void __launch_bounds__(256)
dynconv_step_kernel(float* __restrict__ window, const float* __restrict__ x_new,
                    const float* __restrict__ z, long long z_b_stride, float* __restrict__ out,
                    __nv_bfloat16* __restrict__ out16, int B, int C, int H, int K, int softmax) {
  pdl_prologue();
  const long long BC = static_cast<long long>(B) * C;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= BC) return;
  const int b = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(b) * C);
  const int h = c / (C / H);
  const float* zr = z + b * z_b_stride + h * K;
  float w[DC_MAXK];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < DC_MAXK; ++k) {
    w[k] = k < K ? __ldg(zr + k) : -INFINITY;
    m = fmaxf(m, w[k]);
  }
  if (softmax) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < DC_MAXK; ++k) {
      w[k] = k < K ? expf(w[k] - m) : 0.f;
      sum += w[k];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < DC_MAXK; ++k) w[k] *= inv;
  } else {
#pragma unroll
    for (int k = 0; k < DC_MAXK; ++k) w[k] = k < K ? w[k] : 0.f;
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < DC_MAXK - 1; ++k) {
    if (k < K - 1) {
      const float v = window[k * BC + idx];
      acc += w[k] * v;
      if (k > 0) window[(k - 1) * BC + idx] = v;      // shift: row k becomes row k-1
    }
  }
  const float xn = __ldg(x_new + idx);
#pragma unroll
  for (int k = 0; k < DC_MAXK; ++k)
    if (k == K - 1) acc += w[k] * xn;
  if (K > 1) window[(K - 2) * BC + idx] = xn;
  out[idx] = acc;
  if (out16) out16[idx] = __float2bfloat16_rn(acc);
}

// Same step for heads of 64 channels (every shipped decoder: C 1024 / 16 heads): ONE WARP per (b, h).
// Lane k owns tap k for the softmax (one expf per tap per head instead of one per tap per CHANNEL:
// the thread-per-channel form above spends 64x the exponentials and was compute-bound at K = 31),
// then every lane owns two adjacent channels of the head (float2: 256 contiguous bytes per warp row)
// and the tap weights are broadcast by shuffle while the window is read, shifted and appended.
__global__ void __launch_bounds__(256)
dynconv_step_warp_kernel(float* __restrict__ window, const float* __restrict__ x_new,
                         const float* __restrict__ z, long long z_b_stride, float* __restrict__ out,
                         __nv_bfloat16* __restrict__ out16, int B, int C, int H, int K, int softmax) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (wid >= static_cast<long long>(B) * H) return;          // warp-uniform
  const int b = static_cast<int>(wid / H), h = static_cast<int>(wid - static_cast<long long>(b) * H);
  const long long BC = static_cast<long long>(B) * C;
  float w = lane < K ? __ldg(z + b * z_b_stride + h * K + lane) : -INFINITY;
  if (softmax) {
    const float m = warp_max(w);
    const float e = lane < K ? expf(w - m) : 0.f;
    const float sum = warp_sum(e);
    w = e / sum;
  } else if (lane >= K) {
    w = 0.f;
  }
  const long long idx = static_cast<long long>(b) * C + h * 64 + 2 * lane;
  float2 acc = make_float2(0.f, 0.f);
  float2 cur = K > 1 ? *reinterpret_cast<const float2*>(window + idx) : make_float2(0.f, 0.f);
  for (int k = 0; k < K - 1; ++k) {
    const float wk = __shfl_sync(0xffffffffu, w, k);
    // prefetch the next row before this one is written back one slot earlier
    float2 nxt = make_float2(0.f, 0.f);
    if (k + 1 < K - 1) nxt = *reinterpret_cast<const float2*>(window + (k + 1) * BC + idx);
    acc.x = fmaf(wk, cur.x, acc.x);
    acc.y = fmaf(wk, cur.y, acc.y);
    if (k > 0) *reinterpret_cast<float2*>(window + (k - 1) * BC + idx) = cur;     // shift: row k -> k-1
    cur = nxt;
  }
  const float2 xn = __ldg(reinterpret_cast<const float2*>(x_new + idx));
  const float wl = __shfl_sync(0xffffffffu, w, K - 1);
  acc.x = fmaf(wl, xn.x, acc.x);
  acc.y = fmaf(wl, xn.y, acc.y);
  if (K > 1) *reinterpret_cast<float2*>(window + (K - 2) * BC + idx) = xn;
  *reinterpret_cast<float2*>(out + idx) = acc;
  if (out16) *reinterpret_cast<uint32_t*>(out16 + idx) = pack_bf16(acc.x, acc.y);
}

struct DynConvBwdArgs {
  const float* dout;  // [T,B,C]
  const float* x;     // [T,B,C]
  const float* probs; // [T,B,H,K] (softmax output; or raw weights when softmax == 0)
  float* dx;          // [T,B,C]
  float* dz;          // [T,B,H,K]
  int T, B, C, H, K;
  int softmax;
  float p_drop;
  unsigned long long seed;
  const unsigned long long* step_ptr;
  __nv_bfloat16* dz16;  // optional bf16 twin of dz (operand of the filter projection's backward GEMMs)
  long long lddz16;     // its row pitch (>= H*K, multiple of 8)
};

// dp[t,h,k] = sum_c dout[t,c] x[t-K+1+k,c];  dz = softmax_bwd(dp * mask/(1-q));
// dx[s,c]   = sum_k w[s+K-1-k,h,k] dout[s+K-1-k,c]     (SURVEY 3.4 backward formulas)
__global__ void __launch_bounds__(256)
dynconv_bwd_kernel(DynConvBwdArgs a) {
  pdl_prologue();
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ float xs[DC_ROWS][DC_LD];   // x rows   t0-K+1 .. t0+TT-1
  __shared__ float ds[DC_ROWS][DC_LD];   // dout rows t0 .. t0+TT+K-2
  __shared__ float ws[DC_ROWS][DC_MAXK]; // w rows    t0 .. t0+TT+K-2
  const int R = a.C / a.H;
  const int K = a.K;
  const int bh = blockIdx.x;
  const int b = bh / a.H, h = bh - b * a.H;
  const int t0 = blockIdx.y * DC_TT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = DC_TT + K - 1;
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int i = threadIdx.x; i < nrows * R; i += blockDim.x) {
    const int row = i / R, c = i - row * R;
    const int tx = t0 - (K - 1) + row;
    const int td = t0 + row;
    float v = 0.f, d = 0.f;
    if (tx >= 0 && tx < a.T)
      v = __ldg(a.x + (static_cast<long long>(tx) * a.B + b) * a.C + h * R + c);
    if (td < a.T) d = __ldg(a.dout + (static_cast<long long>(td) * a.B + b) * a.C + h * R + c);
    xs[row][c] = v;
    ds[row][c] = d;
  }
  for (int i = threadIdx.x; i < nrows * K; i += blockDim.x) {
    const int row = i / K, k = i - row * K;
    const int t = t0 + row;
    float w = 0.f;
    if (t < a.T) {
      const long long widx = ((static_cast<long long>(t) * a.B + b) * a.H + h) * K + k;
      w = __ldg(a.probs + widx);
      if (a.p_drop > 0.f)
        w *= dropout_scale(a.seed, static_cast<unsigned long long>(widx), a.p_drop, inv_keep);
    }
    ws[row][k] = w;
  }
  __syncthreads();
  for (int tt = warp; tt < DC_TT; tt += (blockDim.x >> 5)) {
    const int t = t0 + tt;
    if (t >= a.T) break;
    const long long tb = static_cast<long long>(t) * a.B + b;
    // ---- dz: lane k owns tap k
    float dp = 0.f;
    if (lane < K) {
      for (int c = 0; c < R; ++c) dp += ds[tt][c] * xs[tt + lane][c];
    }
    const long long widx = (tb * a.H + h) * K + lane;
    float dw = dp;
    if (a.p_drop > 0.f && lane < K)
      dw *= dropout_scale(a.seed, static_cast<unsigned long long>(widx), a.p_drop, inv_keep);
    float dzk = dw;
    if (a.softmax) {
      const float pk = lane < K ? __ldg(a.probs + widx) : 0.f;
      const float dot = warp_sum(lane < K ? pk * dw : 0.f);
      dzk = pk * (dw - dot);
    }
    if (lane < K) {
      a.dz[widx] = dzk;
      if (a.dz16) a.dz16[tb * a.lddz16 + h * K + lane] = __float2bfloat16_rn(dzk);
    }
    // ---- dx for time s = t
    for (int c = lane; c < R; c += 32) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const int row = tt + (K - 1) - k;  // t' - t0, t' = s + K-1-k
        acc += ws[row][k] * ds[row][c];
      }
      a.dx[tb * a.C + h * R + c] = acc;
    }
  }
}

}  // namespace tt

using namespace tt;

static int dynconv_check(int T, int B, int C, int H, int K) {
  TT_REQUIRE(T >= 0 && B > 0 && C > 0 && H > 0 && K > 0, "dynconv: bad shape");
  TT_REQUIRE(C % H == 0 && C / H <= DC_MAXR, "dynconv: channels per head must be <= %d (got %d)",
             DC_MAXR, C / H);
  TT_REQUIRE(K <= DC_MAXK, "dynconv: kernel size must be <= %d (got %d)", DC_MAXK, K);
  return TT_OK;
}

extern "C" int tt_dynconv_fwd_tw(const float* x, const float* z, long long z_tb_stride, float* out,
                                 float* probs, int T, int B, int C, int H, int K, int softmax,
                                 float p_drop, unsigned long long seed, void* out16, void* stream) {
  TT_REQUIRE(x && z && out, "tt_dynconv_fwd: null pointer");
  int rc = dynconv_check(T, B, C, H, K);
  if (rc != TT_OK) return rc;
  if (T == 0) return TT_OK;
  DynConvArgs a{x, z, z_tb_stride, out, probs, T, B, C, H, K, softmax, p_drop, seed, rng_step_ptr(),
                reinterpret_cast<__nv_bfloat16*>(out16)};
  dim3 grid(B * H, ceil_div(T, DC_TT));
  launch_k(dynconv_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a);
  return check_launch("dynconv_fwd_kernel");
}
extern "C" int tt_dynconv_fwd(const float* x, const float* z, long long z_tb_stride, float* out,
                              float* probs, int T, int B, int C, int H, int K, int softmax,
                              float p_drop, unsigned long long seed, void* stream) {
  return tt_dynconv_fwd_tw(x, z, z_tb_stride, out, probs, T, B, C, H, K, softmax, p_drop, seed, nullptr,
                           stream);
}

extern "C" int tt_dynconv_bwd_tw(const float* dout, const float* x, const float* probs, float* dx,
                                 float* dz, int T, int B, int C, int H, int K, int softmax,
                                 float p_drop, unsigned long long seed, void* dz16, long long lddz16,
                                 void* stream) {
  TT_REQUIRE(dout && x && probs && dx && dz, "tt_dynconv_bwd: null pointer");
  TT_REQUIRE(dz16 == nullptr || lddz16 >= static_cast<long long>(H) * K, "tt_dynconv_bwd: lddz16 < H*K");
  int rc = dynconv_check(T, B, C, H, K);
  if (rc != TT_OK) return rc;
  if (T == 0) return TT_OK;
  DynConvBwdArgs a{dout, x, probs, dx, dz, T, B, C, H, K, softmax, p_drop, seed, rng_step_ptr(),
                   reinterpret_cast<__nv_bfloat16*>(dz16), lddz16};
  dim3 grid(B * H, ceil_div(T, DC_TT));
  launch_k(dynconv_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a);
  return check_launch("dynconv_bwd_kernel");
}
extern "C" int tt_dynconv_bwd(const float* dout, const float* x, const float* probs, float* dx,
                              float* dz, int T, int B, int C, int H, int K, int softmax,
                              float p_drop, unsigned long long seed, void* stream) {
  return tt_dynconv_bwd_tw(dout, x, probs, dx, dz, T, B, C, H, K, softmax, p_drop, seed, nullptr, 0, stream);
}

extern "C" int tt_dynconv_step_tw(float* window, const float* x_new, const float* z, long long z_b_stride,
                                  float* out, int B, int C, int H, int K, int softmax, void* out16v,
                                  void* stream) {
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(out16v);
  TT_REQUIRE(x_new && z && out && (window || K == 1), "tt_dynconv_step: null pointer");
  int rc = dynconv_check(1, B, C, H, K);
  if (rc != TT_OK) return rc;
  const long long BC = static_cast<long long>(B) * C;
  if (BC == 0) return TT_OK;
  if (C == H * 64 && K <= 32 && (reinterpret_cast<uintptr_t>(x_new) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 7) == 0 && (reinterpret_cast<uintptr_t>(window) & 7) == 0) {
    const long long warps = static_cast<long long>(B) * H;
    launch_k(dynconv_step_warp_kernel, dim3(static_cast<unsigned>(ceil_div_ll(warps, 8))), dim3(256), 0,
             (cudaStream_t)stream, window, x_new, z, z_b_stride, out, out16, B, C, H, K, softmax);
    return check_launch("dynconv_step_warp_kernel");
  }
  launch_k(dynconv_step_kernel, dim3(static_cast<unsigned>(ceil_div_ll(BC, 256))), dim3(256), 0,
           (cudaStream_t)stream, window, x_new, z, z_b_stride, out, out16, B, C, H, K, softmax);
  return check_launch("dynconv_step_kernel");
}
extern "C" int tt_dynconv_step(float* window, const float* x_new, const float* z, long long z_b_stride,
                               float* out, int B, int C, int H, int K, int softmax, void* stream) {
  return tt_dynconv_step_tw(window, x_new, z, z_b_stride, out, B, C, H, K, softmax, nullptr, stream);
}
