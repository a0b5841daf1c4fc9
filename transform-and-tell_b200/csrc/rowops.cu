// Row-wise HBM-bound kernels of the decoder layer: residual + dropout + LayerNorm (fwd/bwd), GLU
// (fwd/bwd), dropout, weight-norm (fwd/bwd), NaN-row masking, vector helpers.
// One warp per row, float4 accesses, warp-shuffle reductions, no shared memory.
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA

static inline int row_grid(long long rows) {
  long long g = ceil_div_ll(rows, ROW_WARPS);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// x = res + dropout(h) (written back into h), y = LN(x) * gamma + beta.
// decoder_faces_objects.py:263-266, :283-287, :361-364 (dropout, residual add, post-LayerNorm).
__global__ void ln_fwd_kernel(float* __restrict__ h, const float* __restrict__ res,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ y, long long ldy, float* __restrict__ mean,
                              float* __restrict__ rstd, int N, int E, float eps, float p,
                              unsigned long long seed, const unsigned long long* step_ptr) {
  pdl_prologue();
  seed = mix_seed(seed, step_ptr);
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int E4 = E >> 2;
  for (int r = warp_global; r < N; r += nwarps) {
    float4* hr = reinterpret_cast<float4*>(h + static_cast<long long>(r) * E);
    const float4* rr = res ? reinterpret_cast<const float4*>(res + static_cast<long long>(r) * E)
                           : nullptr;
    float s = 0.f;
    for (int c = lane; c < E4; c += 32) {
      float4 v = hr[c];
      if (p > 0.f) {
        const unsigned long long base = static_cast<unsigned long long>(r) * E + 4ull * c;
        v.x *= dropout_scale(seed, base, p, inv_keep);
        v.y *= dropout_scale(seed, base + 1, p, inv_keep);
        v.z *= dropout_scale(seed, base + 2, p, inv_keep);
        v.w *= dropout_scale(seed, base + 3, p, inv_keep);
      }
      if (rr) {
        const float4 q = __ldg(rr + c);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      hr[c] = v;
      s += (v.x + v.y) + (v.z + v.w);
    }
    s = warp_sum(s);
    const float mu = s / E;
    float ss = 0.f;
    for (int c = lane; c < E4; c += 32) {
      const float4 v = hr[c];
      const float a = v.x - mu, b = v.y - mu, d = v.z - mu, e = v.w - mu;
      ss += (a * a + b * b) + (d * d + e * e);
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / E + eps);
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
    float4* yr = reinterpret_cast<float4*>(y + static_cast<long long>(r) * ldy);
    for (int c = lane; c < E4; c += 32) {
      const float4 v = hr[c];
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v.x - mu) * rs * gm.x + bt.x;
      o.y = (v.y - mu) * rs * gm.y + bt.y;
      o.z = (v.z - mu) * rs * gm.z + bt.z;
      o.w = (v.w - mu) * rs * gm.w + bt.w;
      yr[c] = o;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ; dh = dx * dropmask/(1-p).
// dgamma/dbeta are accumulated with one atomicAdd per (CTA-warp, column) at the end.
template <int MAXC>  // float4 chunks per lane: E <= 128 * MAXC
__global__ void ln_bwd_kernel(const float* __restrict__ dy, long long lddy,
                              const float* __restrict__ x, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma,
                              float* __restrict__ dx, float* __restrict__ dh,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int E,
                              float p, unsigned long long seed,
                              const unsigned long long* step_ptr) {
  pdl_prologue();
  seed = mix_seed(seed, step_ptr);
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int E4 = E >> 2;
  float4 acc_g[MAXC], acc_b[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = warp_global; r < N; r += nwarps) {
    const float4* dyr = reinterpret_cast<const float4*>(dy + static_cast<long long>(r) * lddy);
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * E);
    const float mu = mean[r], rs = rstd[r];
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
        xh.x = (xv.x - mu) * rs; xh.y = (xv.y - mu) * rs;
        xh.z = (xv.z - mu) * rs; xh.w = (xv.w - mu) * rs;
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
        o.x = rs * (g4[i].x - s1 - xh4[i].x * s2);
        o.y = rs * (g4[i].y - s1 - xh4[i].y * s2);
        o.z = rs * (g4[i].z - s1 - xh4[i].z * s2);
        o.w = rs * (g4[i].w - s1 - xh4[i].w * s2);
        if (dx) reinterpret_cast<float4*>(dx + static_cast<long long>(r) * E)[c] = o;
        if (dh) {
          if (p > 0.f) {
            const unsigned long long base = static_cast<unsigned long long>(r) * E + 4ull * c;
            o.x *= dropout_scale(seed, base, p, inv_keep);
            o.y *= dropout_scale(seed, base + 1, p, inv_keep);
            o.z *= dropout_scale(seed, base + 2, p, inv_keep);
            o.w *= dropout_scale(seed, base + 3, p, inv_keep);
          }
          reinterpret_cast<float4*>(dh + static_cast<long long>(r) * E)[c] = o;
        }
      }
    }
  }
  // CTA-level reduction: every warp parks its register accumulators in shared memory (plain
  // float4 stores, one quantity at a time), the CTA adds the 8 copies and issues ONE global atomic
  // per column.
  extern __shared__ float ln_red[];   // [ROW_WARPS][E]
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float* dst = q == 0 ? dgamma : dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < E4) reinterpret_cast<float4*>(ln_red + warp * E)[c] = q == 0 ? acc_g[i] : acc_b[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) t += ln_red[w * E + i];
      atomicAdd(dst + i, t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// GLU over the last dim: out[n,c] = h[n,c] * sigmoid(h[n,C+c])   (nn.GLU, decoder_faces_objects.py:193-195,259-260)
__global__ void glu_fwd_kernel(const float* __restrict__ h, float* __restrict__ out, long long n4,
                               int C4) {
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
  }
}
__global__ void glu_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                               float* __restrict__ dh, long long n4, int C4) {
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
  }
}

// ------------------------------------------------------------------------------------------------
// y = x * dropmask/(1-p)  (F.dropout; same call regenerates the mask for the backward).
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                               float p, unsigned long long seed,
                               const unsigned long long* step_ptr) {
  pdl_prologue();
  seed = mix_seed(seed, step_ptr);
  const float inv_keep = 1.f / (1.f - p);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, static_cast<unsigned long long>(i), p, inv_keep);
}

// y = a*x + b*y (vector), used for the RoBERTa layer mix and gradient accumulation.
__global__ void axpby_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                             float a, float b) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}

// ------------------------------------------------------------------------------------------------
// Weight norm (nn.utils.weight_norm, dim=0):  w[o,:] = g[o] * v[o,:] / ||v[o,:]||   linear.py:30-34
__global__ void wnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                 float* __restrict__ w, float* __restrict__ norm, int O, int I) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int o = warp_global; o < O; o += nwarps) {
    const float* vr = v + static_cast<long long>(o) * I;
    float s = 0.f;
    for (int i = lane; i < I; i += 32) s += vr[i] * vr[i];
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    const float sc = g[o] / nrm;
    if (lane == 0 && norm) norm[o] = nrm;
    float* wr = w + static_cast<long long>(o) * I;
    for (int i = lane; i < I; i += 32) wr[i] = vr[i] * sc;
  }
}
// dg[o] = <dw,v>/||v|| ; dv = g/||v|| * (dw - v <dw,v>/||v||^2)
__global__ void wnorm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                 const float* __restrict__ g, const float* __restrict__ norm,
                                 float* __restrict__ dv, float* __restrict__ dg, int O, int I) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int o = warp_global; o < O; o += nwarps) {
    const float* vr = v + static_cast<long long>(o) * I;
    const float* dr = dw + static_cast<long long>(o) * I;
    float s = 0.f;
    for (int i = lane; i < I; i += 32) s += dr[i] * vr[i];
    s = warp_sum(s);
    const float nrm = norm[o];
    const float gg = g[o];
    if (lane == 0) dg[o] = s / nrm;
    const float a = gg / nrm, b = s / (nrm * nrm);
    float* o_ = dv + static_cast<long long>(o) * I;
    for (int i = lane; i < I; i += 32) o_[i] = a * (dr[i] - vr[i] * b);
  }
}

// ------------------------------------------------------------------------------------------------
// mask[r] = any(isnan(x[r,:])); NaN rows are zeroed in place.  transformer_faces_objects.py:374-379
__global__ void nan_rows_kernel(float* __restrict__ x, uint8_t* __restrict__ mask, int R, int D) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int r = warp_global; r < R; r += nwarps) {
    float* xr = x + static_cast<long long>(r) * D;
    int bad = 0;
    for (int i = lane; i < D; i += 32) bad |= (xr[i] != xr[i]) ? 1 : 0;
    bad = __any_sync(0xffffffffu, bad);
    if (bad)
      for (int i = lane; i < D; i += 32) xr[i] = 0.f;
    if (lane == 0) mask[r] = bad ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Wide-row variants (E = 1024 with 256 threads): ONE CTA PER ROW, the row lives in registers (one
// float4 per thread), block reductions through shared memory.  The decoder has only T*B = 800
// rows: warp-per-row left 5 warps per SM and a chain of 8 dependent loads per pass; here every row
// is a single round of memory latency and all SMs are busy.
__device__ __forceinline__ float2 block_sum2_256(float a, float b, float2* red) {
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

__global__ void __launch_bounds__(256) ln_fwd_row_kernel(
    float* __restrict__ h, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ y, long long ldy, float* __restrict__ mean,
    float* __restrict__ rstd, int N, float eps, float p, unsigned long long seed,
    const unsigned long long* step_ptr) {
  pdl_prologue();
  __shared__ float2 red[8];
  constexpr int E = 1024;
  seed = mix_seed(seed, step_ptr);
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int c = threadIdx.x;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
  for (int r = blockIdx.x; r < N; r += gridDim.x) {
    float4* hr = reinterpret_cast<float4*>(h + static_cast<long long>(r) * E);
    float4 v = hr[c];
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) q = __ldg(reinterpret_cast<const float4*>(res + static_cast<long long>(r) * E) + c);
    if (p > 0.f) {
      const unsigned long long base = static_cast<unsigned long long>(r) * E + 4ull * c;
      v.x *= dropout_scale(seed, base, p, inv_keep);
      v.y *= dropout_scale(seed, base + 1, p, inv_keep);
      v.z *= dropout_scale(seed, base + 2, p, inv_keep);
      v.w *= dropout_scale(seed, base + 3, p, inv_keep);
    }
    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    hr[c] = v;
    const float mu = block_sum2_256((v.x + v.y) + (v.z + v.w), 0.f, red).x * (1.f / E);
    const float a = v.x - mu, b = v.y - mu, d = v.z - mu, e = v.w - mu;
    const float var = block_sum2_256((a * a + b * b) + (d * d + e * e), 0.f, red).x * (1.f / E);
    const float rs = rsqrtf(var + eps);
    if (c == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
    float4 o;
    o.x = a * rs * gm.x + bt.x;
    o.y = b * rs * gm.y + bt.y;
    o.z = d * rs * gm.z + bt.z;
    o.w = e * rs * gm.w + bt.w;
    reinterpret_cast<float4*>(y + static_cast<long long>(r) * ldy)[c] = o;
  }
}

static inline int flat_grid(long long n) {
  long long g = ceil_div_ll(n, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace tt

using namespace tt;

extern "C" int tt_ln_fwd(float* h, const float* res, const float* gamma, const float* beta,
                         float* y, long long ldy, float* mean, float* rstd, int N, int E, float eps,
                         float p_drop, unsigned long long seed, void* stream) {
  TT_REQUIRE(h && gamma && beta && y, "tt_ln_fwd: null pointer");
  TT_REQUIRE(E > 0 && E % 4 == 0 && ldy % 4 == 0, "tt_ln_fwd: E and ldy must be multiples of 4");
  TT_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "tt_ln_fwd: bad dropout p");
  if (N <= 0) return TT_OK;
  if (E == 1024) {      // production width: one CTA per row
    const int grid = N < num_sms() * 8 ? N : num_sms() * 8;
    launch_k(ln_fwd_row_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
        h, res, gamma, beta, y, ldy, mean, rstd, N, eps, p_drop, seed, rng_step_ptr());
    return check_launch("ln_fwd_row_kernel");
  }
  if (E < 1024) {       // same CTA-per-row arithmetic at every width (twin.cu), so that the single and
    TtLnFwdMulti a{};   // the multi-context launches agree bit for bit
    a.h[0] = h; a.gamma[0] = gamma; a.beta[0] = beta; a.mean[0] = mean; a.rstd[0] = rstd; a.seed[0] = seed;
    a.res = res; a.y = y; a.ldy = ldy; a.n = 1; a.N = N; a.E = E; a.eps = eps; a.p_drop = p_drop;
    return tt_ln_fwd_multi(&a, stream);
  }
  launch_k(ln_fwd_kernel, dim3(row_grid(N)), dim3(ROW_WARPS * 32), 0, (cudaStream_t)stream, 
      h, res, gamma, beta, y, ldy, mean, rstd, N, E, eps, p_drop, seed, rng_step_ptr());
  return check_launch("ln_fwd_kernel");
}

extern "C" int tt_ln_bwd(const float* dy, long long lddy, const float* x, const float* mean,
                         const float* rstd, const float* gamma, float* dx, float* dh,
                         float* dgamma, float* dbeta, int N, int E, float p_drop,
                         unsigned long long seed, void* stream) {
  TT_REQUIRE(dy && x && mean && rstd && gamma, "tt_ln_bwd: null pointer");
  TT_REQUIRE(E > 0 && E % 4 == 0 && E <= 1024 && lddy % 4 == 0,
             "tt_ln_bwd: E must be a multiple of 4 and <= 1024 (got %d)", E);
  if (N <= 0) return TT_OK;
  // fewer, fatter warps so the dgamma/dbeta atomics stay cheap
  int grid = row_grid(N);
  if (grid > num_sms()) grid = num_sms();
  const size_t smem = ROW_WARPS * static_cast<size_t>(E) * sizeof(float);
  if (E <= 256)
    launch_k(ln_bwd_kernel<2>, dim3(grid), dim3(ROW_WARPS * 32), smem, (cudaStream_t)stream, 
        dy, lddy, x, mean, rstd, gamma, dx, dh, dgamma, dbeta, N, E, p_drop, seed, rng_step_ptr());
  else
    launch_k(ln_bwd_kernel<8>, dim3(grid), dim3(ROW_WARPS * 32), smem, (cudaStream_t)stream, 
        dy, lddy, x, mean, rstd, gamma, dx, dh, dgamma, dbeta, N, E, p_drop, seed, rng_step_ptr());
  return check_launch("ln_bwd_kernel");
}

extern "C" int tt_glu_fwd(const float* h, float* out, long long N, int C, void* stream) {
  TT_REQUIRE(h && out, "tt_glu_fwd: null pointer");
  TT_REQUIRE(C > 0 && C % 4 == 0, "tt_glu_fwd: C must be a multiple of 4");
  if (N <= 0) return TT_OK;
  const long long n4 = N * (C / 4);
  launch_k(glu_fwd_kernel, dim3(flat_grid(n4)), dim3(256), 0, (cudaStream_t)stream, h, out, n4, C / 4);
  return check_launch("glu_fwd_kernel");
}

extern "C" int tt_glu_bwd(const float* dout, const float* h, float* dh, long long N, int C,
                          void* stream) {
  TT_REQUIRE(dout && h && dh, "tt_glu_bwd: null pointer");
  TT_REQUIRE(C > 0 && C % 4 == 0, "tt_glu_bwd: C must be a multiple of 4");
  if (N <= 0) return TT_OK;
  const long long n4 = N * (C / 4);
  launch_k(glu_bwd_kernel, dim3(flat_grid(n4)), dim3(256), 0, (cudaStream_t)stream, dout, h, dh, n4, C / 4);
  return check_launch("glu_bwd_kernel");
}

extern "C" int tt_dropout(const float* x, float* y, long long n, float p,
                          unsigned long long seed, void* stream) {
  TT_REQUIRE(x && y, "tt_dropout: null pointer");
  TT_REQUIRE(p >= 0.f && p < 1.f, "tt_dropout: bad p");
  if (n <= 0) return TT_OK;
  launch_k(dropout_kernel, dim3(flat_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, y, n, p, seed, rng_step_ptr());
  return check_launch("dropout_kernel");
}

extern "C" int tt_axpby(const float* x, float* y, long long n, float a, float b, void* stream) {
  TT_REQUIRE(x && y, "tt_axpby: null pointer");
  if (n <= 0) return TT_OK;
  launch_k(axpby_kernel, dim3(flat_grid(n)), dim3(256), 0, (cudaStream_t)stream, x, y, n, a, b);
  return check_launch("axpby_kernel");
}

extern "C" int tt_wnorm_fwd(const float* v, const float* g, float* w, float* norm, int O, int I,
                            void* stream) {
  TT_REQUIRE(v && g && w, "tt_wnorm_fwd: null pointer");
  if (O <= 0 || I <= 0) return TT_OK;
  launch_k(wnorm_fwd_kernel, dim3(row_grid(O)), dim3(ROW_WARPS * 32), 0, (cudaStream_t)stream, v, g, w, norm, O, I);
  return check_launch("wnorm_fwd_kernel");
}

extern "C" int tt_wnorm_bwd(const float* dw, const float* v, const float* g, const float* norm,
                            float* dv, float* dg, int O, int I, void* stream) {
  TT_REQUIRE(dw && v && g && norm && dv && dg, "tt_wnorm_bwd: null pointer");
  if (O <= 0 || I <= 0) return TT_OK;
  launch_k(wnorm_bwd_kernel, dim3(row_grid(O)), dim3(ROW_WARPS * 32), 0, (cudaStream_t)stream, dw, v, g, norm, dv, dg,
                                                                           O, I);
  return check_launch("wnorm_bwd_kernel");
}

extern "C" int tt_nan_rows(float* x, uint8_t* mask, int R, int D, void* stream) {
  TT_REQUIRE(x && mask, "tt_nan_rows: null pointer");
  if (R <= 0) return TT_OK;
  if (D <= 0) {
    cudaMemsetAsync(mask, 0, R, (cudaStream_t)stream);
    return TT_OK;
  }
  launch_k(nan_rows_kernel, dim3(row_grid(R)), dim3(ROW_WARPS * 32), 0, (cudaStream_t)stream, x, mask, R, D);
  return check_launch("nan_rows_kernel");
}

// ------------------------------------------------------------------------------------------------
namespace tt {
// out[c] += scale * sum_r x[r,c]   (bias gradients).  HBM-bound: x is read once (4 B/element).
// Vector kernel: a CTA owns a 128-column strip (32 float4 lanes) x a slab of rows; its 8 row lanes
// keep 4 independent float4 accumulators in flight, partial sums are combined in shared memory
// and ONE atomicAdd per column per CTA goes to `out` (zeroed first unless accumulating).  The row
// slab is sized so that the grid has ~4 CTAs per SM: the old 32 x 256 decomposition issued 32
// dependent loads per thread from 128 CTAs and sat at 17 us whatever the matrix size.
constexpr int COLSUM_ROWS = 256;
__global__ void colsum_kernel(const float* __restrict__ x, long long ld, int M, int N,
                              float* __restrict__ out, float scale) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * COLSUM_ROWS;
  const int r1 = min(M, r0 + COLSUM_ROWS);
  float s = 0.f;
  if (c < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += x[r * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t * scale);
  }
}
__global__ void __launch_bounds__(256) colsum_vec_kernel(const float* __restrict__ x, long long ld, int M,
                                                         int N, float* __restrict__ out, float scale,
                                                         int rows_per_cta) {
  pdl_prologue();
  __shared__ float4 red[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  if (c < N) {
    const float* xc = x + c;
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + (long long)r * ld));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(xc + (long long)(r + 8) * ld));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(xc + (long long)(r + 16) * ld));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(xc + (long long)(r + 24) * ld));
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < r1; r += 8) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + (long long)r * ld));
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = make_float4(a0.x + a1.x + a2.x + a3.x, a0.y + a1.y + a2.y + a3.y,
                                              a0.z + a1.z + a2.z + a3.z, a0.w + a1.w + a2.w + a3.w);
  __syncthreads();
  // 128 columns of the strip: thread (ty, tx) with ty < 4 reduces component ty of lane tx
  if (threadIdx.y < 4 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += reinterpret_cast<const float*>(&red[i][threadIdx.x])[threadIdx.y];
    atomicAdd(out + c + threadIdx.y, t * scale);
  }
}
// Same tiling for a bf16 matrix (the bf16 dL/d(k|v) slab of the cross-attention backward): 4 bf16
// (8 bytes) per thread per row, fp32 accumulation.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                          int M, int N, float* __restrict__ out,
                                                          float scale, int rows_per_cta) {
  pdl_prologue();
  __shared__ float4 red[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  auto add = [](float4& a, uint2 u) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    a.x += __low2float(h[0]); a.y += __high2float(h[0]);
    a.z += __low2float(h[1]); a.w += __high2float(h[1]);
  };
  if (c < N) {
    const __nv_bfloat16* xc = x + c;
    int r = r0 + threadIdx.y;
    for (; r + 8 < r1; r += 16) {
      const uint2 v0 = __ldg(reinterpret_cast<const uint2*>(xc + (long long)r * ld));
      const uint2 v1 = __ldg(reinterpret_cast<const uint2*>(xc + (long long)(r + 8) * ld));
      add(a0, v0);
      add(a1, v1);
    }
    for (; r < r1; r += 8) add(a0, __ldg(reinterpret_cast<const uint2*>(xc + (long long)r * ld)));
  }
  red[threadIdx.y][threadIdx.x] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
  __syncthreads();
  if (threadIdx.y < 4 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += reinterpret_cast<const float*>(&red[i][threadIdx.x])[threadIdx.y];
    atomicAdd(out + c + threadIdx.y, t * scale);
  }
}
// dx = dy * (y > 0)
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                float* __restrict__ dx, long long n) {
  pdl_prologue();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// transformer_faces_objects.py:355-364: out = sum_l softmax(w)[l] * hiddens[l]   (bf16 hiddens).
// 8 elements (16 B of bf16) per thread per layer; n must be a multiple of 8.
__global__ void layer_mix_fwd_kernel(const __nv_bfloat16* __restrict__ hid, long long layer_stride,
                                     const float* __restrict__ w, int L, long long n8,
                                     float* __restrict__ out) {
  pdl_prologue();
  __shared__ float sw[64];
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int l = 0; l < L; ++l) m = fmaxf(m, w[l]);
    float s = 0.f;
    for (int l = 0; l < L; ++l) { sw[l] = expf(w[l] - m); s += sw[l]; }
    for (int l = 0; l < L; ++l) sw[l] /= s;
  }
  __syncthreads();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int l = 0; l < L; ++l) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(hid + l * layer_stride) + i);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
      const float wl = sw[l];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += wl * __low2float(h2[j]);
        acc[2 * j + 1] += wl * __high2float(h2[j]);
      }
    }
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}
// dots[l] += sum_i dout[i] * hiddens[l][i]: dout is read ONCE, the L partial sums live in registers.
template <int MAXL>
__global__ void layer_mix_bwd_kernel(const __nv_bfloat16* __restrict__ hid, long long layer_stride,
                                     const float* __restrict__ dout, int L, long long n8,
                                     float* __restrict__ dots) {
  pdl_prologue();
  __shared__ float red[32];
  float acc[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) acc[l] = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dout) + 2 * i);
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(dout) + 2 * i + 1);
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
      if (l < L) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(hid + l * layer_stride) + i);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
        acc[l] += d0.x * __low2float(h2[0]) + d0.y * __high2float(h2[0]) +
                  d0.z * __low2float(h2[1]) + d0.w * __high2float(h2[1]) +
                  d1.x * __low2float(h2[2]) + d1.y * __high2float(h2[2]) +
                  d1.z * __low2float(h2[3]) + d1.w * __high2float(h2[3]);
      }
    }
  }
  for (int l = 0; l < L && l < MAXL; ++l) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j)
      if (j == l) v = acc[j];
    v = block_sum(v, red);
    if (threadIdx.x == 0) atomicAdd(dots + l, v);
  }
}
// dw[l] = p[l] * (dots[l] - sum_j p[j] dots[j]),  p = softmax(w)
__global__ void softmax_bwd_small_kernel(const float* __restrict__ w, const float* __restrict__ dots,
                                         int L, float* __restrict__ dw) {
  pdl_prologue();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int l = 0; l < L; ++l) m = fmaxf(m, w[l]);
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += expf(w[l] - m);
    float dot = 0.f;
    for (int l = 0; l < L; ++l) dot += expf(w[l] - m) / s * dots[l];
    for (int l = 0; l < L; ++l) dw[l] = expf(w[l] - m) / s * (dots[l] - dot);
  }
}
}  // namespace tt

extern "C" int tt_colsum(const float* x, long long ld, int M, int N, float* out, float scale,
                         int accumulate, void* stream) {
  TT_REQUIRE(x && out, "tt_colsum: null pointer");
  if (N <= 0) return TT_OK;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, (cudaStream_t)stream);
  if (M <= 0) return TT_OK;
  const bool vec = (N % 4 == 0) && (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec) {
    const int strips = ceil_div(N, 128);
    int slabs = ceil_div(4 * num_sms(), strips);            // ~4 CTAs per SM
    const int max_slabs = ceil_div(M, 32);                  // at least 32 rows (4 per row lane) each
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    int rows_per_cta = ceil_div(M, slabs);
    rows_per_cta = ceil_div(rows_per_cta, 8) * 8;
    slabs = ceil_div(M, rows_per_cta);
    launch_k(colsum_vec_kernel, dim3(strips, slabs), dim3(32, 8), 0, (cudaStream_t)stream, x, ld, M, N, out,
             scale, rows_per_cta);
    return check_launch("colsum_vec_kernel");
  }
  dim3 block(32, 8), grid(ceil_div(N, 32), ceil_div(M, COLSUM_ROWS));
  launch_k(colsum_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, x, ld, M, N, out, scale);
  return check_launch("colsum_kernel");
}

extern "C" int tt_colsum_bf16(const void* x, long long ld, int M, int N, float* out, float scale,
                              int accumulate, void* stream) {
  TT_REQUIRE(x && out, "tt_colsum_bf16: null pointer");
  TT_REQUIRE(N % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
             "tt_colsum_bf16: N and ld must be multiples of 4, x 8-byte aligned");
  if (N <= 0) return TT_OK;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, (cudaStream_t)stream);
  if (M <= 0) return TT_OK;
  const int strips = ceil_div(N, 128);
  int slabs = ceil_div(4 * num_sms(), strips);
  const int max_slabs = ceil_div(M, 32);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  int rows_per_cta = ceil_div(M, slabs);
  rows_per_cta = ceil_div(rows_per_cta, 8) * 8;
  slabs = ceil_div(M, rows_per_cta);
  launch_k(colsum_bf16_kernel, dim3(strips, slabs), dim3(32, 8), 0, (cudaStream_t)stream,
           reinterpret_cast<const __nv_bfloat16*>(x), ld, M, N, out, scale, rows_per_cta);
  return check_launch("colsum_bf16_kernel");
}

namespace tt {
__global__ void scalar_mul_kernel(const float* a, const float* b, float* out) {
  pdl_prologue(); *out = *a * *b; }
}  // namespace tt
extern "C" int tt_scalar_mul(const float* a, const float* b, float* out, void* stream) {
  TT_REQUIRE(a && b && out, "tt_scalar_mul: null pointer");
  launch_k(scalar_mul_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, a, b, out);
  return check_launch("scalar_mul_kernel");
}

extern "C" int tt_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream) {
  TT_REQUIRE(dy && y && dx, "tt_relu_bwd: null pointer");
  if (n <= 0) return TT_OK;
  launch_k(relu_bwd_kernel, dim3(flat_grid(n)), dim3(256), 0, (cudaStream_t)stream, dy, y, dx, n);
  return check_launch("relu_bwd_kernel");
}

extern "C" int tt_layer_mix_fwd(const void* hiddens, long long layer_stride, const float* w, int L,
                                long long n, float* out, void* stream) {
  TT_REQUIRE(hiddens && w && out, "tt_layer_mix_fwd: null pointer");
  TT_REQUIRE(L > 0 && L <= 64, "tt_layer_mix_fwd: L must be in [1,64]");
  TT_REQUIRE(n % 8 == 0 && layer_stride % 8 == 0, "tt_layer_mix_fwd: n and layer_stride must be multiples of 8");
  if (n <= 0) return TT_OK;
  launch_k(layer_mix_fwd_kernel, dim3(flat_grid(n / 8)), dim3(256), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(hiddens), layer_stride, w, L, n / 8, out);
  return check_launch("layer_mix_fwd_kernel");
}

extern "C" int tt_layer_mix_bwd(const void* hiddens, long long layer_stride, const float* w,
                                const float* dout, int L, long long n, float* dots, float* dw,
                                void* stream) {
  TT_REQUIRE(hiddens && w && dout && dots && dw, "tt_layer_mix_bwd: null pointer");
  TT_REQUIRE(L > 0 && L <= 32, "tt_layer_mix_bwd: L must be in [1,32]");
  cudaMemsetAsync(dots, 0, sizeof(float) * L, (cudaStream_t)stream);
  TT_REQUIRE(n % 8 == 0 && layer_stride % 8 == 0, "tt_layer_mix_bwd: n and layer_stride must be multiples of 8");
  if (n > 0) {
    int grid = flat_grid(n / 8);
    if (grid > 4 * num_sms()) grid = 4 * num_sms();
    launch_k(layer_mix_bwd_kernel<32>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(hiddens), layer_stride, dout, L, n / 8, dots);
    int rc = check_launch("layer_mix_bwd_kernel");
    if (rc != TT_OK) return rc;
  }
  launch_k(softmax_bwd_small_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, w, dots, L, dw);
  return check_launch("softmax_bwd_small_kernel");
}
