// Tensor-core version of the decoder cross-attention (same math and argument block as
// attention.cu; multi_head.py:355-466): bf16 operands on mma.sync.m16n8k16 with fp32 accumulation,
// fp32 softmax statistics.  Used in the 'bf16' throughput mode; the fp32 SIMT kernels of
// attention.cu remain the parity-mode ('bf16x3') path.  D = 64.
//   forward : CTA = (b, h, 64 queries), 4 warps x 16 query rows, loop over 64-key tiles
//   dQ      : same tiling; S = QK^T, dP = dO V^T, dS = P*(dP*drop - D), dQ += dS K
//   dK/dV   : CTA = (b, h, 64 keys), 4 warps x 16 key rows, loop over 64-query tiles using the
//             transposed products S^T = K Q^T, dP^T = V dO^T so that P^T / dS^T are A operands:
//             dV += P^T dO, dK += dS^T Q
// q/k/v/dO arrive as fp32 (row strides honoured) and are converted to bf16 while being staged in
// shared memory; outputs are fp32.
#include "attention_args.cuh"
#include "common.cuh"
#include "runtime.h"

namespace tt {

#ifndef TT_ATTN_MINB
#define TT_ATTN_MINB 2      // CTAs per SM the tensor-core attention kernels are compiled for (3 / 4 spill: measured slower)
#endif
constexpr int TC_BM = 64, TC_BN = 64, TC_D = 64, TC_LD = TC_D + 8;

constexpr int TC_THREADS = 128;   // every kernel here runs 4 warps
constexpr int TC_ST = 64 * 16 / TC_THREADS;   // float4 loads per thread per 64x64 tile

// rows [r0, r0+64) x 64 dims of an fp32 matrix -> registers (rows >= nrows are zero).
// src row r lives at base + (r * B + b) * ld + h*64.  All TC_ST loads are issued back to back.
__device__ __forceinline__ void load_tile(const float* __restrict__ base, long long ld, int B, int b,
                                          int h, int r0, int nrows, float4 (&reg)[TC_ST]) {
#pragma unroll
  for (int it = 0; it < TC_ST; ++it) {
    const int i = threadIdx.x + it * TC_THREADS;
    const int r = i >> 4, c4 = i & 15;
    reg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < nrows)
      reg[it] = __ldg(reinterpret_cast<const float4*>(base + (static_cast<long long>(r0 + r) * B + b) * ld +
                                                      h * TC_D) + c4);
  }
}
__device__ __forceinline__ void store_tile(const float4 (&reg)[TC_ST], __nv_bfloat16 (*dst)[TC_LD]) {
#pragma unroll
  for (int it = 0; it < TC_ST; ++it) {
    const int i = threadIdx.x + it * TC_THREADS;
    const int r = i >> 4, c4 = i & 15;
    uint2 u;
    u.x = pack_bf16(reg[it].x, reg[it].y);
    u.y = pack_bf16(reg[it].z, reg[it].w);
    *reinterpret_cast<uint2*>(&dst[r][c4 * 4]) = u;
  }
}
__device__ __forceinline__ void stage_tile(const float* __restrict__ base, long long ld, int B, int b,
                                           int h, int r0, int nrows,
                                           __nv_bfloat16 (*dst)[TC_LD]) {
  float4 reg[TC_ST];
  load_tile(base, ld, B, b, h, r0, nrows, reg);
  store_tile(reg, dst);
}

// Key/value tile j0..j0+63 of the extended key set [ctx rows ; bias row ; zero row] + additive mask,
// split into a load phase (global -> registers; lets the caller prefetch the next tile while the
// tensor cores work on the current one) and a store phase (registers -> bf16 smem).
// KV16: the projected keys|values live in HBM as bf16 (the K|V projection GEMM writes bf16
// directly): half the bytes of this HBM-bound kernel, no conversion on the way to shared memory.
template <bool KV16> struct KvRegs {
  float4 k[TC_ST], v[TC_ST];
  float mask;     // threads 0..63: additive mask of key j0 + threadIdx.x
};
template <> struct KvRegs<true> {
  uint4 k[TC_ST / 2], v[TC_ST / 2];    // 8 bf16 per 16-byte chunk
  float mask;
};
__device__ __forceinline__ float kv_mask(const AttnArgs& a, int b, int j0, int L) {
  float m = 0.f;
  if (threadIdx.x < 64) {
    const int j = j0 + threadIdx.x;
    bool ok;
    if (j < a.S) ok = !(a.mask && a.mask[static_cast<long long>(b) * a.S + j]);
    else ok = j < L;
    m = ok ? 0.f : -INFINITY;
  }
  return m;
}
__device__ __forceinline__ void load_kv(const AttnArgs& a, int b, int h, int j0, int L, KvRegs<false>& R) {
  const bool has_bias = a.bias_k != nullptr;
#pragma unroll
  for (int it = 0; it < TC_ST; ++it) {
    const int i = threadIdx.x + it * TC_THREADS;
    const int r = i >> 4, c4 = i & 15;
    const int j = j0 + r;
    R.k[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    R.v[it] = R.k[it];
    if (j < a.S) {
      const long long off = (static_cast<long long>(j) * a.B + b) * a.ldkv + h * TC_D;
      R.k[it] = __ldg(reinterpret_cast<const float4*>(a.k + off) + c4);
      R.v[it] = __ldg(reinterpret_cast<const float4*>(a.v + off) + c4);
    } else if (has_bias && j == a.S) {
      R.k[it] = __ldg(reinterpret_cast<const float4*>(a.bias_k + h * TC_D) + c4);
      R.v[it] = __ldg(reinterpret_cast<const float4*>(a.bias_v + h * TC_D) + c4);
    }
  }
  R.mask = kv_mask(a, b, j0, L);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* p) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(p));
  const float4 y = __ldg(reinterpret_cast<const float4*>(p) + 1);
  uint4 u;
  u.x = pack_bf16(x.x, x.y); u.y = pack_bf16(x.z, x.w);
  u.z = pack_bf16(y.x, y.y); u.w = pack_bf16(y.z, y.w);
  return u;
}
__device__ __forceinline__ void load_kv(const AttnArgs& a, int b, int h, int j0, int L, KvRegs<true>& R) {
  const bool has_bias = a.bias_k != nullptr;
  const __nv_bfloat16* k16 = reinterpret_cast<const __nv_bfloat16*>(a.k);
  const __nv_bfloat16* v16 = reinterpret_cast<const __nv_bfloat16*>(a.v);
#pragma unroll
  for (int it = 0; it < TC_ST / 2; ++it) {
    const int i = threadIdx.x + it * TC_THREADS;
    const int r = i >> 3, c8 = i & 7;
    const int j = j0 + r;
    R.k[it] = make_uint4(0u, 0u, 0u, 0u);
    R.v[it] = R.k[it];
    if (j < a.S) {
      const long long off = (static_cast<long long>(j) * a.B + b) * a.ldkv + h * TC_D;
      R.k[it] = __ldg(reinterpret_cast<const uint4*>(k16 + off) + c8);
      R.v[it] = __ldg(reinterpret_cast<const uint4*>(v16 + off) + c8);
    } else if (has_bias && j == a.S) {
      R.k[it] = pack8_bf16(a.bias_k + h * TC_D + c8 * 8);
      R.v[it] = pack8_bf16(a.bias_v + h * TC_D + c8 * 8);
    }
  }
  R.mask = kv_mask(a, b, j0, L);
}
__device__ __forceinline__ void store_kv(const KvRegs<false>& R, __nv_bfloat16 (*sK)[TC_LD],
                                         __nv_bfloat16 (*sV)[TC_LD], float* sMask) {
  store_tile(R.k, sK);
  store_tile(R.v, sV);
  if (threadIdx.x < 64) sMask[threadIdx.x] = R.mask;
}
__device__ __forceinline__ void store_kv(const KvRegs<true>& R, __nv_bfloat16 (*sK)[TC_LD],
                                         __nv_bfloat16 (*sV)[TC_LD], float* sMask) {
#pragma unroll
  for (int it = 0; it < TC_ST / 2; ++it) {
    const int i = threadIdx.x + it * TC_THREADS;
    const int r = i >> 3, c8 = i & 7;
    *reinterpret_cast<uint4*>(&sK[r][c8 * 8]) = R.k[it];
    *reinterpret_cast<uint4*>(&sV[r][c8 * 8]) = R.v[it];
  }
  if (threadIdx.x < 64) sMask[threadIdx.x] = R.mask;
}

// A fragments (16 rows x 64 k) of the warp's row block from a [64][TC_LD] tile.
__device__ __forceinline__ void load_a_frags(const __nv_bfloat16 (*tile)[TC_LD], int warp, int lane,
                                             uint32_t (&f)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(f[ks][0], f[ks][1], f[ks][2], f[ks][3],
                &tile[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
}
// C[16 x 64] += A(frags, 16 x 64) . Bt^T where Bt is a [64 n][64 k] row-major tile (B[k][n] = Bt[n][k]).
__device__ __forceinline__ void mma_a_bt(float (&c)[8][4], const uint32_t (&af)[4][4],
                                         const __nv_bfloat16 (*bt)[TC_LD], int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(b0, b1, b2, b3,
                  &bt[np * 16 + (lane & 7) + ((lane >> 4) << 3)][ks * 16 + ((lane >> 3) & 1) * 8]);
      mma_bf16_16816(c[2 * np], af[ks], b0, b1);
      mma_bf16_16816(c[2 * np + 1], af[ks], b2, b3);
    }
  }
}
// C[16 x 64] += P(frags, 16 x 64 k) . Bm where Bm is a [64 k][64 n] row-major tile.
__device__ __forceinline__ void mma_p_b(float (&c)[8][4], const uint32_t (&pf)[4][4],
                                        const __nv_bfloat16 (*bm)[TC_LD], int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(b0, b1, b2, b3,
                        &bm[kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]);
      mma_bf16_16816(c[2 * np], pf[kk], b0, b1);
      mma_bf16_16816(c[2 * np + 1], pf[kk], b2, b3);
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int n = 0; n < 8; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Key tiles that consist of trailing padding only (an article shorter than the batch maximum: keys
// [kv_len[b], S) are masked, their probabilities exactly 0) are skipped: the tile walk jumps from the
// last tile holding a valid key to the tile holding key S (the bias / zero rows).
struct TileWalk {
  int skip_from, skip_to;       // tile starts in [skip_from, skip_to) are all padding
  __device__ __forceinline__ TileWalk(const AttnArgs& a, int b) {
    const int Sv = a.kv_len != nullptr ? min(a.S, __ldg(a.kv_len + b)) : a.S;
    skip_from = max((Sv + TC_BN - 1) / TC_BN * TC_BN, TC_BN);      // tile 0 is always walked
    skip_to = a.S / TC_BN * TC_BN;
  }
  __device__ __forceinline__ int next(int j0) const {
    const int n = j0 + TC_BN;
    return (n >= skip_from && n < skip_to) ? skip_to : n;
  }
  __device__ __forceinline__ bool skipped(int j0) const { return j0 >= skip_from && j0 < skip_to; }
};

// ------------------------------------------------------------------------------------------ forward
template <bool KV16>
__global__ void __launch_bounds__(128, TT_ATTN_MINB)
attn_fwd_tc_kernel(const __grid_constant__ AttnArgsN args) {
  pdl_prologue();
  AttnArgs a = args.a[blockIdx.z];
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ __align__(16) __nv_bfloat16 sQ[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sK[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sV[64][TC_LD];
  __shared__ float sMask[64];
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.y * TC_BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  KvRegs<KV16> nxt;
  {
    float4 qreg[TC_ST];
    load_tile(a.q, a.ldq, a.B, b, h, q0, a.T, qreg);
    load_kv(a, b, h, 0, L, nxt);              // first K/V tile in flight together with Q
    store_tile(qreg, sQ);
  }
  store_kv(nxt, sK, sV, sMask);
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(sQ, warp, lane, qf);
  float o[8][4];
  zero_acc(o);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int t0 = q0 + warp * 16 + g, t1 = t0 + 8;
  const unsigned long long base0 = (static_cast<unsigned long long>(bh) * a.T + t0) * L;
  const unsigned long long base1 = (static_cast<unsigned long long>(bh) * a.T + t1) * L;
  const TileWalk walk(a, b);
  for (int j0 = 0; j0 < L; j0 = walk.next(j0)) {
    if (j0 > 0) {
      __syncthreads();                        // everyone is done with the previous tile
      store_kv(nxt, sK, sV, sMask);
      __syncthreads();
    }
    if (walk.next(j0) < L) load_kv(a, b, h, walk.next(j0), L, nxt);   // prefetch under the MMAs below
    float s[8][4];
    zero_acc(s);
    mma_a_bt(s, qf, sK, lane);
    float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float k0 = sMask[n * 8 + 2 * tg], k1 = sMask[n * 8 + 2 * tg + 1];
      s[n][0] += k0; s[n][1] += k1; s[n][2] += k0; s[n][3] += k1;
      tm0 = fmaxf(tm0, fmaxf(s[n][0], s[n][1]));
      tm1 = fmaxf(tm1, fmaxf(s[n][2], s[n][3]));
    }
    tm0 = quad_max(tm0);
    tm1 = quad_max(tm1);
    const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
    const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mn0);
    const float c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mn1);
    const float b0 = (mn0 == -INFINITY) ? 0.f : mn0, b1 = (mn1 == -INFINITY) ? 0.f : mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float p0 = __expf(s[n][0] - b0), p1 = __expf(s[n][1] - b0);
      float p2 = __expf(s[n][2] - b1), p3 = __expf(s[n][3] - b1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      if (a.p_drop > 0.f) {
        const int j = j0 + n * 8 + 2 * tg;
        p0 *= dropout_scale(a.seed, base0 + j, a.p_drop, inv_keep);
        p1 *= dropout_scale(a.seed, base0 + j + 1, a.p_drop, inv_keep);
        p2 *= dropout_scale(a.seed, base1 + j, a.p_drop, inv_keep);
        p3 *= dropout_scale(a.seed, base1 + j + 1, a.p_drop, inv_keep);
      }
      pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l0 = l0 * c0 + quad_sum(rs0);
    l1 = l1 * c1 + quad_sum(rs1);
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1;
    }
    mma_p_b(o, pf, sV, lane);
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = h * TC_D + n * 8 + 2 * tg;
    if (t0 < a.T) {
      *reinterpret_cast<float2*>(a.out + (static_cast<long long>(t0) * a.B + b) * a.ldo + col) =
          make_float2(o[n][0] * i0, o[n][1] * i0);
      if (a.out16)
        *reinterpret_cast<uint32_t*>(a.out16 + (static_cast<long long>(t0) * a.B + b) * a.ldo16 + col) =
            pack_bf16(o[n][0] * i0, o[n][1] * i0);
    }
    if (t1 < a.T) {
      *reinterpret_cast<float2*>(a.out + (static_cast<long long>(t1) * a.B + b) * a.ldo + col) =
          make_float2(o[n][2] * i1, o[n][3] * i1);
      if (a.out16)
        *reinterpret_cast<uint32_t*>(a.out16 + (static_cast<long long>(t1) * a.B + b) * a.ldo16 + col) =
            pack_bf16(o[n][2] * i1, o[n][3] * i1);
    }
  }
  if (tg == 0 && a.lse) {
    if (t0 < a.T) a.lse[static_cast<long long>(bh) * a.T + t0] = m0 + logf(l0);
    if (t1 < a.T) a.lse[static_cast<long long>(bh) * a.T + t1] = m1 + logf(l1);
  }
}

// D[t] = sum_c dO[t,c] * O[t,c] and lse[t] for the 64 queries q0.. into shared memory.
// Thread pair (2r, 2r+1) owns row r, 32 columns each: 16 independent float4 loads per thread, one
// round of memory latency (the previous warp-per-row loop paid 16 dependent rounds per warp).
template <bool WRITE = false>
__device__ __forceinline__ void stage_row_stats(const AttnArgs& a, int b, int h, int bh, int q0,
                                                float* sD, float* sLse) {
  const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int t = q0 + r;
  float d = 0.f;
  if (t < a.T) {
    const long long off = (static_cast<long long>(t) * a.B + b) * a.ldo + h * TC_D + half * 32;
    const float4* pd = reinterpret_cast<const float4*>(a.dout + off);
    const float4* po = reinterpret_cast<const float4*>(a.out + off);
    float4 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = __ldg(pd + i); y[i] = __ldg(po + i); }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      d += (x[i].x * y[i].x + x[i].y * y[i].y) + (x[i].z * y[i].z + x[i].w * y[i].w);
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  if (half == 0) {
    sD[r] = d;
    sLse[r] = t < a.T ? a.lse[static_cast<long long>(bh) * a.T + t] : INFINITY;  // p := 0 beyond T
    if (WRITE && a.dsum != nullptr && t < a.T) a.dsum[static_cast<long long>(bh) * a.T + t] = d;
  }
}
// The same two row vectors when the dQ kernel (launched just before) has left D in a.dsum.
__device__ __forceinline__ void stage_row_stats_cached(const AttnArgs& a, int bh, int q0, float* sD, float* sLse) {
  if (threadIdx.x < 64) {
    const int t = q0 + threadIdx.x;
    const bool ok = t < a.T;
    sD[threadIdx.x] = ok ? __ldg(a.dsum + static_cast<long long>(bh) * a.T + t) : 0.f;
    sLse[threadIdx.x] = ok ? a.lse[static_cast<long long>(bh) * a.T + t] : INFINITY;
  }
}

// ------------------------------------------------------------------------------------------ dQ
template <bool KV16>
__global__ void __launch_bounds__(128, TT_ATTN_MINB)
attn_bwd_dq_tc_kernel(const __grid_constant__ AttnArgsN args) {
  pdl_prologue();
  AttnArgs a = args.a[blockIdx.z];
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ __align__(16) __nv_bfloat16 sQ[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sdO[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sK[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sV[64][TC_LD];
  __shared__ float sMask[64], sD[64], sLse[64];
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.y * TC_BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  KvRegs<KV16> nxt;
  {
    float4 qreg[TC_ST], doreg[TC_ST];
    load_tile(a.q, a.ldq, a.B, b, h, q0, a.T, qreg);
    load_tile(a.dout, a.ldo, a.B, b, h, q0, a.T, doreg);
    load_kv(a, b, h, 0, L, nxt);              // first K/V tile in flight together with Q, dO
    store_tile(qreg, sQ);
    store_tile(doreg, sdO);
  }
  stage_row_stats<true>(a, b, h, bh, q0, sD, sLse);
  store_kv(nxt, sK, sV, sMask);
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(sQ, warp, lane, qf);
  load_a_frags(sdO, warp, lane, dof);
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const float lse0 = sLse[r0], lse1 = sLse[r1], D0 = sD[r0], D1 = sD[r1];
  const unsigned long long base0 = (static_cast<unsigned long long>(bh) * a.T + q0 + r0) * L;
  const unsigned long long base1 = (static_cast<unsigned long long>(bh) * a.T + q0 + r1) * L;
  float dq[8][4];
  zero_acc(dq);
  const TileWalk walk(a, b);
  for (int j0 = 0; j0 < L; j0 = walk.next(j0)) {
    if (j0 > 0) {
      __syncthreads();
      store_kv(nxt, sK, sV, sMask);
      __syncthreads();
    }
    if (walk.next(j0) < L) load_kv(a, b, h, walk.next(j0), L, nxt);   // prefetch under the MMAs below
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_bt(s, qf, sK, lane);
    mma_a_bt(dp, dof, sV, lane);
    uint32_t dsf[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int jj = n * 8 + 2 * tg;
      const float k0 = sMask[jj], k1 = sMask[jj + 1];
      float sc0 = 1.f, sc1 = 1.f, sc2 = 1.f, sc3 = 1.f;
      if (a.p_drop > 0.f) {
        sc0 = dropout_scale(a.seed, base0 + j0 + jj, a.p_drop, inv_keep);
        sc1 = dropout_scale(a.seed, base0 + j0 + jj + 1, a.p_drop, inv_keep);
        sc2 = dropout_scale(a.seed, base1 + j0 + jj, a.p_drop, inv_keep);
        sc3 = dropout_scale(a.seed, base1 + j0 + jj + 1, a.p_drop, inv_keep);
      }
      const float p0 = __expf(s[n][0] + k0 - lse0), p1 = __expf(s[n][1] + k1 - lse0);
      const float p2 = __expf(s[n][2] + k0 - lse1), p3 = __expf(s[n][3] + k1 - lse1);
      dsf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0 * (dp[n][0] * sc0 - D0), p1 * (dp[n][1] * sc1 - D0));
      dsf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2 * (dp[n][2] * sc2 - D1), p3 * (dp[n][3] * sc3 - D1));
    }
    mma_p_b(dq, dsf, sK, lane);
  }
  const int t0 = q0 + r0, t1 = q0 + r1;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = h * TC_D + n * 8 + 2 * tg;
    if (t0 < a.T) {
      *reinterpret_cast<float2*>(a.dq + (static_cast<long long>(t0) * a.B + b) * a.ldq + col) =
          make_float2(dq[n][0], dq[n][1]);
      if (a.dq16)
        *reinterpret_cast<uint32_t*>(a.dq16 + (static_cast<long long>(t0) * a.B + b) * a.ldq16 + col) =
            pack_bf16(dq[n][0], dq[n][1]);
    }
    if (t1 < a.T) {
      *reinterpret_cast<float2*>(a.dq + (static_cast<long long>(t1) * a.B + b) * a.ldq + col) =
          make_float2(dq[n][2], dq[n][3]);
      if (a.dq16)
        *reinterpret_cast<uint32_t*>(a.dq16 + (static_cast<long long>(t1) * a.B + b) * a.ldq16 + col) =
            pack_bf16(dq[n][2], dq[n][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------ dK, dV
// CTA = (b, h, a.dkv_tpc consecutive key tiles).  The query side of the problem (Q, dO, lse, D) is the
// same for every key tile of a (b, h): it is staged once per CTA and, when the whole query range fits
// one tile (T <= 64: every training shape), reused for all a.dkv_tpc key tiles -- the query-side loads
// were 4x the key-side bytes of a one-tile CTA.  D = rowsum(dO * O) comes from the dQ kernel (a.dsum).
// Key tiles per CTA: enough CTAs for two rounds of the resident slots (2 CTAs x 148 SMs), at most 9.
static int g_dkv_tpc = 0;       // tt_attn_set_dkv_tiles_per_cta: 0 = automatic
static inline int dkv_tiles_per_cta(int L, int BH) {
  if (g_dkv_tpc > 0) return g_dkv_tpc;
  const long long tiles = static_cast<long long>(ceil_div(L, TC_BN)) * BH;
  long long t = (tiles + 295) / 592;
  return static_cast<int>(t < 1 ? 1 : (t > 9 ? 9 : t));
}
template <bool KV16>
__global__ void __launch_bounds__(128, TT_ATTN_MINB)
attn_bwd_dkv_tc_kernel(const __grid_constant__ AttnArgsN args) {
  pdl_prologue();
  AttnArgs a = args.a[blockIdx.z];
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ __align__(16) __nv_bfloat16 sQ[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sdO[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sK[64][TC_LD];
  __shared__ __align__(16) __nv_bfloat16 sV[64][TC_LD];
  __shared__ float sMask[64], sD[64], sLse[64];
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const bool has_bias = a.bias_k != nullptr;
  const int L = a.S + (has_bias ? 1 : 0) + (a.zero_row ? 1 : 0);
  const int jbeg = blockIdx.y * (a.dkv_tpc * TC_BN);
  if (jbeg >= L) return;                                                // grid.y covers the longest context
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const TileWalk walk(a, b);
  bool q_staged = false;        // sQ / sdO / sD / sLse hold query tile 0 (valid across key tiles when T <= 64)
  for (int tile = 0; tile < a.dkv_tpc; ++tile) {
    const int j0 = jbeg + tile * TC_BN;
    if (j0 >= L) break;                                                 // CTA-uniform
    if (walk.skipped(j0)) {
      // a tile of trailing padding: every probability is 0, so dK = dV = 0 (the rows are still written:
      // the projection's dW GEMM reads the whole slab)
      for (int i = threadIdx.x; i < 64 * 8; i += 128) {
        const int r = i >> 3, c8 = i & 7;
        const long long off = (static_cast<long long>(j0 + r) * a.B + b) * a.ldkv + h * TC_D + c8 * 8;
        if (KV16) {
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dk) + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dv) + off) = make_uint4(0, 0, 0, 0);
        } else {
          *reinterpret_cast<float4*>(a.dk + off) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(a.dk + off + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(a.dv + off) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(a.dv + off + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      continue;
    }
    if (tile > 0) __syncthreads();          // the previous tile's readers of sK / sV / sQ / sdO are done
    {
      KvRegs<KV16> kv;
      load_kv(a, b, h, j0, L, kv);                      // all loads of the round in flight
      if (!q_staged) {
        float4 qreg[TC_ST], doreg[TC_ST];
        load_tile(a.q, a.ldq, a.B, b, h, 0, a.T, qreg);
        load_tile(a.dout, a.ldo, a.B, b, h, 0, a.T, doreg);
        store_tile(qreg, sQ);
        store_tile(doreg, sdO);
      }
      store_kv(kv, sK, sV, sMask);
    }
    if (!q_staged) {
      if (a.dsum != nullptr) stage_row_stats_cached(a, bh, 0, sD, sLse);
      else stage_row_stats(a, b, h, bh, 0, sD, sLse);
    }
    q_staged = a.T <= TC_BM;                // with more query tiles the loop below restages them
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(sK, warp, lane, kf);
    load_a_frags(sV, warp, lane, vf);
    const int r0 = warp * 16 + g, r1 = r0 + 8;       // key rows of this thread
    const float km0 = sMask[r0], km1 = sMask[r1];
    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);
    for (int q0 = 0; q0 < a.T; q0 += TC_BM) {
      if (q0 > 0) {
        __syncthreads();
        stage_tile(a.q, a.ldq, a.B, b, h, q0, a.T, sQ);
        stage_tile(a.dout, a.ldo, a.B, b, h, q0, a.T, sdO);
        if (a.dsum != nullptr) stage_row_stats_cached(a, bh, q0, sD, sLse);
        else stage_row_stats(a, b, h, bh, q0, sD, sLse);
        __syncthreads();
      }
      float st[8][4], dpt[8][4];      // S^T and dP^T: rows = keys, cols = queries
      zero_acc(st);
      zero_acc(dpt);
      mma_a_bt(st, kf, sQ, lane);
      mma_a_bt(dpt, vf, sdO, lane);
      uint32_t pf[4][4], dsf[4][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int tt = n * 8 + 2 * tg;             // query column (and tt+1)
        const float lse_a = sLse[tt], lse_b = sLse[tt + 1], Da = sD[tt], Db = sD[tt + 1];
        float sc0 = 1.f, sc1 = 1.f, sc2 = 1.f, sc3 = 1.f;
        if (a.p_drop > 0.f) {
          const unsigned long long ba = (static_cast<unsigned long long>(bh) * a.T + q0 + tt) * L;
          const unsigned long long bb = ba + L;
          sc0 = dropout_scale(a.seed, ba + j0 + r0, a.p_drop, inv_keep);
          sc1 = dropout_scale(a.seed, bb + j0 + r0, a.p_drop, inv_keep);
          sc2 = dropout_scale(a.seed, ba + j0 + r1, a.p_drop, inv_keep);
          sc3 = dropout_scale(a.seed, bb + j0 + r1, a.p_drop, inv_keep);
        }
        const float p0 = __expf(st[n][0] + km0 - lse_a), p1 = __expf(st[n][1] + km0 - lse_b);
        const float p2 = __expf(st[n][2] + km1 - lse_a), p3 = __expf(st[n][3] + km1 - lse_b);
        pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0 * sc0, p1 * sc1);
        pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2 * sc2, p3 * sc3);
        dsf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0 * (dpt[n][0] * sc0 - Da), p1 * (dpt[n][1] * sc1 - Db));
        dsf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2 * (dpt[n][2] * sc2 - Da), p3 * (dpt[n][3] * sc3 - Db));
      }
      mma_p_b(dv, pf, sdO, lane);
      mma_p_b(dk, dsf, sQ, lane);
    }
    const int ja = j0 + r0, jb = j0 + r1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int c = n * 8 + 2 * tg;
      if (ja < a.S) {
        const long long off = (static_cast<long long>(ja) * a.B + b) * a.ldkv + h * TC_D + c;
        if (KV16) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.dk) + off) = pack_bf16(dk[n][0], dk[n][1]);
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.dv) + off) = pack_bf16(dv[n][0], dv[n][1]);
        } else {
          *reinterpret_cast<float2*>(a.dk + off) = make_float2(dk[n][0], dk[n][1]);
          *reinterpret_cast<float2*>(a.dv + off) = make_float2(dv[n][0], dv[n][1]);
        }
      } else if (has_bias && ja == a.S) {
        if (a.dbias_k) { atomicAdd(a.dbias_k + h * TC_D + c, dk[n][0]); atomicAdd(a.dbias_k + h * TC_D + c + 1, dk[n][1]); }
        if (a.dbias_v) { atomicAdd(a.dbias_v + h * TC_D + c, dv[n][0]); atomicAdd(a.dbias_v + h * TC_D + c + 1, dv[n][1]); }
      }
      if (jb < a.S) {
        const long long off = (static_cast<long long>(jb) * a.B + b) * a.ldkv + h * TC_D + c;
        if (KV16) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.dk) + off) = pack_bf16(dk[n][2], dk[n][3]);
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.dv) + off) = pack_bf16(dv[n][2], dv[n][3]);
        } else {
          *reinterpret_cast<float2*>(a.dk + off) = make_float2(dk[n][2], dk[n][3]);
          *reinterpret_cast<float2*>(a.dv + off) = make_float2(dv[n][2], dv[n][3]);
        }
      } else if (has_bias && jb == a.S) {
        if (a.dbias_k) { atomicAdd(a.dbias_k + h * TC_D + c, dk[n][2]); atomicAdd(a.dbias_k + h * TC_D + c + 1, dk[n][3]); }
        if (a.dbias_v) { atomicAdd(a.dbias_v + h * TC_D + c, dv[n][2]); atomicAdd(a.dbias_v + h * TC_D + c + 1, dv[n][3]); }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ decode (T = 1)
// One query row per (b, h): the 64-query tensor-core tile would spend 63/64 of its MMAs and
// exponentials on padding rows, so incremental decoding gets a bandwidth-shaped kernel instead.
// CTA = (b, group of HG heads), one WARP per head, everything per head is warp-local (no block
// barrier).  The projected keys|values of one (key, batch) row hold all heads side by side, so the
// HG warps of a CTA read HG*128 contiguous bytes per key: DRAM-friendly bursts instead of isolated
// 128-byte lines 16 KB apart.
//   phase 1: lane = key (32 keys per pass), 64-dim dot with q (shared memory broadcast)
//   phase 2: warp softmax over the L scores kept in shared memory
//   phase 3: lane = 2 of the 64 value dims, loop over keys (coalesced 128-byte row reads)
// K and V are read exactly once: 2*L*64*sizeof(kv) bytes per (b, h).
__device__ __forceinline__ float dot4(float acc, float x0, float x1, float x2, float x3, const float* q) {
  return fmaf(x3, q[3], fmaf(x2, q[2], fmaf(x1, q[1], fmaf(x0, q[0], acc))));
}
constexpr int DEC_HG = 8;      // heads (warps) per CTA
template <bool KV16>
__global__ void __launch_bounds__(DEC_HG * 32)
attn_decode_kernel(const __grid_constant__ AttnArgsN args) {
  pdl_prologue();
  const AttnArgs& a = args.a[blockIdx.z];
  const int Lp = a.Lp;
  extern __shared__ float dsm[];        // [DEC_HG][Lp] scores -> probabilities
  __shared__ float sq_all[DEC_HG][TC_D];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.y * DEC_HG + warp;
  if (h >= a.H) return;                 // warp-uniform; no block barrier below
  const bool has_bias = a.bias_k != nullptr;
  const int L = a.S + (has_bias ? 1 : 0) + (a.zero_row ? 1 : 0);
  // trailing padding of this sample's context (an article shorter than the batch maximum): those keys
  // are masked, their probability is exactly 0 -- their K and V rows are never read
  const int Sv = a.kv_len != nullptr ? min(a.S, __ldg(a.kv_len + b)) : a.S;
  float* sp = dsm + warp * Lp;
  float* sq = sq_all[warp];
  {
    const float2 q2 = *reinterpret_cast<const float2*>(a.q + static_cast<long long>(b) * a.ldq + h * TC_D + 2 * lane);
    sq[2 * lane] = q2.x;
    sq[2 * lane + 1] = q2.y;
  }
  __syncwarp();
  // ---- phase 1: scores
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    float sc;
    if (j < a.S && j >= Sv) {
      sc = -INFINITY;
    } else if (j < a.S) {
      const long long off = j * a.kv_j_stride + b * a.kv_b_stride + h * a.kv_h_stride;
      float acc = 0.f;
      if (KV16) {
        const uint4* kr = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.k) + off);
        uint4 u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __ldg(kr + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) {       // same association order as the fp32-storage branch
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u[i]);
          acc = dot4(acc, __low2float(h2[0]), __high2float(h2[0]), __low2float(h2[1]), __high2float(h2[1]),
                     sq + 8 * i);
          acc = dot4(acc, __low2float(h2[2]), __high2float(h2[2]), __low2float(h2[3]), __high2float(h2[3]),
                     sq + 8 * i + 4);
        }
      } else {
        const float4* kr = reinterpret_cast<const float4*>(a.k + off);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const float4 v = __ldg(kr + i);
          acc = dot4(acc, v.x, v.y, v.z, v.w, sq + 4 * i);
        }
      }
      sc = (a.mask && a.mask[static_cast<long long>(b) * a.S + j]) ? -INFINITY : acc;
    } else if (has_bias && j == a.S) {
      float acc = 0.f;
      for (int i = 0; i < TC_D; ++i) acc = fmaf(__ldg(a.bias_k + h * TC_D + i), sq[i], acc);
      sc = acc;
    } else {
      sc = 0.f;                       // add_zero_attn: an all-zero key
    }
    sp[j] = sc;
    m = fmaxf(m, sc);
  }
  m = warp_max(m);
  // ---- phase 2: softmax
  const float base = (m == -INFINITY) ? 0.f : m;
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = __expf(sp[j] - base);
    sp[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // ---- phase 3: out = sum_j p_j v_j.  8 lanes cover one 64-dim value row with 16-byte loads, the
  // warp takes 4 keys per pass (x4 unrolled: 16 rows in flight), the 4 key groups meet by shuffle.
  const int kq = lane >> 3, c8 = lane & 7;         // key slot within a pass, 8-dim chunk of the row
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  const long long vstep = a.kv_j_stride;
  const long long vbase = b * a.kv_b_stride + h * a.kv_h_stride + c8 * 8;
  if (KV16) {
    const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(a.v) + vbase;
#pragma unroll 4
    for (int j = kq; j < Sv; j += 4) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(vp + j * vstep));
      const float pj = sp[j];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[2 * e] = fmaf(pj, __low2float(h2[e]), o[2 * e]);
        o[2 * e + 1] = fmaf(pj, __high2float(h2[e]), o[2 * e + 1]);
      }
    }
  } else {
    const float* vp = a.v + vbase;
#pragma unroll 4
    for (int j = kq; j < Sv; j += 4) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(vp + j * vstep));
      const float4 y = __ldg(reinterpret_cast<const float4*>(vp + j * vstep) + 1);
      const float pj = sp[j];
      o[0] = fmaf(pj, x.x, o[0]); o[1] = fmaf(pj, x.y, o[1]); o[2] = fmaf(pj, x.z, o[2]); o[3] = fmaf(pj, x.w, o[3]);
      o[4] = fmaf(pj, y.x, o[4]); o[5] = fmaf(pj, y.y, o[5]); o[6] = fmaf(pj, y.z, o[6]); o[7] = fmaf(pj, y.w, o[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (kq == 0) {
    const float inv = 1.f / sum;
    float bvv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bvv[i] = has_bias ? sp[a.S] * __ldg(a.bias_v + h * TC_D + c8 * 8 + i) : 0.f;
    float4* op = reinterpret_cast<float4*>(a.out + static_cast<long long>(b) * a.ldo + h * TC_D + c8 * 8);
    op[0] = make_float4((o[0] + bvv[0]) * inv, (o[1] + bvv[1]) * inv, (o[2] + bvv[2]) * inv, (o[3] + bvv[3]) * inv);
    op[1] = make_float4((o[4] + bvv[4]) * inv, (o[5] + bvv[5]) * inv, (o[6] + bvv[6]) * inv, (o[7] + bvv[7]) * inv);
    if (a.out16) {
      uint4 u;
      u.x = pack_bf16((o[0] + bvv[0]) * inv, (o[1] + bvv[1]) * inv);
      u.y = pack_bf16((o[2] + bvv[2]) * inv, (o[3] + bvv[3]) * inv);
      u.z = pack_bf16((o[4] + bvv[4]) * inv, (o[5] + bvv[5]) * inv);
      u.w = pack_bf16((o[6] + bvv[6]) * inv, (o[7] + bvv[7]) * inv);
      *reinterpret_cast<uint4*>(a.out16 + static_cast<long long>(b) * a.ldo16 + h * TC_D + c8 * 8) = u;
    }
  }
  if (lane == 0 && a.lse) a.lse[b * a.H + h] = base + logf(sum);
}

// Token-major projected keys|values ([S*B, ld] rows, one 64-dim slice per head) -> head-major decode
// cache K, V [B, H, S, 64]: one (batch, head) slice becomes S*128 contiguous bytes, so the decode
// kernel streams it instead of touching isolated 128-byte lines 16 KB apart.  One 16-byte chunk per
// thread, done once per generate() call.
__global__ void kv_repack_heads_kernel(const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                                       long long ld, __nv_bfloat16* __restrict__ ko,
                                       __nv_bfloat16* __restrict__ vo, int S, int B, int H) {
  pdl_prologue();
  const long long total = static_cast<long long>(S) * B * H * 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i & 7);
    long long t = i >> 3;
    const int h = static_cast<int>(t % H); t /= H;
    const int b = static_cast<int>(t % B);
    const int j = static_cast<int>(t / B);
    const long long src = (static_cast<long long>(j) * B + b) * ld + h * TC_D + c8 * 8;
    const long long dst = ((static_cast<long long>(b) * H + h) * S + j) * TC_D + c8 * 8;
    *reinterpret_cast<uint4*>(ko + dst) = __ldg(reinterpret_cast<const uint4*>(k + src));
    *reinterpret_cast<uint4*>(vo + dst) = __ldg(reinterpret_cast<const uint4*>(v + src));
  }
}

static int tc_check(const AttnArgs& a, int D, bool kv16 = false) {
  TT_REQUIRE(D == TC_D, "attention (tensor-core path): head_dim must be %d (got %d)", TC_D, D);
  TT_REQUIRE(a.ldq % 4 == 0 && a.ldkv % (kv16 ? 8 : 4) == 0 && a.ldo % 4 == 0,
             "attention (tensor-core path): row strides must be multiples of 4 (8 for bf16 k/v)");
  TT_REQUIRE(!kv16 || ((reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0,
             "attention (tensor-core path): bf16 k/v must be 16-byte aligned");
  return TT_OK;
}

static int attn_fwd_tc_impl(const float* q, const void* k, const void* v, const float* bias_k,
                            const float* bias_v, const uint8_t* key_padding_mask, float* out,
                            float* lse, int T, int B, int S, int H, int D, long long ldq,
                            long long ldkv, long long ldo, int zero_row, float p_drop,
                            unsigned long long seed, void* stream, bool kv16) {
  TT_REQUIRE(q && out && lse, "tt_attn_fwd_tc: null pointer");
  TT_REQUIRE(S == 0 || (k && v), "tt_attn_fwd_tc: null k/v with S > 0");
  TT_REQUIRE((bias_k == nullptr) == (bias_v == nullptr), "tt_attn_fwd_tc: bias_k/bias_v mismatch");
  TT_REQUIRE(S + (bias_k ? 1 : 0) + (zero_row ? 1 : 0) > 0, "tt_attn_fwd_tc: empty key set");
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = reinterpret_cast<const float*>(k); a.v = reinterpret_cast<const float*>(v);
  a.bias_k = bias_k; a.bias_v = bias_v; a.mask = key_padding_mask;
  a.out = out; a.lse = lse; a.T = T; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row;
  a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldo;
  a.p_drop = p_drop; a.seed = seed; a.step_ptr = rng_step_ptr();
  int rc = tc_check(a, D, kv16 && S > 0);
  if (rc != TT_OK) return rc;
  a.kv_j_stride = static_cast<long long>(B) * ldkv; a.kv_b_stride = ldkv; a.kv_h_stride = TC_D;
  if (T == 1 && p_drop == 0.f) {      // incremental decoding: one query row per (b, h)
    const int L = S + (bias_k ? 1 : 0) + (zero_row ? 1 : 0);
    const int Lp = (L + 3) & ~3;
    const size_t smem = static_cast<size_t>(DEC_HG) * Lp * sizeof(float);
    if (smem <= 40 * 1024) {
      const dim3 grid(B, ceil_div(H, DEC_HG));
      a.Lp = Lp;
      AttnArgsN an{};
      an.a[0] = a;
      if (kv16) launch_k(attn_decode_kernel<true>, grid, dim3(DEC_HG * 32), smem, (cudaStream_t)stream, an);
      else launch_k(attn_decode_kernel<false>, grid, dim3(DEC_HG * 32), smem, (cudaStream_t)stream, an);
      return check_launch("attn_decode_kernel");
    }
  }
  dim3 grid(B * H, ceil_div(T, TC_BM));
  AttnArgsN an{};
  an.a[0] = a;
  if (kv16) launch_k(attn_fwd_tc_kernel<true>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_fwd_tc_kernel<false>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, an);
  return check_launch("attn_fwd_tc_kernel");
}

static int attn_bwd_tc_impl(const float* dout, const float* q, const void* k, const void* v,
                            const float* bias_k, const float* bias_v,
                            const uint8_t* key_padding_mask, const float* out, const float* lse,
                            float* dq, void* dk, void* dv, float* dbias_k, float* dbias_v,
                            int T, int B, int S, int H, int D, long long ldq, long long ldkv,
                            long long ldo, int zero_row, float p_drop, unsigned long long seed,
                            void* stream, bool kv16) {
  TT_REQUIRE(dout && q && out && lse && dq, "tt_attn_bwd_tc: null pointer");
  TT_REQUIRE(S == 0 || (k && v && dk && dv), "tt_attn_bwd_tc: null k/v/dk/dv with S > 0");
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = reinterpret_cast<const float*>(k); a.v = reinterpret_cast<const float*>(v);
  a.bias_k = bias_k; a.bias_v = bias_v; a.mask = key_padding_mask;
  a.out = const_cast<float*>(out); a.lse = const_cast<float*>(lse);
  a.T = T; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row; a.p_drop = p_drop; a.seed = seed;
  a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldo; a.step_ptr = rng_step_ptr();
  a.dout = dout; a.dq = dq; a.dk = reinterpret_cast<float*>(dk); a.dv = reinterpret_cast<float*>(dv);
  a.dbias_k = dbias_k; a.dbias_v = dbias_v;
  int rc = tc_check(a, D, kv16 && S > 0);
  if (rc != TT_OK) return rc;
  const int L = S + (bias_k ? 1 : 0) + (zero_row ? 1 : 0);
  dim3 grid(B * H, ceil_div(T, TC_BM));
  AttnArgsN an{};
  an.a[0] = a;
  if (kv16) launch_k(attn_bwd_dq_tc_kernel<true>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_bwd_dq_tc_kernel<false>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, an);
  rc = check_launch("attn_bwd_dq_tc_kernel");
  if (rc != TT_OK) return rc;
  an.a[0].dkv_tpc = dkv_tiles_per_cta(L, B * H);
  dim3 grid2(B * H, ceil_div(L, TC_BN * an.a[0].dkv_tpc));
  if (kv16) launch_k(attn_bwd_dkv_tc_kernel<true>, dim3(grid2), dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_bwd_dkv_tc_kernel<false>, dim3(grid2), dim3(128), 0, (cudaStream_t)stream, an);
  return check_launch("attn_bwd_dkv_tc_kernel");
}

}  // namespace tt

using namespace tt;

extern "C" int tt_attn_fwd_tc(const float* q, const float* k, const float* v, const float* bias_k,
                              const float* bias_v, const uint8_t* key_padding_mask, float* out,
                              float* lse, int T, int B, int S, int H, int D, long long ldq,
                              long long ldkv, long long ldo, int zero_row, float p_drop,
                              unsigned long long seed, void* stream) {
  return attn_fwd_tc_impl(q, k, v, bias_k, bias_v, key_padding_mask, out, lse, T, B, S, H, D, ldq, ldkv,
                          ldo, zero_row, p_drop, seed, stream, false);
}

extern "C" int tt_attn_fwd_tc_kv16(const float* q, const void* k16, const void* v16, const float* bias_k,
                                   const float* bias_v, const uint8_t* key_padding_mask, float* out,
                                   float* lse, int T, int B, int S, int H, int D, long long ldq,
                                   long long ldkv, long long ldo, int zero_row, float p_drop,
                                   unsigned long long seed, void* stream) {
  return attn_fwd_tc_impl(q, k16, v16, bias_k, bias_v, key_padding_mask, out, lse, T, B, S, H, D, ldq,
                          ldkv, ldo, zero_row, p_drop, seed, stream, true);
}

extern "C" int tt_attn_bwd_tc(const float* dout, const float* q, const float* k, const float* v,
                              const float* bias_k, const float* bias_v,
                              const uint8_t* key_padding_mask, const float* out, const float* lse,
                              float* dq, float* dk, float* dv, float* dbias_k, float* dbias_v,
                              int T, int B, int S, int H, int D, long long ldq, long long ldkv,
                              long long ldo, int zero_row, float p_drop, unsigned long long seed,
                              void* stream) {
  return attn_bwd_tc_impl(dout, q, k, v, bias_k, bias_v, key_padding_mask, out, lse, dq, dk, dv, dbias_k,
                          dbias_v, T, B, S, H, D, ldq, ldkv, ldo, zero_row, p_drop, seed, stream, false);
}

extern "C" int tt_attn_bwd_tc_kv16(const float* dout, const float* q, const void* k16, const void* v16,
                                   const float* bias_k, const float* bias_v,
                                   const uint8_t* key_padding_mask, const float* out, const float* lse,
                                   float* dq, void* dk16, void* dv16, float* dbias_k, float* dbias_v,
                                   int T, int B, int S, int H, int D, long long ldq, long long ldkv,
                                   long long ldo, int zero_row, float p_drop, unsigned long long seed,
                                   void* stream) {
  return attn_bwd_tc_impl(dout, q, k16, v16, bias_k, bias_v, key_padding_mask, out, lse, dq, dk16, dv16,
                          dbias_k, dbias_v, T, B, S, H, D, ldq, ldkv, ldo, zero_row, p_drop, seed, stream,
                          true);
}

extern "C" int tt_kv_repack_heads(const void* k16, const void* v16, long long ldkv, void* k_out,
                                  void* v_out, int S, int B, int H, int D, void* stream) {
  TT_REQUIRE(k16 && v16 && k_out && v_out, "tt_kv_repack_heads: null pointer");
  TT_REQUIRE(D == TC_D && ldkv % 8 == 0, "tt_kv_repack_heads: head_dim must be %d, ldkv a multiple of 8", TC_D);
  const long long total = static_cast<long long>(S) * B * H * 8;
  if (total <= 0) return TT_OK;
  long long g = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  launch_k(kv_repack_heads_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, (cudaStream_t)stream,
           reinterpret_cast<const __nv_bfloat16*>(k16), reinterpret_cast<const __nv_bfloat16*>(v16), ldkv,
           reinterpret_cast<__nv_bfloat16*>(k_out), reinterpret_cast<__nv_bfloat16*>(v_out), S, B, H);
  return check_launch("kv_repack_heads_kernel");
}

extern "C" int tt_attn_decode_hm(const float* q, const void* k_hm, const void* v_hm, const float* bias_k,
                                 const float* bias_v, const uint8_t* key_padding_mask, float* out,
                                 float* lse, int B, int S, int H, int D, long long ldq, long long ldo,
                                 int zero_row, void* stream) {
  TT_REQUIRE(q && out && k_hm && v_hm && S > 0, "tt_attn_decode_hm: null pointer / empty cache");
  TT_REQUIRE(D == TC_D, "tt_attn_decode_hm: head_dim must be %d (got %d)", TC_D, D);
  TT_REQUIRE((bias_k == nullptr) == (bias_v == nullptr), "tt_attn_decode_hm: bias_k/bias_v mismatch");
  TT_REQUIRE(ldq % 2 == 0 && ldo % 4 == 0, "tt_attn_decode_hm: ldq must be even, ldo a multiple of 4");
  if (B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = reinterpret_cast<const float*>(k_hm); a.v = reinterpret_cast<const float*>(v_hm);
  a.bias_k = bias_k; a.bias_v = bias_v; a.mask = key_padding_mask;
  a.out = out; a.lse = lse; a.T = 1; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row;
  a.ldq = ldq; a.ldo = ldo; a.ldkv = 0;
  a.kv_j_stride = TC_D; a.kv_b_stride = static_cast<long long>(H) * S * TC_D;
  a.kv_h_stride = static_cast<long long>(S) * TC_D;
  const int L = S + (bias_k ? 1 : 0) + (zero_row ? 1 : 0);
  const int Lp = (L + 3) & ~3;
  const size_t smem = static_cast<size_t>(DEC_HG) * Lp * sizeof(float);
  TT_REQUIRE(smem <= 40 * 1024, "tt_attn_decode_hm: key set too long for the decode kernel (L=%d)", L);
  a.Lp = Lp;
  AttnArgsN an{};
  an.a[0] = a;
  launch_k(attn_decode_kernel<true>, dim3(B, ceil_div(H, DEC_HG)), dim3(DEC_HG * 32), smem,
           (cudaStream_t)stream, an);
  return check_launch("attn_decode_kernel");
}


// ---------------------------------------------------------------------------- multi-context launches
namespace tt {
static int fill_ctx(AttnArgs& a, const TtAttnCtx& c, int T, int B, int H, int D, int zero_row, float p_drop,
                    bool kv16, bool backward, const char* who) {
  TT_REQUIRE(c.q && c.out && c.lse, "%s: null q/out/lse", who);
  TT_REQUIRE(c.S >= 0 && (c.S == 0 || (c.k && c.v)), "%s: null k/v with S > 0", who);
  TT_REQUIRE((c.bias_k == nullptr) == (c.bias_v == nullptr), "%s: bias_k/bias_v mismatch", who);
  TT_REQUIRE(c.S + (c.bias_k ? 1 : 0) + (zero_row ? 1 : 0) > 0, "%s: empty key set", who);
  a = AttnArgs{};
  a.q = c.q; a.k = reinterpret_cast<const float*>(c.k); a.v = reinterpret_cast<const float*>(c.v);
  a.bias_k = c.bias_k; a.bias_v = c.bias_v; a.mask = c.mask;
  a.out = c.out; a.lse = c.lse; a.T = T; a.B = B; a.S = c.S; a.H = H; a.zero_row = zero_row;
  a.ldq = c.ldq; a.ldkv = c.ldkv; a.ldo = c.ldo;
  a.p_drop = p_drop; a.seed = c.seed; a.step_ptr = rng_step_ptr();
  a.kv_j_stride = static_cast<long long>(B) * c.ldkv; a.kv_b_stride = c.ldkv; a.kv_h_stride = TC_D;
  a.kv_len = c.kv_len;
  TT_REQUIRE(c.out16 == nullptr || (c.ldo16 % 8 == 0 && (reinterpret_cast<uintptr_t>(c.out16) & 15) == 0),
             "%s: out16 must be 16-byte aligned with a row pitch that is a multiple of 8", who);
  a.out16 = reinterpret_cast<__nv_bfloat16*>(c.out16); a.ldo16 = c.ldo16;
  if (backward) {
    TT_REQUIRE(c.dq16 == nullptr || (c.ldq16 % 2 == 0 && (reinterpret_cast<uintptr_t>(c.dq16) & 3) == 0),
               "%s: dq16 must be 4-byte aligned with an even row pitch", who);
    a.dq16 = reinterpret_cast<__nv_bfloat16*>(c.dq16); a.ldq16 = c.ldq16;
    a.dsum = c.dsum;
    TT_REQUIRE(c.dout && c.dq && (c.S == 0 || (c.dk && c.dv)), "%s: null dout/dq/dk/dv", who);
    a.dout = c.dout; a.dq = c.dq; a.dk = reinterpret_cast<float*>(c.dk); a.dv = reinterpret_cast<float*>(c.dv);
    a.dbias_k = c.dbias_k; a.dbias_v = c.dbias_v;
  }
  return tc_check(a, D, kv16 && c.S > 0);
}
}  // namespace tt

extern "C" void tt_attn_set_dkv_tiles_per_cta(int n) { tt::g_dkv_tpc = n > 0 ? (n > 64 ? 64 : n) : 0; }

extern "C" int tt_attn_fwd_tc_multi(const TtAttnCtx* ctx, int n, int T, int B, int H, int D, int zero_row,
                                    float p_drop, int kv16, void* stream) {
  TT_REQUIRE(ctx && n >= 1 && n <= ATTN_MAX_CTX, "tt_attn_fwd_tc_multi: 1..%d contexts", ATTN_MAX_CTX);
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgsN an{};
  for (int i = 0; i < n; ++i) {
    const int rc = fill_ctx(an.a[i], ctx[i], T, B, H, D, zero_row, p_drop, kv16 != 0, false, "tt_attn_fwd_tc_multi");
    if (rc != TT_OK) return rc;
  }
  if (T == 1 && p_drop == 0.f) {      // incremental decoding: one query row per (b, h), as tt_attn_fwd_tc
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
      const int L = ctx[i].S + (ctx[i].bias_k ? 1 : 0) + (zero_row ? 1 : 0);
      an.a[i].Lp = (L + 3) & ~3;
      const size_t need = static_cast<size_t>(DEC_HG) * an.a[i].Lp * sizeof(float);
      smem = need > smem ? need : smem;
    }
    if (smem <= 40 * 1024) {
      const dim3 grid(B, ceil_div(H, DEC_HG), n);
      if (kv16) launch_k(attn_decode_kernel<true>, grid, dim3(DEC_HG * 32), smem, (cudaStream_t)stream, an);
      else launch_k(attn_decode_kernel<false>, grid, dim3(DEC_HG * 32), smem, (cudaStream_t)stream, an);
      return check_launch("attn_decode_kernel");
    }
  }
  const dim3 grid(B * H, ceil_div(T, TC_BM), n);
  if (kv16) launch_k(attn_fwd_tc_kernel<true>, grid, dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_fwd_tc_kernel<false>, grid, dim3(128), 0, (cudaStream_t)stream, an);
  return check_launch("attn_fwd_tc_kernel");
}

extern "C" int tt_attn_bwd_tc_multi(const TtAttnCtx* ctx, int n, int T, int B, int H, int D, int zero_row,
                                    float p_drop, int kv16, void* stream) {
  TT_REQUIRE(ctx && n >= 1 && n <= ATTN_MAX_CTX, "tt_attn_bwd_tc_multi: 1..%d contexts", ATTN_MAX_CTX);
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgsN an{};
  int Lmax = 0;
  for (int i = 0; i < n; ++i) {
    const int rc = fill_ctx(an.a[i], ctx[i], T, B, H, D, zero_row, p_drop, kv16 != 0, true, "tt_attn_bwd_tc_multi");
    if (rc != TT_OK) return rc;
    const int L = ctx[i].S + (ctx[i].bias_k ? 1 : 0) + (zero_row ? 1 : 0);
    Lmax = L > Lmax ? L : Lmax;
  }
  const dim3 grid(B * H, ceil_div(T, TC_BM), n);
  if (kv16) launch_k(attn_bwd_dq_tc_kernel<true>, grid, dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_bwd_dq_tc_kernel<false>, grid, dim3(128), 0, (cudaStream_t)stream, an);
  int rc = check_launch("attn_bwd_dq_tc_kernel");
  if (rc != TT_OK) return rc;
  const int tpc = dkv_tiles_per_cta(Lmax, B * H);
  for (int i = 0; i < n; ++i) an.a[i].dkv_tpc = tpc;
  const dim3 grid2(B * H, ceil_div(Lmax, TC_BN * tpc), n);
  if (kv16) launch_k(attn_bwd_dkv_tc_kernel<true>, grid2, dim3(128), 0, (cudaStream_t)stream, an);
  else launch_k(attn_bwd_dkv_tc_kernel<false>, grid2, dim3(128), 0, (cudaStream_t)stream, an);
  return check_launch("attn_bwd_dkv_tc_kernel");
}

extern "C" int tt_attn_decode_hm_multi(const TtAttnCtx* ctx, int n, int B, int H, int D, int zero_row, void* stream) {
  TT_REQUIRE(ctx && n >= 1 && n <= ATTN_MAX_CTX, "tt_attn_decode_hm_multi: 1..%d contexts", ATTN_MAX_CTX);
  TT_REQUIRE(D == TC_D, "tt_attn_decode_hm_multi: head_dim must be %d (got %d)", TC_D, D);
  if (B <= 0) return TT_OK;
  AttnArgsN an{};
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    const TtAttnCtx& c = ctx[i];
    TT_REQUIRE(c.q && c.out && (c.S == 0 || (c.k && c.v)), "tt_attn_decode_hm_multi: null pointer");
    TT_REQUIRE((c.bias_k == nullptr) == (c.bias_v == nullptr), "tt_attn_decode_hm_multi: bias_k/bias_v mismatch");
    TT_REQUIRE(c.ldq % 2 == 0 && c.ldo % 4 == 0, "tt_attn_decode_hm_multi: ldq must be even, ldo a multiple of 4");
    AttnArgs& a = an.a[i];
    a.q = c.q; a.k = reinterpret_cast<const float*>(c.k); a.v = reinterpret_cast<const float*>(c.v);
    a.bias_k = c.bias_k; a.bias_v = c.bias_v; a.mask = c.mask;
    a.out = c.out; a.lse = c.lse; a.T = 1; a.B = B; a.S = c.S; a.H = H; a.zero_row = zero_row;
    a.ldq = c.ldq; a.ldo = c.ldo; a.ldkv = 0;
    a.kv_j_stride = TC_D; a.kv_b_stride = static_cast<long long>(H) * c.S * TC_D;
    a.kv_h_stride = static_cast<long long>(c.S) * TC_D;
    a.kv_len = c.kv_len;
    TT_REQUIRE(c.out16 == nullptr || (c.ldo16 % 8 == 0 && (reinterpret_cast<uintptr_t>(c.out16) & 15) == 0),
               "tt_attn_decode_hm_multi: out16 must be 16-byte aligned with a row pitch that is a multiple of 8");
    a.out16 = reinterpret_cast<__nv_bfloat16*>(c.out16); a.ldo16 = c.ldo16;
    const int L = c.S + (c.bias_k ? 1 : 0) + (zero_row ? 1 : 0);
    TT_REQUIRE(L > 0, "tt_attn_decode_hm_multi: empty key set");
    a.Lp = (L + 3) & ~3;
    const size_t need = static_cast<size_t>(DEC_HG) * a.Lp * sizeof(float);
    smem = need > smem ? need : smem;
  }
  TT_REQUIRE(smem <= 40 * 1024, "tt_attn_decode_hm_multi: key set too long for the decode kernel");
  launch_k(attn_decode_kernel<true>, dim3(B, ceil_div(H, DEC_HG), n), dim3(DEC_HG * 32), smem, (cudaStream_t)stream, an);
  return check_launch("attn_decode_kernel");
}
