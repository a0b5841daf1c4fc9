// Multi-head cross-attention core of tell/modules/attention/multi_head.py:355-466 in the
// static_kv=True, incremental_state=None mode used by every decoder layer
// (decoder_faces_objects.py:272-352):
//   keys   = [ K(ctx) rows 0..S-1 ; bias_k row ; zero row ]      (add_bias_kv, add_zero_attn)
//   scores = q . k^T  (q already scaled by d^-0.5), key_padding_mask -> -inf on ctx rows only
//   P      = softmax_fp32(scores);  P = dropout(P);  out = P . V
// fp32 SIMT, flash-style (online softmax over 64-key tiles, nothing of size T x S touches HBM).
// Forward saves the row log-sum-exp; backward is two kernels (dQ: grid over query tiles,
// dK/dV: grid over key tiles) that recompute P from q, k and lse -- no atomics except for the
// shared bias_k / bias_v rows.
// Layout: q/out [T,B,E], k/v [S,B,E], head h = columns [h*D, (h+1)*D), D = 64 (or 16/32 in tests).
#include "attention_args.cuh"
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int AT_KT = 64;     // keys per tile
constexpr int AT_QT = 16;     // queries per CTA in fwd / dq kernels (4 per warp)
constexpr int AT_THREADS = 128;
constexpr int AT_QW = 4;      // queries per warp

// Key tile loader: rows j0..j0+63 of the extended key sequence into smem (pitch D+1).
template <int D>
__device__ __forceinline__ void load_kv_tile(const AttnArgs& a, int b, int h, int j0, int L,
                                             float (*ks)[D + 1], float (*vs)[D + 1],
                                             float* kvalid) {
  const int E = a.H * D;
  const int has_bias = a.bias_k != nullptr;
  for (int i = threadIdx.x; i < AT_KT * D; i += blockDim.x) {
    const int jj = i / D, c = i - jj * D;
    const int j = j0 + jj;
    float kv = 0.f, vv = 0.f;
    if (j < a.S) {
      const long long off = (static_cast<long long>(j) * a.B + b) * a.ldkv + h * D + c;
      kv = __ldg(a.k + off);
      if (vs) vv = __ldg(a.v + off);
    } else if (has_bias && j == a.S) {
      kv = __ldg(a.bias_k + h * D + c);
      if (vs) vv = __ldg(a.bias_v + h * D + c);
    }
    ks[jj][c] = kv;
    if (vs) vs[jj][c] = vv;
  }
  for (int jj = threadIdx.x; jj < AT_KT; jj += blockDim.x) {
    const int j = j0 + jj;
    float ok = 0.f;
    if (j < a.S) ok = (a.mask && a.mask[static_cast<long long>(b) * a.S + j]) ? 0.f : 1.f;
    else if (j < L) ok = 1.f;
    kvalid[jj] = ok;
  }
}

// ------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(AttnArgs a) {
  pdl_prologue();
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ float qs[AT_QT][D];
  __shared__ float ks[AT_KT][D + 1];
  __shared__ float vs[AT_KT][D + 1];
  __shared__ float ps[AT_THREADS / 32][AT_QW][AT_KT];
  __shared__ float kvalid[AT_KT];
  const int E = a.H * D;
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.y * AT_QT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;

  for (int i = threadIdx.x; i < AT_QT * D; i += blockDim.x) {
    const int qi = i / D, c = i - qi * D;
    const int t = q0 + qi;
    qs[qi][c] = t < a.T ? __ldg(a.q + (static_cast<long long>(t) * a.B + b) * a.ldq + h * D + c) : 0.f;
  }
  float m[AT_QW], l[AT_QW], o[AT_QW][(D + 31) / 32];
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int u = 0; u < (D + 31) / 32; ++u) o[i][u] = 0.f;
  }

  for (int j0 = 0; j0 < L; j0 += AT_KT) {
    __syncthreads();
    load_kv_tile<D>(a, b, h, j0, L, ks, vs, kvalid);
    __syncthreads();
    // scores: lane owns keys lane and lane+32 for the warp's 4 queries
    float s[AT_QW][2];
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 8
    for (int c = 0; c < D; ++c) {
      const float k0 = ks[lane][c], k1 = ks[lane + 32][c];
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) {
        const float qv = qs[warp * AT_QW + i][c];
        s[i][0] += qv * k0;
        s[i][1] += qv * k1;
      }
    }
    const bool ok0 = kvalid[lane] != 0.f, ok1 = kvalid[lane + 32] != 0.f;
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) {
      const float s0 = ok0 ? s[i][0] : -INFINITY, s1 = ok1 ? s[i][1] : -INFINITY;
      const float tmax = warp_max(fmaxf(s0, s1));
      const float mnew = fmaxf(m[i], tmax);
      // mnew can only be -inf if every key so far is masked; exp(-inf - -inf) guarded below
      const float corr = (m[i] == -INFINITY) ? 0.f : expf(m[i] - mnew);
      float p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - mnew);
      float p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - mnew);
      l[i] = l[i] * corr + warp_sum(p0 + p1);
      m[i] = mnew;
#pragma unroll
      for (int u = 0; u < (D + 31) / 32; ++u) o[i][u] *= corr;
      if (a.p_drop > 0.f) {
        const int t = q0 + warp * AT_QW + i;
        const unsigned long long base =
            (static_cast<unsigned long long>(bh) * a.T + t) * static_cast<unsigned long long>(L);
        p0 *= dropout_scale(a.seed, base + j0 + lane, a.p_drop, inv_keep);
        p1 *= dropout_scale(a.seed, base + j0 + lane + 32, a.p_drop, inv_keep);
      }
      ps[warp][i][lane] = p0;
      ps[warp][i][lane + 32] = p1;
    }
    __syncwarp();
    // out += P . V : lane owns output dims lane (+32)
    const int jmax = min(AT_KT, L - j0);
    for (int j = 0; j < jmax; ++j) {
#pragma unroll
      for (int u = 0; u < (D + 31) / 32; ++u) {
        const int c = lane + 32 * u;
        const float vv = (c < D) ? vs[j][c] : 0.f;
#pragma unroll
        for (int i = 0; i < AT_QW; ++i) o[i][u] += ps[warp][i][j] * vv;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    const int t = q0 + warp * AT_QW + i;
    if (t >= a.T) continue;
    const float inv_l = 1.f / l[i];
#pragma unroll
    for (int u = 0; u < (D + 31) / 32; ++u) {
      const int c = lane + 32 * u;
      if (c < D) a.out[(static_cast<long long>(t) * a.B + b) * a.ldo + h * D + c] = o[i][u] * inv_l;
    }
    if (lane == 0 && a.lse) a.lse[static_cast<long long>(bh) * a.T + t] = m[i] + logf(l[i]);
  }
}

// ------------------------------------------------------------------------------ head-averaged weights
// avg_w[b,t,j] = (1/H) sum_h softmax(q_h k_h^T)[t,j]   (multi_head.py:478-483, eval only, no dropout)
template <int D>
__global__ void __launch_bounds__(AT_THREADS)
attn_weights_kernel(const AttnArgs a) {
  pdl_prologue();
  __shared__ float qs[AT_QT][D];
  __shared__ float ks[AT_KT][D + 1];
  __shared__ float kvalid[AT_KT];
  const int E = a.H * D;
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.y * AT_QT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  for (int i = threadIdx.x; i < AT_QT * D; i += blockDim.x) {
    const int qi = i / D, c = i - qi * D;
    const int t = q0 + qi;
    qs[qi][c] = t < a.T ? __ldg(a.q + (static_cast<long long>(t) * a.B + b) * a.ldq + h * D + c) : 0.f;
  }
  const float inv_h = 1.f / a.H;
  for (int j0 = 0; j0 < L; j0 += AT_KT) {
    __syncthreads();
    load_kv_tile<D>(a, b, h, j0, L, ks, nullptr, kvalid);
    __syncthreads();
    float s[AT_QW][2];
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 8
    for (int c = 0; c < D; ++c) {
      const float k0 = ks[lane][c], k1 = ks[lane + 32][c];
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) {
        const float qv = qs[warp * AT_QW + i][c];
        s[i][0] += qv * k0;
        s[i][1] += qv * k1;
      }
    }
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) {
      const int t = q0 + warp * AT_QW + i;
      if (t >= a.T) continue;
      const float lse = a.lse[static_cast<long long>(bh) * a.T + t];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = lane + 32 * u, j = j0 + jj;
        if (j < L && kvalid[jj] != 0.f)
          atomicAdd(a.avg_w + (static_cast<long long>(b) * a.T + t) * L + j,
                    expf(s[i][u] - lse) * inv_h);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward dQ
// dP = dO . V^T ; dS = P * (dP*mask/(1-p) - Drow) ; dQ = dS . K ;  Drow = dO . O
template <int D>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(AttnArgs a) {
  pdl_prologue();
  a.seed = mix_seed(a.seed, a.step_ptr);
  __shared__ float qs[AT_QT][D];
  __shared__ float dos[AT_QT][D];
  __shared__ float ks[AT_KT][D + 1];
  __shared__ float vs[AT_KT][D + 1];
  __shared__ float dss[AT_THREADS / 32][AT_QW][AT_KT];
  __shared__ float kvalid[AT_KT];
  __shared__ float drow[AT_QT];
  const int E = a.H * D;
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.y * AT_QT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int i = threadIdx.x; i < AT_QT * D; i += blockDim.x) {
    const int qi = i / D, c = i - qi * D;
    const int t = q0 + qi;
    const long long tb = static_cast<long long>(t) * a.B + b;
    qs[qi][c] = t < a.T ? __ldg(a.q + tb * a.ldq + h * D + c) : 0.f;
    dos[qi][c] = t < a.T ? __ldg(a.dout + tb * a.ldo + h * D + c) : 0.f;
  }
  __syncthreads();
  // Drow[t] = sum_c dO[t,c] O[t,c]
  for (int qi = warp; qi < AT_QT; qi += AT_THREADS / 32) {
    const int t = q0 + qi;
    float d = 0.f;
    if (t < a.T)
      for (int c = lane; c < D; c += 32)
        d += dos[qi][c] * __ldg(a.out + (static_cast<long long>(t) * a.B + b) * a.ldo + h * D + c);
    d = warp_sum(d);
    if (lane == 0) drow[qi] = d;
  }
  float lse[AT_QW], dq[AT_QW][(D + 31) / 32];
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    const int t = q0 + warp * AT_QW + i;
    lse[i] = t < a.T ? a.lse[static_cast<long long>(bh) * a.T + t] : 0.f;
#pragma unroll
    for (int u = 0; u < (D + 31) / 32; ++u) dq[i][u] = 0.f;
  }
  for (int j0 = 0; j0 < L; j0 += AT_KT) {
    __syncthreads();
    load_kv_tile<D>(a, b, h, j0, L, ks, vs, kvalid);
    __syncthreads();
    float s[AT_QW][2], dp[AT_QW][2];
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) s[i][0] = s[i][1] = dp[i][0] = dp[i][1] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float k0 = ks[lane][c], k1 = ks[lane + 32][c];
      const float v0 = vs[lane][c], v1 = vs[lane + 32][c];
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) {
        const float qv = qs[warp * AT_QW + i][c], dv_ = dos[warp * AT_QW + i][c];
        s[i][0] += qv * k0; s[i][1] += qv * k1;
        dp[i][0] += dv_ * v0; dp[i][1] += dv_ * v1;
      }
    }
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) {
      const int t = q0 + warp * AT_QW + i;
      const unsigned long long base =
          (static_cast<unsigned long long>(bh) * a.T + t) * static_cast<unsigned long long>(L);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = lane + 32 * u;
        float dsv = 0.f;
        if (kvalid[jj] != 0.f && t < a.T) {
          const float p = expf(s[i][u] - lse[i]);
          float dpv = dp[i][u];
          if (a.p_drop > 0.f) dpv *= dropout_scale(a.seed, base + j0 + jj, a.p_drop, inv_keep);
          dsv = p * (dpv - drow[warp * AT_QW + i]);
        }
        dss[warp][i][jj] = dsv;
      }
    }
    __syncwarp();
    const int jmax = min(AT_KT, L - j0);
    for (int j = 0; j < jmax; ++j) {
#pragma unroll
      for (int u = 0; u < (D + 31) / 32; ++u) {
        const int c = lane + 32 * u;
        const float kv = (c < D) ? ks[j][c] : 0.f;
#pragma unroll
        for (int i = 0; i < AT_QW; ++i) dq[i][u] += dss[warp][i][j] * kv;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    const int t = q0 + warp * AT_QW + i;
    if (t >= a.T) continue;
#pragma unroll
    for (int u = 0; u < (D + 31) / 32; ++u) {
      const int c = lane + 32 * u;
      if (c < D) a.dq[(static_cast<long long>(t) * a.B + b) * a.ldq + h * D + c] = dq[i][u];
    }
  }
}

// ------------------------------------------------------------------------------------------ backward dK,dV
// One CTA per (b,h, 64-key tile); loops over query tiles of 16.
//   dV[j,:] = sum_t Pdrop[t,j] dO[t,:]     dK[j,:] = sum_t dS[t,j] q[t,:]
// thread layout for the accumulation: thread = (key jj = tid/2, half = tid&1) owns D/2 dims.
template <int D>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dkv_kernel(AttnArgs a) {
  pdl_prologue();
  a.seed = mix_seed(a.seed, a.step_ptr);
  extern __shared__ float dkv_smem[];
  float (*qs)[D] = reinterpret_cast<float (*)[D]>(dkv_smem);
  float (*dos)[D] = reinterpret_cast<float (*)[D]>(dkv_smem + AT_QT * D);
  float (*ks)[D + 1] = reinterpret_cast<float (*)[D + 1]>(dkv_smem + 2 * AT_QT * D);
  float (*vs)[D + 1] = reinterpret_cast<float (*)[D + 1]>(dkv_smem + 2 * AT_QT * D + AT_KT * (D + 1));
  float* tail = dkv_smem + 2 * AT_QT * D + 2 * AT_KT * (D + 1);
  float (*pss)[AT_KT + 1] = reinterpret_cast<float (*)[AT_KT + 1]>(tail);                       // dropped P
  float (*dss)[AT_KT + 1] = reinterpret_cast<float (*)[AT_KT + 1]>(tail + AT_QT * (AT_KT + 1));  // dS
  float* kvalid = tail + 2 * AT_QT * (AT_KT + 1);
  float* drow = kvalid + AT_KT;
  float* lses = drow + AT_QT;
  const int E = a.H * D;
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  const int j0 = blockIdx.y * AT_KT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int has_bias = a.bias_k != nullptr;
  const int L = a.S + has_bias + (a.zero_row ? 1 : 0);
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  load_kv_tile<D>(a, b, h, j0, L, ks, vs, kvalid);
  constexpr int HD = D / 2;
  const int kj = threadIdx.x >> 1, half = threadIdx.x & 1;
  float dk[HD], dv[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) dk[c] = dv[c] = 0.f;

  for (int q0 = 0; q0 < a.T; q0 += AT_QT) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_QT * D; i += blockDim.x) {
      const int qi = i / D, c = i - qi * D;
      const int t = q0 + qi;
      const long long tb = static_cast<long long>(t) * a.B + b;
      qs[qi][c] = t < a.T ? __ldg(a.q + tb * a.ldq + h * D + c) : 0.f;
      dos[qi][c] = t < a.T ? __ldg(a.dout + tb * a.ldo + h * D + c) : 0.f;
    }
    __syncthreads();
    for (int qi = warp; qi < AT_QT; qi += AT_THREADS / 32) {
      const int t = q0 + qi;
      float d = 0.f;
      if (t < a.T)
        for (int c = lane; c < D; c += 32)
          d += dos[qi][c] * __ldg(a.out + (static_cast<long long>(t) * a.B + b) * a.ldo + h * D + c);
      d = warp_sum(d);
      if (lane == 0) {
        drow[qi] = d;
        lses[qi] = t < a.T ? a.lse[static_cast<long long>(bh) * a.T + t] : 0.f;
      }
    }
    __syncthreads();
    // scores for (query = warp*4+i, keys lane, lane+32)
    float s[AT_QW][2], dp[AT_QW][2];
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) s[i][0] = s[i][1] = dp[i][0] = dp[i][1] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float k0 = ks[lane][c], k1 = ks[lane + 32][c];
      const float v0 = vs[lane][c], v1 = vs[lane + 32][c];
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) {
        const float qv = qs[warp * AT_QW + i][c], dv_ = dos[warp * AT_QW + i][c];
        s[i][0] += qv * k0; s[i][1] += qv * k1;
        dp[i][0] += dv_ * v0; dp[i][1] += dv_ * v1;
      }
    }
#pragma unroll
    for (int i = 0; i < AT_QW; ++i) {
      const int qi = warp * AT_QW + i;
      const int t = q0 + qi;
      const unsigned long long base =
          (static_cast<unsigned long long>(bh) * a.T + t) * static_cast<unsigned long long>(L);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int jj = lane + 32 * u;
        float pd = 0.f, dsv = 0.f;
        if (kvalid[jj] != 0.f && t < a.T) {
          const float p = expf(s[i][u] - lses[qi]);
          float sc = 1.f;
          if (a.p_drop > 0.f) sc = dropout_scale(a.seed, base + j0 + jj, a.p_drop, inv_keep);
          pd = p * sc;
          dsv = p * (dp[i][u] * sc - drow[qi]);
        }
        pss[qi][jj] = pd;
        dss[qi][jj] = dsv;
      }
    }
    __syncthreads();
    // accumulate: this thread's key kj, dims [half*HD, half*HD+HD)
#pragma unroll 4
    for (int qi = 0; qi < AT_QT; ++qi) {
      const float pd = pss[qi][kj], dsv = dss[qi][kj];
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dv[c] += pd * dos[qi][half * HD + c];
        dk[c] += dsv * qs[qi][half * HD + c];
      }
    }
  }
  const int j = j0 + kj;
  if (j < a.S) {
    const long long off = (static_cast<long long>(j) * a.B + b) * a.ldkv + h * D + half * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      if (a.dk) a.dk[off + c] = dk[c];
      if (a.dv) a.dv[off + c] = dv[c];
    }
  } else if (has_bias && j == a.S) {
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      if (a.dbias_k) atomicAdd(a.dbias_k + h * D + half * HD + c, dk[c]);
      if (a.dbias_v) atomicAdd(a.dbias_v + h * D + half * HD + c, dv[c]);
    }
  }
}

template <int D>
static int attn_launch(const AttnArgs& a, int mode, cudaStream_t s) {
  const int L = a.S + (a.bias_k ? 1 : 0) + (a.zero_row ? 1 : 0);
  if (mode == 0) {
    dim3 grid(a.B * a.H, ceil_div(a.T, AT_QT));
    launch_k(attn_fwd_kernel<D>, dim3(grid), dim3(AT_THREADS), 0, s, a);
    return check_launch("attn_fwd_kernel");
  } else if (mode == 1) {
    dim3 grid(a.B * a.H, ceil_div(a.T, AT_QT));
    launch_k(attn_bwd_dq_kernel<D>, dim3(grid), dim3(AT_THREADS), 0, s, a);
    int rc = check_launch("attn_bwd_dq_kernel");
    if (rc != TT_OK) return rc;
    dim3 grid2(a.B * a.H, ceil_div(L, AT_KT));
    constexpr int kDkvSmem =
        (2 * AT_QT * D + 2 * AT_KT * (D + 1) + 2 * AT_QT * (AT_KT + 1) + AT_KT + 2 * AT_QT) *
        static_cast<int>(sizeof(float));
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kDkvSmem);
      attr_set = true;
    }
    launch_k(attn_bwd_dkv_kernel<D>, dim3(grid2), dim3(AT_THREADS), kDkvSmem, s, a);
    return check_launch("attn_bwd_dkv_kernel");
  } else {
    dim3 grid(a.B * a.H, ceil_div(a.T, AT_QT));
    launch_k(attn_weights_kernel<D>, dim3(grid), dim3(AT_THREADS), 0, s, a);
    return check_launch("attn_weights_kernel");
  }
}

static int attn_dispatch(const AttnArgs& a, int D, int mode, cudaStream_t s) {
  switch (D) {
    case 64: return attn_launch<64>(a, mode, s);
    case 32: return attn_launch<32>(a, mode, s);
    case 16: return attn_launch<16>(a, mode, s);
    default:
      set_error("attention: head_dim %d unsupported (16, 32, 64)", D);
      return TT_ERR_INVALID;
  }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_attn_fwd(const float* q, const float* k, const float* v, const float* bias_k,
                           const float* bias_v, const uint8_t* key_padding_mask, float* out,
                           float* lse, int T, int B, int S, int H, int D, long long ldq,
                           long long ldkv, long long ldo, int zero_row, float p_drop,
                           unsigned long long seed, void* stream) {
  TT_REQUIRE(q && out && lse, "tt_attn_fwd: null pointer");
  TT_REQUIRE(S == 0 || (k && v), "tt_attn_fwd: null k/v with S > 0");
  TT_REQUIRE((bias_k == nullptr) == (bias_v == nullptr), "tt_attn_fwd: bias_k/bias_v mismatch");
  TT_REQUIRE(S + (bias_k ? 1 : 0) + (zero_row ? 1 : 0) > 0, "tt_attn_fwd: empty key set");
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.bias_k = bias_k; a.bias_v = bias_v; a.mask = key_padding_mask;
  a.out = out; a.lse = lse; a.T = T; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row;
  a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldo;
  a.p_drop = p_drop; a.seed = seed; a.step_ptr = rng_step_ptr();
  return attn_dispatch(a, D, 0, (cudaStream_t)stream);
}

extern "C" int tt_attn_bwd(const float* dout, const float* q, const float* k, const float* v,
                           const float* bias_k, const float* bias_v,
                           const uint8_t* key_padding_mask, const float* out, const float* lse,
                           float* dq, float* dk, float* dv, float* dbias_k, float* dbias_v, int T,
                           int B, int S, int H, int D, long long ldq, long long ldkv,
                           long long ldo, int zero_row, float p_drop, unsigned long long seed,
                           void* stream) {
  TT_REQUIRE(dout && q && out && lse && dq, "tt_attn_bwd: null pointer");
  TT_REQUIRE(S == 0 || (k && v), "tt_attn_bwd: null k/v with S > 0");
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.bias_k = bias_k; a.bias_v = bias_v; a.mask = key_padding_mask;
  a.out = const_cast<float*>(out); a.lse = const_cast<float*>(lse);
  a.T = T; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row; a.p_drop = p_drop; a.seed = seed;
  a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldo;
  a.step_ptr = rng_step_ptr();
  a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.dbias_k = dbias_k; a.dbias_v = dbias_v;
  return attn_dispatch(a, D, 1, (cudaStream_t)stream);
}

extern "C" int tt_attn_avg_weights(const float* q, const float* k, const float* bias_k,
                                   const uint8_t* key_padding_mask, const float* lse,
                                   float* avg_w, int T, int B, int S, int H, int D, long long ldq,
                                   long long ldkv, int zero_row, void* stream) {
  TT_REQUIRE(q && lse && avg_w, "tt_attn_avg_weights: null pointer");
  if (T <= 0 || B <= 0) return TT_OK;
  AttnArgs a{};
  a.q = q; a.k = k; a.v = k; a.bias_k = bias_k; a.bias_v = bias_k; a.mask = key_padding_mask;
  a.lse = const_cast<float*>(lse); a.avg_w = avg_w;
  a.T = T; a.B = B; a.S = S; a.H = H; a.zero_row = zero_row;
  a.ldq = ldq; a.ldkv = ldkv; a.ldo = ldq;
  return attn_dispatch(a, D, 2, (cudaStream_t)stream);
}
