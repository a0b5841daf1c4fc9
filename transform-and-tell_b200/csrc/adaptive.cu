// Adaptive softmax / adaptive loss pieces (tell/modules/softmax.py:144-222,
// tell/modules/criteria/adaptive_loss.py:27-73) and row gather/scatter helpers.
// The logit GEMMs run in gemm.cu; this file holds the data-dependent parts the reference does
// with boolean masks and nonzero() host syncs: cluster target remap + ordered compaction on the
// device, row-wise log-sum-exp cross entropy with ignore_index, its backward, and the
// full-vocabulary log-prob / argmax used by greedy decoding.
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int AD_MAX_CLUSTERS = 8;

struct AdaptiveCut {
  int n_clusters;                  // head + tails
  int cutoff[AD_MAX_CLUSTERS];     // cutoff[0] = head words, ..., cutoff[n_clusters-1] = vocab
};

// softmax.py:144-167 adapt_target.  Single CTA, ordered (deterministic) compaction.
//   head_target[n] = id                       if id < cutoff[0]
//                  = cutoff[0] + i            if cutoff[i] <= id < cutoff[i+1]
//   tail i: idx[i][slot] = n, local[i][slot] = id - cutoff[i]  for rows in band i, in row order.
__global__ void adaptive_prepare_kernel(const long long* __restrict__ target, int N,
                                        AdaptiveCut cut, int pad_idx,
                                        int* __restrict__ head_target, int* __restrict__ tail_idx,
                                        int* __restrict__ tail_local, int* __restrict__ tail_count,
                                        int* __restrict__ ntokens) {
  pdl_prologue();
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // head targets + non-pad count
  int cnt = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const long long id = target[n];
    int ht = static_cast<int>(id);
    for (int i = 0; i + 1 < cut.n_clusters; ++i)
      if (id >= cut.cutoff[i] && id < cut.cutoff[i + 1]) ht = cut.cutoff[0] + i;
    head_target[n] = ht;
    cnt += (id != pad_idx) ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) warp_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < nw; ++w) s += warp_cnt[w];
    if (ntokens) *ntokens = s;
  }
  // ordered compaction per tail
  for (int i = 0; i + 1 < cut.n_clusters; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int n0 = 0; n0 < N; n0 += blockDim.x) {
      const int n = n0 + threadIdx.x;
      long long id = -1;
      if (n < N) id = target[n];
      const bool in = (id >= cut.cutoff[i] && id < cut.cutoff[i + 1]);
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (lane == 0) warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = base;
      for (int w = 0; w < warp; ++w) off += warp_cnt[w];
      if (in) {
        const int slot = off + __popc(bal & ((1u << lane) - 1u));
        tail_idx[static_cast<long long>(i) * N + slot] = n;
        tail_local[static_cast<long long>(i) * N + slot] = static_cast<int>(id - cut.cutoff[i]);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int s = base;
        for (int w = 0; w < nw; ++w) s += warp_cnt[w];
        base = s;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) tail_count[i] = base;
  }
}

// dst[i,:] = src[idx[i],:] for i < count, else 0.   (X.index_select(0, target_idxs[i]), softmax.py:185)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                   const int* __restrict__ count_ptr, float* __restrict__ dst,
                                   int cap, int E4) {
  pdl_prologue();
  const int count = count_ptr ? min(*count_ptr, cap) : cap;
  const long long total = static_cast<long long>(cap) * E4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / E4), c = static_cast<int>(i - static_cast<long long>(r) * E4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < count) v = __ldg(reinterpret_cast<const float4*>(src) + static_cast<long long>(idx[r]) * E4 + c);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}
// dst[idx[i],:] += src[i,:] for i < count (atomic: idx may repeat for embedding grads).
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                        const int* __restrict__ count_ptr, float* __restrict__ dst,
                                        int cap, int E) {
  pdl_prologue();
  const int count = count_ptr ? min(*count_ptr, cap) : cap;
  const long long total = static_cast<long long>(count) * E;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / E), c = static_cast<int>(i - static_cast<long long>(r) * E);
    const int d = idx[r];
    if (d >= 0) atomicAdd(dst + static_cast<long long>(d) * E + c, src[i]);
  }
}

// F.cross_entropy(logits, target, ignore_index, reduction='sum') per row + saved lse.
// VEC: 16-byte loads (row pitch a multiple of 4 floats, aligned base -- ops.f32_padded); the
// 30265-wide tail cluster has only a few dozen active rows, so per-thread trip count is what
// bounds it.
template <bool VEC>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ logits, long long ld, const int* __restrict__ target,
              const int* __restrict__ count_ptr, int M, int V, int ignore_index,
              float* __restrict__ lse, float* __restrict__ row_loss) {
  pdl_prologue();
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int count = count_ptr ? min(*count_ptr, M) : M;
  if (r >= count) {
    if (threadIdx.x == 0) { row_loss[r] = 0.f; lse[r] = 0.f; }
    return;
  }
  const float* x = logits + r * ld;
  const int nv = VEC ? (V >> 2) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float m = -INFINITY;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const float4 v = x4[j];
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int j = nv * 4 + threadIdx.x; j < V; j += blockDim.x) m = fmaxf(m, x[j]);
  m = block_max(m, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const float4 v = x4[j];
    s += (expf(v.x - m) + expf(v.y - m)) + (expf(v.z - m) + expf(v.w - m));
  }
  for (int j = nv * 4 + threadIdx.x; j < V; j += blockDim.x) s += expf(x[j] - m);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float l = m + logf(s);
    lse[r] = l;
    const int t = target[r];
    row_loss[r] = (t == ignore_index) ? 0.f : (l - x[t]);
  }
}

// In place: logits <- (softmax - onehot(target)) * scale ; ignored / out-of-range rows <- 0.
__global__ void __launch_bounds__(256)
ce_bwd_kernel(float* __restrict__ logits, long long ld, const int* __restrict__ target,
              const int* __restrict__ count_ptr, int M, int V, int ignore_index,
              const float* __restrict__ lse, const float* __restrict__ scale_ptr) {
  pdl_prologue();
  const int r = blockIdx.x;
  const int count = count_ptr ? min(*count_ptr, M) : M;
  float* x = logits + r * ld;
  const int t = r < count ? target[r] : ignore_index;
  if (r >= count || t == ignore_index) {
    for (int j = threadIdx.x; j < V; j += blockDim.x) x[j] = 0.f;
    return;
  }
  const float l = lse[r];
  const float sc = scale_ptr ? *scale_ptr : 1.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    float g = expf(x[j] - l);
    if (j == t) g -= 1.f;
    x[j] = g * sc;
  }
}

// Same gradient written straight as the bf16 GEMM operand of the backward GEMMs (throughput mode):
// out16 <- bf16((softmax - onehot) * scale), logits untouched.  Replaces the in-place fp32 pass plus
// a separate fp32 -> bf16 cast of the whole [M, V] matrix.  With zero_round > 0 and a device-side
// row count, rows >= round_up(count, zero_round) are left unwritten: the row-limited GEMMs that
// consume the operand (m_limit / k_limit = the same count) never read them, which saves ~145 MB of
// zero fill per step for the two tail clusters.
template <bool VEC>
__global__ void __launch_bounds__(256)
ce_bwd16_kernel(const float* __restrict__ logits, long long ld, const int* __restrict__ target,
                const int* __restrict__ count_ptr, int M, int V, int ignore_index,
                const float* __restrict__ lse, const float* __restrict__ scale_ptr,
                __nv_bfloat16* __restrict__ out, long long ld16, int zero_round) {
  pdl_prologue();
  const int r = blockIdx.x;
  const int count = count_ptr ? min(*count_ptr, M) : M;
  const float* x = logits + r * ld;
  __nv_bfloat16* o = out + r * ld16;
  const int nv = VEC ? (V >> 2) : 0;
  const int t = r < count ? target[r] : ignore_index;
  if (r >= count || t == ignore_index) {
    // at least one zero_round of rows is always zeroed: a k_limit GEMM reads one k-block even
    // when the count is 0
    if (r >= count && zero_round > 0 && count_ptr != nullptr &&
        r >= max((count + zero_round - 1) / zero_round, 1) * zero_round)
      return;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) reinterpret_cast<uint2*>(o)[j] = make_uint2(0u, 0u);
    for (int j = nv * 4 + threadIdx.x; j < V; j += blockDim.x) o[j] = __float2bfloat16_rn(0.f);
    return;
  }
  const float l = lse[r];
  const float sc = scale_ptr ? *scale_ptr : 1.f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    const float4 v = x4[j];
    float g[4] = {expf(v.x - l), expf(v.y - l), expf(v.z - l), expf(v.w - l)};
    const int d = t - 4 * j;
    if (d >= 0 && d < 4) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e == d) g[e] -= 1.f;
    }
    uint2 u;
    u.x = pack_bf16(g[0] * sc, g[1] * sc);
    u.y = pack_bf16(g[2] * sc, g[3] * sc);
    reinterpret_cast<uint2*>(o)[j] = u;
  }
  for (int j = nv * 4 + threadIdx.x; j < V; j += blockDim.x) {
    float g = expf(x[j] - l);
    if (j == t) g -= 1.f;
    o[j] = __float2bfloat16_rn(g * sc);
  }
}

// loss = sum(row_loss[0..n)) / ln2 / ntokens ; scale = upstream / (ln2 * ntokens)
// transformer_faces_objects.py:85-90.  Single CTA, fixed summation order (deterministic).
__global__ void loss_finalize_kernel(const float* __restrict__ row_loss, long long n,
                                     const int* __restrict__ ntokens, float* __restrict__ loss,
                                     float* __restrict__ scale) {
  pdl_prologue();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float denom = 0.69314718055994530942f * static_cast<float>(*ntokens);
    if (loss) *loss = s / denom;
    if (scale) *scale = 1.f / denom;
  }
}

// softmax.py:193-222 get_log_prob (+ topk(1) of transformer_faces_objects.py:443-464).
// One CTA per row.  head [M, c0 + n_tails], tails[i] [M, V_i].
struct LogProbArgs {
  const float* head; long long ld_head;
  const float* tail[AD_MAX_CLUSTERS - 1]; long long ld_tail[AD_MAX_CLUSTERS - 1];
  AdaptiveCut cut;
  float* log_probs;      // [M, vocab] or null
  long long* argmax_id;  // [M] or null
  float* argmax_lp;      // [M] or null
  int M;
};
// Block-wide arg-max of (value, index) pairs, ties -> lowest index (torch.topk on CPU returns the
// first maximal element).  Result valid in every thread.
__device__ __forceinline__ void block_argmax_256(float& v, int& i, float* sv, int* si) {
  __syncthreads();
  sv[threadIdx.x] = v;
  si[threadIdx.x] = i;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v2 = sv[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
        sv[threadIdx.x] = v2;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  v = sv[0];
  i = si[0];
}

// One cluster of logits x[0..n): maximum, arg-max (lowest index among ties) and sum exp(x - max).
// The arg-max of the log-probabilities of a cluster is the arg-max of its raw logits (the
// normaliser is constant within the cluster), so greedy decoding needs two passes, not three.
// 16-byte loads when the row is aligned (the logits buffers have a padded pitch).
__device__ __forceinline__ void cluster_stats(const float* __restrict__ x, int n, float* red, float* sv,
                                              int* si, float& mx, int& amax, float& sumexp) {
  float v = -INFINITY;
  int vi = 0x7fffffff;
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int n4 = al ? (n >> 2) : 0;
  for (int j = threadIdx.x; j < n4; j += blockDim.x) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(x) + j);
    if (q.x > v) { v = q.x; vi = 4 * j; }
    if (q.y > v) { v = q.y; vi = 4 * j + 1; }
    if (q.z > v) { v = q.z; vi = 4 * j + 2; }
    if (q.w > v) { v = q.w; vi = 4 * j + 3; }
  }
  for (int j = 4 * n4 + threadIdx.x; j < n; j += blockDim.x) {
    const float q = x[j];
    if (q > v) { v = q; vi = j; }
  }
  block_argmax_256(v, vi, sv, si);
  mx = v;
  amax = vi;
  float s = 0.f;
  for (int j = threadIdx.x; j < n4; j += blockDim.x) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(x) + j);
    s += (expf(q.x - v) + expf(q.y - v)) + (expf(q.z - v) + expf(q.w - v));
  }
  for (int j = 4 * n4 + threadIdx.x; j < n; j += blockDim.x) s += expf(x[j] - v);
  sumexp = block_sum(s, red);
}

__global__ void __launch_bounds__(256)
adaptive_logprob_kernel(const LogProbArgs a) {
  pdl_prologue();
  __shared__ float red[32];
  __shared__ float best_v[256];
  __shared__ int best_i[256];
  const int r = blockIdx.x;
  const int c0 = a.cut.cutoff[0];
  const int n_tails = a.cut.n_clusters - 1;
  const int vocab = a.cut.cutoff[a.cut.n_clusters - 1];
  const int head_n = c0 + n_tails;
  const float* h = a.head + r * a.ld_head;
  // head: the softmax runs over the c0 words AND the cluster priors; the arg-max candidates are
  // the words only
  float wm, ws_unused;
  int wi;
  cluster_stats(h, c0, red, best_v, best_i, wm, wi, ws_unused);
  float m = wm;
  for (int i = 0; i < n_tails; ++i) m = fmaxf(m, h[c0 + i]);
  float s = 0.f;
  {
    const bool al = (reinterpret_cast<uintptr_t>(h) & 15) == 0;
    const int n4 = al ? (head_n >> 2) : 0;
    for (int j = threadIdx.x; j < n4; j += blockDim.x) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(h) + j);
      s += (expf(q.x - m) + expf(q.y - m)) + (expf(q.z - m) + expf(q.w - m));
    }
    for (int j = 4 * n4 + threadIdx.x; j < head_n; j += blockDim.x) s += expf(h[j] - m);
  }
  s = block_sum(s, red);
  const float head_lse = m + logf(s);
  float bv = wm - head_lse;
  int bi = wi;
  if (a.log_probs)
    for (int j = threadIdx.x; j < c0; j += blockDim.x)
      a.log_probs[static_cast<long long>(r) * vocab + j] = h[j] - head_lse;
  for (int i = 0; i < n_tails; ++i) {
    const int Vi = a.cut.cutoff[i + 1] - a.cut.cutoff[i];
    const float* t = a.tail[i] + r * a.ld_tail[i];
    float tm, ts;
    int ti;
    cluster_stats(t, Vi, red, best_v, best_i, tm, ti, ts);
    const float prior = h[c0 + i] - head_lse;
    const float off = prior - (tm + logf(ts));
    const float lp = tm + off;
    if (lp > bv) { bv = lp; bi = a.cut.cutoff[i] + ti; }      // ties keep the lower (earlier) id
    if (a.log_probs)
      for (int j = threadIdx.x; j < Vi; j += blockDim.x)
        a.log_probs[static_cast<long long>(r) * vocab + a.cut.cutoff[i] + j] = t[j] + off;
  }
  if (threadIdx.x == 0) {
    if (a.argmax_id) a.argmax_id[r] = bi;
    if (a.argmax_lp) a.argmax_lp[r] = bv;
  }
}

static inline int flat_grid2(long long n) {
  long long g = ceil_div_ll(n, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

static int make_cut(const int* cutoffs, int n_clusters, AdaptiveCut* cut) {
  TT_REQUIRE(cutoffs && n_clusters >= 1 && n_clusters <= AD_MAX_CLUSTERS,
             "adaptive: n_clusters must be in [1,%d]", AD_MAX_CLUSTERS);
  cut->n_clusters = n_clusters;
  for (int i = 0; i < n_clusters; ++i) cut->cutoff[i] = cutoffs[i];
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_adaptive_prepare(const long long* target, int N, const int* cutoffs,
                                   int n_clusters, int pad_idx, int* head_target, int* tail_idx,
                                   int* tail_local, int* tail_count, int* ntokens, void* stream) {
  TT_REQUIRE(target && head_target, "tt_adaptive_prepare: null pointer");
  AdaptiveCut cut;
  int rc = make_cut(cutoffs, n_clusters, &cut);
  if (rc != TT_OK) return rc;
  TT_REQUIRE(n_clusters == 1 || (tail_idx && tail_local && tail_count),
             "tt_adaptive_prepare: null tail buffers");
  if (N <= 0) return TT_OK;
  launch_k(adaptive_prepare_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, 
      target, N, cut, pad_idx, head_target, tail_idx, tail_local, tail_count, ntokens);
  return check_launch("adaptive_prepare_kernel");
}

extern "C" int tt_gather_rows(const float* src, const int* idx, const int* count_ptr, float* dst,
                              int cap, int E, void* stream) {
  TT_REQUIRE(src && idx && dst, "tt_gather_rows: null pointer");
  TT_REQUIRE(E % 4 == 0, "tt_gather_rows: E must be a multiple of 4");
  if (cap <= 0) return TT_OK;
  launch_k(gather_rows_kernel, dim3(flat_grid2(static_cast<long long>(cap) * (E / 4))), dim3(256), 0, (cudaStream_t)stream, src, idx, count_ptr, dst, cap, E / 4);
  return check_launch("gather_rows_kernel");
}

extern "C" int tt_scatter_add_rows(const float* src, const int* idx, const int* count_ptr,
                                   float* dst, int cap, int E, void* stream) {
  TT_REQUIRE(src && idx && dst, "tt_scatter_add_rows: null pointer");
  if (cap <= 0) return TT_OK;
  launch_k(scatter_add_rows_kernel, dim3(flat_grid2(static_cast<long long>(cap) * E)), dim3(256), 0, (cudaStream_t)stream, src, idx, count_ptr, dst, cap, E);
  return check_launch("scatter_add_rows_kernel");
}

extern "C" int tt_ce_fwd(const float* logits, long long ld, const int* target,
                         const int* count_ptr, int M, int V, int ignore_index, float* lse,
                         float* row_loss, void* stream) {
  TT_REQUIRE(logits && target && lse && row_loss, "tt_ce_fwd: null pointer");
  if (M <= 0) return TT_OK;
  const bool vec = (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (vec)
    launch_k(ce_fwd_kernel<true>, dim3(M), dim3(256), 0, (cudaStream_t)stream, logits, ld, target, count_ptr,
             M, V, ignore_index, lse, row_loss);
  else
    launch_k(ce_fwd_kernel<false>, dim3(M), dim3(256), 0, (cudaStream_t)stream, logits, ld, target, count_ptr,
             M, V, ignore_index, lse, row_loss);
  return check_launch("ce_fwd_kernel");
}

extern "C" int tt_ce_bwd_bf16(const float* logits, long long ld, const int* target,
                              const int* count_ptr, int M, int V, int ignore_index,
                              const float* lse, const float* scale_ptr, void* out16,
                              long long ld16, int zero_round, void* stream) {
  TT_REQUIRE(logits && target && lse && out16, "tt_ce_bwd_bf16: null pointer");
  TT_REQUIRE(ld >= V && ld16 >= V && zero_round >= 0, "tt_ce_bwd_bf16: bad pitch / zero_round");
  if (M <= 0) return TT_OK;
  const bool vec = (ld % 4 == 0) && (ld16 % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out16) & 7) == 0;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out16);
  if (vec)
    launch_k(ce_bwd16_kernel<true>, dim3(M), dim3(256), 0, (cudaStream_t)stream, logits, ld, target,
             count_ptr, M, V, ignore_index, lse, scale_ptr, o, ld16, zero_round);
  else
    launch_k(ce_bwd16_kernel<false>, dim3(M), dim3(256), 0, (cudaStream_t)stream, logits, ld, target,
             count_ptr, M, V, ignore_index, lse, scale_ptr, o, ld16, zero_round);
  return check_launch("ce_bwd16_kernel");
}

extern "C" int tt_ce_bwd(float* logits, long long ld, const int* target, const int* count_ptr,
                         int M, int V, int ignore_index, const float* lse, const float* scale_ptr,
                         void* stream) {
  TT_REQUIRE(logits && target && lse, "tt_ce_bwd: null pointer");
  if (M <= 0) return TT_OK;
  launch_k(ce_bwd_kernel, dim3(M), dim3(256), 0, (cudaStream_t)stream, logits, ld, target, count_ptr, M, V,
                                                     ignore_index, lse, scale_ptr);
  return check_launch("ce_bwd_kernel");
}

extern "C" int tt_loss_finalize(const float* row_loss, long long n, const int* ntokens,
                                float* loss, float* scale, void* stream) {
  TT_REQUIRE(row_loss && ntokens, "tt_loss_finalize: null pointer");
  launch_k(loss_finalize_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, row_loss, n, ntokens, loss, scale);
  return check_launch("loss_finalize_kernel");
}

extern "C" int tt_adaptive_logprob(const float* head, long long ld_head, const float* const* tails,
                                   const long long* ld_tails, const int* cutoffs, int n_clusters,
                                   int M, float* log_probs, long long* argmax_id, float* argmax_lp,
                                   void* stream) {
  TT_REQUIRE(head, "tt_adaptive_logprob: null pointer");
  LogProbArgs a{};
  int rc = make_cut(cutoffs, n_clusters, &a.cut);
  if (rc != TT_OK) return rc;
  a.head = head; a.ld_head = ld_head;
  for (int i = 0; i + 1 < n_clusters; ++i) {
    TT_REQUIRE(tails && tails[i] && ld_tails, "tt_adaptive_logprob: null tail %d", i);
    a.tail[i] = tails[i];
    a.ld_tail[i] = ld_tails[i];
  }
  a.log_probs = log_probs; a.argmax_id = argmax_id; a.argmax_lp = argmax_lp; a.M = M;
  if (M <= 0) return TT_OK;
  launch_k(adaptive_logprob_kernel, dim3(M), dim3(256), 0, (cudaStream_t)stream, a);
  return check_launch("adaptive_logprob_kernel");
}
