// BertAdam optimizer step for all trainable tensors in THREE launches.
//
// The reference trains with `type: bert_adam` (expt/nytimes/9_transformer_objects/config.yaml:126-136
// -> allennlp 0.9 -> pytorch-pretrained-bert 0.6.2 `BertAdam.step`, called at
// tell/training/callback_apex_trainer.py:238).  Per parameter tensor and step it does
//     clip_grad_norm_(p, max_grad_norm)                  # PER-TENSOR l2 clip, eps 1e-6
//     m = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g       # no bias correction
//     update = m / (sqrt(v) + e) + weight_decay * p
//     p -= lr * schedule(step / t_total, warmup) * update ;  step += 1
// as ~10 torch kernels per tensor (x ~190 tensors).  Here:
//   1. adam_sumsq_kernel   : partial sum of g^2 per 8192-element chunk (deterministic, no atomics)
//   2. adam_prepare_kernel : one warp per tensor adds its partials in fixed order -> clip
//                            coefficient; thread 0 evaluates the lr schedule from the DEVICE step
//                            counter (CUDA-graph replays advance it), the NaN-loss skip flag
//                            (callback_apex_trainer.py:225-227) and bumps the counter
//   3. adam_update_kernel  : fused m/v/p update, float4, one chunk per CTA iteration
// HBM-bound: 4 B (g, pass 1) + 16 B read + 12 B written per element = 32 B/element.
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int ADAM_CHUNK = 8192;     // elements per work item (256 threads x 8 float4)

__device__ __forceinline__ int adam_find_seg(const TtAdamSeg* __restrict__ segs, int nsegs, int chunk) {
  int lo = 0, hi = nsegs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float red[8];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();                 // red[] may still be read from the previous call
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (w == 0) {
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
  }
  return t;                        // valid in thread 0
}

__global__ void __launch_bounds__(256) adam_sumsq_kernel(const TtAdamSeg* __restrict__ segs, int nsegs,
                                                         int total_chunks, float* __restrict__ partial) {
  pdl_prologue();
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const TtAdamSeg sg = segs[adam_find_seg(segs, nsegs, chunk)];
    const long long e0 = static_cast<long long>(chunk - sg.chunk0) * ADAM_CHUNK;
    const long long n = sg.n - e0 < ADAM_CHUNK ? sg.n - e0 : ADAM_CHUNK;
    const float* g = sg.g + e0;
    float s = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const int n4 = static_cast<int>(n >> 2);
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + i * 256;
        v[i] = idx < n4 ? __ldg(reinterpret_cast<const float4*>(g) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) s += g[i] * g[i];
    } else {
      for (int i = threadIdx.x; i < n; i += 256) s += g[i] * g[i];
    }
    s = block_sum_256(s);
    if (threadIdx.x == 0) partial[chunk] = s;
  }
}

// scratch layout (floats): [0] lr * schedule(step), [1] skip flag (1 = NaN loss: leave everything
// untouched), [2 .. 2+nsegs) per-tensor clip coefficient.
__global__ void __launch_bounds__(256) adam_prepare_kernel(const TtAdamSeg* __restrict__ segs, int nsegs,
                                                           const float* __restrict__ partial,
                                                           float* __restrict__ scratch, TtAdamHyper h,
                                                           long long* __restrict__ step,
                                                           const float* __restrict__ loss) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int sidx = warp; sidx < nsegs; sidx += nwarps) {
    const TtAdamSeg sg = segs[sidx];
    const int nch = static_cast<int>((sg.n + ADAM_CHUNK - 1) / ADAM_CHUNK);
    float s = 0.f;
    for (int i = lane; i < nch; i += 32) s += partial[sg.chunk0 + i];
    s = warp_sum(s);
    if (lane == 0) {
      float coef = 1.f;
      if (h.max_grad_norm > 0.f) {
        const float c = h.max_grad_norm / (sqrtf(s) + 1e-6f);     // torch clip_grad_norm_
        if (c < 1.f) coef = c;
      }
      scratch[2 + sidx] = coef;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const bool skip = loss != nullptr && isnan(*loss);
    const long long st = *step;
    double f = 1.0;
    if (h.t_total > 0 && h.schedule != TT_SCHED_NONE) {
      const double x = static_cast<double>(st) / static_cast<double>(h.t_total);
      const double w = static_cast<double>(h.warmup);
      if (x < w) f = x / w;
      else if (h.schedule == TT_SCHED_WARMUP_LINEAR) f = fmax((x - 1.0) / (w - 1.0), 0.0);
      else f = 1.0;                                               // warmup_constant
    }
    scratch[0] = static_cast<float>(static_cast<double>(h.lr) * f);
    scratch[1] = skip ? 1.f : 0.f;
    if (!skip) *step = st + 1;
  }
}

__global__ void __launch_bounds__(256) adam_update_kernel(const TtAdamSeg* __restrict__ segs, int nsegs,
                                                          int total_chunks,
                                                          const float* __restrict__ scratch,
                                                          TtAdamHyper h) {
  pdl_prologue();
  if (scratch[1] != 0.f) return;
  const float lr = scratch[0];
  const float b1 = h.b1, b2 = h.b2, c1 = 1.f - h.b1, c2 = 1.f - h.b2, eps = h.e, wd = h.weight_decay;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int sidx = adam_find_seg(segs, nsegs, chunk);
    const TtAdamSeg sg = segs[sidx];
    const float coef = scratch[2 + sidx];
    const long long e0 = static_cast<long long>(chunk - sg.chunk0) * ADAM_CHUNK;
    const long long n = sg.n - e0 < ADAM_CHUNK ? sg.n - e0 : ADAM_CHUNK;
    float* p = sg.p + e0;
    const float* g = sg.g + e0;
    float* m = sg.m + e0;
    float* v = sg.v + e0;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
      gg *= coef;
      mm = b1 * mm + c1 * gg;
      vv = b2 * vv + c2 * gg * gg;
      float u = mm / (sqrtf(vv) + eps);
      if (wd > 0.f) u += wd * pp;
      pp -= lr * u;
    };
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                      reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (al) {
      const int n4 = static_cast<int>(n >> 2);
#pragma unroll 2
      for (int idx = threadIdx.x; idx < n4; idx += 256) {
        float4 P = reinterpret_cast<float4*>(p)[idx];
        const float4 G = __ldg(reinterpret_cast<const float4*>(g) + idx);
        float4 M = reinterpret_cast<float4*>(m)[idx];
        float4 V = reinterpret_cast<float4*>(v)[idx];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y);
        upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        reinterpret_cast<float4*>(p)[idx] = P;
        reinterpret_cast<float4*>(m)[idx] = M;
        reinterpret_cast<float4*>(v)[idx] = V;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) upd(p[i], g[i], m[i], v[i]);
    } else {
      for (int i = threadIdx.x; i < n; i += 256) upd(p[i], g[i], m[i], v[i]);
    }
  }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_bertadam_chunk(void) { return ADAM_CHUNK; }

extern "C" int tt_bertadam_step(const TtAdamSeg* segs_dev, int nsegs, int total_chunks,
                                const TtAdamHyper* hyper, long long* step_dev, const float* loss_dev,
                                float* partial, float* scratch, void* stream) {
  TT_REQUIRE(segs_dev && hyper && step_dev && partial && scratch, "tt_bertadam_step: null pointer");
  TT_REQUIRE(nsegs > 0 && total_chunks >= nsegs, "tt_bertadam_step: bad segment table (%d segments, %d chunks)",
             nsegs, total_chunks);
  TT_REQUIRE(hyper->b1 >= 0.f && hyper->b1 < 1.f && hyper->b2 >= 0.f && hyper->b2 < 1.f,
             "tt_bertadam_step: betas must be in [0, 1)");
  TT_REQUIRE(hyper->lr >= 0.f && hyper->e >= 0.f, "tt_bertadam_step: negative lr or eps");
  TT_REQUIRE(hyper->schedule == TT_SCHED_NONE || hyper->t_total <= 0 ||
                 (hyper->warmup >= 0.f && hyper->warmup < 1.f),
             "tt_bertadam_step: warmup must be in [0, 1)");
  const TtAdamHyper h = *hyper;
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = num_sms() * 8;
  const int grid = total_chunks < cap ? total_chunks : cap;
  launch_k(adam_sumsq_kernel, dim3(grid), dim3(256), 0, st, segs_dev, nsegs, total_chunks, partial);
  int rc = check_launch("adam_sumsq_kernel");
  if (rc != TT_OK) return rc;
  launch_k(adam_prepare_kernel, dim3(ceil_div(nsegs, 8)), dim3(256), 0, st, segs_dev, nsegs,
           (const float*)partial, scratch, h, step_dev, loss_dev);
  rc = check_launch("adam_prepare_kernel");
  if (rc != TT_OK) return rc;
  launch_k(adam_update_kernel, dim3(grid), dim3(256), 0, st, segs_dev, nsegs, total_chunks,
           (const float*)scratch, h);
  return check_launch("adam_update_kernel");
}
