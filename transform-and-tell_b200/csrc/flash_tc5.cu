// RoBERTa self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), packed
// (variable-length) layout.  Replaces the mma.sync flash kernel of encoders.cu on the article
// encoder (fairseq RoBERTa self-attention, called at transformer_faces_objects.py:352-353).
//
//   qkv [R, 3E] bf16 (q pre-scaled by d^-0.5), sample b = rows cu[b] .. cu[b+1]; out [R, E] bf16.
//   CTA = 128 queries of one (b, h); loop over 128-key tiles.  D = 64.
//
// Per key tile:
//   S  = Q . K^T     tcgen05.mma 128x128x16 x4, A = Q tile, B = K tile (both K-major, 128B swizzle,
//                    TMA boxes of 64 dims x 128 rows straight out of the qkv matrix) -> TMEM cols 0..127
//   softmax          two threads per query row (64 keys each): tcgen05.ld the scores, online max /
//                    ex2 / sum in registers, P (bf16) written to shared memory in the K-major
//                    swizzled layout
//   PV = P . V       tcgen05.mma 128x64x16 x8, A = P (K-major), B = V tile as it sits in the qkv
//                    matrix (keys x dims = MN-major operand) -> TMEM cols 128..191
//   O  = O*alpha+PV  thread = row, O lives in registers (no TMEM read-modify-write)
// Software pipeline inside a CTA: S(t+1) is issued as soon as every thread holds S(t) in registers and
// PV(t) is collected in the middle of tile t+1, so both MMAs (and their commit -> mbarrier latency)
// run under the softmax arithmetic.  V tiles are double buffered, the K tile single buffered (its
// buffer is free once S(t) has retired).  Two CTAs per SM (97 KB smem, 256 TMEM columns each).
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int FT_BM = 128, FT_BN = 128, FT_D = 64;
constexpr int FT_TILE = FT_BM * FT_D * 2;            // 16 KB: 128 rows x 128 B
constexpr int FT_OFF_Q = 0;
constexpr int FT_OFF_K = FT_TILE;
constexpr int FT_OFF_V = 2 * FT_TILE;                // 2 stages
constexpr int FT_OFF_P = 4 * FT_TILE;                // 2 k-blocks of 64 keys: 32 KB
constexpr int FT_OFF_BAR = 6 * FT_TILE;              // barriers + TMEM slot
constexpr int FT_SMEM = 6 * FT_TILE + 64 + 1024 + 1024;   // + barriers, row exchange, alignment slack
constexpr int FT_TMEM_COLS = 256;                    // S: 128, PV: 64 (power of two >= 192)

// Barrier wait that parks in hardware: try_wait with a suspend-time hint instead of a tight poll
// loop, so the waiting lanes of one CTA do not take issue slots from the other CTA's softmax.
__device__ __forceinline__ void ft_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if ((clock64() - t0) > 4000000000LL) __trap();     // protocol bug -> trapped kernel, never a hang
  }
}
__device__ __forceinline__ void ft_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) ft_wait(bar, parity);
  __syncwarp();
}

__device__ __forceinline__ float ex2_approx(float x) {     // ex2.approx.ftz(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 256 threads: warp w works on TMEM lane quarter (w & 3) -- query rows 32*(w&3) .. +31, one row per
// lane -- and on column half (w >> 2): keys [64*half, +64) of the score tile, dims [32*half, +32) of
// the output.  The two threads of a row exchange their partial row maximum through shared memory.
__global__ void __launch_bounds__(256, 2)       // 2 CTAs per SM: <= 128 registers per thread
flash_tc5_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out,
                 const int* __restrict__ cu, int H) {
  extern __shared__ uint8_t ft_raw[];
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int q0 = blockIdx.x * FT_BM;
  const int row0 = cu[b];
  const int len = cu[b + 1] - row0;
  if (q0 >= len) return;                              // CTA-uniform, before any barrier / TMEM use
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + FT_OFF_Q;
  uint8_t* sK = smem + FT_OFF_K;
  uint8_t* sV = smem + FT_OFF_V;
  uint8_t* sP = smem + FT_OFF_P;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + FT_OFF_BAR);
  uint64_t* bar_k = bar_q + 1;
  uint64_t* bar_v = bar_q + 2;        // [2]
  uint64_t* bar_s = bar_q + 4;
  uint64_t* bar_pv = bar_q + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 6);
  float* xch = reinterpret_cast<float*>(smem + FT_OFF_BAR + 64);    // [2][128] row-statistic exchange
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int E = H * FT_D;
  const int ntiles = (len + FT_BN - 1) / FT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(&bar_v[0], 1);
    mbar_init(&bar_v[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_pv, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, FT_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);   // this warp's lanes

  const int E3k = E + h * FT_D, E3v = 2 * E + h * FT_D;
  if (threadIdx.x == 0) {       // Q, K(0), V(0), V(1)
    mbar_arrive_expect_tx(bar_q, FT_TILE);
    tma_load_2d(sQ, &tm_qkv, bar_q, h * FT_D, row0 + q0);
    mbar_arrive_expect_tx(bar_k, FT_TILE);
    tma_load_2d(sK, &tm_qkv, bar_k, E3k, row0);
    mbar_arrive_expect_tx(&bar_v[0], FT_TILE);
    tma_load_2d(sV, &tm_qkv, &bar_v[0], E3v, row0);
    if (ntiles > 1) {
      mbar_arrive_expect_tx(&bar_v[1], FT_TILE);
      tma_load_2d(sV + FT_TILE, &tm_qkv, &bar_v[1], E3v, row0 + FT_BN);
    }
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(FT_BM, FT_BN, false, false);
  constexpr uint32_t idesc_pv = umma_idesc_bf16(FT_BM, FT_D, false, true);
  const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ));
  const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK));
  const uint64_t dp = umma_desc_kmajor_sw128(smem_u32(sP));
  const uint64_t dv0 = umma_desc_mnmajor_sw128(smem_u32(sV));
  constexpr float LOG2E = 1.4426950408889634f;

  const int r = quarter * 32 + lane;         // query row of this thread within the tile
  float o[FT_D / 2];                         // this thread's 32 output dims
#pragma unroll
  for (int i = 0; i < FT_D / 2; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;              // running max (log2 units), partial sum (own columns)

  auto issue_s = [&]() {                     // thread 0: S = Q . K^T into TMEM cols [0, 128)
#pragma unroll
    for (int k = 0; k < FT_D / 16; ++k) umma_bf16(tmem_base, dq + k * 2, dk + k * 2, idesc_s, k != 0 ? 1u : 0u);
    umma_commit(bar_s);
  };
  auto add_pv = [&]() {                      // o += PV (this thread's 32 dims of its row)
    uint32_t pv[32];
    tmem_ld_32x32(tmem_row + FT_BN + half * 32, pv);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(pv[i]);
  };

  if (threadIdx.x == 0) {
    ft_wait(bar_q, 0);
    ft_wait(bar_k, 0);
    tcgen05_fence_after();
    issue_s();
  }
  // Software pipeline: the MMAs of the neighbouring tiles run under this tile's softmax --
  //   S(t+1) is issued as soon as every thread holds S(t) in registers,
  //   PV(t-1) is collected in the middle of tile t, PV(t) is issued at its end.
  for (int t = 0; t < ntiles; ++t) {
    const uint32_t ph = t & 1;               // phase of the once-per-tile barriers
    ft_wait_warp(bar_s, ph);
    tcgen05_fence_after();
    if (threadIdx.x == 0 && t + 1 < ntiles) {          // S(t) retired: the K buffer is free
      mbar_arrive_expect_tx(bar_k, FT_TILE);
      tma_load_2d(sK, &tm_qkv, bar_k, E3k, row0 + (t + 1) * FT_BN);
    }
    // ---- softmax: this thread's 64 scores of row r (2 chunks of 32), kept in registers
    const int valid = len - t * FT_BN - half * 64;     // own columns [0, valid) exist
    uint32_t sr[2][32];
    tmem_ld_32x32(tmem_row + half * 64, sr[0]);
    tmem_ld_32x32(tmem_row + half * 64 + 32, sr[1]);
    tmem_ld_wait();
    float tm = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float sv = (c * 32 + i < valid) ? __uint_as_float(sr[c][i]) : -INFINITY;
        sr[c][i] = __float_as_uint(sv);
        tm = fmaxf(tm, sv);
      }
    }
    xch[half * 128 + r] = tm;
    tcgen05_fence_before();
    __syncthreads();                         // every thread holds S(t): the S columns are free
    if (threadIdx.x == 0 && t + 1 < ntiles) {
      ft_wait(bar_k, ph ^ 1);              // K(t+1) landed
      tcgen05_fence_after();
      issue_s();
    }
    const float mn = fmaxf(m, fmaxf(tm, xch[(half ^ 1) * 128 + r]) * LOG2E);   // finite: key 0 of
    const float alpha = ex2_approx(m - mn);                                     // the tile is valid
    if (t > 0) {                             // collect PV(t-1) (o is still in the scale of m_{t-1})
      ft_wait_warp(bar_pv, ph ^ 1);
      tcgen05_fence_after();
      add_pv();
      if (threadIdx.x == 0 && t + 1 < ntiles) {        // PV(t-1) retired: its V buffer takes V(t+1)
        const int vs = (t + 1) & 1;
        mbar_arrive_expect_tx(&bar_v[vs], FT_TILE);
        tma_load_2d(sV + vs * FT_TILE, &tm_qkv, &bar_v[vs], E3v, row0 + (t + 1) * FT_BN);
      }
    }
#pragma unroll
    for (int i = 0; i < FT_D / 2; ++i) o[i] *= alpha;
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {       // 8 keys -> one 16-byte chunk of the swizzled P row
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          p[e] = ex2_approx(fmaf(__uint_as_float(sr[c][i8 * 8 + e]), LOG2E, -mn));   // masked: -inf -> 0
          rs += p[e];
        }
        uint4 u;
        u.x = pack_bf16(p[0], p[1]); u.y = pack_bf16(p[2], p[3]);
        u.z = pack_bf16(p[4], p[5]); u.w = pack_bf16(p[6], p[7]);
        const int cc = c * 4 + i8;           // 16-byte chunk within this half's 64-key block
        *reinterpret_cast<uint4*>(sP + half * FT_TILE + r * 128 + ((cc ^ (r & 7)) << 4)) = u;
      }
    }
    l = l * alpha + rs;
    m = mn;
    fence_proxy_async();                     // P (generic-proxy stores) -> visible to the MMA
    tcgen05_fence_before();
    __syncthreads();
    // ---- PV(t) = P . V(t), collected during the next tile (or after the loop)
    if (threadIdx.x == 0) {
      ft_wait(&bar_v[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      const uint64_t dv = dv0 + static_cast<uint64_t>(((t & 1) * FT_TILE) >> 4);
#pragma unroll
      for (int k = 0; k < FT_BN / 16; ++k) {
        const uint64_t da = dp + static_cast<uint64_t>(((k >> 2) * FT_TILE + (k & 3) * 32) >> 4);
        umma_bf16(tmem_base + FT_BN, da, dv + static_cast<uint64_t>((k * 2048) >> 4), idesc_pv, k != 0 ? 1u : 0u);
      }
      umma_commit(bar_pv);
    }
  }
  ft_wait_warp(bar_pv, (ntiles - 1) & 1);
  tcgen05_fence_after();
  add_pv();
  // ---- epilogue: the two threads of a row add their partial sums, each writes its 32 dims
  xch[half * 128 + r] = l;
  __syncthreads();
  l += xch[(half ^ 1) * 128 + r];
  if (q0 + r < len) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* op = out + static_cast<long long>(row0 + q0 + r) * E + h * FT_D + half * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(o[8 * i] * inv, o[8 * i + 1] * inv);
      u.y = pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
      u.z = pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
      u.w = pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
      reinterpret_cast<uint4*>(op)[i] = u;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, FT_TMEM_COLS);
  }
}

}  // namespace tt

using namespace tt;

extern "C" int tt_flash_self_attn_varlen_tc5(const void* qkv, const int* cu_seqlens, void* out, int B,
                                             int S_max, int H, int D, long long rows, void* stream) {
  TT_REQUIRE(qkv && out && cu_seqlens, "tt_flash_self_attn_varlen_tc5: null pointer");
  TT_REQUIRE(D == FT_D, "tt_flash_self_attn_varlen_tc5: head_dim must be %d (got %d)", FT_D, D);
  TT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "tt_flash_self_attn_varlen_tc5: qkv / out must be 16-byte aligned");
  if (B <= 0 || S_max <= 0 || rows <= 0) return TT_OK;
  const long long E = static_cast<long long>(H) * D;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, qkv, static_cast<uint64_t>(3 * E), static_cast<uint64_t>(rows),
                             static_cast<uint64_t>(3 * E), FT_D, FT_BM);
  if (rc != TT_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(flash_tc5, %d B): %s", FT_SMEM, cudaGetErrorString(e));
      return TT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(S_max, FT_BM), B * H);
  launch_k(flash_tc5_kernel, grid, dim3(256), FT_SMEM, (cudaStream_t)stream, tm,
           reinterpret_cast<__nv_bfloat16*>(out), cu_seqlens, H);
  return check_launch("flash_tc5_kernel");
}
