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
#include <cstdlib>
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

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100): one issue slot for two lanes of the softmax
// arithmetic.  Same rounding as the scalar instructions (.rn per element).
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1,
                                      float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
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
    for (int i = 0; i < 32; i += 2)
      fadd2(o[i], o[i + 1], o[i], o[i + 1], __uint_as_float(pv[i]), __uint_as_float(pv[i + 1]));
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
    if (valid < 64) {                          // warp-uniform: only the last key tile of a sample masks
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;     // -inf
      }
    }
    float tm = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        tm = fmaxf(tm, fmaxf(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])));
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
    for (int i = 0; i < FT_D / 2; i += 2) fmul2(o[i], o[i + 1], o[i], o[i + 1], alpha, alpha);
    float rs0 = 0.f, rs1 = 0.f;
    const float nmn = -mn;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {       // 8 keys -> one 16-byte chunk of the swizzled P row
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sr[c][i8 * 8 + e]), __uint_as_float(sr[c][i8 * 8 + e + 1]), LOG2E,
                LOG2E, nmn, nmn);
          p[e] = ex2_approx(x0);             // masked: -inf -> 0
          p[e + 1] = ex2_approx(x1);
          fadd2(rs0, rs1, rs0, rs1, p[e], p[e + 1]);
        }
        uint4 u;
        u.x = pack_bf16(p[0], p[1]); u.y = pack_bf16(p[2], p[3]);
        u.z = pack_bf16(p[4], p[5]); u.w = pack_bf16(p[6], p[7]);
        const int cc = c * 4 + i8;           // 16-byte chunk within this half's 64-key block
        *reinterpret_cast<uint4*>(sP + half * FT_TILE + r * 128 + ((cc ^ (r & 7)) << 4)) = u;
      }
    }
    l = l * alpha + (rs0 + rs1);
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

// ------------------------------------------------------------------------------------------------
// Warp-specialised, fully double-buffered form (the default).  Three warpgroups:
//   warps 0-7  softmax / output: quarter = warp & 3 owns TMEM lanes (query rows) 32*quarter .. +31,
//              half = warp >> 2 owns 32 of the 64 keys of a tile and 32 of the 64 output dims;
//              104 registers (setmaxnreg)
//   warp  8    control: ONE lane issues every TMA load and every tcgen05.mma; warps 9-11 only fill
//              its warpgroup (setmaxnreg works per warpgroup: 32 registers each)
// Key tiles are 64 keys wide and EVERYTHING a tile touches exists twice -- K and V stages in shared
// memory, the score tile S and the P.V product in TMEM (4 x 64 columns), the probability tile P in
// shared memory -- so no role ever waits for the latency of another one in steady state:
//   S(t+2) is issued as soon as the softmax warps hold S(t) in registers (its K stage was refilled
//          a whole tile earlier), i.e. the scores of the next two tiles are always ready;
//   PV(t)  is issued when P(t) is announced; the softmax warps fold PV(t-1) into their registers at the
//          END of tile t, a whole exponential phase after it was issued;
//   P(t)   goes into the buffer PV(t-2) read, and PV(t-2) was collected in tile t-1.
// ncu on the single-buffered kernel above showed the softmax warps asleep on the S-ready and
// PV-ready barriers (MMA + TMA latency in the per-tile dependency chain) with 31 % of the issue slots
// used; here the chain is the softmax arithmetic alone.
// The softmax warps synchronise pairwise only (the two warps that share 32 rows exchange their row
// maxima through a 64-thread named barrier, double-buffered slots) and talk to the control warp through
// mbarriers (8 warp arrivals): S(t) consumed, P(t) written.
constexpr int FTW_THREADS = 384;
constexpr int FTW_BN = 64;                            // keys per tile
constexpr int FTW_KV = FTW_BN * FT_D * 2;             // 8 KB: 64 keys x 128 B
constexpr int FTW_OFF_Q = 0;                          // 16 KB
constexpr int FTW_OFF_K = FT_TILE;                    // 2 x 8 KB
constexpr int FTW_OFF_V = FT_TILE + 2 * FTW_KV;       // 2 x 8 KB
constexpr int FTW_OFF_P = FT_TILE + 4 * FTW_KV;       // 2 x 16 KB (128 rows x 64 keys bf16)
constexpr int FTW_OFF_BAR = FTW_OFF_P + 2 * FT_TILE;  // 80 KB
constexpr int FTW_SMEM = FTW_OFF_BAR + 256 + 2048 + 1024;   // barriers, row exchange, alignment slack
constexpr int FTW_TMEM_COLS = 256;                    // S0 S1 PV0 PV1, 64 columns each

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
               : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// mbarrier wait: plain try_wait polling with a trap instead of a hang on a protocol bug.
__device__ __forceinline__ void ftw_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void ftw_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) ftw_wait(bar, parity);
  __syncwarp();
}

__global__ void __launch_bounds__(FTW_THREADS, 2)
flash_tc5_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    __nv_bfloat16* __restrict__ out, const int* __restrict__ cu, int H) {
  extern __shared__ uint8_t ft_raw[];
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int q0 = blockIdx.x * FT_BM;
  const int row0 = cu[b];
  const int len = cu[b + 1] - row0;
  if (q0 >= len) return;                              // CTA-uniform, before any barrier / TMEM use
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + FTW_OFF_Q;
  uint8_t* sK = smem + FTW_OFF_K;        // [2]
  uint8_t* sV = smem + FTW_OFF_V;        // [2]
  uint8_t* sP = smem + FTW_OFF_P;        // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FTW_OFF_BAR);
  uint64_t* bar_q = bars;                // TMA: Q landed
  uint64_t* bar_k = bars + 1;            // [2] TMA: K(t) landed in stage t & 1
  uint64_t* bar_v = bars + 3;            // [2] TMA: V(t) landed in stage t & 1
  uint64_t* bar_s = bars + 5;            // [2] MMA commit: S(t) in TMEM buffer t & 1
  uint64_t* bar_pv = bars + 7;           // [2] MMA commit: PV(t) in TMEM buffer t & 1
  uint64_t* bar_sfree = bars + 9;        // [2] 8 softmax warps: S(t) is in registers
  uint64_t* bar_p = bars + 11;           // [2] 8 softmax warps: P(t) is in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const uint32_t xch = smem_u32(smem + FTW_OFF_BAR + 256);   // [2 parities][2 halves][128 rows] floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = H * FT_D;
  const int ntiles = (len + FTW_BN - 1) / FTW_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_q, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_k[i], 1);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_pv[i], 1);
      mbar_init(&bar_sfree[i], 8);
      mbar_init(&bar_p[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, FTW_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    // ------------------------------------------------------------------ control warp (one lane)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");     // hand the registers to the softmax warps
    if (warp == 8 && lane == 0) {
      const int ck = E + h * FT_D, cv = 2 * E + h * FT_D;     // column of this head's K / V slice
      auto load_k = [&](int t) {
        mbar_arrive_expect_tx(&bar_k[t & 1], FTW_KV);
        tma_load_2d(sK + (t & 1) * FTW_KV, &tm_kv, &bar_k[t & 1], ck, row0 + t * FTW_BN);
      };
      auto load_v = [&](int t) {
        mbar_arrive_expect_tx(&bar_v[t & 1], FTW_KV);
        tma_load_2d(sV + (t & 1) * FTW_KV, &tm_kv, &bar_v[t & 1], cv, row0 + t * FTW_BN);
      };
      mbar_arrive_expect_tx(bar_q, FT_TILE);
      tma_load_2d(sQ, &tm_q, bar_q, h * FT_D, row0 + q0);
      load_k(0);
      if (ntiles > 1) load_k(1);
      load_v(0);
      if (ntiles > 1) load_v(1);
      constexpr uint32_t idesc = umma_idesc_bf16(FT_BM, FTW_BN, false, false);      // S: 128 x 64 keys
      constexpr uint32_t idesc_pv = umma_idesc_bf16(FT_BM, FT_D, false, true);     // PV: 128 x 64 dims
      const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK));
      const uint64_t dp = umma_desc_kmajor_sw128(smem_u32(sP));
      const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(sV));
      auto issue_s = [&](int t) {              // S(t) = Q . K(t)^T -> TMEM columns [64 (t&1), +64)
        const uint64_t dkt = dk + static_cast<uint64_t>(((t & 1) * FTW_KV) >> 4);
#pragma unroll
        for (int k = 0; k < FT_D / 16; ++k)
          umma_bf16(tmem_base + (t & 1) * FTW_BN, dq + k * 2, dkt + k * 2, idesc, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[t & 1]);
      };
      ftw_wait(bar_q, 0);
      ftw_wait(&bar_k[0], 0);
      tcgen05_fence_after();
      issue_s(0);
      if (ntiles > 1) {
        ftw_wait(&bar_k[1], 0);
        tcgen05_fence_after();
        issue_s(1);
      }
      if (ntiles > 2) {                        // S(0) retired: its K stage takes K(2)
        ftw_wait(&bar_s[0], 0);
        load_k(2);
      }
      for (int t = 0; t < ntiles; ++t) {
        const int bf = t & 1;
        const uint32_t par = (t >> 1) & 1;
        if (t >= 1 && t + 1 < ntiles) {        // PV(t-1) retired: its V stage takes V(t+1)
          ftw_wait(&bar_pv[bf ^ 1], ((t - 1) >> 1) & 1);
          load_v(t + 1);
        }
        if (t + 3 < ntiles) {                  // S(t+1) retired: its K stage takes K(t+3)
          ftw_wait(&bar_s[bf ^ 1], ((t + 1) >> 1) & 1);
          load_k(t + 3);
        }
        if (t + 2 < ntiles) {                  // S(t) sits in registers: its columns take S(t+2)
          ftw_wait(&bar_sfree[bf], par);
          ftw_wait(&bar_k[bf], par ^ 1);       // K(t+2): requested a whole tile ago
          tcgen05_fence_after();
          issue_s(t + 2);
        }
        ftw_wait(&bar_p[bf], par);             // P(t) in shared memory (and PV(t-2) folded by everyone)
        ftw_wait(&bar_v[bf], par);
        tcgen05_fence_after();
        const uint64_t dpt = dp + static_cast<uint64_t>((bf * FT_TILE) >> 4);
        const uint64_t dvt = dv + static_cast<uint64_t>((bf * FTW_KV) >> 4);
#pragma unroll
        for (int k = 0; k < FTW_BN / 16; ++k)
          umma_bf16(tmem_base + 2 * FTW_BN + bf * FT_D, dpt + k * 2, dvt + static_cast<uint64_t>((k * 2048) >> 4),
                    idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(&bar_pv[bf]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int r = quarter * 32 + lane;         // query row of this thread within the tile
    const uint32_t sP_row = smem_u32(sP) + r * 128;
    constexpr float LOG2E = 1.4426950408889634f;
    float o[FT_D / 2];
#pragma unroll
    for (int i = 0; i < FT_D / 2; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int bf = t & 1;
      const uint32_t par = (t >> 1) & 1;
      ftw_wait_warp(&bar_s[bf], par);
      tcgen05_fence_after();
      uint32_t sr[32];                          // this thread's 32 scores of row r
      tmem_ld_32x32(tmem_row + bf * FTW_BN + half * 32, sr);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_sfree[bf]);        // this warp's share of S(t) is in registers
      const int valid = len - t * FTW_BN - half * 32;    // own columns [0, valid) exist
      if (valid < 32) {                                  // warp-uniform: only a sample's last key tile
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) sr[i] = 0xff800000u;           // -inf
      }
      float tm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        tm = fmaxf(tm, fmaxf(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])));
      const uint32_t xslot = xch + bf * 1024;
      st_shared_f32(xslot + (half * 128 + r) * 4, tm);
      named_bar_sync(1 + quarter, 64);                   // the two warps that share these 32 rows
      const float other = ld_shared_f32(xslot + ((half ^ 1) * 128 + r) * 4);
      const float mn = fmaxf(m, fmaxf(tm, other) * LOG2E);       // finite: key 0 of the tile is valid
      const float alpha = ex2_approx(m - mn);
      const float nmn = -mn;
      float rs0 = 0.f, rs1 = 0.f;
      const uint32_t prow = sP_row + bf * FT_TILE;
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {         // 8 keys -> one 16-byte chunk of the swizzled P row
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sr[i8 * 8 + e]), __uint_as_float(sr[i8 * 8 + e + 1]), LOG2E, LOG2E,
                nmn, nmn);
          p[e] = ex2_approx(x0);               // masked: -inf -> 0
          p[e + 1] = ex2_approx(x1);
          fadd2(rs0, rs1, rs0, rs1, p[e], p[e + 1]);
        }
        uint4 u;
        u.x = pack_bf16(p[0], p[1]); u.y = pack_bf16(p[2], p[3]);
        u.z = pack_bf16(p[4], p[5]); u.w = pack_bf16(p[6], p[7]);
        const int cc = half * 4 + i8;          // 16-byte chunk within the 64-key row
        st_shared_v4(prow + ((cc ^ (r & 7)) << 4), u);
      }
      l = l * alpha + (rs0 + rs1);
      m = mn;
      fence_proxy_async();                     // P (generic-proxy stores) -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[bf]);  // PV(t) may be issued
      // ---- fold PV(t-1) (issued one exponential phase ago) while o is in the scale of m_{t-1}, rescale
      if (t > 0) {
        ftw_wait_warp(&bar_pv[bf ^ 1], ((t - 1) >> 1) & 1);
        tcgen05_fence_after();
        uint32_t pv[32];
        tmem_ld_32x32(tmem_row + 2 * FTW_BN + (bf ^ 1) * FT_D + half * 32, pv);
        tmem_ld_wait();
        tcgen05_fence_before();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          fadd2(o[i], o[i + 1], o[i], o[i + 1], __uint_as_float(pv[i]), __uint_as_float(pv[i + 1]));
          fmul2(o[i], o[i + 1], o[i], o[i + 1], alpha, alpha);
        }
      }
    }
    {
      const int tl = ntiles - 1;
      ftw_wait_warp(&bar_pv[tl & 1], (tl >> 1) & 1);
      tcgen05_fence_after();
      uint32_t pv[32];
      tmem_ld_32x32(tmem_row + 2 * FTW_BN + (tl & 1) * FT_D + half * 32, pv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        fadd2(o[i], o[i + 1], o[i], o[i + 1], __uint_as_float(pv[i]), __uint_as_float(pv[i + 1]));
    }
    // ---- epilogue: the two threads of a row add their partial sums, each writes its 32 dims
    const uint32_t xslot = xch + (ntiles & 1) * 1024;
    st_shared_f32(xslot + (half * 128 + r) * 4, l);
    named_bar_sync(1 + quarter, 64);
    l += ld_shared_f32(xslot + ((half ^ 1) * 128 + r) * 4);
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
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, FTW_TMEM_COLS);
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
  static bool ws = true;        // warp-specialised kernel (default); TT_FLASH_WS=0: the single-role kernel
  if (!attr_set) {
    const char* env = getenv("TT_FLASH_WS");
    ws = !(env && env[0] == '0');
    cudaError_t e = cudaFuncSetAttribute(flash_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(flash_tc5_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FTW_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(flash_tc5, %d B): %s", FTW_SMEM, cudaGetErrorString(e));
      return TT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(S_max, FT_BM), B * H);
  if (ws) {
    CUtensorMap tm_kv;          // K / V boxes: 64 dims x 64 keys
    rc = make_tmap_bf16_2d(&tm_kv, qkv, static_cast<uint64_t>(3 * E), static_cast<uint64_t>(rows),
                           static_cast<uint64_t>(3 * E), FT_D, FTW_BN);
    if (rc != TT_OK) return rc;
    launch_k(flash_tc5_ws_kernel, grid, dim3(FTW_THREADS), FTW_SMEM, (cudaStream_t)stream, tm, tm_kv,
             reinterpret_cast<__nv_bfloat16*>(out), cu_seqlens, H);
    return check_launch("flash_tc5_ws_kernel");
  }
  launch_k(flash_tc5_kernel, grid, dim3(256), FT_SMEM, (cudaStream_t)stream, tm,
           reinterpret_cast<__nv_bfloat16*>(out), cu_seqlens, H);
  return check_launch("flash_tc5_kernel");
}
