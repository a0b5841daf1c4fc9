// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = act(alpha * (A[M,K] . B[N,K]^T + bias) + residual)
//
// Persistent, warp-specialised (one CTA per SM):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128x64 (A) and BNx64 (B) bf16 tiles, 128B swizzle
//   warp 1      MMA issuer     : one thread issues tcgen05.mma 128xBNx16 (kind::f16, fp32 accum in TMEM)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       : tcgen05.ld accumulator -> registers -> alpha/bias/act/residual -> HBM
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, the
// accumulator is double buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "gemm_common.cuh"
#include "runtime.h"

namespace tt {

constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps

template <int BN, bool STAGED = false>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // double-buffered accumulator
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES + 512;   // staged-epilogue tiles (512 B aligned)
  // + barriers + align slack (+ the staged epilogue's tiles)
  static constexpr int SMEM_BYTES = STAGED ? EPI_OFF + EPI_SMEM_BYTES + 1024 : STAGES * STAGE_BYTES + 256 + 1024;
};

// STAGED: the instantiation that carries the staged (TMA store) epilogue of gemm_common.cuh; the
// plain one is compiled without it -- with both epilogues in one kernel the register allocation of
// the row-per-thread path degraded (+1.4 us per launch, +3 us per tile on the fp32-output decoder
// GEMMs, measured against the previous build on the same box: profiles/r2_gemm_small_ab.txt).
// MODE 0: plain (register epilogue), 1: staged epilogue, 2: staged epilogue + column statistics.
template <int BN, bool TA, bool TB, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const GemmArgs g) {
  constexpr bool STAGED = MODE != 0;
  constexpr bool STATS = MODE == 2;
  using Cfg = GemmCfg<BN, STAGED>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  [[maybe_unused]] uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // one per epilogue warp
  [[maybe_unused]] uint8_t* epi = smem + Cfg::EPI_OFF;

  // shuffled so the compiler knows the role index is warp-uniform (uniform-datapath code)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // the next kernel may start its own prologue
  if (threadIdx.x == 0) trace_stamp(g, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (STAGED) {
      tma_prefetch_desc(&tmC);
      if (g.tma_epi == 2) tma_prefetch_desc(&tmR);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    if constexpr (STAGED) {
      for (int s = 0; s < 8; ++s) mbar_init(&res_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], GEMM_THREADS - 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(g, 1);
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of
  // the previous kernel; from here on we touch its outputs.
  pdl_wait();
  int M = g.M;
  if (g.m_limit != nullptr) M = min(M, __ldg(g.m_limit));
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (g.N + BN - 1) / BN;
  const int num_mn = num_m * num_n;
  int Keff = g.K;
  if (g.k_limit != nullptr) Keff = min(Keff, max(__ldg(g.k_limit), 1));   // >= 1 block: zeros give zeros
  const int num_k = (Keff + BK - 1) / BK;
  // split-K: tile index = (k split, n block, m block); split ks owns k-blocks [ks*kper, +kper).
  // Few-tile / long-K problems (the adaptive-softmax tail back-projections: one m block x 16 n
  // blocks x 473 k-blocks on 16 SMs) spread over the machine; partials meet in C through atomics.
  const int splits = g.splits > 1 ? g.splits : 1;
  const int kper = (num_k + splits - 1) / splits;
  const int nbatch = g.nbatch > 1 ? g.nbatch : 1;      // batched launches never split K
  const int num_tiles = num_mn * splits * nbatch;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the (warp-uniform) loop so that addresses / coordinates live in uniform
    // registers; one elected lane issues.  A divergent `if (lane == 0)` region costs ~100 dependent
    // SASS instructions (R2UR waterfall loops) per k-block = 0.2-0.3 us, which for narrow tiles is
    // longer than the MMAs themselves.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), full_u = smem_u32(full_bar);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int kz = tile / num_mn, mn = tile - kz * num_mn;
      const int ks = nbatch > 1 ? 0 : kz, z = nbatch > 1 ? kz : 0;
      const int m_blk = mn % num_m, n_blk = mn / num_m;
      const int kb_end = min(num_k, (ks + 1) * kper);
      const int a0 = z * g.a_off0, a1 = z * g.a_off1, b0 = z * g.b_off0, b1 = z * g.b_off1;
      for (int kb = ks * kper; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const uint32_t bar = full_u + stage * 8;
          mbar_arrive_expect_tx_u(bar, Cfg::STAGE_BYTES);
          if (!TA) {
            tma_load_2d_u(sA_u + stage * Cfg::A_BYTES, &tmA, bar, kb * BK + a0, m_blk * BM + a1);
          } else {  // MN-major: one 64(m) x 64(k) box per 64-row chunk of the tile
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d_u(sA_u + stage * Cfg::A_BYTES + c * 8192, &tmA, bar, m_blk * BM + c * 64 + a0, kb * BK + a1);
          }
          if (!TB) {
            tma_load_2d_u(sB_u + stage * Cfg::B_BYTES, &tmB, bar, kb * BK + b0, n_blk * BN + b1);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d_u(sB_u + stage * Cfg::B_BYTES + c * 8192, &tmB, bar, n_blk * BN + c * 64 + b0, kb * BK + b1);
          }
          if (kb == ks * kper && tile == blockIdx.x) trace_stamp(g, 2);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, TA, TB);
    // descriptor templates: the start-address field (bits 0-13, 16-byte units) is advanced by adds
    const uint64_t da0 = TA ? umma_desc_mnmajor_sw128(smem_u32(sA)) : umma_desc_kmajor_sw128(smem_u32(sA));
    const uint64_t db0 = TB ? umma_desc_mnmajor_sw128(smem_u32(sB)) : umma_desc_kmajor_sw128(smem_u32(sB));
    // K-major: 16 k = 32 B inside the 128 B swizzle row; MN-major: 16 k-rows = 2048 B
    constexpr uint32_t KA = (TA ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    constexpr uint32_t KB = (TB ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    const uint32_t empty_u = smem_u32(empty_bar), tfull_u = smem_u32(tmem_full);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = nbatch > 1 ? 0 : tile / num_mn;
      const int kb0 = ks * kper, kb_end = min(num_k, (ks + 1) * kper);
      if (kb0 >= kb_end) continue;            // empty split (num_k not a multiple): no accumulator
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = kb0; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = da0 + static_cast<uint64_t>((stage * Cfg::A_BYTES) >> 4);
        const uint64_t db = db0 + static_cast<uint64_t>((stage * Cfg::B_BYTES) >> 4);
        if (elect_one()) {
          if (kb == kb0 && tile == blockIdx.x) trace_stamp(g, 3);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d_tmem, da + k * KA, db + k * KB, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit_u(empty_u + stage * 8);  // frees the smem slot when these MMAs retire
          if (kb == kb_end - 1) {
            umma_commit_u(tfull_u + acc * 8);
            if (tile == blockIdx.x) trace_stamp(g, 4);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // A warp may only touch TMEM lanes [32*(warp%4), +32): two warps share each lane quarter and
    // split the tile's 32-column chunks between them (even / odd).  Every thread owns one row of
    // the chunk; bias and residual are fetched with 16-byte loads issued BEFORE waiting on the
    // TMEM load so their latency overlaps it.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    [[maybe_unused]] EpiWarp ew;
    if constexpr (STAGED) {
      ew.st_out = smem_u32(epi) + static_cast<uint32_t>(warp - 4) * EPI_WARP_BYTES;
      ew.st_res = ew.st_out + 8 * EPI_WARP_BYTES;
      ew.res_bar = smem_u32(&res_bar[warp - 4]);
      ew.res_phase = 0;
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int kz = tile / num_mn, mn = tile - kz * num_mn;
      const int ks = nbatch > 1 ? 0 : kz, z = nbatch > 1 ? kz : 0;
      const int m_blk = mn % num_m, n_blk = mn / num_m;
      if (ks * kper >= min(num_k, (ks + 1) * kper)) continue;     // empty split
      const int row0 = m_blk * BM + q * 32;
      // staged (TMA) epilogue for warps whose 32 rows are all valid (never with split-K: the host
      // only picks the STAGED instantiation for plain bf16-output problems)
      bool staged = false;
      if constexpr (STAGED) {
        staged = STATS ? row0 < M : row0 + 32 <= M;
        if (staged && g.tma_epi == 2 && lane == 0 && n_blk * BN + half * 32 < g.N)
          epi_request_residual(&tmR, ew, n_blk * BN + half * 32, row0);   // lands under this tile's MMAs
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      if (threadIdx.x == 128 && tile == blockIdx.x) trace_stamp(g, 5);
      const long long row = row0 + lane;
      const bool row_ok = row < M;
      if constexpr (STAGED) {
        if (staged)
          epilogue_chunks_tma<BN, STATS>(g, &tmC, &tmR, tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                                            static_cast<uint32_t>(acc * BN),
                                         half, row0, n_blk * BN, ew, min(32, M - row0));
      }
      if (staged) {
      } else if (splits > 1 || nbatch > 1) {
        GemmArgs ge = g;
        if (splits > 1) {                     // partial product: bias / residual enter once (split 0)
          ge.atomic = 1;
          if (ks > 0) { ge.bias = nullptr; ge.residual = nullptr; ge.residual16 = nullptr; }
        } else {                              // batch z: its block of the shared output / bias buffers
          if (ge.C != nullptr) ge.C += z * g.c_off;
          if (ge.C16 != nullptr) ge.C16 += z * g.c_off;
          if (ge.bias != nullptr) ge.bias += z * g.bias_off;
        }
        epilogue_chunks<BN>(ge, tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                    static_cast<uint32_t>(acc * BN),
                            half, row, row_ok, n_blk * BN);
      } else {
        epilogue_chunks<BN>(g, tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                          static_cast<uint32_t>(acc * BN),
                                   half, row, row_ok, n_blk * BN);
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (threadIdx.x == 128 && tile == blockIdx.x) trace_stamp(g, 6);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if constexpr (STAGED) {
      if (lane == 0) bulk_wait_group0();   // staged stores performed before exit
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    if (lane == 0) trace_stamp(g, 7);
  }
}

template <int BN, bool TA, bool TB, int MODE = 0>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                         const CUtensorMap& tmR, const GemmArgs& g, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, MODE != 0>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, TA, TB, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm BN=%d, %d B): %s", BN, Cfg::SMEM_BYTES,
                cudaGetErrorString(e));
      return TT_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_k(gemm_bf16_tn_kernel<BN, TA, TB, MODE>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA,
           tmB, tmC, tmR, g);
  return check_launch("gemm_bf16_tn_kernel");
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       const CUtensorMap& tmR, const GemmArgs& g, int grid, bool ta, bool tb,
                       cudaStream_t stream) {
  if (ta) {
    if (tb) {
      if constexpr (BN >= 64) return launch_gemm_t<BN, true, true>(tmA, tmB, tmC, tmR, g, grid, stream);
    } else {
      return launch_gemm_t<BN, true, false>(tmA, tmB, tmC, tmR, g, grid, stream);
    }
  } else {
    if (tb) {
      if constexpr (BN >= 64) return launch_gemm_t<BN, false, true>(tmA, tmB, tmC, tmR, g, grid, stream);
    } else {
      if (g.tma_epi != 0 && g.col_stats != nullptr)
        return launch_gemm_t<BN, false, false, 2>(tmA, tmB, tmC, tmR, g, grid, stream);
      if (g.tma_epi != 0) return launch_gemm_t<BN, false, false, 1>(tmA, tmB, tmC, tmR, g, grid, stream);
      return launch_gemm_t<BN, false, false>(tmA, tmB, tmC, tmR, g, grid, stream);
    }
  }
  set_error("tt_gemm_bf16_tn: transposed B needs a tile width >= 64");
  return TT_ERR_INVALID;
}

// Tile configuration: a small cost model fitted to measurements (tools/gemm_tile_sweep.py,
// profiles/r2_gemm_tile_sweep.txt: 30 signatures of the step x 6 configurations on a B200).
//   cost = waves * (fixed + k-blocks * t_k(config) + columns * t_epilogue)        [microseconds]
// waves = tiles / CTAs (148) or tiles / CTA pairs (74); t_k is what one k-block costs a CTA: the
// larger of its share of the L2 -> shared-memory traffic and its MMAs (the CTA-pair kernel halves
// the B traffic per SM, the 256-wide single-CTA tile is shared-memory-port bound).  It replaces
// "widest tile that fills 87 % of the SMs", which left 5-45 % on the table for the short-K / many-
// row convolution GEMMs (12544 x 128 x 1152: 14.3 -> 9.6 us) and the N = 2048 decoder GEMMs.
// Occupancy weight w of the tile model's objective: cost * (fraction of the SMs a launch holds)^w.
// w = 0 minimises the launch's own duration (a latency-bound chain running alone: greedy decoding);
// w = 1 minimises SMs x time -- what counts when several streams share the machine: the overlapped
// train step is bound by SM occupancy (sum over all kernels of SMs held x duration / 148 = 10.3 ms of
// the 10.6 ms step), and an M = 800 GEMM on 112 CTAs of 64 columns holds 870 SM.us where 56 CTAs of
// 128 columns hold 529 for 22 % more latency.
static double g_occ_weight = -1.0;
struct TileChoice {
  bool pair;
  int bn;
};
static TileChoice choose_tile(int M, int N, int K, bool pair_ok, bool tb, int sms, int nbatch = 1) {
  static int forced = -1, forced2 = -1, pair_enabled = -1;
  if (forced < 0) {
    const char* e = getenv("TT_GEMM_BN");          // experiments: force the single-CTA tile width
    forced = e ? atoi(e) : 0;
    e = getenv("TT_GEMM2_BN");                     // ... or the CTA-pair kernel with this width
    forced2 = e ? atoi(e) : 0;
    e = getenv("TT_GEMM_2CTA");
    pair_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (pair_ok && pair_enabled && (forced2 == 128 || forced2 == 256) && N >= forced2) return {true, forced2};
  if (forced == 32 || forced == 64 || forced == 128 || forced == 256) return {false, (tb && forced < 64) ? 64 : forced};
  if (N <= 32 && !tb) return {false, 32};
  if (pair_ok && pair_enabled && N >= 256) {
    // At least two full waves of 256 x 256 pair tiles: the CTA-pair kernel at its widest tile is
    // the most efficient configuration there is (half the B traffic per SM, 1.1-1.28 PFLOP/s), and
    // unlike the single-CTA kernels it keeps that rate when the other encoder runs beside it.
    // 128-wide pair tiles only when they save whole idle waves (15 % slower per FLOP, measured).
    const long long t256 = static_cast<long long>(ceil_div(M, 2 * BM)) * ceil_div(N, 256);
    if (t256 >= 2 * (sms / 2)) {
      const long long t128 = static_cast<long long>(ceil_div(M, 2 * BM)) * ceil_div(N, 128);
      const double c256 = static_cast<double>(ceil_div_ll(t256, sms / 2)) * 256.0;
      const double c128 = static_cast<double>(ceil_div_ll(t128, sms / 2)) * 128.0 * 1.15;
      return {true, c128 < c256 ? 128 : 256};
    }
  }
  if (g_occ_weight < 0.0) {
    const char* e = getenv("TT_GEMM_OCC_WEIGHT");
    g_occ_weight = e ? atof(e) : 0.0;
  }
  const double occ_weight = g_occ_weight;
  const int num_k = ceil_div(K, BK);
  struct Cand { bool pair; int bn; double fixed, tk; };
  const Cand cands[5] = {{false, 64, 0.6, 0.17}, {false, 128, 0.6, 0.22}, {false, 256, 0.6, 0.55},
                         {true, 128, 1.0, 0.2125}, {true, 256, 1.0, 0.418}};
  TileChoice best = {false, 64};
  double best_cost = 1e30;
  for (const Cand& c : cands) {
    if (c.pair && !(pair_ok && pair_enabled)) continue;
    if (N < c.bn && c.bn > 64) continue;
    const long long tiles = static_cast<long long>(ceil_div(M, c.pair ? 2 * BM : BM)) * ceil_div(N, c.bn) * nbatch;
    const long long slots = c.pair ? sms / 2 : sms;
    const double waves = static_cast<double>(ceil_div_ll(tiles, slots));
    double cost = waves * (c.fixed + num_k * c.tk + 0.02 * (N < c.bn ? N : c.bn));
    if (!c.pair && c.bn >= 128 && pair_ok && pair_enabled) cost *= 1.05;   // ties go to the pair kernel
    if (occ_weight > 0.0) {
      // occupancy-aware objective: SMs held x time (w = 1) instead of time alone (w = 0)
      const double used = static_cast<double>(tiles < slots ? tiles : slots) * (c.pair ? 2 : 1) / sms;
      cost *= pow(used, occ_weight);
    }
    if (cost < best_cost) {
      best_cost = cost;
      best = {c.pair, c.bn};
    }
  }
  return best;
}

int gemm2_launch(const TtGemmParams* p, const GemmArgs& g, const CUtensorMap& tmC, const CUtensorMap& tmR, int bn,
                 cudaStream_t stream);  // gemm2.cu

// SM cap of the persistent GEMM grids (0 = all SMs).  A persistent GEMM on every SM holds the whole
// machine for its 30-50 us, so kernels of concurrent streams (the decoder's and the ResNet's short
// latency-bound chains beside RoBERTa's large GEMMs) queue behind it; launched on fewer SMs the large
// GEMM runs proportionally longer but the other streams keep flowing -- measured: RoBERTa's GEMMs on
// 64 of 148 SMs make the overlapped train step 4-5 % FASTER.  The caller sets it
// around the launches (or graph capture) of the stream that should yield.
static int g_sm_cap = -1;
int gemm_sm_cap() {
  if (g_sm_cap < 0) {
    const char* e = getenv("TT_GEMM_SM_CAP");
    g_sm_cap = e ? atoi(e) : 0;
  }
  return g_sm_cap;
}
static int g_staged = -1;    // staged (TMA) epilogue switch: -1 = read TT_GEMM_TMA_EPI on first use
static long long* g_trace = nullptr;
long long* gemm_trace_ptr() { return g_trace; }

}  // namespace tt

extern "C" void tt_gemm_set_trace(long long* dev_ptr) { tt::g_trace = dev_ptr; }
extern "C" void tt_gemm_set_staged_epilogue(int on) { tt::g_staged = on ? 1 : 0; }
extern "C" void tt_gemm_set_sm_cap(int sms) { tt::g_sm_cap = sms > 0 ? sms : 0; }
extern "C" void tt_gemm_set_occupancy_weight(float w) { tt::g_occ_weight = w > 0.f ? w : 0.0; }

extern "C" int tt_gemm_bf16_tn(const TtGemmParams* p, void* stream) {
  using namespace tt;
  TT_REQUIRE(p != nullptr, "tt_gemm_bf16_tn: null params");
  TT_REQUIRE(p->M >= 0 && p->N > 0 && p->K > 0, "tt_gemm_bf16_tn: bad shape M=%d N=%d K=%d", p->M,
             p->N, p->K);
  if (p->M == 0) return TT_OK;
  TT_REQUIRE(p->A && p->B, "tt_gemm_bf16_tn: null operand");
  TT_REQUIRE(p->C || p->C16, "tt_gemm_bf16_tn: no output buffer");
  const bool ta = p->trans_a != 0, tb = p->trans_b != 0;
  TT_REQUIRE(p->nbatch <= 1 || (p->nbatch <= 16 && p->a_off0 >= 0 && p->a_off1 >= 0 && p->b_off0 >= 0 &&
                                p->b_off1 >= 0 && p->c_off >= 0 && p->bias_off >= 0),
             "tt_gemm_bf16_tn: bad batch description");
  TT_REQUIRE(p->lda % 8 == 0 && p->ldb % 8 == 0 && p->lda >= (ta ? p->M : p->K) &&
                 p->ldb >= (tb ? p->N : p->K),
             "tt_gemm_bf16_tn: lda/ldb must be multiples of 8 and cover a stored row "
             "(lda=%lld ldb=%lld M=%d N=%d K=%d ta=%d tb=%d)",
             p->lda, p->ldb, p->M, p->N, p->K, (int)ta, (int)tb);
  TT_REQUIRE((reinterpret_cast<uintptr_t>(p->A) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(p->B) & 15) == 0,
             "tt_gemm_bf16_tn: operands must be 16-byte aligned");
  TT_REQUIRE(!(p->accumulate && !p->C), "tt_gemm_bf16_tn: accumulate needs the fp32 output");

  const int sms = num_sms();
  // rows the kernel will really compute (a host-side hint for device-limited problems)
  const int m_eff = (p->m_limit != nullptr && p->m_hint > 0 && p->m_hint < p->M) ? p->m_hint : p->M;
  const int nb = p->nbatch > 1 ? p->nbatch : 1;
  const TileChoice choice = choose_tile(m_eff, p->N, p->K, !ta && !tb && nb == 1, tb, sms, nb);
  int bn = choice.bn;           // (an MN-major B tile is built from 64-column TMA boxes: bn >= 64)

  CUtensorMap tmA, tmB;
  // K-major operand: rows = M (or N), inner = K, box 64(k) x rows.  MN-major operand (stored
  // [K, M] or [K, N]): rows = K, inner = M (or N), box 64(mn) x 64(k).
  // (a batched launch addresses batch z at coordinate + z * off: the maps span all batches)
  const uint64_t ax0 = static_cast<uint64_t>(nb - 1) * p->a_off0, ax1 = static_cast<uint64_t>(nb - 1) * p->a_off1;
  const uint64_t bx0 = static_cast<uint64_t>(nb - 1) * p->b_off0, bx1 = static_cast<uint64_t>(nb - 1) * p->b_off1;
  int rc = ta ? make_tmap_bf16_2d(&tmA, p->A, (uint64_t)p->M + ax0, (uint64_t)p->K + ax1, (uint64_t)p->lda, 64, BK)
              : make_tmap_bf16_2d(&tmA, p->A, (uint64_t)p->K + ax0, (uint64_t)p->M + ax1, (uint64_t)p->lda, BK, BM);
  if (rc != TT_OK) return rc;
  rc = tb ? make_tmap_bf16_2d(&tmB, p->B, (uint64_t)p->N + bx0, (uint64_t)p->K + bx1, (uint64_t)p->ldb, 64, BK)
          : make_tmap_bf16_2d(&tmB, p->B, (uint64_t)p->K + bx0, (uint64_t)p->N + bx1, (uint64_t)p->ldb, BK, bn);
  if (rc != TT_OK) return rc;

  GemmArgs g;
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.C = p->C; g.ldc = p->ldc;
  g.C16 = reinterpret_cast<__nv_bfloat16*>(p->C16); g.ldc16 = p->ldc16;
  g.bias = p->bias;
  g.residual = p->residual; g.ldr = p->ldr;
  g.residual16 = reinterpret_cast<const __nv_bfloat16*>(p->residual16); g.ldr16 = p->ldr16;
  g.alpha = p->alpha; g.act = p->act; g.accumulate = p->accumulate;
  g.m_limit = p->m_limit;
  g.k_limit = p->k_limit;
  g.col_stats = p->col_stats;
  TT_REQUIRE(p->col_stats == nullptr || (p->C16 != nullptr && p->C == nullptr && !p->accumulate &&
                                         p->m_limit == nullptr && p->bias == nullptr && p->residual16 == nullptr),
             "tt_gemm_bf16_tn: col_stats needs a plain bf16-only output (no bias / residual / row limit)");
  bool vec = true;
  if (p->C) vec = vec && (reinterpret_cast<uintptr_t>(p->C) & 15) == 0 && (p->ldc % 4 == 0);
  if (p->C16) vec = vec && (reinterpret_cast<uintptr_t>(p->C16) & 15) == 0 && (p->ldc16 % 8 == 0);
  if (p->residual)
    vec = vec && (reinterpret_cast<uintptr_t>(p->residual) & 15) == 0 && (p->ldr % 4 == 0);
  if (p->residual16)
    vec = vec && (reinterpret_cast<uintptr_t>(p->residual16) & 15) == 0 && (p->ldr16 % 8 == 0);
  if (p->bias) vec = vec && (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0;
  g.vec_ok = vec ? 1 : 0;
  g.trace = gemm_trace_ptr();
  g.splits = 1;
  g.atomic = 0;
  g.nbatch = p->nbatch > 1 ? p->nbatch : 1;
  g.a_off0 = p->a_off0; g.a_off1 = p->a_off1; g.b_off0 = p->b_off0; g.b_off1 = p->b_off1;
  g.bias_off = p->bias_off; g.c_off = p->c_off;
  if (g.nbatch > 1) {
    TT_REQUIRE(p->residual == nullptr && p->residual16 == nullptr && p->m_limit == nullptr &&
                   p->k_limit == nullptr && !p->accumulate && p->col_stats == nullptr,
               "tt_gemm_bf16_tn: a batched launch takes plain problems (bias / activation only)");
    TT_REQUIRE((ta || p->a_off0 == 0 || p->K % BK == 0) && (tb || p->b_off0 == 0 || p->K % BK == 0),
               "tt_gemm_bf16_tn: batches laid side by side along K need K %% 64 == 0");
    if (p->bias) vec = vec && (p->bias_off % 4 == 0);
    if (p->C) vec = vec && (p->c_off % 4 == 0);
    if (p->C16) vec = vec && (p->c_off % 8 == 0);
    g.vec_ok = vec ? 1 : 0;
  }

  // Staged epilogue (gemm_common.cuh): bf16-only output whose rows the TMA engine can address.
  CUtensorMap tmC = tmA, tmR = tmA;
  g.tma_epi = 0;
  {
    if (g_staged < 0) {
      const char* e = getenv("TT_GEMM_TMA_EPI");        // experiments: 0 = register epilogue everywhere
      g_staged = (e && e[0] == '0') ? 0 : 1;
    }
    if (g_staged && nb == 1 && !ta && !tb && p->C16 != nullptr && p->C == nullptr && p->residual == nullptr &&
        !p->accumulate && vec && p->N % 32 == 0 && p->M >= 32) {
      rc = make_tmap_bf16_2d_sw(&tmC, p->C16, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->ldc16, 32, 32, 64);
      if (rc != TT_OK) return rc;
      g.tma_epi = 1;
      if (p->residual16 != nullptr) {
        rc = make_tmap_bf16_2d_sw(&tmR, p->residual16, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->ldr16, 32, 32,
                                  64);
        if (rc != TT_OK) return rc;
        g.tma_epi = 2;
      }
    }
  }

  TT_REQUIRE(p->col_stats == nullptr || g.tma_epi != 0,
             "tt_gemm_bf16_tn: col_stats needs a staged-epilogue problem (K-major operands, N %% 32 == 0, "
             "16-byte aligned bf16 output rows, M >= 32)");

  if (choice.pair)   // CTA-pair kernel (gemm2.cu)
    return gemm2_launch(p, g, tmC, tmR, choice.bn, reinterpret_cast<cudaStream_t>(stream));

  int tiles = ceil_div(p->M, BM) * ceil_div(p->N, bn) * nb;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {
    // split-K for few-tile, long-K problems with a plain fp32 epilogue
    static int sk = -1;
    if (sk < 0) {
      const char* e = getenv("TT_GEMM_SPLITK");
      sk = (e && e[0] == '0') ? 0 : 1;
    }
    const int num_k = ceil_div(p->K, BK);
    if (p->m_limit != nullptr && p->m_hint > 0 && p->m_hint < p->M)     // expected rows (a hint)
      tiles = ceil_div(p->m_hint, BM) * ceil_div(p->N, bn);
    // Opt-in through m_hint (row-limited training GEMMs): atomics make the fp32 summation order
    // run-dependent, which the backward already tolerates (LayerNorm / column-sum / scatter
    // atomics) but greedy decoding must not -- its GEMMs never set a hint and stay bit-reproducible.
    if (sk && nb == 1 && p->m_limit != nullptr && p->m_hint > 0 && p->C != nullptr && p->C16 == nullptr &&
        p->act == TT_ACT_NONE && !p->accumulate && tiles * 2 <= sms && num_k >= 32 && p->col_stats == nullptr) {
      int sp = sms / tiles;
      if (sp > num_k / 8) sp = num_k / 8;     // at least 8 k-blocks per split
      if (sp > 1) {
        g.splits = sp;
        cudaMemset2DAsync(p->C, static_cast<size_t>(p->ldc) * sizeof(float), 0,
                          static_cast<size_t>(p->N) * sizeof(float), static_cast<size_t>(p->M), s);
        tiles *= sp;
      }
    }
    if (g.splits == 1) tiles = ceil_div(p->M, BM) * ceil_div(p->N, bn) * nb;
    else tiles = ceil_div(p->M, BM) * ceil_div(p->N, bn) * g.splits;
  }
  int grid = tiles < sms ? tiles : sms;
  if (gemm_sm_cap() > 0 && grid > gemm_sm_cap() && p->M >= 2048) grid = gemm_sm_cap();   // large problems only
  switch (bn) {
    case 256: return launch_gemm<256>(tmA, tmB, tmC, tmR, g, grid, ta, tb, s);
    case 128: return launch_gemm<128>(tmA, tmB, tmC, tmR, g, grid, ta, tb, s);
    case 64: return launch_gemm<64>(tmA, tmB, tmC, tmR, g, grid, ta, tb, s);
    default: return launch_gemm<32>(tmA, tmB, tmC, tmR, g, grid, ta, tb, s);
  }
}
