// 2-CTA (cta_group::2) variant of the tcgen05 GEMM for the large K-major GEMMs of the frozen
// encoders:  C[M,N] = act(alpha * (A[M,K] . B[N,K]^T + bias) + residual).
//
// Why: with one CTA per MMA, every SM must stream a full B tile (BN x 64) per k-block through its
// own shared memory; the 128 B/clk shared-memory port (TMA fill + UMMA operand reads) then caps
// the tensor pipe near 60-65 %.  A CTA pair (two SMs of one TPC, cluster 2x1) computes a 256 x BN
// tile with ONE tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128 rows of A but only
// HALF of B (BN/2 rows); the tensor cores of both SMs read both halves.  Per SM and k-block the
// shared-memory traffic drops from 48 KB to 32 KB for the same 128 x 256 outputs.
//
// Roles per CTA (384 threads): warp 0 TMA producer (both CTAs; the loads signal the LEADER's full
// barrier), warp 1 MMA issuer (leader CTA only; commits are multicast to both CTAs' barriers),
// warp 2 TMEM allocator (cta_group::2, both CTAs), warps 4-11 epilogue (each CTA drains its own
// 128 TMEM lanes = its 128 rows of the tile and reports to the leader's tmem_empty barrier).
#include <cstdlib>

#include "common.cuh"
#include "gemm_common.cuh"
#include "runtime.h"

namespace tt {

constexpr int G2_THREADS = 384;

template <int BN, bool STAGED>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;            // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;               // double-buffered 128 x BN accumulator
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES + 512;   // staged-epilogue tiles (512 B aligned)
  static constexpr int SMEM_BYTES = STAGED ? EPI_OFF + EPI_SMEM_BYTES + 1024 : STAGES * STAGE_BYTES + 256 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
// TMA tile load issued by either CTA of the pair; completes on the LEADER CTA's mbarrier
// (peer bit of the shared::cluster barrier address cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1) {
  const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_u(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                                   int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mcast_u(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives (once all MMAs issued so far by this thread retire) on the barrier at this smem offset in
// BOTH CTAs of the pair.
__device__ __forceinline__ void umma2_commit_mcast(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// mbarrier arrive on the barrier at the same smem offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int BN, int MODE>     // MODE as in gemm.cu: 0 plain, 1 staged epilogue, 2 staged + column statistics
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                  const GemmArgs g) {
  constexpr bool STAGED = MODE != 0;
  constexpr bool STATS = MODE == 2;
  using Cfg = Gemm2Cfg<BN, STAGED>;
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

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  pdl_launch_dependents();
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
      mbar_init(&full_bar[s], 1);    // leader: one arrive.expect_tx + bytes of both CTAs
      mbar_init(&empty_bar[s], 1);   // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);              // one multicast commit per tile
      mbar_init(&tmem_empty[s], 2 * 8);         // leader: 8 epilogue warps of each CTA
    }
    if constexpr (STAGED) {
      for (int s = 0; s < 8; ++s) mbar_init(&res_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();     // peer barriers initialised and peer TMEM allocated before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  int M = g.M;
  if (g.m_limit != nullptr) M = min(M, __ldg(g.m_limit));
  const int num_m = (M + 2 * BM - 1) / (2 * BM);   // 256-row tiles
  const int num_n = (g.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  int Keff = g.K;
  if (g.k_limit != nullptr) Keff = min(Keff, max(__ldg(g.k_limit), 1));
  const int num_k = (Keff + BK - 1) / BK;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // Whole warp in the loop (uniform registers), one elected lane issues -- see gemm.cu.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    const uint32_t full_u = smem_u32(full_bar);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int row_a = m_blk * 2 * BM + static_cast<int>(rank) * BM;
      const int row_b = n_blk * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const uint32_t bar = full_u + stage * 8;
          if (leader) mbar_arrive_expect_tx_u(bar, 2 * Cfg::STAGE_BYTES);
          tma_load_2d_pair_u(sA_u + stage * Cfg::A_BYTES, &tmA, bar, kb * BK, row_a);
          tma_load_2d_pair_u(sB_u + stage * Cfg::B_BYTES, &tmB, bar, kb * BK, row_b);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(sA));
      const uint64_t db0 = umma_desc_kmajor_sw128(smem_u32(sB));
      constexpr uint32_t KS = (UMMA_K * 2) >> 4;   // 16 k = 32 B inside the swizzle row
      const uint32_t empty_u = smem_u32(empty_bar), tfull_u = smem_u32(tmem_full);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = da0 + static_cast<uint64_t>((stage * Cfg::A_BYTES) >> 4);
          const uint64_t db = db0 + static_cast<uint64_t>((stage * Cfg::B_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma2_bf16(d_tmem, da + k * KS, db + k * KS, idesc, (kb | k) != 0 ? 1u : 0u);
            umma2_commit_mcast_u(empty_u + stage * 8);
            if (kb == num_k - 1) umma2_commit_mcast_u(tfull_u + acc * 8);
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
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps, both CTAs) =====================
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
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int row0 = m_blk * 2 * BM + static_cast<int>(rank) * BM + q * 32;
      // staged (TMA) epilogue for warps whose 32 rows are all valid; the ragged last rows of a
      // row-limited problem keep the register path (rows beyond the limit stay untouched)
      bool staged = false;
      if constexpr (STAGED) {
        staged = STATS ? row0 < M : row0 + 32 <= M;
        if (staged && g.tma_epi == 2 && lane == 0 && n_blk * BN + half * 32 < g.N)
          epi_request_residual(&tmR, ew, n_blk * BN + half * 32, row0);   // lands under this tile's MMAs
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      if constexpr (STAGED) {
        if (staged) epilogue_chunks_tma<BN, STATS>(g, &tmC, &tmR, tacc, half, row0, n_blk * BN, ew, min(32, M - row0));
      }
      if (staged) {
      } else {
        const long long row = static_cast<long long>(row0) + lane;
        epilogue_chunks<BN>(g, tacc, half, row, row < M, n_blk * BN);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);   // report to the leader's barrier
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
  cluster_sync_all();   // the peer may still be reading our shared memory / TMEM through the pair MMA
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int MODE>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                        const CUtensorMap& tmR, const GemmArgs& g, int pairs, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, MODE != 0>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<BN, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm2 BN=%d): %s", BN, cudaGetErrorString(e));
      return TT_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_k(gemm2_bf16_kernel<BN, MODE>, dim3(2 * pairs), dim3(G2_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB,
           tmC, tmR, g);
  return check_launch("gemm2_bf16_kernel");
}

int gemm_sm_cap();   // gemm.cu

// Called by tt_gemm_bf16_tn when its tile model (gemm.cu: choose_tile) picks the CTA-pair kernel
// with tile width bn (128 or 256) for a K-major problem.
int gemm2_launch(const TtGemmParams* p, const GemmArgs& g, const CUtensorMap& tmC, const CUtensorMap& tmR, int bn,
                 cudaStream_t stream) {
  int pairs_max = num_sms() / 2;
  if (gemm_sm_cap() > 0 && gemm_sm_cap() / 2 < pairs_max) pairs_max = gemm_sm_cap() / 2 > 0 ? gemm_sm_cap() / 2 : 1;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, p->A, (uint64_t)p->K, (uint64_t)p->M, (uint64_t)p->lda, BK, BM);
  if (rc != TT_OK) return rc;
  rc = make_tmap_bf16_2d(&tmB, p->B, (uint64_t)p->K, (uint64_t)p->N, (uint64_t)p->ldb, BK, bn / 2);
  if (rc != TT_OK) return rc;
  const int tiles = ceil_div(p->M, 2 * BM) * ceil_div(p->N, bn);
  const int pairs = tiles < pairs_max ? tiles : pairs_max;
  if (g.tma_epi != 0 && g.col_stats != nullptr)
    rc = (bn == 256) ? launch_gemm2<256, 2>(tmA, tmB, tmC, tmR, g, pairs, stream)
                     : launch_gemm2<128, 2>(tmA, tmB, tmC, tmR, g, pairs, stream);
  else if (g.tma_epi != 0)
    rc = (bn == 256) ? launch_gemm2<256, 1>(tmA, tmB, tmC, tmR, g, pairs, stream)
                     : launch_gemm2<128, 1>(tmA, tmB, tmC, tmR, g, pairs, stream);
  else
    rc = (bn == 256) ? launch_gemm2<256, 0>(tmA, tmB, tmC, tmR, g, pairs, stream)
                     : launch_gemm2<128, 0>(tmA, tmB, tmC, tmR, g, pairs, stream);
  return rc;
}

}  // namespace tt
