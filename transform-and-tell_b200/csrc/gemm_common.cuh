// Pieces shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 GEMM kernels: argument block
// and the register-level epilogue.
#pragma once
#include "common.cuh"

namespace tt {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;

struct GemmArgs {
  int M, N, K;
  float* C;
  long long ldc;
  __nv_bfloat16* C16;
  long long ldc16;
  const float* bias;
  const float* residual;
  long long ldr;
  const __nv_bfloat16* residual16;
  long long ldr16;
  float alpha;
  int act;
  int accumulate;
  const int* m_limit;
  const int* k_limit;   // device int, optional: operand rows/cols k >= *k_limit are known to be zero
  int splits;           // split-K factor (1-CTA kernel): partial products are added atomically into C
  int atomic;           // epilogue adds into C with atomics (set per tile by the split-K kernel)
  // Batched launch (single-CTA kernel, register epilogue): nbatch same-shape problems whose operands
  // are blocks of shared buffers; batch z adds z * off to the TMA coordinates of A / B, z * c_off
  // elements to the output pointers and z * bias_off to the bias pointer.
  int nbatch;
  int a_off0, a_off1, b_off0, b_off1, bias_off;
  long long c_off;
  int vec_ok;  // all fp32/bf16 row pointers 16-byte aligned for 32-column chunks
  float* col_stats;   // optional [2N]: += column sums / sums of squares of the bf16-rounded outputs
  int tma_epi;  // staged epilogue (bf16-only output, N % 32 == 0): 1 = TMA store of C16, 2 = + TMA load of residual16
  long long* trace;  // debug: [grid, 8] globaltimer stamps (tt_gemm_set_trace), normally null
};

__device__ __forceinline__ void trace_stamp(const GemmArgs& g, int slot) {
  if (g.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[blockIdx.x * 8 + slot] = static_cast<long long>(t);
  }
}
long long* gemm_trace_ptr();

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == TT_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TT_ACT_GELU) return gelu_erf(v);
  return v;
}

// Epilogue of one 128 x BN accumulator (one thread = one row, 32 columns per tcgen05.ld).
// tmem_acc: TMEM address of column 0 of the accumulator, already offset to this warp's lane quarter.
// Two warps share a lane quarter: `half` selects the even / odd 32-column chunks.
// Result = act(alpha * (acc + bias) + residual) -> fp32 and/or bf16, optional += into C.
template <int BN>
__device__ __forceinline__ void epilogue_chunks(const GemmArgs& g, uint32_t tmem_acc, int half,
                                                long long row, bool row_ok, int n0) {
  const int lane = threadIdx.x & 31;
  (void)lane;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int col0 = n0 + c * 32;
        const bool in_n = col0 < g.N;                       // warp-uniform
        const bool full = in_n && (col0 + 32 <= g.N) && g.vec_ok;
        float4 bv[8];
        float4 rv[8];
        uint4 rh[4];
        if (full) {
          if (g.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + j);
          }
          if (row_ok && g.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              rv[j] = __ldg(reinterpret_cast<const float4*>(g.residual + row * g.ldr + col0) + j);
          }
          if (row_ok && g.residual16 != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              rh[j] = __ldg(reinterpret_cast<const uint4*>(g.residual16 + row * g.ldr16 + col0) + j);
          }
        }
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + static_cast<uint32_t>(c * 32);
        tmem_ld_32x32(taddr, r);
        tmem_ld_wait();
        if (!in_n || !row_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (full) {
          if (g.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += bv[j].x; v[4 * j + 1] += bv[j].y; v[4 * j + 2] += bv[j].z; v[4 * j + 3] += bv[j].w;
            }
          }
          if (g.alpha != 1.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= g.alpha;
          }
          if (g.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += rv[j].x; v[4 * j + 1] += rv[j].y; v[4 * j + 2] += rv[j].z; v[4 * j + 3] += rv[j].w;
            }
          }
          if (g.residual16 != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
                v[8 * j + 2 * e] += __low2float(h2);
                v[8 * j + 2 * e + 1] += __high2float(h2);
              }
            }
          }
          if (g.act == TT_ACT_GELU && g.C == nullptr) {        // bf16-only output: bf16-accurate GELU
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_bf16(v[j]);
          } else if (g.act != TT_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], g.act);
          }
          if (g.C != nullptr) {
            float4* cp = reinterpret_cast<float4*>(g.C + row * g.ldc + col0);
            if (g.atomic) {               // split-K partial: C was zeroed by the host wrapper
#pragma unroll
              for (int j = 0; j < 8; ++j)
                atomicAdd(cp + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            } else {
              if (g.accumulate) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = cp[j];
                  v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j)
                cp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          if (g.C16 != nullptr) {
            uint4* hp = reinterpret_cast<uint4*>(g.C16 + row * g.ldc16 + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
              hp[j] = u;
            }
          }
        } else {
          // ragged N or unaligned pointers: scalar path
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < g.N) {
              float o = v[j];
              if (g.bias != nullptr) o += __ldg(g.bias + col);
              o *= g.alpha;
              if (g.residual != nullptr) o += __ldg(g.residual + row * g.ldr + col);
              if (g.residual16 != nullptr) o += __bfloat162float(g.residual16[row * g.ldr16 + col]);
              o = apply_act(o, g.act);
              if (g.C != nullptr) {
                float* cp = g.C + row * g.ldc + col;
                if (g.atomic) {
                  atomicAdd(cp, o);
                } else {
                  if (g.accumulate) o += *cp;
                  *cp = o;
                }
              }
              if (g.C16 != nullptr) g.C16[row * g.ldc16 + col] = __float2bfloat16_rn(o);
            }
          }
        }
      }
}

// ------------------------------------------------------------------------------------------------
// Staged epilogue for bf16-only outputs (g.tma_epi != 0; host guarantees N % 32 == 0, 16-byte aligned
// rows, no fp32 output / fp32 residual / accumulate / atomics).
//
// The register-level epilogue above is row-per-thread: every 16-byte global load / store instruction
// of a warp touches 32 different cache lines, and that LSU work (not the round trip) is what the
// large GEMMs and the ResNet 1x1 convolutions were bound by (profiles/r1_resnet_gemm_analysis.txt).
// Here each epilogue warp owns a 32-row x 64-byte staging tile in shared memory (+ one for the
// residual): the thread still owns one row of the chunk, but it writes its 4 x 16 bytes to SHARED
// memory (SWIZZLE_64B: chunk j of row t at j ^ ((t >> 1) & 3) -> a quarter-warp covers all 32 banks)
// and one lane hands the 2 KB tile to the TMA engine (cp.async.bulk.tensor store).  The bf16 residual
// tile arrives the same way (TMA load -> swizzled ld.shared), requested one chunk ahead -- the first
// chunk of a tile while that tile's MMAs are still running.
constexpr int EPI_WARP_BYTES = 32 * 64;                 // 32 rows x 32 bf16
constexpr int EPI_SMEM_BYTES = 2 * 8 * EPI_WARP_BYTES;  // output + residual staging of 8 warps

struct EpiWarp {
  uint32_t st_out, st_res, res_bar;   // shared-window addresses of this warp's tiles / barrier
  uint32_t res_phase;
};

__device__ __forceinline__ void epi_request_residual(const CUtensorMap* tmR, const EpiWarp& ew, int col0,
                                                     int row0) {
  mbar_arrive_expect_tx_u(ew.res_bar, EPI_WARP_BYTES);
  tma_load_2d_u(ew.st_res, tmR, ew.res_bar, col0, row0);
}

// Whole-warp, warp-uniform control flow.  row0: first of the 32 rows this warp owns (all < M).
// STATS: compiled only into the instantiations that serve col_stats problems (the extra loop costs
// the plain staged epilogue registers: +150 B of spills, 10-35 % on the epilogue-bound N = 64 shapes).
// rows_valid: how many of the warp's 32 rows exist (< 32 only in STATS mode, where the ragged last
// rows of the problem also take this path: they enter the tile and the sums as zeros and the TMA
// store clips them -- statistics problems carry no device-side row limit).
template <int BN, bool STATS>
__device__ __forceinline__ void epilogue_chunks_tma(const GemmArgs& g, const CUtensorMap* tmC,
                                                    const CUtensorMap* tmR, uint32_t tmem_acc, int half,
                                                    int row0, int n0, EpiWarp& ew, int rows_valid = 32) {
  const int lane = threadIdx.x & 31;
  const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
  const uint32_t my_out = ew.st_out + static_cast<uint32_t>(lane) * 64u;
  const uint32_t my_res = ew.st_res + static_cast<uint32_t>(lane) * 64u;
  const bool has_res = g.tma_epi == 2;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += 2) {
    const int col0 = n0 + c * 32;
    if (col0 >= g.N) break;
    float4 bv[8];
    if (g.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + j);
    }
    uint4 rh[4];
    if (has_res) {
      mbar_wait_u(ew.res_bar, ew.res_phase);
      ew.res_phase ^= 1u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = my_res + ((static_cast<uint32_t>(j) ^ sw) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rh[j].x), "=r"(rh[j].y), "=r"(rh[j].z), "=r"(rh[j].w) : "r"(a) : "memory");
      }
      // The refill below must not overtake the reads above: its coordinate is made to depend on
      // the last ld.shared (every lane executes the dependent instruction before the warp barrier).
      uint32_t zero;
      asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(rh[3].w));
      __syncwarp();
      const int cn = c + 2;
      if (lane == 0 && cn < BN / 32 && n0 + cn * 32 < g.N)
        epi_request_residual(tmR, ew, n0 + cn * 32 + static_cast<int>(zero), row0);
    }
    // tcgen05.ld writes its destination registers asynchronously until wait::ld: NOTHING may sit
    // between the two (with other code in between the compiler is free to spill / move the
    // not-yet-written registers -- observed as rare stale 16-byte pieces in the output).
    uint32_t r[32];
    tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c * 32), r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if constexpr (STATS) {
      if (lane >= rows_valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
    }
    if (g.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[4 * j] += bv[j].x; v[4 * j + 1] += bv[j].y; v[4 * j + 2] += bv[j].z; v[4 * j + 3] += bv[j].w;
      }
    }
    if (g.alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= g.alpha;
    }
    if (has_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {rh[j].x, rh[j].y, rh[j].z, rh[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
          v[8 * j + 2 * e] += __low2float(h2);
          v[8 * j + 2 * e + 1] += __high2float(h2);
        }
      }
    }
    if (g.act == TT_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_bf16(v[j]);
    } else if (g.act == TT_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    // the previous chunk's store must have finished reading the staging tile
    if (lane == 0) bulk_wait_group_read0();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t a = my_out + ((static_cast<uint32_t>(j) ^ sw) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                   "r"(pack_bf16(v[8 * j], v[8 * j + 1])), "r"(pack_bf16(v[8 * j + 2], v[8 * j + 3])),
                   "r"(pack_bf16(v[8 * j + 4], v[8 * j + 5])), "r"(pack_bf16(v[8 * j + 6], v[8 * j + 7]))
                   : "memory");
    }
    fence_proxy_async();                  // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      tma_store_2d_u(tmC, ew.st_out, col0, row0);
      bulk_commit_group();
    }
    if constexpr (STATS) {
      // lane c sums column c of the staged bf16 tile (32 lanes read one 64-byte row per step: one
      // wavefront), then one atomic per column per 32 rows
      const uint32_t jc = static_cast<uint32_t>(lane >> 3), off = static_cast<uint32_t>(lane & 7) * 2u;
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const uint32_t a = ew.st_out + static_cast<uint32_t>(rr) * 64u + ((jc ^ static_cast<uint32_t>((rr >> 1) & 3)) << 4) + off;
        unsigned short h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a));
        const float x = __uint_as_float(static_cast<uint32_t>(h) << 16);
        s += x;
        q = fmaf(x, x, q);
      }
      atomicAdd(g.col_stats + col0 + lane, s);
      atomicAdd(g.col_stats + g.N + col0 + lane, q);
    }
  }
}

}  // namespace tt
