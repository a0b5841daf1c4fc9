// Argument block shared by the fp32 SIMT (attention.cu) and bf16 tensor-core (attention_tc.cu)
// cross-attention kernels.
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>

namespace tt {

struct AttnArgs {
  const float* q;
  const float* k;
  const float* v;
  const float* bias_k;  // [E] or null
  const float* bias_v;
  const uint8_t* mask;  // [B,S] 1 = padding, or null
  float* out;           // [T,B,E]
  float* lse;           // [B,H,T]
  int T, B, S, H;
  long long ldq, ldkv, ldo;  // row strides (elements) of q/dq, k/v/dk/dv, out/dout
  int zero_row;         // add_zero_attn
  float p_drop;
  unsigned long long seed;
  const unsigned long long* step_ptr;
  // backward
  const float* dout;
  float* dq;
  float* dk;
  float* dv;
  float* dbias_k;
  float* dbias_v;
  // head-averaged weights (eval): [B,T,L]
  float* avg_w;
  // decode kernel only: element strides of k/v between keys, batch rows and heads
  // (token-major projection buffer: B*ldkv, ldkv, 64; head-major decode cache: 64, H*S*64, S*64)
  long long kv_j_stride, kv_b_stride, kv_h_stride;
  int Lp;               // decode kernel: padded key count of THIS context (scores row pitch in smem)
  const int* kv_len;    // decode kernel, optional [B]: keys j >= kv_len[b] are all padding (not loaded)
  // optional bf16 twins (tensor-core kernels only): out16 mirrors out (operand of out_proj),
  // dq16 mirrors dq (operand of the query projection's backward GEMMs); row pitches in elements
  __nv_bfloat16* out16;
  __nv_bfloat16* dq16;
  long long ldo16, ldq16;
  // backward, optional [B,H,T]: D = rowsum(dO * O), written by the dQ kernel and read by the dK|dV
  // kernel (which would otherwise recompute it in every key-tile CTA: 2 x 16 KB of fp32 per CTA)
  float* dsum;
  int dkv_tpc;          // dK|dV kernel: consecutive key tiles per CTA (set by the launcher)
};

// Up to four contexts per launch (image / article / faces / objects of one decoder layer,
// decoder_faces_objects.py:272-352): blockIdx.z selects the context, so a layer's cross-attention
// is ONE launch per kernel type instead of one per context.
constexpr int ATTN_MAX_CTX = 4;
struct AttnArgsN {
  AttnArgs a[ATTN_MAX_CTX];
};

}  // namespace tt
