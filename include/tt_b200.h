/* tt_b200.h -- C ABI of libtt_b200.so: the sm_100a kernels behind the Transform-and-Tell
 * caption hot path (train forward+backward and greedy decode).
 *
 * Conventions:
 *   - every buffer (inputs, outputs, workspace) is owned by the caller; the library never
 *     allocates, frees or synchronises device memory and launches only on `stream`
 *     (a cudaStream_t passed as void*);
 *   - plain pointers and sizes only, no torch types;
 *   - return value: 0 (TT_OK) or a negative TtStatus; tt_last_error() gives the thread-local
 *     message; no C++ exception crosses the ABI;
 *   - "rows" are flattened (time, batch) positions; feature dimension is contiguous.
 *
 * Each entry point cites the reference call site (relative to the reference repository root)
 * whose torch ops it replaces.
 */
#ifndef TT_B200_H_
#define TT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TT_OK = 0,
  TT_ERR_INVALID = -1, /* bad argument / unsupported shape */
  TT_ERR_CUDA = -2,    /* CUDA runtime or driver error */
  TT_ERR_NO_DEVICE = -3
} TtStatus;

typedef enum { TT_ACT_NONE = 0, TT_ACT_RELU = 1, TT_ACT_GELU = 2 } TtActivation;

const char* tt_last_error(void);
/* Library/ABI version; bumped when a signature changes. */
int tt_abi_version(void);
/* Number of kernel launches issued by this library in this process (bench `gpu_launches`). */
long long tt_launch_count(void);
void tt_reset_launch_count(void);
/* Dropout masks are a pure function of (seed argument, element index, *step) where `step` is an
 * optional device-resident counter registered here (NULL = none).  Advancing it inside a captured
 * CUDA graph gives every replay fresh masks while forward and backward of one step still agree. */
void tt_set_rng_step_ptr(const unsigned long long* dev_ptr);
int tt_rng_step_advance(unsigned long long* dev_ptr, void* stream);

/* ------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = act(alpha * (A[M,K] . B[N,K]^T + bias[N]) + residual[M,N])
 * bf16 operands (both K-contiguous, i.e. nn.Linear's x @ W^T), fp32 accumulation in TMEM
 * (tcgen05.mma), TMA operand staging.  Replaces every F.linear / addmm on the path:
 *   tell/modules/linear.py:8-34 (GehringLinear), tell/modules/attention/multi_head.py:491-518
 *   (in_proj_q/k/v), :476 (out_proj), tell/modules/softmax.py:169-191 (head/tail logits),
 *   tell/modules/token_embedders/adaptive.py:61-76 (band projections),
 *   tell/modules/convolutions/dynamic.py:300 (filter logits).
 * lda/ldb/ldc/... are row strides in ELEMENTS; lda, ldb multiples of 8; A, B 16-byte aligned.
 * With trans_a/trans_b the same kernel reads MN-major operands (UMMA descriptor major bits).
 * C (fp32) and/or C16 (bf16) receive the result; accumulate!=0 adds into C (fp32 only).
 * m_limit (device int, optional): only the first min(M, *m_limit) rows are computed.
 */
typedef struct {
  int M, N, K;
  const void* A;
  long long lda;
  const void* B;
  long long ldb;
  float* C;
  long long ldc;
  void* C16;
  long long ldc16;
  const float* bias;
  const float* residual;   /* fp32 [M,N], optional */
  long long ldr;
  const void* residual16;  /* bf16 [M,N], optional (ResNet identity branch) */
  long long ldr16;
  float alpha;
  int act;
  int accumulate;
  const int* m_limit;
  /* Operand storage.  0: [M,K] / [N,K] row-major (K contiguous, "K-major").
   * 1: stored transposed, [K,M] / [K,N] row-major (M or N contiguous, "MN-major"); lda/ldb are
   * then the row strides of that storage.  Lets the backward GEMMs (dX = dY.W, dW = dY^T.X)
   * consume the forward's operands in place instead of materialising transposes. */
  int trans_a;
  int trans_b;
  /* Host-side estimate of *m_limit (0 = unknown).  Scheduling hint only: it picks the tile shape
   * whose tile count quantises best into waves of the machine; results never depend on it. */
  int m_hint;
  /* Device int, optional: the caller guarantees that both operands are ZERO for contraction
   * indices k >= *k_limit (e.g. the rows of a gradient beyond a device-side row count), so the
   * kernel may stop its K loop there (rounded up to whole 64-wide k-blocks, at least one). */
  const int* k_limit;
  /* Optional fp32 [2*N], zeroed by the caller: the kernel ADDS the per-column sum of the (bf16-rounded)
   * outputs to col_stats[0..N) and the sum of their squares to col_stats[N..2N) -- the batch statistics
   * a train-mode BatchNorm behind this convolution needs (tt_bn_apply_bf16), accumulated in the epilogue
   * instead of by a separate pass over the output.  Needs the bf16 output (C16). */
  float* col_stats;
  /* Batched launch: nbatch (<= 16; 0 / 1 = a single problem) same-shape problems in ONE launch -- the
   * four out-projections of a decoder layer's cross-attentions (multi_head.py:476), their dX and
   * their dW GEMMs.  Operands are blocks of shared buffers: batch z reads A at TMA coordinate
   * (c0 + z*a_off0, c1 + z*a_off1) [c0 = the contiguous axis of the STORED matrix, c1 = its rows],
   * likewise B; writes C / C16 at + z*c_off elements and reads bias at + z*bias_off.  lda / ldb / ldc
   * are the pitches of the shared buffers.  Plain problems only (bias / activation; no residual, row
   * limit, accumulate, col_stats); K-major operands batched along K need K % 64 == 0. */
  int nbatch;
  int a_off0, a_off1, b_off0, b_off1, bias_off;
  long long c_off;
} TtGemmParams;
int tt_gemm_bf16_tn(const TtGemmParams* p, void* stream);
/* Debug hook: when non-NULL, every CTA of the 1-CTA GEMM kernel writes 8 %globaltimer stamps (
 * (entry, setup done, first TMA issued, first tile landed, last MMA issued, accumulator ready,
 * epilogue done, exit) to dev_ptr[cta*8 ..]; NULL (default) disables it. */
void tt_gemm_set_trace(long long* dev_ptr);
/* Experiment / test switch: 0 forces the register-level (row-per-thread) epilogue for every problem,
 * non-zero (default; also env TT_GEMM_TMA_EPI) lets bf16-only outputs with N % 32 == 0 take the staged
 * epilogue (shared-memory tile + cp.async.bulk.tensor store, bf16 residual by TMA load).  Both paths
 * produce bit-identical results. */
void tt_gemm_set_staged_epilogue(int on);
/* Scheduling switch (also env TT_GEMM_SM_CAP): 0 (default) = persistent GEMM grids use every SM;
 * n > 0 = large GEMMs (the CTA-pair kernel, and single-CTA problems with M >= 2048) launch on at most
 * n SMs, leaving the rest to kernels of concurrent streams.  Set it around the launches / graph capture
 * of the stream that should yield (the frozen encoders).  Results identical. */
void tt_gemm_set_sm_cap(int sms);
/* Tile-choice objective (also env TT_GEMM_OCC_WEIGHT): 0 (default) = the configuration with the shortest
 * launch; w > 0 = cost * (fraction of SMs held)^w, w = 1 being SMs x time -- for steps whose streams
 * share the machine (throughput bound by SM occupancy, not by one chain's latency).  Results identical. */
void tt_gemm_set_occupancy_weight(float w);
/* Programmatic dependent launch for every kernel of the library (also env TT_PDL; default off): the next
 * kernel's CTAs are scheduled while the current one still runs and wait at their first instruction.  Worth
 * it for single latency-bound chains (decoder forward/backward graph, decode step); graphs with parallel
 * branches lose.  Returns the previous setting.  Results identical. */
int tt_set_pdl(int on);

/* fp32 [rows, cols] (row stride ld_src) -> bf16 operand for tt_gemm_bf16_tn.
 *   transpose==0: dst is [rows, cols*rep]   transpose!=0: dst is [cols, rows*rep]
 *   split==0 (rep=1): plain round-to-nearest bf16 (throughput mode)
 *   split==1 (rep=3): error-compensated operand "A side": [hi | lo | hi] along K
 *   split==2 (rep=3): error-compensated operand "B side": [hi | hi | lo] along K
 * so that A'.B'^T = hi.hi + lo.hi + hi.lo (parity mode, ~2^-16 relative error).
 * ld_dst in elements.  seg_stride: distance (elements) between the three K segments; 0 = the
 * contraction length of this call (cols, or rows when transposed).  A larger value lets several
 * calls assemble ONE operand whose K axis is a concatenation (e.g. the band projections). */
int tt_cast_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, int rows,
                 int cols, int transpose, int split, long long seg_stride, void* stream);
/* ------------------------------------------------------------------------------------------
 * Weight bank: every weight operand of the decoder prepared by ONE launch per step, every
 * weight-norm backward by ONE launch (bank.cu).  Replaces the per-module nn.utils.weight_norm
 * recompute of GehringLinear (tell/modules/linear.py:30-34) and the per-call fp32->bf16 operand
 * casts.  Tables live in DEVICE memory (built once per model; pointers are stable), rows of all
 * segments are numbered consecutively: row0 = exclusive prefix sum of rows, ascending.
 */
typedef struct {
  const float* src;  /* [rows, cols] fp32 (weight_v, or a plain weight), row stride ld_src */
  const float* g;    /* weight_g [rows], or NULL: plain cast */
  float* w32;        /* optional contiguous fp32 effective weight g*v/||v|| [rows, cols], or NULL */
  float* norm;       /* optional ||v|| per row (kept for the backward), or NULL */
  void* dst16;       /* bf16 operand [rows, cols], row stride ld_dst */
  long long ld_src, ld_dst;
  int rows, cols;
  int row0;
  int pad_;
} TtPrepSeg;
int tt_weight_prep(const TtPrepSeg* segs_dev, int nsegs, int total_rows, void* stream);
typedef struct {
  const float* dw;   /* dL/dw [rows, cols] contiguous */
  const float* v;
  const float* g;
  const float* norm;
  float* dv;         /* [rows, cols] */
  float* dg;         /* [rows] */
  int rows, cols;
  int row0;
  int pad_;
} TtWnormBwdSeg;
int tt_wnorm_bwd_multi(const TtWnormBwdSeg* segs_dev, int nsegs, int total_rows, void* stream);

/* out[0] = a[0] * b[0] (device scalars; folds the upstream loss gradient without a host sync). */
int tt_scalar_mul(const float* a, const float* b, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise decoder-layer kernels (fp32 activations, one warp per row).
 */
/* x = res + dropout(h) is written back into h (it is what the backward needs); y = LN(x).
 * decoder_faces_objects.py:263-266, 283-287, 361-364 (dropout -> residual -> post-LayerNorm).
 * res may be NULL; mean/rstd [N] saved for tt_ln_bwd.  ldy = row stride of y (concat buffers). */
int tt_ln_fwd(float* h, const float* res, const float* gamma, const float* beta, float* y,
              long long ldy, float* mean, float* rstd, int N, int E, float eps, float p_drop,
              unsigned long long seed, void* stream);
/* dx = dL/d(pre-norm sum) (goes to the residual); dh = dx * dropmask/(1-p) (goes to the branch);
 * dgamma/dbeta [E] are ACCUMULATED (atomicAdd) -- zero them first.  Any output may be NULL. */
int tt_ln_bwd(const float* dy, long long lddy, const float* x, const float* mean,
              const float* rstd, const float* gamma, float* dx, float* dh, float* dgamma,
              float* dbeta, int N, int E, float p_drop, unsigned long long seed, void* stream);
/* nn.GLU over the last dim: h [N,2C] -> out [N,C].  decoder_faces_objects.py:193-195,259-260 */
int tt_glu_fwd(const float* h, float* out, long long N, int C, void* stream);
int tt_glu_bwd(const float* dout, const float* h, float* dh, long long N, int C, void* stream);
/* F.dropout with a counter-based mask: y[i] = x[i] * keep(seed,i)/(1-p).  Calling it again on the
 * gradient with the same seed is the backward.  decoder_faces_objects.py:106,258 */
int tt_dropout(const float* x, float* y, long long n, float p, unsigned long long seed,
               void* stream);
/* y = a*x + b*y */
int tt_axpby(const float* x, float* y, long long n, float a, float b, void* stream);
/* nn.utils.weight_norm (dim=0) of GehringLinear, linear.py:30-34: w[o,:] = g[o] v[o,:]/||v[o,:]|| */
int tt_wnorm_fwd(const float* v, const float* g, float* w, float* norm, int O, int I,
                 void* stream);
int tt_wnorm_bwd(const float* dw, const float* v, const float* g, const float* norm, float* dv,
                 float* dg, int O, int I, void* stream);
/* mask[r] = any(isnan(x[r,:])), NaN rows zeroed in place. transformer_faces_objects.py:374-379 */
int tt_nan_rows(float* x, uint8_t* mask, int R, int D, void* stream);

/* ------------------------------------------------------------------------------------------
 * "Twin" producers (twin.cu): the kernels above that also write their result as the bf16 operand
 * of the tcgen05 GEMM that consumes it (nn.Linear / GehringLinear of decoder_faces_objects.py:255-365
 * and their backward), so no standalone fp32 -> bf16 cast pass runs between a row kernel and a GEMM.
 * fp32 outputs are bit-identical to the single-output entries; *16 outputs are bf16 (round to
 * nearest even), contiguous unless a pitch is given, and optional (NULL = not written).
 */
int tt_dropout_tw(const float* x, float* y, void* y16, long long n, float p, unsigned long long seed,
                  void* stream);                                   /* n % 4 == 0; y or y16 may be NULL */
int tt_relu_bwd_tw(const float* dy, const float* y, float* dx, void* dx16, long long n, void* stream);
int tt_glu_fwd_tw(const float* h, float* out, void* out16, long long N, int C, void* stream);
int tt_glu_bwd_tw(const float* dout, const float* h, float* dh, void* dh16, long long N, int C,
                  void* stream);
/* n <= 4 LayerNorms sharing their residual in ONE launch: the parallel context branches of a decoder
 * layer (decoder_faces_objects.py:272-352: x_c = LN_c(X + dropout(attn_c)) for image / article / faces
 * / objects, then torch.cat :354) -- context c reads h[c] [N,E] (overwritten with the pre-norm sum),
 * writes Y[:, c*E:(c+1)*E] as fp32 (y, ldy) and/or bf16 (y16, ldy16) and mean[c] / rstd[c] [N].
 * n = 1 is tt_ln_fwd with an operand twin.  E % 4 == 0, E <= 1024. */
typedef struct {
  float* h[4];
  const float* gamma[4];
  const float* beta[4];
  float* mean[4];
  float* rstd[4];
  unsigned long long seed[4];
  const float* res;            /* shared residual [N,E] or NULL */
  float* y;                    /* [N, >= n*E] fp32 or NULL */
  long long ldy;
  void* y16;                   /* [N, >= n*E] bf16 or NULL */
  long long ldy16;
  int n, N, E;
  float eps, p_drop;
} TtLnFwdMulti;
int tt_ln_fwd_multi(const TtLnFwdMulti* p, void* stream);
/* Backward of tt_ln_fwd_multi in ONE launch: dy [N, n*E] (lddy); per context dh[c] (fp32 [N,E],
 * optional) and dh16[:, c*E:(c+1)*E] (bf16, optional) = dx_c * dropmask_c/(1-p); dx [N,E] = sum_c dx_c
 * (the gradient of the shared residual, optional); dgamma[c] / dbeta[c] ACCUMULATED (zero them). */
typedef struct {
  const float* x[4];           /* pre-norm sums saved by the forward (its h[c]) */
  const float* mean[4];
  const float* rstd[4];
  const float* gamma[4];
  float* dgamma[4];
  float* dbeta[4];
  float* dh[4];
  unsigned long long seed[4];
  const float* dy;
  long long lddy;
  float* dx;
  void* dh16;
  long long lddh16;
  int n, N, E;
  float p_drop;
} TtLnBwdMulti;
int tt_ln_bwd_multi(const TtLnBwdMulti* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * DynamicConv1dTBC core, tell/modules/convolutions/dynamic.py:285-336 (_forward_expanded).
 * x/out [T,B,C]; z = filter logits [T,B,H,K] (z_tb_stride = H*K) or [H,K] broadcast
 * (z_tb_stride = 0: LightweightConv1dTBC, lightweight.py:88-240); K <= 32, C/H <= 64.
 * probs [T,B,H,K] (softmax output before DropConnect) is saved for the backward.
 */
int tt_dynconv_fwd(const float* x, const float* z, long long z_tb_stride, float* out, float* probs,
                   int T, int B, int C, int H, int K, int softmax, float p_drop,
                   unsigned long long seed, void* stream);
int tt_dynconv_bwd(const float* dout, const float* x, const float* probs, float* dx, float* dz,
                   int T, int B, int C, int H, int K, int softmax, float p_drop,
                   unsigned long long seed, void* stream);
/* One incremental decoding step of DynamicConv1dTBC / LightweightConv1dTBC (dynamic.py:95-116,
 * lightweight.py incremental branch) for T = 1: window [K-1,B,C] is the time-ordered input buffer
 * (updated in place: shifted by one step, x_new appended), z [B, H*K] the new row's tap logits
 * (z_b_stride = H*K, or 0 for taps shared by the batch), out [B,C]. */
int tt_dynconv_step(float* window, const float* x_new, const float* z, long long z_b_stride,
                    float* out, int B, int C, int H, int K, int softmax, void* stream);
/* The same three with bf16 twins: out16 mirrors out (operand of linear2, decoder_faces_objects.py:262),
 * dz16 [T*B, lddz16] mirrors dz (operand of the filter projection's backward GEMMs, dynamic.py:300). */
int tt_dynconv_fwd_tw(const float* x, const float* z, long long z_tb_stride, float* out, float* probs,
                      int T, int B, int C, int H, int K, int softmax, float p_drop,
                      unsigned long long seed, void* out16, void* stream);
int tt_dynconv_bwd_tw(const float* dout, const float* x, const float* probs, float* dx, float* dz,
                      int T, int B, int C, int H, int K, int softmax, float p_drop,
                      unsigned long long seed, void* dz16, long long lddz16, void* stream);
int tt_dynconv_step_tw(float* window, const float* x_new, const float* z, long long z_b_stride,
                       float* out, int B, int C, int H, int K, int softmax, void* out16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cross-attention core, tell/modules/attention/multi_head.py:355-466 (static_kv=True path).
 * q/out [T,B,H*D] (q pre-scaled by D^-0.5), k/v [S,B,H*D]; extended key set is
 * [k rows ; bias_k row (if bias_k) ; zero row (if zero_row)]; key_padding_mask [B,S] (1 = pad).
 * ldq/ldkv/ldo: row strides (elements) of q & dq, k/v & dk/dv, out & dout, so that fused
 * projection outputs ([N,4E] queries, [S*B,2E] key|value) are consumed without copies.
 * lse [B,H,T] saved for the backward.  dbias_k/dbias_v [H*D] are ACCUMULATED (atomicAdd).
 * tt_attn_avg_weights accumulates head-averaged probabilities into avg_w [B,T,L] (zero it first).
 */
int tt_attn_fwd(const float* q, const float* k, const float* v, const float* bias_k,
                const float* bias_v, const uint8_t* key_padding_mask, float* out, float* lse,
                int T, int B, int S, int H, int D, long long ldq, long long ldkv, long long ldo,
                int zero_row, float p_drop, unsigned long long seed, void* stream);
int tt_attn_bwd(const float* dout, const float* q, const float* k, const float* v,
                const float* bias_k, const float* bias_v, const uint8_t* key_padding_mask,
                const float* out, const float* lse, float* dq, float* dk, float* dv,
                float* dbias_k, float* dbias_v, int T, int B, int S, int H, int D, long long ldq,
                long long ldkv, long long ldo, int zero_row, float p_drop,
                unsigned long long seed, void* stream);
int tt_attn_avg_weights(const float* q, const float* k, const float* bias_k,
                        const uint8_t* key_padding_mask, const float* lse, float* avg_w, int T,
                        int B, int S, int H, int D, long long ldq, long long ldkv, int zero_row,
                        void* stream);

/* Tensor-core (bf16 mma.sync, fp32 accumulate/softmax) variants of tt_attn_fwd / tt_attn_bwd with
 * identical arguments; head_dim must be 64.  Used in the bf16 throughput mode. */
int tt_attn_fwd_tc(const float* q, const float* k, const float* v, const float* bias_k,
                   const float* bias_v, const uint8_t* key_padding_mask, float* out, float* lse,
                   int T, int B, int S, int H, int D, long long ldq, long long ldkv, long long ldo,
                   int zero_row, float p_drop, unsigned long long seed, void* stream);
int tt_attn_bwd_tc(const float* dout, const float* q, const float* k, const float* v,
                   const float* bias_k, const float* bias_v, const uint8_t* key_padding_mask,
                   const float* out, const float* lse, float* dq, float* dk, float* dv,
                   float* dbias_k, float* dbias_v, int T, int B, int S, int H, int D, long long ldq,
                   long long ldkv, long long ldo, int zero_row, float p_drop,
                   unsigned long long seed, void* stream);
/* Variants whose projected keys|values are bf16 in HBM (the K|V projection GEMM writes bf16
 * directly: half the bytes of these HBM-bound kernels) and whose dL/dk, dL/dv are written as bf16
 * -- exactly the operand the projection's weight-gradient GEMM consumes.  ldkv in bf16 elements,
 * multiple of 8; k16/v16 16-byte aligned. */
int tt_attn_fwd_tc_kv16(const float* q, const void* k16, const void* v16, const float* bias_k,
                        const float* bias_v, const uint8_t* key_padding_mask, float* out,
                        float* lse, int T, int B, int S, int H, int D, long long ldq,
                        long long ldkv, long long ldo, int zero_row, float p_drop,
                        unsigned long long seed, void* stream);
int tt_attn_bwd_tc_kv16(const float* dout, const float* q, const void* k16, const void* v16,
                        const float* bias_k, const float* bias_v,
                        const uint8_t* key_padding_mask, const float* out, const float* lse,
                        float* dq, void* dk16, void* dv16, float* dbias_k, float* dbias_v, int T,
                        int B, int S, int H, int D, long long ldq, long long ldkv, long long ldo,
                        int zero_row, float p_drop, unsigned long long seed, void* stream);
/* One decoder layer's cross-attention over up to 4 contexts (image / article / faces / objects,
 * decoder_faces_objects.py:272-352; multi_head.py:355-466 per context) as ONE launch per kernel type:
 * blockIdx.z walks the contexts.  Same math as tt_attn_fwd_tc / tt_attn_bwd_tc(_kv16) per context; T, B,
 * H, D, zero_row, p_drop are shared, everything else is per context.  kv16: k / v / dk / dv are bf16.
 * tt_attn_decode_hm_multi: the T = 1 incremental step over head-major caches (k, v = [B,H,S,64] bf16;
 * S = 0 with null k / v is an empty context: only bias_k / the zero row). */
typedef struct {
  const float* q;              /* [T*B, ldq] already scaled */
  const void* k;               /* fp32 or bf16 (kv16) [S*B, ldkv] */
  const void* v;
  const float* bias_k;         /* [E] or NULL */
  const float* bias_v;
  const unsigned char* mask;   /* [B,S] 1 = padding, or NULL */
  float* out;                  /* [T*B, ldo] */
  float* lse;                  /* [B,H,T] */
  int S;
  long long ldq, ldkv, ldo;
  unsigned long long seed;     /* dropout stream of this context */
  const float* dout;           /* backward only */
  float* dq;
  void* dk;
  void* dv;
  float* dbias_k;
  float* dbias_v;
  const int* kv_len;           /* optional [B]: keys j >= kv_len[b] of sample b are all padding (mask = 1): the
                                * kernels skip those rows / whole key tiles (results identical) */
  void* out16;                 /* optional bf16 twin of out [T*B, ldo16]: the operand of out_proj (multi_head.py:476) */
  void* dq16;                  /* optional bf16 twin of dq [T*B, ldq16]: the operand of in_proj_q's backward */
  long long ldo16, ldq16;
  float* dsum;                 /* backward, optional scratch [B,H,T]: D = rowsum(dO * O), written by the dQ kernel and
                                * read by the dK|dV kernel instead of recomputing it in every key-tile CTA */
} TtAttnCtx;
int tt_attn_fwd_tc_multi(const TtAttnCtx* ctx, int n, int T, int B, int H, int D, int zero_row, float p_drop,
                         int kv16, void* stream);
int tt_attn_bwd_tc_multi(const TtAttnCtx* ctx, int n, int T, int B, int H, int D, int zero_row, float p_drop,
                         int kv16, void* stream);
int tt_attn_decode_hm_multi(const TtAttnCtx* ctx, int n, int B, int H, int D, int zero_row, void* stream);
/* Key tiles (of 64 keys) one CTA of the dK|dV kernel walks with the query side staged once; 0 = chosen from
 * the problem size (the default).  Results do not depend on it (tests force 1 / 2 / 5). */
void tt_attn_set_dkv_tiles_per_cta(int n);
/* Incremental decoding (transformer_faces_objects.py:399-494 recomputes every K|V projection per
 * step; here they are projected once and cached).  tt_kv_repack_heads turns the token-major bf16
 * projection ([S*B, ldkv] rows, key j of batch b at row j*B+b) into head-major K, V [B,H,S,64];
 * tt_attn_decode_hm is the T = 1 attention over that cache: q [B, H*64] (row stride ldq, already
 * scaled), extended key set [S keys ; bias row ; zero row] as in multi_head.py:355-425,
 * out [B, H*64] (row stride ldo), lse [B,H] optional. */
int tt_kv_repack_heads(const void* k16, const void* v16, long long ldkv, void* k_out, void* v_out,
                       int S, int B, int H, int D, void* stream);
int tt_attn_decode_hm(const float* q, const void* k_hm, const void* v_hm, const float* bias_k,
                      const float* bias_v, const uint8_t* key_padding_mask, float* out, float* lse,
                      int B, int S, int H, int D, long long ldq, long long ldo, int zero_row,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Adaptive softmax / loss, tell/modules/softmax.py:144-222, criteria/adaptive_loss.py:27-73.
 * cutoffs = {c0, c1, ..., vocab} (n_clusters entries: head + tails).
 */
/* adapt_target (softmax.py:144-167) without host syncs: head_target [N]; per tail i an ordered
 * list of member rows tail_idx[i*N + slot], their local ids tail_local[i*N + slot], and
 * tail_count[i]; ntokens = #(target != pad_idx) (adaptive_loss.py:64-65). */
int tt_adaptive_prepare(const long long* target, int N, const int* cutoffs, int n_clusters,
                        int pad_idx, int* head_target, int* tail_idx, int* tail_local,
                        int* tail_count, int* ntokens, void* stream);
/* dst[i,:] = src[idx[i],:] for i < *count_ptr (all `cap` rows when NULL), else 0. */
int tt_gather_rows(const float* src, const int* idx, const int* count_ptr, float* dst, int cap,
                   int E, void* stream);
/* dst[idx[i],:] += src[i,:] for i < *count_ptr (atomicAdd; idx < 0 skipped). */
int tt_scatter_add_rows(const float* src, const int* idx, const int* count_ptr, float* dst,
                        int cap, int E, void* stream);
/* F.cross_entropy(reduction='sum', ignore_index) per row: row_loss [M], lse [M]; rows >= *count_ptr
 * contribute 0. */
int tt_ce_fwd(const float* logits, long long ld, const int* target, const int* count_ptr, int M,
              int V, int ignore_index, float* lse, float* row_loss, void* stream);
/* in place: logits <- (softmax - onehot) * *scale_ptr ; ignored / inactive rows <- 0. */
int tt_ce_bwd(float* logits, long long ld, const int* target, const int* count_ptr, int M, int V,
              int ignore_index, const float* lse, const float* scale_ptr, void* stream);
/* The same gradient (adaptive_loss.py:47-58 backward of F.cross_entropy) written as the bf16 operand
 * of the backward GEMMs: out16 [M, ld16] <- bf16((softmax - onehot) * *scale_ptr); logits are not
 * modified.  zero_round > 0 with count_ptr: rows >= max(round_up(*count_ptr, zero_round), zero_round) are NOT
 * written (their consumers are limited to *count_ptr rows); zero_round = 0 zeroes every inactive row. */
int tt_ce_bwd_bf16(const float* logits, long long ld, const int* target, const int* count_ptr,
                   int M, int V, int ignore_index, const float* lse, const float* scale_ptr,
                   void* out16, long long ld16, int zero_round, void* stream);
/* loss = sum(row_loss)/ln2/ntokens, scale = 1/(ln2*ntokens). transformer_faces_objects.py:85-90 */
int tt_loss_finalize(const float* row_loss, long long n, const int* ntokens, float* loss,
                     float* scale, void* stream);
/* get_log_prob (softmax.py:193-222) + greedy top-1 (transformer_faces_objects.py:443-464).
 * head [M, c0+n_tails], tails[i] [M, V_i].  Any of log_probs [M,vocab], argmax_id, argmax_lp
 * may be NULL. */
int tt_adaptive_logprob(const float* head, long long ld_head, const float* const* tails,
                        const long long* ld_tails, const int* cutoffs, int n_clusters, int M,
                        float* log_probs, long long* argmax_id, float* argmax_lp, void* stream);

/* ------------------------------------------------------------------------------------------
 * Embedding front end: adaptive.py:61-76, positional.py:167-268.
 * ids [B,T] int64; output row n = t*B+b when tbc != 0 (decoder works in T x B x C), else b*T+t.
 */
/* out [N, n_bands*E]: token's embedding row in its band's slot, zeros elsewhere. */
int tt_embed_gather(const long long* ids, int B, int T, int tbc, const int* cutoffs, int n_bands,
                    const float* const* tables, int E, float* out, void* stream);
/* grads[band][local,:] += dA[n, band*E:(band+1)*E], skipping local == padding_idx. */
int tt_embed_scatter_grad(const long long* ids, int B, int T, int tbc, const int* cutoffs,
                          int n_bands, float* const* grads, int E, int padding_idx,
                          const float* dA, void* stream);
/* make_positions (positional.py:231-268) + incremental start_pos (:196-198). */
int tt_make_positions(const long long* ids, int B, int T, int pad, int left_pad, int start_pos,
                      int tbc, int* pos, void* stream);
/* Same with a device-resident running position added to start_pos (incremental decoding inside a
 * captured CUDA graph: positional.py:170-176 keeps the position in the incremental state). */
int tt_make_positions_at(const long long* ids, int B, int T, int pad, int left_pad, int start_pos,
                         const int* start_dev, int tbc, int* pos, void* stream);
/* [A,B,C] -> [B,A,C] */
int tt_transpose01(const float* in, float* out, int A, int B, int C, void* stream);

/* out[c] = scale * sum_r x[r,c] (+ out[c] when accumulate): bias gradients. */
int tt_colsum(const float* x, long long ld, int M, int N, float* out, float scale, int accumulate,
              void* stream);
/* Same for a bf16 matrix (fp32 accumulation); N, ld multiples of 4. */
int tt_colsum_bf16(const void* x, long long ld, int M, int N, float* out, float scale,
                   int accumulate, void* stream);
/* dx = dy * (y > 0): backward of the ReLU fused into fc1's GEMM epilogue. */
int tt_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream);
/* RoBERTa layer mix, transformer_faces_objects.py:355-364: out = sum_l softmax(w)[l] * h_l.
 * hiddens: bf16, layer l at hiddens + l*layer_stride, n elements each.  bwd: dw [L] (= d bert_weight);
 * dots [L] is scratch. */
int tt_layer_mix_fwd(const void* hiddens, long long layer_stride, const float* w, int L,
                     long long n, float* out, void* stream);
int tt_layer_mix_bwd(const void* hiddens, long long layer_stride, const float* w,
                     const float* dout, int L, long long n, float* dots, float* dw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Frozen encoders (inference only, bf16 activations).
 * ResNet-152, tell/models/resnet.py:92-117 (torchvision Bottleneck, BatchNorm folded into the
 * conv weights/bias at load time): each convolution = im2col (NHWC) + tt_gemm_bf16_tn whose
 * epilogue adds the folded bias, the identity branch (residual16) and applies ReLU.
 */
/* out[(b,ho,wo), (kh,kw,c)] bf16, K zero-padded to Kp (multiple of 8); in: NHWC bf16. */
int tt_im2col_nhwc(const void* in, void* out, int B, int H, int W, int C, int KH, int KW,
                   int stride, int pad, int Kp, void* stream);
/* same from the NCHW fp32 input image (first 7x7/2 convolution). */
int tt_im2col_nchw_f32(const float* in, void* out, int B, int H, int W, int C, int KH, int KW,
                       int stride, int pad, int Kp, void* stream);
int tt_maxpool3x3s2_nhwc(const void* in, void* out, int B, int H, int W, int C, void* stream);
int tt_bf16_to_f32(const void* in, float* out, long long n, void* stream);
/* RoBERTa-large (fairseq hub `roberta.large`, extract_features(return_all_hiddens=True)); call
 * site transformer_faces_objects.py:352-353.  x = tok[ids] + pos[pad + #non-pad so far];
 * is_pad [B*S] flags padding rows. */
int tt_roberta_embed(const long long* ids, const float* tok, const float* pos, float* x,
                     uint8_t* is_pad, int B, int S, int E, int pad, void* stream);
/* y16 (bf16) = LayerNorm(x fp32) * gamma + beta; rows with row_zero[r] != 0 are zeroed. */
int tt_ln_fwd16(const float* x, const float* gamma, const float* beta, void* y16,
                const uint8_t* row_zero, int N, int E, float eps, void* stream);
/* Flash self-attention on bf16 tensor cores: qkv [B*S, 3*H*D] (q pre-scaled), mask [B,S] (1 = pad)
 * -> out [B*S, H*D] bf16.  D must be 64. */
int tt_flash_self_attn(const void* qkv, const uint8_t* key_padding_mask, void* out, int B, int S,
                       int H, int D, void* stream);
/* Variable-length ("packed") RoBERTa forward: articles are padded to S with `pad`
 * (tell/data/token_indexers/roberta_indexer.py:185-200) and every consumer of the encoder output
 * masks the padding (transformer_faces_objects.py:366 article_padding_mask), so the encoder only
 * has to compute the real tokens.  tt_varlen_prepare builds inv_map[B*S] (padded row -> packed row,
 * -1 = padding) and cu_seqlens[B+1] (first packed row of each sample; cu_seqlens[B] = number of
 * real tokens, the m_limit of the encoder GEMMs).  tt_flash_self_attn_varlen is the self-attention
 * over the packed layout (sample b = rows cu[b]..cu[b+1]); tt_ln_fwd16_varlen is the LayerNorm that
 * moves between the layouts (see encoders.cu) and writes zeros to the padding rows of the padded
 * hidden-state buffer; its input rows are fp32 (embedding sum) or bf16 (x_bf16 != 0: the pre-norm
 * residual sums the encoder GEMMs write). */
int tt_varlen_prepare(const long long* ids, int B, int S, int pad, int* inv_map, int* cu_seqlens,
                      void* stream);
int tt_flash_self_attn_varlen(const void* qkv, const int* cu_seqlens, void* out, int B, int S_max,
                              int H, int D, void* stream);
/* The same self-attention on tcgen05 tensor cores (TMEM accumulators, TMA tiles straight out of the
 * qkv matrix, flash_tc5.cu).  `rows` = allocated rows of qkv/out (>= cu_seqlens[B]); rows of qkv
 * beyond cu_seqlens[B] must hold finite values (masked keys still enter P.V with weight 0). */
int tt_flash_self_attn_varlen_tc5(const void* qkv, const int* cu_seqlens, void* out, int B, int S_max,
                                  int H, int D, long long rows, void* stream);
int tt_ln_fwd16_varlen(const void* x, int x_bf16, int x_packed, const float* gamma, const float* beta,
                       void* y_packed, void* y_padded, const int* inv_map, const int* count_ptr,
                       int R, int E, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step: BertAdam as configured by
 * expt/nytimes/9_transformer_objects/config.yaml:126-136 (allennlp 0.9 "bert_adam" ->
 * pytorch-pretrained-bert 0.6.2 BertAdam.step, called at
 * tell/training/callback_apex_trainer.py:238): per-TENSOR gradient clip to max_grad_norm, Adam
 * moments without bias correction, decoupled-style weight decay added to the update, lr scaled by
 * schedule(step / t_total).  All tensors of a parameter group in three launches.
 * Segment table (device memory): one entry per tensor; chunk0 = exclusive prefix sum of
 * ceil(n / tt_bertadam_chunk()) over the entries, ascending.
 * step_dev: device int64 step counter (BertAdam's state['step'], shared by all tensors), read for
 * the schedule and then incremented -- CUDA-graph replays advance it.
 * loss_dev (optional): device scalar; if it is NaN the step is skipped entirely (parameters,
 * moments and step untouched), the NaN-batch skip of callback_apex_trainer.py:225-227.
 * partial: float[total_chunks]; scratch: float[2 + nsegs] (out: [0] = lr used, [1] = skipped flag,
 * [2+i] = clip coefficient of tensor i).
 */
typedef struct {
  float* p;        /* parameter (fp32 master copy), updated in place */
  const float* g;  /* gradient */
  float* m;        /* next_m */
  float* v;        /* next_v */
  long long n;     /* elements */
  int chunk0;
  int pad_;
} TtAdamSeg;
typedef enum { TT_SCHED_NONE = 0, TT_SCHED_WARMUP_LINEAR = 1, TT_SCHED_WARMUP_CONSTANT = 2 } TtSchedule;
typedef struct {
  float lr, b1, b2, e, weight_decay, max_grad_norm, warmup;
  int schedule;        /* TtSchedule */
  long long t_total;   /* <= 0: constant lr */
} TtAdamHyper;
int tt_bertadam_chunk(void);
int tt_bertadam_step(const TtAdamSeg* segs_dev, int nsegs, int total_chunks,
                     const TtAdamHyper* hyper, long long* step_dev, const float* loss_dev,
                     float* partial, float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Face encoders: InceptionResnetV1 "FaceNet"
 * (tell/facenet/inception_resnet_v1.py:184-299) and the MTCNN P/R/O networks
 * (tell/facenet/mtcnn.py:11-159).  Inference only, NHWC bf16 activations; convolutions are
 * tt_im2col_nhwc_hw + tt_gemm_bf16_tn (folded BatchNorm / bias / residual scale / ReLU in the GEMM
 * epilogue, branch outputs written straight into channel slices of the concatenation buffer).
 * "pitch" = elements between consecutive pixels of an NHWC view (>= C, multiple of 8).
 */
/* nn.Conv2d / nn.MaxPool2d output extent (ceil_mode as in torch: the last window starts inside). */
int tt_conv_out_size(int in, int k, int stride, int pad, int ceil_mode);
/* im2col with rectangular kernels and per-axis padding (1x7 / 7x1 / 1x3 / 3x1 of Block17 / Block8,
 * inception_resnet_v1.py:75-78,103-106). */
int tt_im2col_nhwc_hw(const void* in, long long in_pitch, void* out, int B, int H, int W, int C,
                      int KH, int KW, int stride, int pad_h, int pad_w, int Kp, void* stream);
/* nn.MaxPool2d(k, stride, padding, ceil_mode) on NHWC bf16 (inception_resnet_v1.py:121,147,212;
 * mtcnn.py:23,66,69,116,119,122). */
int tt_maxpool_nhwc(const void* in, long long in_pitch, void* out, long long out_pitch, int B, int H,
                    int W, int C, int k, int stride, int pad, int ceil_mode, void* stream);
/* Train-mode nn.BatchNorm2d of the frozen ResNet (tell/models/resnet.py:35,92-117 under model.train(),
 * callback_apex_trainer.py:259; torchvision Bottleneck bn1/bn2/bn3/downsample.1): batch statistics.
 * x: raw convolution output, bf16 NHWC rows [M = B*H*W, C] (row pitch in elements).
 * tt_bn_stats_bf16: stats[0..C) += sum over rows, stats[C..2C) += sum of squares (fp32; the caller
 *   zeroes `stats` -- one memset for all layers of a forward).
 * tt_bn_apply_bf16: in place x = relu?((x - mean) * rsqrt(var_biased + eps) * gamma + beta + residual?);
 *   when running_mean/var are given they move by `momentum` towards the batch mean / UNBIASED batch
 *   variance and *num_batches_tracked (may be NULL) is incremented, as F.batch_norm(training=True). */
/* tt_im2col_nhwc_bn: tt_im2col_nhwc of relu((x - mean) * rsqrt(var + eps) * gamma + beta) where x is the
 *   RAW output of the previous convolution and (mean, var) come from `stats` (its col_stats, over
 *   n_stat rows): the train-mode BatchNorm + ReLU between a 1x1 and a 3x3 convolution applied while
 *   gathering, so the normalised activation is never written.  Padding taps stay zero.  Running
 *   statistics are updated exactly once, as in tt_bn_apply_bf16. */
int tt_im2col_nhwc_bn(const void* in, void* out, int B, int H, int W, int C, int KH, int KW, int stride,
                      int pad, int Kp, const float* stats, long long n_stat, const float* gamma,
                      const float* beta, float eps, float* running_mean, float* running_var,
                      float momentum, long long* num_batches_tracked, void* stream);
int tt_bn_stats_bf16(const void* x, long long pitch, long long M, int C, float* stats, void* stream);
int tt_bn_apply_bf16(void* x, long long pitch, long long M, int C, const float* stats, const float* gamma,
                     const float* beta, float eps, const void* residual, long long rpitch, int relu,
                     float* running_mean, float* running_var, float momentum,
                     long long* num_batches_tracked, void* stream);
/* nn.AdaptiveAvgPool2d(1) (inception_resnet_v1.py:249): [B, HW, C] bf16 -> [B, C] fp32. */
int tt_avgpool_nhwc(const void* in, float* out, int B, int HW, int C, void* stream);
/* nn.PReLU(C) in place on bf16 rows (mtcnn.py:22-27 ...). */
int tt_prelu_bf16(void* x, long long pitch, const float* slope, long long rows, int C, void* stream);
/* F.normalize(x, p=2, dim=1) (inception_resnet_v1.py:296). */
int tt_l2norm_rows(const float* x, float* y, int N, int D, float eps, void* stream);
/* nn.Softmax(dim=1) over the two face / non-face logits stored in columns [c0, c0+2) of fp32 rows
 * (mtcnn.py:29,75,126). */
int tt_softmax2(float* x, long long ld, long long rows, int c0, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batch collation.  The reference pads each field on the host and uploads
 * one tensor per field (allennlp TextField.as_tensor via tell/data/token_indexers/
 * roberta_indexer.py:185-200; ArrayField(padding_value=nan) for face / object features,
 * tell/data/dataset_readers/nytimes_faces_ner_matched.py:213-217).  Here the ragged rows of a field
 * are one flat buffer + row offsets [B+1] and the padding happens on the device:
 *   out[b, i, :] = i < n_b ? flat[(row_offsets[b] + i) * width + :] : fill.
 */
int tt_pad_ragged_f32(const float* flat, const long long* row_offsets, float* out, int B,
                      int max_rows, int width, float fill, void* stream);
int tt_pad_ragged_i64(const long long* flat, const long long* row_offsets, long long* out, int B,
                      int max_rows, long long fill, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H_ */
