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

/* ------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = act(alpha * A[M,K] . B[N,K]^T + bias[N]) + residual[M,N]
 * bf16 operands (both K-contiguous, i.e. nn.Linear's x @ W^T), fp32 accumulation in TMEM
 * (tcgen05.mma), TMA operand staging.  Replaces every F.linear / addmm on the path:
 *   tell/modules/linear.py:8-34 (GehringLinear), tell/modules/attention/multi_head.py:491-518
 *   (in_proj_q/k/v), :476 (out_proj), tell/modules/softmax.py:169-191 (head/tail logits),
 *   tell/modules/token_embedders/adaptive.py:61-76 (band projections),
 *   tell/modules/convolutions/dynamic.py:300 (filter logits).
 * lda/ldb/ldc/... are row strides in ELEMENTS; lda, ldb multiples of 8; A, B 16-byte aligned.
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
  const float* residual;
  long long ldr;
  float alpha;
  int act;
  int accumulate;
  const int* m_limit;
} TtGemmParams;
int tt_gemm_bf16_tn(const TtGemmParams* p, void* stream);

/* fp32 [rows, cols] (row stride ld_src) -> bf16 operand for tt_gemm_bf16_tn.
 *   transpose==0: dst is [rows, cols*rep]   transpose!=0: dst is [cols, rows*rep]
 *   split==0 (rep=1): plain round-to-nearest bf16 (throughput mode)
 *   split==1 (rep=3): error-compensated operand "A side": [hi | lo | hi] along K
 *   split==2 (rep=3): error-compensated operand "B side": [hi | hi | lo] along K
 * so that A'.B'^T = hi.hi + lo.hi + hi.lo (parity mode, ~2^-16 relative error).
 * ld_dst in elements. */
int tt_cast_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, int rows,
                 int cols, int transpose, int split, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H_ */
