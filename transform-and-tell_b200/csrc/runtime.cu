// Error plumbing, launch counter and driver-entry-point lookup shared by all ops.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "runtime.h"

namespace tt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static const unsigned long long* g_step_ptr = nullptr;
const unsigned long long* rng_step_ptr() { return g_step_ptr; }

__global__ void rng_step_advance_kernel(unsigned long long* p) {
  pdl_prologue(); *p += 1ull; }

// Programmatic dependent launch of every kernel of the library (each kernel starts with
// griddepcontrol.launch_dependents + griddepcontrol.wait: the NEXT kernel's CTAs are scheduled while this
// one still runs and wait at their first instruction).  tt_set_pdl / env TT_PDL; off by default.
// Measured on the captured train step: a single latency-bound chain gains (decoder graph 5.05 -> 4.88 ms),
// a graph with parallel branches loses (frozen encoders 6.45 -> 7.3 ms: early-resident waiting CTAs take
// the SMs of the other branch) -- callers switch it on around the capture of single-chain graphs only.
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("TT_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pdl == 1;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TT_ERR_CUDA;
  }
  count_launch(1);
  return TT_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    n = p.multiProcessorCount;
  }
  return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows) {
  return make_tmap_bf16_2d_sw(m, ptr, inner, rows, ld_elems, box_inner, box_rows, 128);
}

int make_tmap_bf16_2d_sw(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows,
                         uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return TT_ERR_NO_DEVICE;
  }
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu rows=%llu ld=%llu box=%ux%u",
              (int)r, ptr, (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)ld_elems, box_inner, box_rows);
    return TT_ERR_CUDA;
  }
  return TT_OK;
}

}  // namespace tt

extern "C" {
const char* tt_last_error(void) { return tt::g_err; }
int tt_abi_version(void) { return 1; }
long long tt_launch_count(void) { return tt::g_launches.load(); }
void tt_reset_launch_count(void) { tt::g_launches.store(0); }
void tt_set_rng_step_ptr(const unsigned long long* dev_ptr) { tt::g_step_ptr = dev_ptr; }
int tt_rng_step_advance(unsigned long long* dev_ptr, void* stream) {
  if (!dev_ptr) { tt::set_error("tt_rng_step_advance: null pointer"); return TT_ERR_INVALID; }
  tt::launch_k(tt::rng_step_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, dev_ptr);
  return tt::check_launch("rng_step_advance_kernel");
}
}

extern "C" int tt_set_pdl(int on) {
  const int prev = tt::pdl_enabled() ? 1 : 0;
  tt::set_pdl(on);
  return prev;
}
