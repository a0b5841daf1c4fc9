// Token embedding front end of the decoder:
//   AdaptiveEmbedding  (tell/modules/token_embedders/adaptive.py:61-76): 3 bands of
//     Embedding(padding_idx=0) -> Linear(E->E, no bias), summed (bands are disjoint) and scaled;
//   SinusoidalPositionalEmbedding (positional.py:167-211, make_positions :231-268).
// The band projections are ONE GEMM: this file gathers each token's embedding row into its band's
// slot of a [N, n_bands*E] operand (zeros elsewhere) so that  out = A . [W_0 | W_1 | W_2]^T.
// Backward scatters dA rows back into the (tied, dense) embedding-table gradients.
#include "common.cuh"
#include "runtime.h"

namespace tt {

constexpr int EM_MAX_BANDS = 8;
struct EmbedBands {
  int n_bands;
  int cutoff[EM_MAX_BANDS];          // upper id bound of each band
  const float* table[EM_MAX_BANDS];  // [cutoff[i]-cutoff[i-1], E]
  float* grad[EM_MAX_BANDS];
};

// ids [B,T] (row-major).  Output row n = t*B + b when tbc != 0, else b*T + t.
__global__ void embed_gather_kernel(const long long* __restrict__ ids, int B, int T, int tbc,
                                    EmbedBands bands, int E4, float* __restrict__ out) {
  pdl_prologue();
  const int n = blockIdx.x;
  const int b = tbc ? n % B : n / T, t = tbc ? n / B : n % T;
  const long long id = ids[static_cast<long long>(b) * T + t];
  int band = -1, local = 0;
  for (int i = 0; i < bands.n_bands; ++i) {
    const int lo = i ? bands.cutoff[i - 1] : 0;
    if (id >= lo && id < bands.cutoff[i]) { band = i; local = static_cast<int>(id - lo); }
  }
  float4* o = reinterpret_cast<float4*>(out) + static_cast<long long>(n) * bands.n_bands * E4;
  for (int i = 0; i < bands.n_bands; ++i) {
    const float4* src = (i == band)
        ? reinterpret_cast<const float4*>(bands.table[i]) + static_cast<long long>(local) * E4
        : nullptr;
    for (int c = threadIdx.x; c < E4; c += blockDim.x)
      o[i * E4 + c] = src ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// grad_table[band][local,:] += dA[n, band*E : (band+1)*E]   (skips nn.Embedding's padding_idx row)
__global__ void embed_scatter_kernel(const long long* __restrict__ ids, int B, int T, int tbc,
                                     EmbedBands bands, int E, int padding_idx,
                                     const float* __restrict__ dA) {
  pdl_prologue();
  const int n = blockIdx.x;
  const int b = tbc ? n % B : n / T, t = tbc ? n / B : n % T;
  const long long id = ids[static_cast<long long>(b) * T + t];
  int band = -1, local = 0;
  for (int i = 0; i < bands.n_bands; ++i) {
    const int lo = i ? bands.cutoff[i - 1] : 0;
    if (id >= lo && id < bands.cutoff[i]) { band = i; local = static_cast<int>(id - lo); }
  }
  if (band < 0 || local == padding_idx || bands.grad[band] == nullptr) return;
  const float* src = dA + (static_cast<long long>(n) * bands.n_bands + band) * E;
  float* dst = bands.grad[band] + static_cast<long long>(local) * E;
  for (int c = threadIdx.x; c < E; c += blockDim.x) atomicAdd(dst + c, src[c]);
}

// positional.py:231-268: non-pad symbols -> pad+1+index (+start_pos), pads stay pad.
// left_pad shifts positions so the last real token sits at the right edge.
__global__ void make_positions_kernel(const long long* __restrict__ ids, int B, int T, int pad,
                                      int left_pad, int start_pos, const int* __restrict__ start_dev,
                                      int tbc, int* __restrict__ pos) {
  pdl_prologue();
  if (start_dev != nullptr) start_pos += *start_dev;   // running position kept on the device
  const int b = blockIdx.x;
  __shared__ int nonpad;
  if (threadIdx.x == 0) nonpad = 0;
  __syncthreads();
  int c = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) c += ids[static_cast<long long>(b) * T + t] != pad;
  atomicAdd(&nonpad, c);
  __syncthreads();
  const int offset = left_pad ? (T - nonpad) : 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const bool real = ids[static_cast<long long>(b) * T + t] != pad;
    const int p = real ? (pad + 1 + t - offset + start_pos) : pad;
    pos[tbc ? (t * B + b) : (b * T + t)] = p;
  }
}

// out[b,a,:] = in[a,b,:]   ([A,B,C] -> [B,A,C]); decoder_faces_objects.py:109,129 transposes.
__global__ void transpose01_kernel(const float* __restrict__ in, float* __restrict__ out, int A,
                                   int B, int C4) {
  pdl_prologue();
  const long long total = static_cast<long long>(A) * B * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long long ab = i / C4;
    const int bb = static_cast<int>(ab % B), aa = static_cast<int>(ab / B);
    reinterpret_cast<float4*>(out)[(static_cast<long long>(bb) * A + aa) * C4 + c] =
        __ldg(reinterpret_cast<const float4*>(in) + i);
  }
}

static int make_bands(const int* cutoffs, int n_bands, const float* const* tables,
                      float* const* grads, EmbedBands* bd) {
  TT_REQUIRE(cutoffs && n_bands >= 1 && n_bands <= EM_MAX_BANDS, "embed: n_bands must be in [1,%d]",
             EM_MAX_BANDS);
  bd->n_bands = n_bands;
  for (int i = 0; i < n_bands; ++i) {
    bd->cutoff[i] = cutoffs[i];
    bd->table[i] = tables ? tables[i] : nullptr;
    bd->grad[i] = grads ? grads[i] : nullptr;
  }
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_embed_gather(const long long* ids, int B, int T, int tbc, const int* cutoffs,
                               int n_bands, const float* const* tables, int E, float* out,
                               void* stream) {
  TT_REQUIRE(ids && tables && out, "tt_embed_gather: null pointer");
  TT_REQUIRE(E % 4 == 0, "tt_embed_gather: E must be a multiple of 4");
  EmbedBands bd;
  int rc = make_bands(cutoffs, n_bands, tables, nullptr, &bd);
  if (rc != TT_OK) return rc;
  if (B * T <= 0) return TT_OK;
  launch_k(embed_gather_kernel, dim3(B * T), dim3(128), 0, (cudaStream_t)stream, ids, B, T, tbc, bd, E / 4, out);
  return check_launch("embed_gather_kernel");
}

extern "C" int tt_embed_scatter_grad(const long long* ids, int B, int T, int tbc,
                                     const int* cutoffs, int n_bands, float* const* grads, int E,
                                     int padding_idx, const float* dA, void* stream) {
  TT_REQUIRE(ids && grads && dA, "tt_embed_scatter_grad: null pointer");
  EmbedBands bd;
  int rc = make_bands(cutoffs, n_bands, nullptr, grads, &bd);
  if (rc != TT_OK) return rc;
  if (B * T <= 0) return TT_OK;
  launch_k(embed_scatter_kernel, dim3(B * T), dim3(128), 0, (cudaStream_t)stream, ids, B, T, tbc, bd, E, padding_idx,
                                                               dA);
  return check_launch("embed_scatter_kernel");
}

extern "C" int tt_make_positions(const long long* ids, int B, int T, int pad, int left_pad,
                                 int start_pos, int tbc, int* pos, void* stream) {
  TT_REQUIRE(ids && pos, "tt_make_positions: null pointer");
  if (B * T <= 0) return TT_OK;
  launch_k(make_positions_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, ids, B, T, pad, left_pad, start_pos,
           (const int*)nullptr, tbc, pos);
  return check_launch("make_positions_kernel");
}

extern "C" int tt_make_positions_at(const long long* ids, int B, int T, int pad, int left_pad,
                                    int start_pos, const int* start_dev, int tbc, int* pos,
                                    void* stream) {
  TT_REQUIRE(ids && pos, "tt_make_positions_at: null pointer");
  if (B * T <= 0) return TT_OK;
  launch_k(make_positions_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, ids, B, T, pad, left_pad, start_pos,
           start_dev, tbc, pos);
  return check_launch("make_positions_kernel");
}

extern "C" int tt_transpose01(const float* in, float* out, int A, int B, int C, void* stream) {
  TT_REQUIRE(in && out, "tt_transpose01: null pointer");
  TT_REQUIRE(C % 4 == 0, "tt_transpose01: C must be a multiple of 4");
  const long long total = static_cast<long long>(A) * B * (C / 4);
  if (total <= 0) return TT_OK;
  long long g = ceil_div_ll(total, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  launch_k(transpose01_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, in, out, A, B, C / 4);
  return check_launch("transpose01_kernel");
}

// ------------------------------------------------------------------------------------------------
// Batch collation: the reference pads every field on the host
// (allennlp TextField / ArrayField(padding_value=nan) as_tensor, roberta_indexer.py:185-200,
// nytimes_faces_ner_matched.py:213-217) and uploads one tensor per field.  Here the ragged rows of
// a field travel as ONE flat pinned buffer + row offsets and are padded on the device.
namespace tt {
// out[b, i, :] = i < n_b ? flat[(off[b] + i) * width + :] : fill     (n_b = off[b+1] - off[b])
template <typename T>
__global__ void pad_ragged_kernel(const T* __restrict__ flat, const long long* __restrict__ off,
                                  T* __restrict__ out, int B, int max_rows, int width, T fill) {
  pdl_prologue();
  const long long per = static_cast<long long>(max_rows) * width;
  const long long total = per * B;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    const long long r = i - static_cast<long long>(b) * per;
    const int row = static_cast<int>(r / width);
    const int c = static_cast<int>(r - static_cast<long long>(row) * width);
    const long long o0 = off[b], n = off[b + 1] - o0;
    out[i] = row < n ? flat[(o0 + row) * width + c] : fill;
  }
}
template <typename T>
static int pad_ragged(const T* flat, const long long* off, T* out, int B, int max_rows, int width,
                      T fill, void* stream, const char* what) {
  if (!(off && out) || (!flat && max_rows > 0 && width > 0)) {
    set_error("%s: null pointer", what);
    return TT_ERR_INVALID;
  }
  const long long total = static_cast<long long>(B) * max_rows * width;
  if (total <= 0) return TT_OK;
  long long g = ceil_div_ll(total, 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  launch_k(pad_ragged_kernel<T>, dim3(static_cast<unsigned>(g)), dim3(256), 0, (cudaStream_t)stream, flat,
           off, out, B, max_rows, width, fill);
  return check_launch(what);
}
}  // namespace tt

extern "C" int tt_pad_ragged_f32(const float* flat, const long long* row_offsets, float* out, int B,
                                 int max_rows, int width, float fill, void* stream) {
  return tt::pad_ragged<float>(flat, row_offsets, out, B, max_rows, width, fill, stream, "tt_pad_ragged_f32");
}
extern "C" int tt_pad_ragged_i64(const long long* flat, const long long* row_offsets, long long* out,
                                 int B, int max_rows, long long fill, void* stream) {
  return tt::pad_ragged<long long>(flat, row_offsets, out, B, max_rows, 1, fill, stream, "tt_pad_ragged_i64");
}
