/*
 * libsfcvit — C ABI of the B200-native hot path for
 * RemcoHoger/Space-Filling-Curves-for-Vision-Transformers.
 *
 * The reference has no FFI layer: its boundary is the Python module API
 * (src/curves, src/tokenizers, src/models, src/training). This header is the plain-C
 * boundary that the Python mirror of that API binds with ctypes; every entry point cites the
 * reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - all pointers are DEVICE pointers owned by the caller (PyTorch tensors); the library never
 *    allocates in the hot path and never synchronises the host;
 *  - every call takes the caller's cudaStream_t (as void*) and is asynchronous;
 *  - return value 0 = ok, non-zero = error, message via sfc_last_error() (thread-local);
 *  - bf16 = raw 16-bit bfloat16; "rows x cols, ld" = row-major with leading dimension in elements.
 */
#ifndef SFCVIT_H_
#define SFCVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFCVIT_ABI_VERSION 1

typedef struct CUstream_st* sfc_stream_t; /* == cudaStream_t */

/* ---- generic ---- */
const char* sfc_last_error(void);
int sfc_abi_version(void);
int sfc_device_sm_count(void);

/* ---- K1: curve permutation (src/curves/space_filling_curves.py:74-251 generators,
 *      :458-491 grid_size/embed_and_prune_sfc; tokenizer flat index multi_hilbert.py:68-72) ----
 * curve_id: 0 hilbert_curve, 1 z_curve (Morton), 2 peano_curve, 3 moore_curve, 4 raster (identity).
 * perm[t] = i*h + j of the t-th in-domain cell along the curve on the w x h grid (i = row < w, j = col < h);
 * inv[i*h + j] = t (may be NULL). Both int32, w*h entries. */
size_t sfc_curve_perm_scratch_bytes(int curve_id, int w, int h);
int sfc_curve_perm(int curve_id, int w, int h, int32_t* perm_dev, int32_t* inv_dev, void* scratch_dev,
                   size_t scratch_bytes, sfc_stream_t stream);

/* ---- K3: bf16 GEMM family with fused epilogue (nn.Linear / MHA projections / einsum:
 *      src/models/vit.py:197-206, 262-266, 289-292; tokenizer proj multi_hilbert.py:66,84) ----
 * D[M,N] = epilogue( alpha * A . B^T )
 *   a_mn_major = 0: A is [M rows][K cols] (K contiguous), lda = row stride
 *   a_mn_major = 1: A is [K rows][M cols] (M contiguous), lda = row stride        (transposed operand, read in place)
 *   b_mn_major = 0: B is [N rows][K cols];  b_mn_major = 1: B is [K rows][N cols]
 * epilogue order: +bias[n] -> (store out_pre) -> act -> dropout -> aux (mask / gelu') -> +residual[m,n] -> store. */
enum { SFC_ACT_NONE = 0, SFC_ACT_RELU = 1, SFC_ACT_GELU = 2 };
enum { SFC_AUX_NONE = 0, SFC_AUX_RELU_MASK = 1, SFC_AUX_GELU_GRAD = 2 };

typedef struct SfcGemmEpilogue {
  const void* bias;      /* bf16 [N] or NULL */
  const void* residual;  /* bf16 [M, ld_res] or NULL */
  const void* aux;       /* bf16 [M, ld_aux]: relu mask source (aux > 0) or gelu pre-activation */
  void* out;             /* bf16 or fp32 [M, ld_out] */
  void* out_pre;         /* optional bf16 [M, ld_out]: value before activation */
  long long ld_out, ld_res, ld_aux;
  float alpha;
  int act;               /* SFC_ACT_* */
  int aux_mode;          /* SFC_AUX_* */
  int out_fp32;          /* 0: out is bf16, 1: out is fp32 */
  int accumulate;        /* split-K only: out += result */
  float drop_p;          /* dropout probability applied after act (0 = off) */
  unsigned long long drop_seed;
} SfcGemmEpilogue;

size_t sfc_gemm_workspace_bytes(int M, int N, int K, int splits);
int sfc_gemm_suggest_splits(int M, int N, int K);
int sfc_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M,
                  int N, int K, const SfcGemmEpilogue* ep, void* workspace, size_t workspace_bytes, int splits,
                  sfc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFCVIT_H_ */
