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

#define SFCVIT_ABI_VERSION 2

typedef struct CUstream_st* sfc_stream_t; /* == cudaStream_t */

/* ---- generic ---- */
const char* sfc_last_error(void);
int sfc_abi_version(void);
int sfc_device_sm_count(void);
/* Dropout under CUDA graphs: kernels that draw dropout masks (GEMM epilogue, LayerNorm backward, activation backward,
 * attention) add *epoch_dev (a device-resident 64-bit counter, may be NULL = 0) to their seed at run time, so a captured
 * step draws fresh masks on every replay when the graph increments the counter. Process-wide; set before capture. */
void sfc_set_dropout_epoch_ptr(const void* epoch_dev);

/* ---- K1: curve permutation (src/curves/space_filling_curves.py:74-251 generators,
 *      :458-491 grid_size/embed_and_prune_sfc; tokenizer flat index multi_hilbert.py:68-72) ----
 * curve_id: 0 hilbert_curve, 1 z_curve (Morton), 2 peano_curve, 3 moore_curve, 4 raster (identity).
 * perm[t] = i*h + j of the t-th in-domain cell along the curve on the w x h grid (i = row < w, j = col < h);
 * inv[i*h + j] = t (may be NULL). Both int32, w*h entries. */
size_t sfc_curve_perm_scratch_bytes(int curve_id, int w, int h);
int sfc_curve_perm(int curve_id, int w, int h, int32_t* perm_dev, int32_t* inv_dev, void* scratch_dev,
                   size_t scratch_bytes, sfc_stream_t stream);

/* ---- host-side curve utilities (C++, no device work; init-time):
 *      block_stitch_sfc (src/curves/space_filling_curves.py:513-591) and find_hamiltonian_path /
 *      refine_curve_to_hamiltonian (:273-455) ----
 * sfc_block_stitch: greedy power-of-base block decomposition of the width x height grid, per block the best of the 8
 * symmetries; writes (i, j) pairs in stitched order and the points per block; returns the point count.
 * sfc_hamiltonian_path: DFS with forced-move and flood-fill pruning; priority[i * height + j] = visiting priority
 * (lower first; width * height for cells the guiding curve does not contain) or NULL; returns width * height on success,
 * 0 if no path exists / the budget of max_steps expansions (0 = unlimited) ran out. */
int sfc_block_stitch(int curve_id, int width, int height, int32_t* out_ij, int cap_pairs, int32_t* block_len, int block_cap,
                     int* n_blocks);
int sfc_hamiltonian_path(int width, int height, const int32_t* priority, int diag, long long max_steps, int32_t* out_ij);

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
  void* colsum_out;      /* optional [M]: colsum_out[m] = sum_k A[m, k] — for a weight gradient dW = dY^T . X this is the bias
                            gradient sum_tokens dY (reference: autograd of every nn.Linear bias), computed as one more
                            accumulator column of the kernel that already streams dY. Needs the split count of
                            sfc_gemm_colsum_splits (> 0) and a workspace of sfc_gemm_workspace_bytes. */
  int colsum_fp32;       /* 0: colsum_out is bf16, 1: fp32 */
} SfcGemmEpilogue;

size_t sfc_gemm_workspace_bytes(int M, int N, int K, int splits);
int sfc_gemm_suggest_splits(int M, int N, int K);
int sfc_gemm_colsum_splits(int M, int N, int K, int a_mn_major, int b_mn_major);   /* 0 = not covered: use sfc_colsum */
int sfc_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, int M,
                  int N, int K, const SfcGemmEpilogue* ep, void* workspace, size_t workspace_bytes, int splits,
                  sfc_stream_t stream);

/* ---- K5: LayerNorm forward / backward and column sums (nn.LayerNorm in the encoder layers, vit.py:197-206 ->
 *      torch TransformerEncoderLayer.norm1/norm2; vit.py:253-254, :303; bias gradients of every nn.Linear) ----
 * x, y, dy, dx, gamma, beta: bf16; mean/rstd: fp32 [rows]; D % 8 == 0, D <= 2048.
 * dgamma/dbeta: bf16 (param_fp32 = 0) or fp32 (1); accumulate != 0 adds to the existing value. */
int sfc_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                      long long rows, int D, float eps, sfc_stream_t stream);
size_t sfc_layernorm_bwd_scratch_bytes(long long rows, int D);
/* optional fused extras: dx_drop = dropout-masked copy of dx (mask of a [rows, D] GEMM output drawn with drop_seed),
 * dcolsum[D] = column sums of dx_drop (of dx when dx_drop is NULL) = bias gradient of the preceding Linear. */
int sfc_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma, void* dx,
                      void* dx_drop, float drop_p, unsigned long long drop_seed, void* dgamma, void* dbeta,
                      void* dcolsum, int param_fp32, int accumulate, void* scratch, size_t scratch_bytes, long long rows,
                      int D, sfc_stream_t stream);
/* out = alpha * dy * f'(aux) * keep(seed, i) / (1 - drop_p): aux_mode SFC_AUX_NONE, SFC_AUX_RELU_MASK (aux > 0) or
 * SFC_AUX_GELU_GRAD (aux = pre-activation); the dropout mask is the one the GEMM epilogue drew for a contiguous
 * [M, N] output with the same seed. bf16, n % 8 == 0. */
int sfc_act_bwd(const void* dy, const void* aux, void* out, long long n, int aux_mode, float alpha, float drop_p,
                unsigned long long drop_seed, sfc_stream_t stream);
size_t sfc_colsum_scratch_bytes(long long rows, int N);
int sfc_colsum(const void* x, long long ld, long long rows, int N, void* out, int out_fp32, int accumulate,
               void* scratch, size_t scratch_bytes, sfc_stream_t stream);

/* ---- K7: linear resampling of a token stream into a column slice of the hierarchical tokenizers' concat buffer
 *      (tokenizers/multiscale/multi_hilbert.py:30-40 HierarchicalHilbertEmbedding.forward and its copies:
 *       F.interpolate(mode="linear", align_corners=False) along the token axis + torch.cat(dim=-1)) ----
 * src: bf16 [B, Ns, D] with row stride ld_src; dst row b*Nd + t has stride ld_dst (the caller pre-offsets dst to
 * the level's column slice). Ns == Nd is a strided copy. bwd is the transposed operator: dsrc = R^T ddst.
 * D, ld_src, ld_dst multiples of 8 elements, pointers 16-byte aligned. */
int sfc_interp_concat_fwd(const void* src, long long ld_src, int B, int Ns, int D, void* dst, long long ld_dst, int Nd,
                          sfc_stream_t stream);
int sfc_interp_concat_bwd(const void* ddst, long long ld_dst, int B, int Nd, int D, void* dsrc, long long ld_src, int Ns,
                          sfc_stream_t stream);

/* ---- K2: fused curve-order patch gather + patch-embedding GEMM
 *      (tokenizers: multiscale/multi_hilbert.py:74-84 SFCEmbedding1D.forward and its morton/peano/moore copies,
 *       _1D/hilbert_embedding1D.py:30-43, _2D/hilbert_embedding.py:80-91, _2D/zigzag_embedding.py:24-30) ----
 * img: NCHW fp32 (img_bf16 = SFC_IMG_F32_NCHW = 0), NCHW bf16 (SFC_IMG_BF16_NCHW = 1) or decoded image bytes, uint8 NHWC
 * (SFC_IMG_U8_NHWC = 2: the stage before the path, main.py:174-178 ToDtype + Normalize, is then folded into Wk / bias by
 * the caller; K axis of Wk ordered (q, p1, p2, c) — the reference's own order — and (p*C) % 8 == 0 required).
 * p = pre-patch size, g = group size, perm = int32 [n_perm] flat
 * pre-patch indices r*(W/p)+c in curve order (n_perm <= (H/p)*(W/p), tokens per image = n_perm / g). Wk: bf16 [D, Kpad], K axis ordered (q, c, p1, p2) and zero padded to
 * Kpad = sfc_patch_embed_kpad(C,p,g). out row of token t of image b: b*rows_per_img + tok_off + t, row stride ld_out
 * (the caller may pre-offset `out` to write a column slice of a wider matrix). pos: optional bf16 [ntok, ld_pos]. */
#define SFC_IMG_F32_NCHW 0
#define SFC_IMG_BF16_NCHW 1
#define SFC_IMG_U8_NHWC 2
int sfc_patch_embed_kpad(int C, int p, int g);
int sfc_patch_embed_fwd(const void* img, int img_bf16, int B, int C, int H, int W, int p, int g, const int32_t* perm,
                        int n_perm, const void* Wk, const void* bias, const void* pos, long long ld_pos, void* out, long long ld_out,
                        int D, int rows_per_img, int tok_off, sfc_stream_t stream);
/* backward helper: A[M, Kpad] bf16 = curve-ordered im2col (same K order); dWk = dOut^T . A via sfc_gemm_bf16 */
int sfc_patch_gather(const void* img, int img_bf16, int B, int C, int H, int W, int p, int g, const int32_t* perm,
                     int n_perm, void* A, sfc_stream_t stream);

/* ---- K4: flash attention forward / backward, head_dim 64, non-causal
 *      (F.scaled_dot_product_attention inside nn.MultiheadAttention: vit.py:197-206 -> torch
 *       TransformerEncoderLayer._sa_block; altvit.py:129-141) ----
 * qkv: bf16 [B*N, 3*D] packed in-projection output (q | k | v, head h = columns h*64..h*64+63 of each third);
 * out: bf16 [B*N, D]; lse: fp32 [B, H, N] (natural log of the softmax denominator, for backward);
 * dqkv: bf16 [B*N, 3*D]; scratch (backward): fp32 [B*N, D] dQ accumulator. drop_p: attention-probability dropout. */
int sfc_attn_fwd(const void* qkv, void* out, float* lse, int B, int H, int N, int D, float scale, float drop_p,
                 unsigned long long drop_seed, sfc_stream_t stream);
size_t sfc_attn_bwd_scratch_bytes(int B, int N, int D);
int sfc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* scratch,
                 size_t scratch_bytes, int B, int H, int N, int D, float scale, float drop_p,
                 unsigned long long drop_seed, sfc_stream_t stream);

/* ---- soft-target cross entropy (SoftTargetCrossEntropy: main.py:45-51,
 *      loss = -(targets * log_softmax(inputs.float(), -1)).sum(-1).mean(), called from src/training/train.py:158-160) ----
 * logits: bf16 or fp32 [B, ld_logits]; targets: fp32 [B, ld_targets] (soft labels, rows need not sum to 1);
 * loss: fp32 [1] = batch mean, reduced in a fixed order; row_lse / row_tsum: fp32 [B], saved for backward.
 * scratch: sfc_softce_scratch_bytes() bytes, zeroed once by the caller, private to the stream.
 * backward: dlogits[b, c] = dloss[0] / B * (softmax(logits)[b, c] * row_tsum[b] - targets[b, c]), in the logits' dtype. */
size_t sfc_softce_scratch_bytes(void);
int sfc_softce_fwd(const void* logits, int logits_fp32, long long ld_logits, const float* targets, long long ld_targets, int B,
                   int C, float* loss, float* row_lse, float* row_tsum, void* scratch, size_t scratch_bytes, sfc_stream_t stream);
int sfc_softce_bwd(const void* logits, int logits_fp32, long long ld_logits, const float* targets, long long ld_targets,
                   const float* row_lse, const float* row_tsum, const float* dloss, int B, int C, void* dlogits,
                   long long ld_dlogits, sfc_stream_t stream);

/* ---- K6: fused grad-norm / clip / AdamW over flat buckets
 *      (torch.nn.utils.clip_grad_norm_ + optim.AdamW.step: src/training/train.py:165-166, main.py:288-289) ----
 * sfc_grad_sumsq: accum[0] += sum(g^2) (caller zeroes accum), reduced in a fixed order (bit-identical on every rank of
 * a data-parallel job); scratch = sfc_grad_sumsq_scratch_bytes() bytes zeroed once by the caller, private to the stream.
 * sfc_adamw_step: decoupled weight decay Adam on n elements; gradient used =
 * g * grad_scale * min(1, max_norm / (sqrt(stats[0]) * grad_scale + 1e-6))
 * (max_norm <= 0 or stats == NULL: no clipping). p/g: bf16 or fp32 (param_fp32); m/v: bf16 or fp32 (state_fp32).
 * hyper_dev: optional DEVICE block of 3 floats {lr, 1 - beta1^t, sqrt(1 - beta2^t)} that overrides lr / step at run
 * time, so a step captured in a CUDA graph follows the host's scheduler (main.py:290-314) without re-capture. */
size_t sfc_grad_sumsq_scratch_bytes(void);
int sfc_grad_sumsq(const void* g, int g_fp32, long long n, float* accum, void* scratch, size_t scratch_bytes,
                   sfc_stream_t stream);
int sfc_adamw_step(void* p, const void* g, void* m, void* v, long long n, int param_fp32, int state_fp32, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, float max_norm,
                   const float* stats, const float* hyper_dev, sfc_stream_t stream);
/* dst[0..3] = {a, b, c, d} in stream order (values passed as kernel arguments): feeds hyper_dev between graph replays */
int sfc_store_f32x4(float* dst, float a, float b, float c, float d, sfc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFCVIT_H_ */
