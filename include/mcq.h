/*
 * mcq.h -- C ABI of libmcq.so: B200 (sm_100a) kernels for the additive multi-codebook
 * encode / refine / decode hot path of danpovey/quantization.
 *
 * The reference is pure Python/PyTorch and has no FFI layer; its "operator interface" for
 * this path is the body of four Python methods (paths relative to
 * /root/reference/quantization/quantization.py):
 *
 *   mcq_prepare          <- Quantizer.get_centers (:77-79), the exp() scale of Quantizer._logits (:278),
 *                           all_centers_sumsq (:411); run once per parameter version
 *   mcq_encode           <- Quantizer.encode (:244-275) = _compute_indexes (:281-305) + byte packing (:266-272)
 *   mcq_refine           <- `iters` x Quantizer._refine_indexes (:308-547) from caller-supplied indexes
 *                           (what compute_loss / QuantizerTrainer.step reach, :212, :652)
 *   mcq_decode           <- Quantizer.decode (:117-148) incl. _maybe_separate_indexes (:551-573)
 *   mcq_decode_backward  <- autograd of decode w.r.t. the scaled centers (used by compute_loss, :213-216)
 *   mcq_class_loss_*     <- the log-softmax / chosen-logprob / mean-probability part of compute_loss (:218-240)
 *   mcq_encode_host      <- the same encode for a caller that holds HOST buffers (what a cgo/JNI/ctypes
 *                           host without its own CUDA plumbing binds);
 *                           mcq_encode_host_ws is its re-entrant form (caller-owned staging buffer and stream)
 *
 * Conventions: plain pointers and sizes only.  Every function returns 0 on success or a negative
 * MCQ_E* code; mcq_last_error() gives the thread's last message.  Unless the name ends in _host,
 * all data pointers are DEVICE pointers, nothing is allocated or freed, nothing synchronises: all
 * work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 */
#ifndef MCQ_H_
#define MCQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCQ_OK 0
#define MCQ_EINVAL -1       /* bad argument (shape not a power of two, null pointer, too-small buffer...) */
#define MCQ_EUNSUPPORTED -2 /* K < 16 with N > 1 (the reference itself raises), K > 256, N > 64 */
#define MCQ_ECUDA -3        /* a CUDA call failed; see mcq_last_error() */
#define MCQ_ENODEVICE -4    /* no sm_100 device: there is no CPU fallback */

/* element types of x / decode output */
#define MCQ_F32 0
#define MCQ_F16 1
#define MCQ_BF16 2
/* element types of index arrays */
#define MCQ_U8 0  /* byte-packed exactly like Quantizer.encode(as_bytes=True) (:266-272) */
#define MCQ_I64 1 /* one int64 per codebook, like as_bytes=False */
#define MCQ_I32 2

int mcq_version(void);
const char *mcq_last_error(void);

/* Number of uint8 columns Quantizer.encode(as_bytes=True) produces (:266-271): N halves while K*K <= 256. */
int mcq_packed_cols(int num_codebooks, int codebook_size);

/* Size in bytes of the prepared blob for a (N, K, D) quantizer. */
size_t mcq_prepared_bytes(int num_codebooks, int codebook_size, int dim);

/* Workspace needed by mcq_encode / mcq_refine for up to `max_frames` frames per call (larger batches
 * are processed in chunks that fit whatever workspace is given, as long as it holds at least
 * mcq_workspace_bytes(1024, ...)). */
size_t mcq_workspace_bytes(int64_t max_frames, int dim, int num_codebooks, int codebook_size);

/*
 * Builds everything the kernels need from the raw parameters (all fp32, device):
 *   centers (N,K,D); centers_scale, logits_scale: 1-element tensors (raw parameters, exp() applied here
 *   as exp(p * scale_speed), :78, :278); to_logits_weight (N*K, D); to_logits_bias (N*K).
 * Must be called again whenever a parameter changes.
 */
int mcq_prepare(const float *centers, const float *centers_scale, const float *to_logits_weight,
                const float *to_logits_bias, const float *logits_scale, float scale_speed, int num_codebooks,
                int codebook_size, int dim, void *prepared, size_t prepared_bytes, void *stream);

/* Quantizer.encode: x (B, D) of x_dtype -> codes.  codes_dtype MCQ_U8: (B, mcq_packed_cols) bytes;
 * MCQ_I64 / MCQ_I32: (B, N).  iters = refine_indexes_iters (reference default 5). */
int mcq_encode(const void *x, int x_dtype, int64_t num_frames, int dim, int num_codebooks, int codebook_size,
               const void *prepared, int iters, void *codes, int codes_dtype, void *workspace,
               size_t workspace_bytes, void *stream);

/* `iters` passes of Quantizer._refine_indexes starting from idx_in (B, N) int64 -> idx_out (B, N) int64
 * (may alias idx_in). */
int mcq_refine(const void *x, int x_dtype, int64_t num_frames, int dim, int num_codebooks, int codebook_size,
               const void *prepared, int iters, const int64_t *idx_in, int64_t *idx_out, void *workspace,
               size_t workspace_bytes, void *stream);

/* Quantizer.decode: codes (B, ncols) of codes_dtype (ncols == N, or a packed column count dividing N)
 * -> out (B, D) of out_dtype.  Sums the selected scaled centers in codebook order n = 0..N-1 in fp32. */
int mcq_decode(const void *codes, int codes_dtype, int64_t num_frames, int ncols, int num_codebooks,
               int codebook_size, int dim, const void *prepared, void *out, int out_dtype, void *stream);

/* Same as mcq_decode, but gathers from a caller-supplied scaled-centers tensor (N, K, D) fp32 instead of the
 * prepared blob (the training path: the tensor autograd differentiates, quantization.py:141). */
int mcq_decode_centers(const void *codes, int codes_dtype, int64_t num_frames, int ncols, int num_codebooks,
                       int codebook_size, int dim, const float *scaled_centers, void *out, int out_dtype,
                       void *stream);

/* Gradient of decode w.r.t. the SCALED centers: grad_scaled_centers[n, idx[b,n], :] += grad_out[b, :].
 * grad_scaled_centers (N,K,D) fp32 must be zeroed by the caller.  idx (B, N) int64. */
int mcq_decode_backward(const float *grad_out, const int64_t *idx, int64_t num_frames, int num_codebooks,
                        int codebook_size, int dim, float *grad_scaled_centers, void *stream);

/*
 * Classifier-side losses of Quantizer.compute_loss (quantization.py:218-240) without its (B, N, K) intermediates.
 * Forward: xw (Bp, N*K) fp32 receives fl(exp(logits_scale*speed) * x) . W^T (the bias-free part of
 * Quantizer._logits, :277-279, from the tcgen05 GEMM; Bp = B rounded up to 128 rows, caller-allocated, kept for
 * the backward pass);  *logprob_sum = sum_{b,n} log_softmax(xw + bias)[b, n, idx[b,n]]  (:221-225 before the mean);
 * prob_sum (N*K) = sum_b softmax(xw + bias)[b, n, k]  (:235 before the mean).  idx (B, N) int64.
 * Backward: grad_logits (B, N*K) = d loss / d logits given g_logprob_sum (1 element) and g_prob_sum (N*K), both
 * DEVICE pointers (no host synchronisation).  part_gx (mcq_class_loss_partials() floats) receives partial sums of
 * grad_logits * xw: their total times scale_speed is d loss / d logits_scale (logits = exp(ls*speed) x W^T + b, so
 * no second GEMM is needed for it).  The weight / bias gradients are products of grad_logits the host layer forms
 * (as the reference's autograd does).
 */
int mcq_class_loss_forward(const void *x, int x_dtype, int64_t num_frames, int dim, int num_codebooks,
                           int codebook_size, const void *prepared, const int64_t *idx, float *xw,
                           float *logprob_sum, float *prob_sum, void *workspace, size_t workspace_bytes,
                           void *stream);
int mcq_class_loss_partials(void);
int mcq_class_loss_backward(const float *xw, int64_t num_frames, int dim, int num_codebooks, int codebook_size,
                            const void *prepared, const int64_t *idx, const float *g_logprob_sum,
                            const float *g_prob_sum, float *grad_logits, float *part_gx, void *stream);

/*
 * The two small reductions left of Quantizer.compute_loss / its backward once the kernels above have run:
 * mcq_index_counts: counts (N*K floats) = per-codebook histogram of the chosen entries, the `counts.mean(dim=0) * B` of
 *   quantization.py:227-231 without the (B, N, K) one-hot tensor (idx (B, N) int64; entries outside [0, K) are not
 *   counted; scratch: N*K uint32; N*K <= 12,288).  Integer counting: the result is independent of the order.
 * mcq_column_sums: out[c] = sum_r x[r][c] for a contiguous (rows, cols) fp32 matrix -- the bias gradient
 *   grad_logits.sum(0) autograd forms for nn.Linear (backward of quantization.py:279); cols a multiple of 4; partials:
 *   mcq_column_sum_partials(cols) floats; fixed summation order (reproducible).
 */
int mcq_index_counts(const int64_t *idx, int64_t num_frames, int num_codebooks, int codebook_size, float *counts,
                     void *scratch, void *stream);
int mcq_column_sum_partials(int cols);
int mcq_column_sums(const float *x, int64_t rows, int cols, float *out, float *partials, void *stream);

/*
 * Reconstruction term of Quantizer.compute_loss (quantization.py:209-216) without its (B, dim) intermediates.
 * Forward: sums[0] = sum_{b,d} (x_hat - x)^2 with x_hat = decode(idx) (scaled centers summed n = 0..N-1), sums[1] =
 * sum_{b,d} (x - mean)^2 (mean (dim) fp32 = Quantizer.get_data_mean()); x (B, dim) fp32 / fp16 / bf16, idx (B, N) int64;
 * partials: mcq_recon_loss_partials() floats of scratch; fixed-order sums (reproducible).
 * Backward: grad_scaled_centers (N, K, dim) += coef[0] * (x_hat - x) scattered to the chosen centers (coef: DEVICE
 * scalar, = 2 * upstream gradient of sums[0]; zero the gradient first).  dim <= 1024.
 */
int mcq_recon_loss_partials(void);
int mcq_recon_loss_forward(const void *x, int x_dtype, const int64_t *idx, int64_t num_frames, int num_codebooks,
                           int codebook_size, int dim, const float *scaled_centers, const float *mean, float *sums,
                           float *partials, void *stream);
int mcq_recon_loss_backward(const void *x, int x_dtype, const int64_t *idx, int64_t num_frames, int num_codebooks,
                            int codebook_size, int dim, const float *scaled_centers, const float *coef,
                            float *grad_scaled_centers, void *stream);

/*
 * Weight-gradient product: out (c1, c2) = a^T . b, reduction over `rows` frames; a (rows, c1) fp32 with row stride lda,
 * b (rows, c2) fp32 / fp16 / bf16 with row stride ldb (elements).  This is what the reference's autograd computes with
 * an fp32 SGEMM for d loss / d to_logits.weight (backward of quantization.py:279) and for the linear1 / linear2 /
 * linear2b weights of JointCodebookLoss (backward of prediction.py:56, :70-76).  Here: transposing fp16x2 split of both
 * operands, split-K tcgen05 product with fp32 accumulation, fixed-order sum of the partial products (reproducible;
 * error <= 2^-18 sum_r |a_r b_r| per output, measured ~2^-20:
 * the class of an fp32 SGEMM over as many rows).  workspace: mcq_gemm_tn_workspace_bytes(rows, c1, c2) bytes.
 */
size_t mcq_gemm_tn_workspace_bytes(int64_t rows, int c1, int c2);
int mcq_gemm_tn(const float *a, int64_t lda, const void *b, int b_dtype, int64_t ldb, int64_t rows, int c1, int c2,
                float *out, void *workspace, size_t workspace_bytes, void *stream);

/*
 * out (m, n) = or += a . b^T: a (m, k) fp32 row stride lda, b (n, k) fp32 row stride ldb, out row stride ldc (a column
 * block of a wider matrix is fine).  The forward / input-gradient products of JointCodebookLoss (prediction.py:56,
 * :70-76 and their backward) -- fp32 SGEMMs in the reference -- as fp32-faithful tcgen05 products: per-row power-of-two
 * scaling, fp16x2 split, three products accumulated in fp32 (the scheme of the encode path).  n must be a multiple of
 * 64, ldc of 4, out 16-byte aligned (MCQ_EUNSUPPORTED otherwise: the host layer then uses a library GEMM).
 */
size_t mcq_gemm_nt_workspace_bytes(int64_t m, int n, int k);
int mcq_gemm_nt(const float *a, int64_t lda, const float *b, int64_t ldb, int64_t m, int n, int k, float *out,
                int64_t ldc, int accumulate, void *workspace, size_t workspace_bytes, void *stream);

/*
 * JointCodebookLoss (reference prediction.py:9-82, class :86-197): the stages between its dense products.
 * Layouts: hidden / grad_hidden (B, H) fp32 = linear1(predictor); codes (B, N) uint8 / int32 / int64 as produced by
 * mcq_encode (negative = padding where the type allows it); embedding ((N-1)*K, H) fp32 = codebook_embedding.weight;
 * act / grad_act (N, B, H) fp32: act[n, b, :] = relu(hidden[b] + sum_{m<n} scale * embedding[m*K + max(codes[b,m],0)])
 * (prediction.py:47-68: embedding gather, concat, cumsum over codebooks, ReLU -- in the reference's summation order),
 * stored codebook-major because it is the left operand of the per-codebook product with linear2_weight[n] (:70-72).
 *   mcq_jcl_hidden_forward    writes act.
 *   mcq_jcl_hidden_backward   given grad_act = d loss / d act: grad_hidden (overwritten) and grad_embedding
 *                             (ACCUMULATED into with atomics: zero it first).
 *   mcq_jcl_cross_entropy     logits (B, N, K) fp32 WITHOUT linear2_bias, bias (N, K): per-row cross entropy of
 *                             softmax(logits + bias) against codes (:79-82) -> row_loss (B, N), 0 where codes ==
 *                             ignore_index; sums[0] = total loss, sums[1] = number of rows counted (for reduction =
 *                             'mean').  With want_grad != 0 the logits are overwritten by d sums[0] / d logits
 *                             (softmax - onehot; 0 on ignored rows).  partials: mcq_jcl_partials() floats of scratch;
 *                             the sums are formed in a fixed order (bitwise reproducible).
 */
int mcq_jcl_hidden_forward(const float *hidden, const void *codes, int codes_dtype, int64_t num_frames,
                           int num_codebooks, int codebook_size, int hidden_channels, const float *embedding,
                           float scale, float *act, void *stream);
int mcq_jcl_hidden_backward(const float *grad_act, const float *act, const void *codes, int codes_dtype,
                            int64_t num_frames, int num_codebooks, int codebook_size, int hidden_channels, float scale,
                            float *grad_hidden, float *grad_embedding, void *stream);
int mcq_jcl_partials(void);
int mcq_jcl_cross_entropy(float *logits, const float *bias, const void *codes, int codes_dtype, int64_t num_frames,
                          int num_codebooks, int codebook_size, int64_t ignore_index, int want_grad, float *row_loss,
                          float *sums, float *partials, void *stream);

/* Pointers into the prepared blob (device): scaled centers (N*K, D) fp32 and the Gram table (N*K, N*K). */
const float *mcq_prepared_scaled_centers(const void *prepared, int num_codebooks, int codebook_size, int dim);
const float *mcq_prepared_gram(const void *prepared, int num_codebooks, int codebook_size, int dim);

/* Stage-level entry points (used by the tests to check each kernel against the CPU model bit for bit):
 *   mcq_xct:    P (B, N*K) = x . scaled_centers^T
 *   mcq_search: `iters` passes of the table-driven refinement given P and the prepared Gram table. */
int mcq_xct(const void *x, int x_dtype, int64_t num_frames, int dim, int num_codebooks, int codebook_size,
            const void *prepared, float *P, void *workspace, size_t workspace_bytes, void *stream);
int mcq_search(const float *P, const float *gram, int64_t num_frames, int num_codebooks, int codebook_size,
               int iters, const int32_t *idx_in, int32_t *idx_out, void *stream);

/* Optional per-kernel timing for benchmarks: while enabled, every kernel the entry points launch is bracketed by
 * CUDA events on its launch stream.  mcq_profile(1) enables and resets, mcq_profile(0) disables.
 * mcq_profile_read synchronises on the recorded events and returns, per kind, total milliseconds and launch count. */
#define MCQ_PROF_OTHER 0  /* staging, arg-max, packing */
#define MCQ_PROF_GEMM 1   /* tcgen05 GEMMs (logits and P) */
#define MCQ_PROF_SEARCH 2 /* the refinement search kernel */
#define MCQ_PROF_DECODE 3
#define MCQ_PROF_KINDS 4
int mcq_profile(int enable);
int mcq_profile_read(double *ms_by_kind, int64_t *launches_by_kind);

/* Work actually done by the refinement search on a workspace: the search kernels add, into a counter block at the
 * start of the workspace they are given, the number of refinement passes they executed (a frame whose pass returns
 * its input has converged and skips the remaining passes -- exact) and the number of frames searched.
 * passes_frames (host, 2 x uint64, may be NULL) receives the totals since the last reset (synchronises `stream`);
 * reset != 0 zeroes them afterwards.  A fresh workspace holds garbage there: reset before relying on the totals.
 * Used by bench.py for the roofline accounting ("count only passes actually executed"). */
int mcq_search_stats(void *workspace, int reset, uint64_t *passes_frames, void *stream);

/*
 * Host-buffer encode: x_host (B, D) and codes_host live in HOST memory (pinned or pageable).  The library
 * stages chunks through its own pinned buffers and device workspace on `device`, overlapping H2D, compute
 * and D2H, and returns when codes_host is complete.  `prepared` is a DEVICE blob on that device.
 */
int mcq_encode_host(const void *x_host, int x_dtype, int64_t num_frames, int dim, int num_codebooks,
                    int codebook_size, const void *prepared, int iters, void *codes_host, int codes_dtype,
                    int device);

/*
 * The re-entrant form of the host-buffer encode (what a multi-threaded host binds): the CALLER owns the device staging
 * buffer (`staging`, at least mcq_encode_host_ws_bytes(...) bytes on the current device; a smaller buffer only shrinks
 * the chunks) and the stream.  Nothing is allocated, no state survives the call, the host is not blocked: the copies
 * and kernels are ordered after everything already on `stream`, and `stream` completes when codes_host is complete
 * (synchronise it, or record an event on it, before reading codes_host or reusing x_host / staging).  x_host and
 * codes_host should be pinned (cudaHostAlloc / cudaHostRegister) for the copies to overlap the kernels.
 * Replaces the same call sites as mcq_encode_host: Quantizer.encode (quantization.py:244-275) on host tensors.
 */
size_t mcq_encode_host_ws_bytes(int64_t num_frames, int dim, int num_codebooks, int codebook_size, int x_dtype,
                                int codes_dtype);
int mcq_encode_host_ws(const void *x_host, int x_dtype, int64_t num_frames, int dim, int num_codebooks,
                       int codebook_size, const void *prepared, int iters, void *codes_host, int codes_dtype,
                       void *staging, size_t staging_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MCQ_H_ */
