// common.cuh -- shared declarations of libmcq.so (sm_100a only; there is no CPU fallback).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mcq.h"

namespace mcq {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define MCQ_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ::mcq::cuda_fail(e__, #call); \
    } while (0)

#define MCQ_LAUNCH_CHECK(what)                                        \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return ::mcq::cuda_fail(e__, what);   \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline bool is_pow2(long v) { return v > 0 && (v & (v - 1)) == 0; }

// Layout of the caller-owned prepared blob (device memory).  All offsets are multiples of 1024 bytes.
struct Prepared {
    int N, K, D, NK;
    int Dp;  // D rounded up to a multiple of 64: row length of the fp16 split operands
    size_t off_cs;     // float [NK*D]   scaled centers  exp(centers_scale*speed) * centers   (quantization.py:77-79)
    size_t off_w;      // float [NK*D]   to_logits.weight (copy)
    size_t off_bias;   // float [NK]     to_logits.bias (copy)
    size_t off_gram;   // float [NK*NK]  G = Cs Cs^T (fp64 accumulation, rounded once)
    size_t off_scal;   // float [4]      {centers scale, logits scale}
    size_t off_csplit; // half  [2][NKp*Dp] two-way fp16 split of the row-scaled cs (operands of the tcgen05 GEMM)
    size_t off_wsplit; // half  [2][NKp*Dp] two-way fp16 split of the row-scaled w
    size_t off_cscale; // float [NKp]  2^-e of each cs row (the GEMM epilogue multiplies by it)
    size_t off_wscale; // float [NKp]  ... of each w row
    size_t bytes;
};

Prepared prepared_layout(int N, int K, int D);

// Per-call workspace layout for a chunk of Bc frames.
struct Workspace {
    int64_t Bc;
    int Mp;            // Bc rounded up to 128
    size_t off_xf;     // float [Mp*D]      x as fp32 (identity for fp32 input is still copied: uniform path)
    size_t off_xsplit; // half  [2][Mp*Dp]  two-way fp16 split of the row-scaled x
    size_t off_lsplit; // half  [2][Mp*Dp]  ... of the row-scaled fl(logits_scale * x)
    size_t off_xscale; // float [Mp]  2^-e of each x row
    size_t off_lscale; // float [Mp]  ... of each fl(logits_scale * x) row
    size_t off_p;      // float [Mp*NK]     P = x Cs^T   (also receives the logits before P is formed)
    size_t off_idx;    // int32 [Mp*N]
    size_t off_ctr;    // uint32 [256]      counter block of the search kernel (always offset 0; see SEARCH_STAT_WORD)
    size_t bytes;
};

Workspace workspace_layout(int64_t Bc, int N, int K, int D);

int check_shape(int N, int K, int D);

// ---- kernel launchers (each enqueues on `st` and returns MCQ_OK or an error) -------------------------------------
int launch_prepare(const float *centers, const float *centers_scale, const float *w, const float *bias,
                   const float *logits_scale, float scale_speed, const Prepared &L, char *blob, cudaStream_t st);
// x (any dtype) -> fp32 copy, row-scaled fp16 splits of x and of fl(lscale * x)
int launch_split_x(const void *x, int x_dtype, int64_t B, const Prepared &L, const char *blob, const Workspace &W,
                   char *ws, bool want_logits_split, cudaStream_t st);
// C (M, NK) = A (M, D) . Bm (NK, D)^T  plain fp32 FFMA version (scaffolding / cross-check of the tcgen05 GEMM)
int launch_gemm_ffma(const float *A, const float *Bm, float *C, int64_t M, int NK, int D, const float *a_scale,
                     cudaStream_t st);
// fp16x2 tcgen05 GEMM: C (M, NK) fp32 = (a0 b0 + a0 b1 + a1 b0) * a_scale[row] * b_scale[col] of the split operands
int launch_gemm_tc(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale, float *C,
                   int64_t Mp, int NK, int Dp, cudaStream_t st);
// logits GEMM with the classifier arg-max fused into the epilogue (K a multiple of 128); scratch: Mp * NK/128 * 8 bytes
bool gemm_tc_argmax_supported(int NK, int K);
int launch_gemm_tc_splitk(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                          float *C, int64_t Mp, int NK, int Dp, int k_splits, cudaStream_t st);
int launch_gemm_tc_general(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                           float *C, int64_t ldc, int64_t m_valid, int64_t Mp, int NK, int Dp, int accumulate,
                           cudaStream_t st, int64_t a_plane_rows = 0);
int launch_gemm_tc_argmax(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                          int64_t Mp, int NK, int Dp, const float *bias, int64_t B, int N, int K, void *scratch,
                          int32_t *idx, cudaStream_t st);
int launch_argmax_init(const float *logits, const float *bias, int64_t B, int N, int K, int32_t *idx, cudaStream_t st);
// work_counter (optional, device, zeroed by the caller on `st`): lets the warps of the search kernel fetch frames
// dynamically instead of striding over the batch (frames take 2..iters passes, so static striding leaves a tail)
// Layout of the counter block `work_counter` points at (the first 1024 bytes of a workspace): word 0 = the frame
// counter of the current launch (zeroed by the caller per launch); bytes 8..23 = two uint64 running totals the search
// kernels add to -- refinement passes executed, frames searched -- which nobody resets (bench.py zeroes / reads them
// to count the passes that actually ran: converged frames stop early).
constexpr int SEARCH_STAT_WORD = 2;  // offset (in 32-bit words) of the two uint64 totals

#ifdef __CUDACC__
__device__ __forceinline__ void search_stats_add(unsigned *work_counter, unsigned npass, unsigned nframes, int lane) {
    if (work_counter != nullptr && lane == 0 && nframes != 0) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(work_counter + SEARCH_STAT_WORD);
        atomicAdd(st, (unsigned long long)npass);
        atomicAdd(st + 1, (unsigned long long)nframes);
    }
}
#endif

int launch_search(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
                  int32_t *idx_out, cudaStream_t st, unsigned *work_counter = nullptr);
// second version of the search (search2.cu): codebook_size 256, 2/4/8 codebooks
bool search2_supports(int N, int K);
int launch_search2(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
                   int32_t *idx_out, cudaStream_t st, unsigned *work_counter);
int64_t max_chunk_frames();
// codebook_size 16, 8 codebooks (trainer phase 1 at bytes_per_frame = 4): Gram table in shared memory, sub-warp selections
bool search_k16_supports(int N, int K);
int launch_search_k16(const float *P, const float *gram, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
                      cudaStream_t st, unsigned *work_counter);
int launch_pack(const int32_t *idx, int64_t B, int N, int K, void *codes, int codes_dtype, cudaStream_t st);
int launch_i64_to_i32(const int64_t *src, int32_t *dst, int64_t n, int K, cudaStream_t st);
int launch_i32_to_i64(const int32_t *src, int64_t *dst, int64_t n, cudaStream_t st);
int launch_decode(const void *codes, int codes_dtype, int64_t B, int ncols, int N, int K, int D, const float *cs,
                  void *out, int out_dtype, cudaStream_t st);
int launch_decode_backward(const float *grad_out, const int64_t *idx, int64_t B, int N, int K, int D, float *grad,
                           cudaStream_t st);

int class_loss_streams(int64_t B, int N, int K);
int launch_class_loss_fwd(const float *xw, const float *bias, const int64_t *idx, int64_t B, int N, int K,
                          float *part_prob, float *part_lp, float *prob_sum, float *logprob_sum, cudaStream_t st);
int class_loss_bwd_partials();
int launch_class_loss_bwd(const float *xw, const float *bias, const int64_t *idx, int64_t B, int N, int K,
                          const float *g_lp, const float *g_prob, float *grad_logits, float *part_gx, cudaStream_t st);

// counts (N*K floats) = histogram of idx (B, N) per codebook; scratch: N*K uint32 (N*K <= 12,288)
int launch_index_counts(const int64_t *idx, int64_t B, int N, int K, float *counts, unsigned *scratch, cudaStream_t st);
// out[c] = sum_r X[r][c] for a contiguous (R, C) fp32 matrix, C a multiple of 4; part: column_sum_partials(C) floats
int column_sum_partials(int C);
int launch_column_sums(const float *X, int64_t R, int C, float *out, float *part, cudaStream_t st);

bool use_tensor_core_gemm();

}  // namespace mcq
