// api.cu -- the extern "C" surface of libmcq.so (include/mcq.h): argument checking, blob/workspace layout,
// chunking of large batches, and the kernel pipeline of each entry point.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mcq {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return MCQ_ECUDA;
}

int check_shape(int N, int K, int D) {
    if (D <= 0 || !is_pow2(N) || !is_pow2(K)) {  // quantization.py:33-36
        set_error("dim must be > 0 and num_codebooks (%d), codebook_size (%d) powers of two", N, K);
        return MCQ_EINVAL;
    }
    if (N > 1 && K < 16) {
        set_error("codebook_size %d < 16 with num_codebooks %d > 1: the reference raises UnboundLocalError "
                  "(quantization.py:453,470,504-507)", K, N);
        return MCQ_EUNSUPPORTED;
    }
    if (K > 256 || N > 64) {
        set_error("codebook_size %d > 256 or num_codebooks %d > 64 not supported", K, N);
        return MCQ_EUNSUPPORTED;
    }
    return MCQ_OK;
}

Prepared prepared_layout(int N, int K, int D) {
    Prepared L;
    L.N = N; L.K = K; L.D = D; L.NK = N * K;
    L.Dp = (int)align_up((size_t)D, 64);
    const size_t NKp = align_up((size_t)L.NK, 128);  // operand rows padded so every GEMM tile is in bounds
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L.off_cs = take(sizeof(float) * (size_t)L.NK * D);
    L.off_w = take(sizeof(float) * (size_t)L.NK * D);
    L.off_bias = take(sizeof(float) * (size_t)L.NK);
    L.off_gram = take(sizeof(float) * ((size_t)L.NK * L.NK + L.NK));  // table followed by its diagonal
    L.off_scal = take(sizeof(float) * 4);
    L.off_csplit = take(sizeof(__half) * 2 * NKp * L.Dp);
    L.off_wsplit = take(sizeof(__half) * 2 * NKp * L.Dp);
    L.off_cscale = take(sizeof(float) * NKp);
    L.off_wscale = take(sizeof(float) * NKp);
    L.bytes = off;
    return L;
}

Workspace workspace_layout(int64_t Bc, int N, int K, int D) {
    Workspace W;
    W.Bc = Bc;
    W.Mp = (int)align_up((size_t)Bc, 128);
    const size_t Dp = align_up((size_t)D, 64), NK = (size_t)N * K;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    W.off_ctr = take(1024);  // first, so that its place does not depend on the chunk size (mcq_search_stats)
    W.off_xf = take(sizeof(float) * (size_t)W.Mp * D);
    W.off_xsplit = take(sizeof(__half) * 2 * (size_t)W.Mp * Dp);
    W.off_lsplit = take(sizeof(__half) * 2 * (size_t)W.Mp * Dp);
    W.off_xscale = take(sizeof(float) * (size_t)W.Mp);
    W.off_lscale = take(sizeof(float) * (size_t)W.Mp);
    W.off_p = take(sizeof(float) * (size_t)W.Mp * NK);
    W.off_idx = take(sizeof(int32_t) * (size_t)W.Mp * N);
    W.bytes = off;
    return W;
}

// Largest chunk (multiple of 128 frames) whose workspace fits in `bytes`.
static int64_t chunk_for(size_t bytes, int64_t B, int N, int K, int D) {
    int64_t hi = (int64_t)align_up((size_t)(B > 0 ? B : 1), 128);
    const int64_t cap = max_chunk_frames();
    if (hi > cap) hi = cap;
    while (hi > 128 && workspace_layout(hi, N, K, D).bytes > bytes) hi -= 128;
    if (workspace_layout(hi, N, K, D).bytes > bytes) return 0;
    return hi;
}

// Frames per chunk of mcq_encode / mcq_refine: a multiple of 148 x 128 (full waves of 128-frame GEMM tiles).
// MCQ_CHUNK_WAVES overrides the number of waves (measurements only).
int64_t max_chunk_frames() {
    static int64_t cap = 0;
    if (cap == 0) {
        int waves = 4;
        const char *e = getenv("MCQ_CHUNK_WAVES");
        if (e && atoi(e) > 0) waves = atoi(e);
        cap = (int64_t)148 * 128 * waves;
    }
    return cap;
}

// MCQ_GEMM=ffma routes the two GEMMs through the CUDA-core kernel (used by the tests to cross-check tcgen05).
bool use_tensor_core_gemm() {
    const char *e = getenv("MCQ_GEMM");
    return !(e && strcmp(e, "ffma") == 0);
}

// P-like GEMM of one chunk: out (Mp, NK) = A . Bm^T, either through the tcgen05 fp16x2 kernel or the FFMA kernel.
static int chunk_gemm(const Prepared &L, const char *blob, const Workspace &W, char *ws, bool logits, int64_t Bc,
                      cudaStream_t st) {
    float *out = (float *)(ws + W.off_p);
    const bool tc = use_tensor_core_gemm() && L.NK % 64 == 0;
    if (tc) {
        const __half *a = (const __half *)(ws + (logits ? W.off_lsplit : W.off_xsplit));
        const __half *b = (const __half *)(blob + (logits ? L.off_wsplit : L.off_csplit));
        const float *as = (const float *)(ws + (logits ? W.off_lscale : W.off_xscale));
        const float *bs = (const float *)(blob + (logits ? L.off_wscale : L.off_cscale));
        return launch_gemm_tc(a, as, b, bs, out, W.Mp, L.NK, L.Dp, st);
    }
    const float *a = (const float *)(ws + W.off_xf);
    const float *b = (const float *)(blob + (logits ? L.off_w : L.off_cs));
    const float *scale = logits ? (const float *)(blob + L.off_scal) + 1 : nullptr;
    return launch_gemm_ffma(a, b, out, Bc, L.NK, L.D, scale, st);
}

static size_t dtype_size(int dt) { return dt == MCQ_F32 ? 4 : 2; }

// ---- optional per-kernel timing (bench.py): CUDA events recorded on the launch stream around every kernel --------
struct Prof {
    std::mutex mu;
    bool on = false;
    std::vector<cudaEvent_t> ev;  // pairs (begin, end)
    std::vector<int> kind;
    size_t used = 0;  // pairs in use
};
static Prof g_prof;

struct ProfScope {
    cudaStream_t st;
    cudaEvent_t end = nullptr;
    ProfScope(int kind, cudaStream_t s) : st(s) {
        if (!g_prof.on) return;
        // timing events cannot be recorded into a CUDA-graph capture (the trainer captures whole steps)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
        std::lock_guard<std::mutex> lock(g_prof.mu);
        if (g_prof.used * 2 + 2 > g_prof.ev.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            g_prof.ev.push_back(a);
            g_prof.ev.push_back(b);
            g_prof.kind.push_back(kind);
        }
        g_prof.kind[g_prof.used] = kind;
        cudaEventRecord(g_prof.ev[g_prof.used * 2], st);
        end = g_prof.ev[g_prof.used * 2 + 1];
        ++g_prof.used;
    }
    ~ProfScope() {
        if (end) cudaEventRecord(end, st);
    }
};
#define PROF(kind, st, call) [&]() { ProfScope ps__(kind, st); return (call); }()

}  // namespace mcq

using namespace mcq;

extern "C" {

int mcq_version(void) { return 1; }

int mcq_profile(int enable) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    g_prof.on = enable != 0;
    g_prof.used = 0;
    return MCQ_OK;
}

int mcq_profile_read(double *ms_by_kind, int64_t *launches_by_kind) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    for (int k = 0; k < MCQ_PROF_KINDS; ++k) {
        ms_by_kind[k] = 0.0;
        launches_by_kind[k] = 0;
    }
    for (size_t i = 0; i < g_prof.used; ++i) {
        MCQ_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float ms = 0.f;
        MCQ_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        ms_by_kind[g_prof.kind[i]] += ms;
        launches_by_kind[g_prof.kind[i]] += 1;
    }
    return MCQ_OK;
}
const char *mcq_last_error(void) { return g_err; }

int mcq_search_stats(void *workspace, int reset, uint64_t *passes_frames, void *stream) {
    if (!workspace) {
        set_error("mcq_search_stats: null workspace");
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *blk = (char *)workspace + sizeof(unsigned) * SEARCH_STAT_WORD;
    if (passes_frames) {
        MCQ_CUDA(cudaMemcpyAsync(passes_frames, blk, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MCQ_CUDA(cudaStreamSynchronize(st));
    }
    if (reset) MCQ_CUDA(cudaMemsetAsync(blk, 0, 2 * sizeof(uint64_t), st));
    return MCQ_OK;
}

int mcq_packed_cols(int N, int K) {
    long k = K;
    int cols = N;
    while (k * k <= 256 && cols >= 2) {  // quantization.py:266-271
        cols /= 2;
        k = k * k;
    }
    return cols;
}

size_t mcq_prepared_bytes(int N, int K, int D) {
    if (check_shape(N, K, D)) return 0;
    return prepared_layout(N, K, D).bytes;
}

size_t mcq_workspace_bytes(int64_t max_frames, int D, int N, int K) {
    if (check_shape(N, K, D)) return 0;
    if (max_frames < 1) max_frames = 1;
    int64_t cap = max_chunk_frames();
    int64_t Bc = (int64_t)align_up((size_t)max_frames, 128);
    if (Bc > cap) Bc = cap;
    return workspace_layout(Bc, N, K, D).bytes;
}

int mcq_prepare(const float *centers, const float *centers_scale, const float *w, const float *bias,
                const float *logits_scale, float scale_speed, int N, int K, int D, void *prepared,
                size_t prepared_bytes, void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (!centers || !centers_scale || !w || !bias || !logits_scale || !prepared) {
        set_error("mcq_prepare: null pointer");
        return MCQ_EINVAL;
    }
    Prepared L = prepared_layout(N, K, D);
    if (prepared_bytes < L.bytes) {
        set_error("mcq_prepare: blob of %zu bytes, need %zu", prepared_bytes, L.bytes);
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // padding rows / columns of the split operands must be zero
    MCQ_CUDA(cudaMemsetAsync((char *)prepared + L.off_csplit, 0, L.bytes - L.off_csplit, st));
    return launch_prepare(centers, centers_scale, w, bias, logits_scale, scale_speed, L, (char *)prepared, st);
}

const float *mcq_prepared_scaled_centers(const void *prepared, int N, int K, int D) {
    if (!prepared || check_shape(N, K, D)) return nullptr;
    return (const float *)((const char *)prepared + prepared_layout(N, K, D).off_cs);
}

const float *mcq_prepared_gram(const void *prepared, int N, int K, int D) {
    if (!prepared || check_shape(N, K, D)) return nullptr;
    return (const float *)((const char *)prepared + prepared_layout(N, K, D).off_gram);
}

int mcq_encode(const void *x, int x_dtype, int64_t B, int D, int N, int K, const void *prepared, int iters,
               void *codes, int codes_dtype, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B < 0 || iters < 0 || x_dtype < 0 || x_dtype > 2 || codes_dtype < 0 || codes_dtype > 2) {
        set_error("mcq_encode: bad argument (B=%lld iters=%d x_dtype=%d codes_dtype=%d)", (long long)B, iters, x_dtype,
                  codes_dtype);
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!x || !prepared || !codes || !workspace) {
        set_error("mcq_encode: null pointer");
        return MCQ_EINVAL;
    }
    const Prepared L = prepared_layout(N, K, D);
    const int64_t Bc = chunk_for(workspace_bytes, B, N, K, D);
    if (Bc <= 0) {
        set_error("mcq_encode: workspace of %zu bytes is too small (need %zu for 128 frames)", workspace_bytes,
                  workspace_layout(128, N, K, D).bytes);
        return MCQ_EINVAL;
    }
    const Workspace W = workspace_layout(Bc, N, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    const char *blob = (const char *)prepared;
    char *ws = (char *)workspace;
    const int ncols = codes_dtype == MCQ_U8 ? mcq_packed_cols(N, K) : N;
    const size_t code_elt = codes_dtype == MCQ_U8 ? 1 : (codes_dtype == MCQ_I64 ? 8 : 4);
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        const int64_t nb = B - b0 < Bc ? B - b0 : Bc;
        const char *xc = (const char *)x + (size_t)b0 * D * dtype_size(x_dtype);
        int32_t *idx = (int32_t *)(ws + W.off_idx);
        float *P = (float *)(ws + W.off_p);
        if ((rc = PROF(MCQ_PROF_OTHER, st, launch_split_x(xc, x_dtype, nb, L, blob, W, ws, true, st)))) return rc;
        // classifier arg-max initialisation (quantization.py:297-301): fused into the GEMM epilogue when the tiles
        // line up with the codebooks (MCQ_ARGMAX=unfused keeps the two-kernel path, for cross-checks)
        const char *am = getenv("MCQ_ARGMAX");
        if (use_tensor_core_gemm() && gemm_tc_argmax_supported(L.NK, K) && !(am && strcmp(am, "unfused") == 0)) {
            if ((rc = PROF(MCQ_PROF_GEMM, st,
                           launch_gemm_tc_argmax((const __half *)(ws + W.off_lsplit),
                                                 (const float *)(ws + W.off_lscale),
                                                 (const __half *)(blob + L.off_wsplit),
                                                 (const float *)(blob + L.off_wscale), W.Mp, L.NK, L.Dp,
                                                 (const float *)(blob + L.off_bias), nb, N, K, ws + W.off_p, idx, st))))
                return rc;
        } else {
            if ((rc = PROF(MCQ_PROF_GEMM, st, chunk_gemm(L, blob, W, ws, true, nb, st)))) return rc;
            if ((rc = PROF(MCQ_PROF_OTHER, st,
                           launch_argmax_init(P, (const float *)(blob + L.off_bias), nb, N, K, idx, st))))
                return rc;
        }
        if (iters > 0) {
            if ((rc = PROF(MCQ_PROF_GEMM, st, chunk_gemm(L, blob, W, ws, false, nb, st)))) return rc;
            unsigned *ctr = (unsigned *)(ws + W.off_ctr);
            MCQ_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned), st));
            if ((rc = PROF(MCQ_PROF_SEARCH, st,
                           launch_search(P, (const float *)(blob + L.off_gram), nb, N, K, iters, idx, idx, st, ctr))))
                return rc;
        }
        if ((rc = PROF(MCQ_PROF_OTHER, st,
                       launch_pack(idx, nb, N, K, (char *)codes + (size_t)b0 * ncols * code_elt, codes_dtype, st))))
            return rc;
    }
    return MCQ_OK;
}

int mcq_refine(const void *x, int x_dtype, int64_t B, int D, int N, int K, const void *prepared, int iters,
               const int64_t *idx_in, int64_t *idx_out, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B < 0 || iters < 0 || x_dtype < 0 || x_dtype > 2) {
        set_error("mcq_refine: bad argument");
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!x || !prepared || !idx_in || !idx_out || !workspace) {
        set_error("mcq_refine: null pointer");
        return MCQ_EINVAL;
    }
    const Prepared L = prepared_layout(N, K, D);
    const int64_t Bc = chunk_for(workspace_bytes, B, N, K, D);
    if (Bc <= 0) {
        set_error("mcq_refine: workspace of %zu bytes is too small", workspace_bytes);
        return MCQ_EINVAL;
    }
    const Workspace W = workspace_layout(Bc, N, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    const char *blob = (const char *)prepared;
    char *ws = (char *)workspace;
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        const int64_t nb = B - b0 < Bc ? B - b0 : Bc;
        const char *xc = (const char *)x + (size_t)b0 * D * dtype_size(x_dtype);
        int32_t *idx = (int32_t *)(ws + W.off_idx);
        float *P = (float *)(ws + W.off_p);
        if ((rc = PROF(MCQ_PROF_OTHER, st, launch_i64_to_i32(idx_in + (size_t)b0 * N, idx, nb * N, K, st)))) return rc;
        if (iters > 0) {
            if ((rc = PROF(MCQ_PROF_OTHER, st, launch_split_x(xc, x_dtype, nb, L, blob, W, ws, false, st)))) return rc;
            if ((rc = PROF(MCQ_PROF_GEMM, st, chunk_gemm(L, blob, W, ws, false, nb, st)))) return rc;
            unsigned *ctr = (unsigned *)(ws + W.off_ctr);
            MCQ_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned), st));
            if ((rc = PROF(MCQ_PROF_SEARCH, st,
                           launch_search(P, (const float *)(blob + L.off_gram), nb, N, K, iters, idx, idx, st, ctr))))
                return rc;
        }
        if ((rc = PROF(MCQ_PROF_OTHER, st, launch_i32_to_i64(idx, idx_out + (size_t)b0 * N, nb * N, st)))) return rc;
    }
    return MCQ_OK;
}

static int decode_common(const void *codes, int codes_dtype, int64_t B, int ncols, int N, int K, int D,
                         const float *cs, void *out, int out_dtype, void *stream) {
    if (D <= 0 || !is_pow2(N) || !is_pow2(K) || K > 256 || N > 64) {
        set_error("decode: unsupported shape N=%d K=%d D=%d", N, K, D);
        return MCQ_EINVAL;
    }
    if (B < 0 || ncols <= 0 || N % ncols != 0) {
        set_error("mcq_decode: %d code columns do not divide num_codebooks %d", ncols, N);
        return MCQ_EINVAL;
    }
    const int r = N / ncols;
    if (!(r == 1 || r == 2 || r == 4 || r == 8 || r == 16)) {  // quantization.py:566
        set_error("mcq_decode: num_codebooks / columns = %d not in {1,2,4,8,16}", r);
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!codes || !cs || !out) {
        set_error("mcq_decode: null pointer");
        return MCQ_EINVAL;
    }
    return PROF(MCQ_PROF_DECODE, (cudaStream_t)stream,
                launch_decode(codes, codes_dtype, B, ncols, N, K, D, cs, out, out_dtype, (cudaStream_t)stream));
}

int mcq_decode(const void *codes, int codes_dtype, int64_t B, int ncols, int N, int K, int D, const void *prepared,
               void *out, int out_dtype, void *stream) {
    if (D <= 0 || !is_pow2(N) || !is_pow2(K) || K > 256 || N > 64) {
        set_error("decode: unsupported shape N=%d K=%d D=%d", N, K, D);
        return MCQ_EINVAL;
    }
    const float *cs = prepared ? (const float *)((const char *)prepared + prepared_layout(N, K, D).off_cs) : nullptr;
    return decode_common(codes, codes_dtype, B, ncols, N, K, D, cs, out, out_dtype, stream);
}

int mcq_decode_centers(const void *codes, int codes_dtype, int64_t B, int ncols, int N, int K, int D,
                       const float *scaled_centers, void *out, int out_dtype, void *stream) {
    return decode_common(codes, codes_dtype, B, ncols, N, K, D, scaled_centers, out, out_dtype, stream);
}

int mcq_decode_backward(const float *grad_out, const int64_t *idx, int64_t B, int N, int K, int D,
                        float *grad_scaled_centers, void *stream) {
    if (B < 0 || N <= 0 || K <= 0 || D <= 0) {
        set_error("mcq_decode_backward: bad shape");
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!grad_out || !idx || !grad_scaled_centers) {
        set_error("mcq_decode_backward: null pointer");
        return MCQ_EINVAL;
    }
    return launch_decode_backward(grad_out, idx, B, N, K, D, grad_scaled_centers, (cudaStream_t)stream);
}

int mcq_class_loss_forward(const void *x, int x_dtype, int64_t B, int D, int N, int K, const void *prepared,
                           const int64_t *idx, float *xw, float *logprob_sum, float *prob_sum, void *workspace,
                           size_t workspace_bytes, void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B <= 0 || x_dtype < 0 || x_dtype > 2) {
        set_error("mcq_class_loss_forward: bad argument (B=%lld)", (long long)B);
        return MCQ_EINVAL;
    }
    if (!x || !prepared || !idx || !xw || !logprob_sum || !prob_sum || !workspace) {
        set_error("mcq_class_loss_forward: null pointer");
        return MCQ_EINVAL;
    }
    const Prepared L = prepared_layout(N, K, D);
    const int64_t Bc = chunk_for(workspace_bytes, B, N, K, D);
    if (Bc <= 0) {
        set_error("mcq_class_loss_forward: workspace too small");
        return MCQ_EINVAL;
    }
    const Workspace W = workspace_layout(Bc, N, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    const char *blob = (const char *)prepared;
    char *ws = (char *)workspace;
    const bool tc = use_tensor_core_gemm() && L.NK % 64 == 0;
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        const int64_t nb = B - b0 < Bc ? B - b0 : Bc;
        const char *xc = (const char *)x + (size_t)b0 * D * dtype_size(x_dtype);
        float *out = xw + (size_t)b0 * L.NK;  // chunk starts are multiples of 128 rows: the GEMM writes in place
        if ((rc = PROF(MCQ_PROF_OTHER, st, launch_split_x(xc, x_dtype, nb, L, blob, W, ws, true, st)))) return rc;
        if (tc) {
            // only the first mp rows of the chunk's split buffer take part; its second fp16 plane starts W.Mp rows
            // (not mp rows) after the first, whatever the size of this -- possibly last, shorter -- chunk
            const int64_t mp = (int64_t)align_up((size_t)nb, 128);
            rc = PROF(MCQ_PROF_GEMM, st,
                      launch_gemm_tc_general((const __half *)(ws + W.off_lsplit), (const float *)(ws + W.off_lscale),
                                             (const __half *)(blob + L.off_wsplit),
                                             (const float *)(blob + L.off_wscale), out, L.NK, nb, mp, L.NK, L.Dp, 0, st,
                                             W.Mp));
        } else {
            rc = PROF(MCQ_PROF_GEMM, st,
                      launch_gemm_ffma((const float *)(ws + W.off_xf), (const float *)(blob + L.off_w), out, nb, L.NK,
                                       L.D, (const float *)(blob + L.off_scal) + 1, st));
        }
        if (rc) return rc;
    }
    // partial sums live in the (now unused) P region of the workspace
    const int ns = class_loss_streams(B, N, K);
    float *part_prob = (float *)(ws + W.off_p);
    const size_t need = sizeof(float) * ((size_t)ns * L.NK + (size_t)ns * N);
    if (need > sizeof(float) * (size_t)W.Mp * L.NK) {
        set_error("mcq_class_loss_forward: workspace too small for %d partial rows", ns);
        return MCQ_EINVAL;
    }
    float *part_lp = part_prob + (size_t)ns * L.NK;
    return PROF(MCQ_PROF_OTHER, st,
                launch_class_loss_fwd(xw, (const float *)(blob + L.off_bias), idx, B, N, K, part_prob, part_lp, prob_sum,
                                      logprob_sum, st));
}

int mcq_class_loss_partials(void) { return class_loss_bwd_partials(); }

int mcq_class_loss_backward(const float *xw, int64_t B, int D, int N, int K, const void *prepared, const int64_t *idx,
                            const float *g_logprob_sum, const float *g_prob_sum, float *grad_logits, float *part_gx,
                            void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B <= 0) {
        set_error("mcq_class_loss_backward: bad argument");
        return MCQ_EINVAL;
    }
    if (!xw || !prepared || !idx || !g_logprob_sum || !g_prob_sum || !grad_logits || !part_gx) {
        set_error("mcq_class_loss_backward: null pointer");
        return MCQ_EINVAL;
    }
    const Prepared L = prepared_layout(N, K, D);
    return PROF(MCQ_PROF_OTHER, (cudaStream_t)stream,
                launch_class_loss_bwd(xw, (const float *)((const char *)prepared + L.off_bias), idx, B, N, K,
                                      g_logprob_sum, g_prob_sum, grad_logits, part_gx, (cudaStream_t)stream));
}

int mcq_index_counts(const int64_t *idx, int64_t B, int N, int K, float *counts, void *scratch, void *stream) {
    if (B < 0 || N <= 0 || K <= 0) {
        set_error("mcq_index_counts: bad shape");
        return MCQ_EINVAL;
    }
    if (!counts || !scratch || (B > 0 && !idx)) {
        set_error("mcq_index_counts: null pointer");
        return MCQ_EINVAL;
    }
    return PROF(MCQ_PROF_OTHER, (cudaStream_t)stream,
                launch_index_counts(idx, B, N, K, counts, (unsigned *)scratch, (cudaStream_t)stream));
}

int mcq_column_sum_partials(int cols) { return cols > 0 ? column_sum_partials(cols) : 0; }

int mcq_column_sums(const float *x, int64_t rows, int cols, float *out, float *partials, void *stream) {
    if (rows <= 0 || cols <= 0 || (cols & 3) || ((uintptr_t)x & 15)) {
        set_error("mcq_column_sums: rows=%lld cols=%d (cols must be a multiple of 4, x 16-byte aligned)", (long long)rows,
                  cols);
        return MCQ_EINVAL;
    }
    if (!x || !out || !partials) {
        set_error("mcq_column_sums: null pointer");
        return MCQ_EINVAL;
    }
    return PROF(MCQ_PROF_OTHER, (cudaStream_t)stream,
                launch_column_sums(x, rows, cols, out, partials, (cudaStream_t)stream));
}

int mcq_xct(const void *x, int x_dtype, int64_t B, int D, int N, int K, const void *prepared, float *P, void *workspace,
            size_t workspace_bytes, void *stream) {
    int rc = check_shape(N, K, D);
    if (rc) return rc;
    if (B <= 0) return MCQ_OK;
    const Prepared L = prepared_layout(N, K, D);
    const int64_t Bc = chunk_for(workspace_bytes, B, N, K, D);
    if (Bc <= 0) {
        set_error("mcq_xct: workspace too small");
        return MCQ_EINVAL;
    }
    const Workspace W = workspace_layout(Bc, N, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
        const int64_t nb = B - b0 < Bc ? B - b0 : Bc;
        const char *xc = (const char *)x + (size_t)b0 * D * dtype_size(x_dtype);
        if ((rc = launch_split_x(xc, x_dtype, nb, L, (const char *)prepared, W, ws, false, st))) return rc;
        if ((rc = chunk_gemm(L, (const char *)prepared, W, ws, false, nb, st))) return rc;
        MCQ_CUDA(cudaMemcpyAsync(P + (size_t)b0 * L.NK, ws + W.off_p, sizeof(float) * (size_t)nb * L.NK,
                                 cudaMemcpyDeviceToDevice, st));
    }
    return MCQ_OK;
}

int mcq_search(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
               int32_t *idx_out, void *stream) {
    int rc = check_shape(N, K, 1);
    if (rc) return rc;
    if (B < 0 || iters < 0) {
        set_error("mcq_search: bad argument");
        return MCQ_EINVAL;
    }
    if (B == 0) return MCQ_OK;
    if (!P || !gram || !idx_in || !idx_out) {
        set_error("mcq_search: null pointer");
        return MCQ_EINVAL;
    }
    if (iters == 0) {
        if (idx_in != idx_out)
            MCQ_CUDA(cudaMemcpyAsync(idx_out, idx_in, sizeof(int32_t) * (size_t)B * N, cudaMemcpyDeviceToDevice,
                                     (cudaStream_t)stream));
        return MCQ_OK;
    }
    return launch_search(P, gram, B, N, K, iters, idx_in, idx_out, (cudaStream_t)stream);
}

}  // extern "C"
