// jcl.cu -- the non-GEMM stages of JointCodebookLoss (reference prediction.py:9-82), the downstream consumer of
// Quantizer.encode's codes.  The reference materialises, per frame, an
// (N, hidden) embedding gather, its concatenation with the projected predictor, a cumsum, a ReLU, and for the loss a
// (N, K) log-softmax plus its autograd graph.  Here:
//   jcl_hidden_fwd   act[n, b, :] = relu(h_b + sum_{m<n} scale * E[m*K + idx[b,m], :])        (:47-68)
//                    one pass: reads h (B, H) and the L2-resident embedding table, writes the (N, B, H) operand of
//                    the per-codebook GEMM directly in the layout that GEMM wants;
//   jcl_hidden_bwd   the transpose of that pass: ReLU mask, reverse running sum over n, scatter-add into the
//                    embedding gradient (vector atomics), projected-predictor gradient;
//   jcl_ce           cross entropy of logits (B, N, K) + bias (N, K) against the codes with ignore_index (:79-82),
//                    per-row losses, fixed-order partial sums (deterministic), and -- in place -- softmax - onehot.
// All three are HBM-bound streaming kernels: coalesced 128-bit accesses, grids sized from the SM count.
// The dense products between them are plain library GEMMs issued by the host layer.
#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <typename T>
__device__ __forceinline__ long long load_code(const void *codes, size_t i) {
    return (long long)reinterpret_cast<const T *>(codes)[i];
}

__device__ __forceinline__ long long code_at(const void *codes, int dt, size_t i) {
    switch (dt) {
        case MCQ_U8: return load_code<uint8_t>(codes, i);
        case MCQ_I32: return load_code<int32_t>(codes, i);
        default: return load_code<int64_t>(codes, i);
    }
}

// Row of the embedding table for codebook m of a frame: clamp(min=0) as the reference (:43-45; padded frames are
// don't-cares), and clamped from above so that a corrupt code cannot read outside the table.
__device__ __forceinline__ long long emb_row(const void *codes, int dt, size_t b, int N, int K, int m) {
    long long c = code_at(codes, dt, b * N + m);
    c = c < 0 ? 0 : (c >= K ? K - 1 : c);
    return c + (long long)m * K;
}

// One warp per frame; lane L owns the float4 chunks L, L+32, ... of the hidden vector.  The running sum follows the
// reference's order exactly (product by `scale` rounded, then added, n ascending: the sequential cumsum).
__global__ void __launch_bounds__(256) jcl_hidden_fwd_kernel(const float *__restrict__ hidden, const void *__restrict__ codes,
                                                             int codes_dtype, int64_t B, int N, int K, int H,
                                                             const float *__restrict__ emb, float scale,
                                                             float *__restrict__ act) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int H4 = H >> 2;
    const size_t plane = (size_t)B * H;
    for (int64_t b = warp; b < B; b += nwarps) {
        // lanes 0..N-2 (and, for N > 33, a second round) hold the embedding rows of this frame
        long long r0 = 0, r1 = 0;
        if (lane < N - 1) r0 = emb_row(codes, codes_dtype, (size_t)b, N, K, lane);
        if (lane + 32 < N - 1) r1 = emb_row(codes, codes_dtype, (size_t)b, N, K, lane + 32);
        for (int c0 = 0; c0 < H4; c0 += 32) {  // warp-uniform trip count: the shuffles below need all lanes
            const int c4 = c0 + lane;
            const bool on = c4 < H4;
            float4 run = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (on) run = __ldg(reinterpret_cast<const float4 *>(hidden + (size_t)b * H) + c4);
            float *o = act + (size_t)b * H + 4 * (size_t)c4;
            for (int n = 0; n < N; ++n) {
                if (on)
                    __stcs(reinterpret_cast<float4 *>(o + (size_t)n * plane),
                           make_float4(fmaxf(run.x, 0.0f), fmaxf(run.y, 0.0f), fmaxf(run.z, 0.0f), fmaxf(run.w, 0.0f)));
                if (n < N - 1) {
                    const long long row = __shfl_sync(FULL, n < 32 ? r0 : r1, n & 31);
                    float4 e = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (on) e = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)row * H) + c4);
                    run.x = __fadd_rn(run.x, __fmul_rn(e.x, scale));
                    run.y = __fadd_rn(run.y, __fmul_rn(e.y, scale));
                    run.z = __fadd_rn(run.z, __fmul_rn(e.z, scale));
                    run.w = __fadd_rn(run.w, __fmul_rn(e.w, scale));
                }
            }
        }
    }
}

// Transpose of the pass above.  grad_act (N, B, H) is d loss / d relu output; act gives the ReLU mask.
__global__ void __launch_bounds__(256) jcl_hidden_bwd_kernel(const float *__restrict__ grad_act, const float *__restrict__ act,
                                                             const void *__restrict__ codes, int codes_dtype, int64_t B,
                                                             int N, int K, int H, float scale,
                                                             float *__restrict__ grad_hidden,
                                                             float *__restrict__ grad_emb) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int H4 = H >> 2;
    const size_t plane = (size_t)B * H;
    for (int64_t b = warp; b < B; b += nwarps) {
        long long r0 = 0, r1 = 0;
        if (lane < N - 1) r0 = emb_row(codes, codes_dtype, (size_t)b, N, K, lane);
        if (lane + 32 < N - 1) r1 = emb_row(codes, codes_dtype, (size_t)b, N, K, lane + 32);
        for (int c0 = 0; c0 < H4; c0 += 32) {
            const int c4 = c0 + lane;
            const bool on = c4 < H4;
            float4 run = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (int n = N - 1; n >= 0; --n) {
                const size_t off = (size_t)n * plane + (size_t)b * H;
                float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f), a = g;
                if (on) {
                    g = __ldcs(reinterpret_cast<const float4 *>(grad_act + off) + c4);
                    a = __ldcs(reinterpret_cast<const float4 *>(act + off) + c4);
                }
                run.x += a.x > 0.0f ? g.x : 0.0f;
                run.y += a.y > 0.0f ? g.y : 0.0f;
                run.z += a.z > 0.0f ? g.z : 0.0f;
                run.w += a.w > 0.0f ? g.w : 0.0f;
                if (n >= 1) {
                    // embedding n-1 entered every position >= n: its gradient is the running sum so far
                    const long long row = __shfl_sync(FULL, (n - 1) < 32 ? r0 : r1, (n - 1) & 31);
                    float4 *dst = reinterpret_cast<float4 *>(grad_emb + (size_t)row * H) + c4;
                    if (on) atomicAdd(dst, make_float4(run.x * scale, run.y * scale, run.z * scale, run.w * scale));
                }
            }
            if (on) reinterpret_cast<float4 *>(grad_hidden + (size_t)b * H)[c4] = run;
        }
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// One warp per (frame, codebook) row of K logits; element k = t*32 + lane.  EPL = ceil(K / 32).
template <int EPL>
__global__ void __launch_bounds__(256) jcl_ce_kernel(float *__restrict__ logits, const float *__restrict__ bias,
                                                     const void *__restrict__ codes, int codes_dtype, int64_t rows,
                                                     int N, int K, long long ignore_index, int want_grad,
                                                     float *__restrict__ row_loss, float *__restrict__ partials) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float loss_acc = 0.0f, cnt_acc = 0.0f;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const int n = (int)(r % N);
        float *z = logits + (size_t)r * K;
        const float *bz = bias + (size_t)n * K;
        float v[EPL];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            const int k = t * 32 + lane;
            v[t] = k < K ? z[k] + __ldg(bz + k) : -INFINITY;
            mx = fmaxf(mx, v[t]);
        }
        mx = warp_max(mx);
        const long long tgt = code_at(codes, codes_dtype, (size_t)r);
        // the reference's cross_entropy accepts ignore_index or 0..K-1 (anything else is a device-side assert there);
        // here any other value is skipped like ignore_index
        const bool ignored = (tgt == ignore_index) || tgt < 0 || tgt >= K;
        float s = 0.0f, zt = 0.0f;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            const float d = v[t] - mx;
            if ((long long)(t * 32 + lane) == tgt) zt = d;
            v[t] = (t * 32 + lane < K) ? expf(d) : 0.0f;
            s += v[t];
        }
        s = warp_sum(s);
        zt = warp_sum(zt);
        const float ls = logf(s);
        const float loss = ignored ? 0.0f : ls - zt;
        if (lane == 0) {
            row_loss[r] = loss;
            loss_acc += loss;
            cnt_acc += ignored ? 0.0f : 1.0f;
        }
        if (want_grad) {
            const float inv = 1.0f / s;
#pragma unroll
            for (int t = 0; t < EPL; ++t) {
                const int k = t * 32 + lane;
                if (k < K) {
                    const float p = v[t] * inv;
                    z[k] = ignored ? 0.0f : p - ((long long)k == tgt ? 1.0f : 0.0f);
                }
            }
        }
    }
    // fixed assignment of rows to warps and a fixed-order second stage: the sums are reproducible run to run
    if (lane == 0) {
        partials[2 * warp] = loss_acc;
        partials[2 * warp + 1] = cnt_acc;
    }
}

__global__ void __launch_bounds__(1024) jcl_ce_reduce_kernel(const float *__restrict__ partials, int nwarps,
                                                             float *__restrict__ sums) {
    __shared__ float sh[2][32];
    float a = 0.0f, c = 0.0f;
    for (int i = threadIdx.x; i < nwarps; i += 1024) {
        a += partials[2 * i];
        c += partials[2 * i + 1];
    }
    a = warp_sum(a);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = a;
        sh[1][threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = warp_sum(sh[0][threadIdx.x]);
        c = warp_sum(sh[1][threadIdx.x]);
        if (threadIdx.x == 0) {
            sums[0] = a;
            sums[1] = c;
        }
    }
}

constexpr int JCL_CE_BLOCKS = 148 * 8;  // 8 CTAs of 8 warps per SM
constexpr int JCL_CE_WARPS = JCL_CE_BLOCKS * 8;

int check_jcl(const char *who, int64_t B, int N, int K, int H, int codes_dtype) {
    if (B < 0 || N < 2 || N > 64 || K < 1 || H < 4 || (H & 3)) {
        set_error("%s: need num_frames >= 0, 2 <= num_codebooks <= 64, codebook_size >= 1, hidden_channels a multiple of 4 "
                  "(got B=%lld N=%d K=%d H=%d)", who, (long long)B, N, K, H);
        return MCQ_EINVAL;
    }
    if (codes_dtype != MCQ_U8 && codes_dtype != MCQ_I32 && codes_dtype != MCQ_I64) {
        set_error("%s: unknown codes dtype %d", who, codes_dtype);
        return MCQ_EINVAL;
    }
    return MCQ_OK;
}

int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

}  // namespace mcq

using namespace mcq;

extern "C" {

int mcq_jcl_hidden_forward(const float *hidden, const void *codes, int codes_dtype, int64_t B, int N, int K, int H,
                           const float *embedding, float scale, float *act, void *stream) {
    int rc = check_jcl("mcq_jcl_hidden_forward", B, N, K, H, codes_dtype);
    if (rc) return rc;
    if (B == 0) return MCQ_OK;
    if (!hidden || !codes || !embedding || !act) {
        set_error("mcq_jcl_hidden_forward: null pointer");
        return MCQ_EINVAL;
    }
    int64_t blocks = (B + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    jcl_hidden_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hidden, codes, codes_dtype, B, N, K, H,
                                                                              embedding, scale, act);
    MCQ_LAUNCH_CHECK("jcl_hidden_fwd_kernel");
    return MCQ_OK;
}

int mcq_jcl_hidden_backward(const float *grad_act, const float *act, const void *codes, int codes_dtype, int64_t B, int N,
                            int K, int H, float scale, float *grad_hidden, float *grad_embedding, void *stream) {
    int rc = check_jcl("mcq_jcl_hidden_backward", B, N, K, H, codes_dtype);
    if (rc) return rc;
    if (B == 0) return MCQ_OK;
    if (!grad_act || !act || !codes || !grad_hidden || !grad_embedding) {
        set_error("mcq_jcl_hidden_backward: null pointer");
        return MCQ_EINVAL;
    }
    int64_t blocks = (B + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    jcl_hidden_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_act, act, codes, codes_dtype, B, N, K, H,
                                                                              scale, grad_hidden, grad_embedding);
    MCQ_LAUNCH_CHECK("jcl_hidden_bwd_kernel");
    return MCQ_OK;
}

int mcq_jcl_partials(void) { return 2 * JCL_CE_WARPS; }

int mcq_jcl_cross_entropy(float *logits, const float *bias, const void *codes, int codes_dtype, int64_t B, int N, int K,
                          int64_t ignore_index, int want_grad, float *row_loss, float *sums, float *partials,
                          void *stream) {
    int rc = check_jcl("mcq_jcl_cross_entropy", B, N, K, 4, codes_dtype);
    if (rc) return rc;
    if (K > 1024) {
        set_error("mcq_jcl_cross_entropy: codebook_size %d > 1024", K);
        return MCQ_EUNSUPPORTED;
    }
    if (!sums || !partials || (B > 0 && (!logits || !bias || !codes || !row_loss))) {
        set_error("mcq_jcl_cross_entropy: null pointer");
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = B * N;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > JCL_CE_BLOCKS) blocks = JCL_CE_BLOCKS;
    if (blocks < 1) blocks = 1;
#define MCQ_CE(EPL)                                                                                                   \
    jcl_ce_kernel<EPL><<<(unsigned)blocks, 256, 0, st>>>(logits, bias, codes, codes_dtype, rows, N, K,                \
                                                         (long long)ignore_index, want_grad, row_loss, partials)
    const int epl = (K + 31) / 32;
    if (epl <= 1) MCQ_CE(1);
    else if (epl <= 2) MCQ_CE(2);
    else if (epl <= 4) MCQ_CE(4);
    else if (epl <= 8) MCQ_CE(8);
    else if (epl <= 16) MCQ_CE(16);
    else MCQ_CE(32);
#undef MCQ_CE
    MCQ_LAUNCH_CHECK("jcl_ce_kernel");
    jcl_ce_reduce_kernel<<<1, 1024, 0, st>>>(partials, (int)blocks * 8, sums);
    MCQ_LAUNCH_CHECK("jcl_ce_reduce_kernel");
    return MCQ_OK;
}

}  // extern "C"
