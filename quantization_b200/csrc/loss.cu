// loss.cu -- the classifier-side losses of Quantizer.compute_loss (quantization.py:218-240) without the
// (B, N, K) intermediates the reference materialises (log_softmax, exp, gather, one-hot counts, their autograd copies):
//   forward : logprob_sum    = sum_{b,n} log_softmax(logits[b,n,:])[idx[b,n]]          (:221-225 before the mean)
//             prob_sum[n,k]  = sum_b softmax(logits[b,n,:])[k]                          (:235 before the mean)
//   backward: grad_logits[b,n,j] = g_lp * (1[j == idx] - s_j) + s_j * (c[n,j] - sum_k c[n,k] s_k),   s = softmax,
//             g_lp = dL/dlogprob_sum, c = dL/dprob_sum.
// logits = xw + bias where xw = fl(exp(logits_scale*speed) * x) . W^T comes from the tcgen05 GEMM (gemm_tc.cu).
// Sums over frames are formed in a fixed order (per-warp partials over a strided frame set, then one pass over the
// partials), so the losses are bit-reproducible from run to run -- no atomics.
#include <math.h>

#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;

// reductions over aligned groups of W lanes (W = 16 or 32)
template <int W>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = v + __shfl_xor_sync(FULL, v, o);
    return v;
}

// One row of K logits is handled by a segment of W lanes, EPL consecutive elements per lane (K = W * EPL, or
// K < 16 = W with EPL = 1 and the upper lanes idle).  Returns softmax s[] and log-softmax ls[] of (xw + bias).
template <int W, int EPL>
__device__ __forceinline__ void row_softmax(const float *__restrict__ xw, const float *__restrict__ bias, int K,
                                            int sl, float (&s)[EPL], float (&ls)[EPL], float (&raw)[EPL]) {
    float l[EPL];
    if constexpr (EPL >= 4) {
#pragma unroll
        for (int t = 0; t < EPL; t += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(xw + sl * EPL + t);
            const float4 c = __ldg(reinterpret_cast<const float4 *>(bias + sl * EPL + t));
            raw[t] = a.x;
            raw[t + 1] = a.y;
            raw[t + 2] = a.z;
            raw[t + 3] = a.w;
            l[t] = a.x + c.x;
            l[t + 1] = a.y + c.y;
            l[t + 2] = a.z + c.z;
            l[t + 3] = a.w + c.w;
        }
    } else {
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            const int k = sl * EPL + t;
            raw[t] = k < K ? xw[k] : 0.0f;
            l[t] = k < K ? raw[t] + __ldg(bias + k) : -INFINITY;
        }
    }
    float m = l[0];
#pragma unroll
    for (int t = 1; t < EPL; ++t) m = fmaxf(m, l[t]);
    m = seg_max<W>(m);
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        s[t] = expf(l[t] - m);  // exp(-inf) = 0 for the idle lanes of K < 16
        sum += s[t];
    }
    sum = seg_sum<W>(sum);
    const float inv = 1.0f / sum, lsum = logf(sum);
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        s[t] = s[t] * inv;
        ls[t] = (l[t] - m) - lsum;
    }
}

// segment gs handles codebook n = gs % N and frames st, st + nstreams, ... with st = gs / N.
template <int W, int EPL>
__global__ void __launch_bounds__(256) class_loss_fwd_kernel(const float *__restrict__ xw, const float *__restrict__ bias,
                                                             const int64_t *__restrict__ idx, int64_t B, int N, int K,
                                                             int nstreams, float *__restrict__ part_prob,
                                                             float *__restrict__ part_lp) {
    constexpr int SPW = 32 / W;  // segments per warp
    const int lane = threadIdx.x & 31, sl = lane % W;
    const int gs = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SPW + lane / W;
    const bool live = gs < nstreams * N;  // dead segments still take part in the shuffles
    const int n = live ? gs % N : 0, st = live ? gs / N : 0;
    const size_t NK = (size_t)N * K;
    float acc[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) acc[t] = 0.0f;
    float lp = 0.0f;
    // all segments of a warp run the same number of iterations (shuffles are warp-wide)
    for (int64_t b0 = 0; b0 < B; b0 += nstreams) {
        const int64_t b = b0 + st;
        const bool on = live && b < B;
        const int64_t br = on ? b : 0;
        float s[EPL], ls[EPL], raw[EPL];
        row_softmax<W, EPL>(xw + (size_t)br * NK + (size_t)n * K, bias + (size_t)n * K, K, sl, s, ls, raw);
        const int kc = (int)idx[(size_t)br * N + n];
        if (on) {
#pragma unroll
            for (int t = 0; t < EPL; ++t) {
                acc[t] += s[t];
                if (sl * EPL + t == kc) lp += ls[t];
            }
        }
    }
    lp = seg_sum<W>(lp);
    if (live) {
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            const int k = sl * EPL + t;
            if (k < K) part_prob[(size_t)st * NK + (size_t)n * K + k] = acc[t];
        }
        if (sl == 0) part_lp[(size_t)st * N + n] = lp;
    }
}

// prob_sum[c] = sum over the partial rows, in a fixed order: 32 groups of threads sum every 32nd row, then the 32
// group sums are added in group order.  Block = 32 columns x 32 groups.
__global__ void __launch_bounds__(1024) class_loss_reduce_kernel(const float *__restrict__ part_prob,
                                                                 const float *__restrict__ part_lp, int nstreams,
                                                                 int N, int K, float *__restrict__ prob_sum,
                                                                 float *__restrict__ logprob_sum) {
    __shared__ float sm[32][33];
    const int NK = N * K;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.0f;
    if (c < NK) {
#pragma unroll 4
        for (int st = ty; st < nstreams; st += 32) s += part_prob[(size_t)st * NK + c];
    }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < NK) {
        float t = 0.0f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += sm[g][tx];
        prob_sum[c] = t;
    }
    if (blockIdx.x == 0 && ty == 1) {  // the scalar: fixed-order per-lane partials, then a shuffle tree
        float t = 0.0f;
        for (int i = tx; i < nstreams * N; i += 32) t += part_lp[i];
        t = seg_sum<32>(t);
        if (tx == 0) *logprob_sum = t;
    }
}

template <int W, int EPL>
__global__ void __launch_bounds__(256) class_loss_bwd_kernel(const float *__restrict__ xw, const float *__restrict__ bias,
                                                             const int64_t *__restrict__ idx, int64_t B, int N, int K,
                                                             const float *__restrict__ g_lp,
                                                             const float *__restrict__ g_prob,
                                                             float *__restrict__ grad_logits,
                                                             float *__restrict__ part_gx) {
    constexpr int SPW = 32 / W;
    const int lane = threadIdx.x & 31, sl = lane % W;
    const int64_t rows = B * N;
    const size_t NK = (size_t)N * K;
    const float glp = *g_lp;
    const int64_t nseg = (int64_t)gridDim.x * (blockDim.x >> 5) * SPW;
    float gx = 0.0f;  // sum of grad_logits * xw over this lane's elements: d loss / d logits_scale up to a factor
    for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SPW; r0 < rows; r0 += nseg) {
        const int64_t r = r0 + lane / W;
        const bool on = r < rows;
        const int64_t rr = on ? r : 0;
        const int64_t b = rr / N;
        const int n = (int)(rr - b * N);
        float s[EPL], ls[EPL], raw[EPL];
        row_softmax<W, EPL>(xw + (size_t)b * NK + (size_t)n * K, bias + (size_t)n * K, K, sl, s, ls, raw);
        const int kc = (int)idx[rr];
        float c[EPL];
        float dotc = 0.0f;
#pragma unroll
        for (int t = 0; t < EPL; ++t) {
            const int k = sl * EPL + t;
            c[t] = k < K ? __ldg(g_prob + (size_t)n * K + k) : 0.0f;
            dotc += c[t] * s[t];
        }
        dotc = seg_sum<W>(dotc);
        float *g = grad_logits + (size_t)b * NK + (size_t)n * K + sl * EPL;
        float o[EPL];
#pragma unroll
        for (int t = 0; t < EPL; ++t) o[t] = glp * ((sl * EPL + t == kc ? 1.0f : 0.0f) - s[t]) + s[t] * (c[t] - dotc);
        if (on) {
#pragma unroll
            for (int t = 0; t < EPL; ++t) gx += o[t] * raw[t];  // (raw is 0 on the idle lanes of K < 16)
            if constexpr (EPL >= 4) {
#pragma unroll
                for (int t = 0; t < EPL; t += 4)
                    *reinterpret_cast<float4 *>(g + t) = make_float4(o[t], o[t + 1], o[t + 2], o[t + 3]);
            } else {
#pragma unroll
                for (int t = 0; t < EPL; ++t)
                    if (sl * EPL + t < K) g[t] = o[t];
            }
        }
    }
    // fixed work assignment + shuffle tree: the per-warp partials (and their sum in warp order) are reproducible
    gx = seg_sum<32>(gx);
    if (lane == 0) part_gx[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = gx;
}

// ---- small codebooks (K = 16 or 32: trainer phase 1): one THREAD per (frame stream, codebook) row.  A row is K*4
// contiguous bytes, a warp reads 32 consecutive rows; no shuffles, the softmax lives in the thread's registers.  Same
// partial-sum layout and frame assignment as the segment kernels above (thread gs: codebook gs % N, frames gs / N,
// gs / N + nstreams, ...), just with many more streams. ----
template <int K>
__device__ __forceinline__ void row_softmax_thread(const float *__restrict__ xw, const float *__restrict__ bias,
                                                   float (&s)[K], float (&raw)[K], float &m, float &lsum) {
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const float4 a = *reinterpret_cast<const float4 *>(xw + 4 * q);
        const float4 c = __ldg(reinterpret_cast<const float4 *>(bias) + q);
        raw[4 * q] = a.x;
        raw[4 * q + 1] = a.y;
        raw[4 * q + 2] = a.z;
        raw[4 * q + 3] = a.w;
        s[4 * q] = a.x + c.x;
        s[4 * q + 1] = a.y + c.y;
        s[4 * q + 2] = a.z + c.z;
        s[4 * q + 3] = a.w + c.w;
    }
    m = s[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, s[k]);
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s[k] = s[k] - m;      // shifted logits; turned into probabilities by the caller
        sum += expf(s[k]);
    }
    lsum = logf(sum);
}

template <int K>
__global__ void __launch_bounds__(256) class_loss_fwd_small_kernel(const float *__restrict__ xw, const float *__restrict__ bias,
                                                                   const int64_t *__restrict__ idx, int64_t B, int N,
                                                                   int nstreams, float *__restrict__ part_prob,
                                                                   float *__restrict__ part_lp) {
    const int64_t gs = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gs < (int64_t)nstreams * N;
    const int n = (int)(gs % N);
    const int64_t st = gs / N;
    const size_t NK = (size_t)N * K;
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0f;
    float lp = 0.0f;
    for (int64_t b = live ? st : B; b < B; b += nstreams) {
        float s[K], raw[K], m, lsum;
        row_softmax_thread<K>(xw + (size_t)b * NK + (size_t)n * K, bias + (size_t)n * K, s, raw, m, lsum);
        const int kc = (int)idx[(size_t)b * N + n];
        const float inv = expf(-lsum);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            acc[k] += expf(s[k]) * inv;
            if (k == kc) lp += s[k] - lsum;
        }
    }
    // block-level sum in a fixed order (thread t: codebook t % N, frame slot t / N; 256 % N == 0): one partial row
    // per BLOCK, so the final pass over the partial rows stays short
    __shared__ float red[256][K + 1];
    __shared__ float redlp[256];
#pragma unroll
    for (int k = 0; k < K; ++k) red[threadIdx.x][k] = acc[k];
    redlp[threadIdx.x] = lp;
    __syncthreads();
    const int slots = 256 / N;
    for (int c = threadIdx.x; c < N * K; c += 256) {
        const int cn = c / K, ck = c % K;
        float t = 0.0f;
        for (int f = 0; f < slots; ++f) t += red[f * N + cn][ck];
        part_prob[(size_t)blockIdx.x * NK + c] = t;
    }
    if (threadIdx.x < N) {
        float t = 0.0f;
        for (int f = 0; f < slots; ++f) t += redlp[f * N + threadIdx.x];
        part_lp[(size_t)blockIdx.x * N + threadIdx.x] = t;
    }
}

template <int K>
__global__ void __launch_bounds__(256) class_loss_bwd_small_kernel(const float *__restrict__ xw, const float *__restrict__ bias,
                                                                   const int64_t *__restrict__ idx, int64_t B, int N,
                                                                   const float *__restrict__ g_lp,
                                                                   const float *__restrict__ g_prob,
                                                                   float *__restrict__ grad_logits,
                                                                   float *__restrict__ part_gx) {
    const int64_t rows = B * N;
    const float glp = *g_lp;
    float gx = 0.0f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(r % N);
        float s[K], raw[K], m, lsum;
        row_softmax_thread<K>(xw + (size_t)r * K, bias + (size_t)n * K, s, raw, m, lsum);
        const int kc = (int)idx[r];
        const float inv = expf(-lsum);
        float dotc = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s[k] = expf(s[k]) * inv;
            dotc += __ldg(g_prob + (size_t)n * K + k) * s[k];
        }
        float *g = grad_logits + (size_t)r * K;
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            float o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int k = 4 * q + t;
                o[t] = glp * ((k == kc ? 1.0f : 0.0f) - s[k]) + s[k] * (__ldg(g_prob + (size_t)n * K + k) - dotc);
                gx += o[t] * raw[k];
            }
            *reinterpret_cast<float4 *>(g + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    // fixed work assignment + shuffle tree: the per-warp partials are reproducible
    gx = seg_sum<32>(gx);
    if ((threadIdx.x & 31) == 0) part_gx[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = gx;
}

// K -> (segment width, elements per lane)
#define MCQ_LOSS_DISPATCH(K, CALL)                 \
    switch (K) {                                   \
        case 256: CALL(32, 8); break;              \
        case 128: CALL(32, 4); break;              \
        case 64: CALL(32, 2); break;               \
        case 32: CALL(32, 1); break;               \
        default: CALL(16, 1); break; /* K <= 16 */ \
    }

// ---- histogram of the chosen entries (quantization.py:227-231: the per-codebook counts behind index_entropy_loss).
// The reference scatters ones into a (B, N, K) tensor and averages it; a scatter_add of 65,536 x 8 ones into 128 bins
// (trainer phase 1) is 92 us of contended global atomics.  Here every CTA counts its frames in shared memory (integer
// atomics) and adds its bins to the global integer histogram once; a second tiny kernel converts to float.  Counts are
// integers, so the result does not depend on the order of the additions.
__global__ void __launch_bounds__(256) index_counts_kernel(const int64_t *__restrict__ idx, int64_t B, int N, int K,
                                                           unsigned *__restrict__ counts) {
    extern __shared__ unsigned hist[];
    const int NK = N * K;
    for (int i = threadIdx.x; i < NK; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t total = B * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int64_t k = idx[i];
        if (k >= 0 && k < K) atomicAdd(&hist[n * K + (int)k], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NK; i += blockDim.x) {
        const unsigned v = hist[i];
        if (v) atomicAdd(&counts[i], v);
    }
}

__global__ void counts_to_float_kernel(const unsigned *__restrict__ counts, int n, float *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (float)counts[i];
}

// ---- column sums of a tall matrix: out[c] = sum_r X[r][c] (the bias gradient grad_logits.sum(0): 65,536 x 128..1024).
// Stage 1: CTA j sums rows j, j + G, j + 2G, ... (thread = column group of 4, fixed order) into part[j][c]; stage 2 adds
// the G partial rows in ascending j: reproducible, no atomics.
constexpr int COLSUM_PARTS = 148 * 4;

__global__ void __launch_bounds__(256) colsum_part_kernel(const float *__restrict__ X, int64_t R, int C,
                                                          float *__restrict__ part) {
    const int C4 = C >> 2;
    // thread -> (column group g, row lane rl): blockDim.x = 256 threads cover min(C4, 256) groups x the rest as row lanes
    const int groups = C4 < 256 ? C4 : 256;
    const int rlanes = 256 / groups;
    const int g0 = threadIdx.x % groups, rl = threadIdx.x / groups;
    __shared__ float4 red[256];
    for (int base = 0; base < C4; base += groups) {  // uniform trip count: the loop holds barriers
        const int g = base + g0;
        const bool on = g < C4 && rl < rlanes;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on)
            for (int64_t r = (int64_t)blockIdx.x * rlanes + rl; r < R; r += (int64_t)gridDim.x * rlanes) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(X + (size_t)r * C) + g);
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
        red[threadIdx.x] = acc;
        __syncthreads();
        if (on && rl == 0) {
            for (int q = 1; q < rlanes; ++q) {  // fixed order over the row lanes
                const float4 v = red[q * groups + g0];
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
            reinterpret_cast<float4 *>(part + (size_t)blockIdx.x * C)[g] = acc;
        }
        __syncthreads();
    }
}

// a warp per column: lane l adds partial rows l, l + 32, ... in ascending order, then a fixed shuffle tree
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float *__restrict__ part, int parts, int C,
                                                            float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    float a = 0.0f;
    for (int j = lane; j < parts; j += 32) a += part[(size_t)j * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[c] = a;
}

}  // namespace

// number of frame streams (and so of partial rows) the forward kernel uses for a batch of B frames
int class_loss_streams(int64_t B, int N, int K) {
    int64_t segs = (int64_t)148 * 4 * 8;  // 4 CTAs of 8 warps per SM
    if (K == 16 || K == 32) segs = (int64_t)148 * 8 * 256 / 4;  // thread-per-row kernels: ~75 k threads
    int64_t st = segs / N;
    if (st > B / 2) st = B / 2;  // the partial rows live in a region of (B rounded up to 128) x N*K floats
    if (st < 1) st = 1;
    return (int)st;
}

int launch_class_loss_fwd(const float *xw, const float *bias, const int64_t *idx, int64_t B, int N, int K,
                          float *part_prob, float *part_lp, float *prob_sum, float *logprob_sum, cudaStream_t st) {
    if (K > 256) {
        set_error("class loss: codebook_size %d > 256", K);
        return MCQ_EUNSUPPORTED;
    }
    const int ns = class_loss_streams(B, N, K);
    if ((K == 16 || K == 32) && 256 % N == 0) {
        const int64_t threads = (int64_t)ns * N;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        if (K == 16)
            class_loss_fwd_small_kernel<16><<<blocks, 256, 0, st>>>(xw, bias, idx, B, N, ns, part_prob, part_lp);
        else
            class_loss_fwd_small_kernel<32><<<blocks, 256, 0, st>>>(xw, bias, idx, B, N, ns, part_prob, part_lp);
        MCQ_LAUNCH_CHECK("class_loss_fwd_small_kernel");
        class_loss_reduce_kernel<<<(N * K + 31) / 32, 1024, 0, st>>>(part_prob, part_lp, (int)blocks, N, K, prob_sum,
                                                                     logprob_sum);
        MCQ_LAUNCH_CHECK("class_loss_reduce_kernel");
        return MCQ_OK;
    }
    const int spw = K >= 32 ? 1 : 2;
    const int warps = (ns * N + spw - 1) / spw;
    const int blocks = (warps + 7) / 8;
#define MCQ_FWD(W, EPL) \
    class_loss_fwd_kernel<W, EPL><<<blocks, 256, 0, st>>>(xw, bias, idx, B, N, K, ns, part_prob, part_lp)
    MCQ_LOSS_DISPATCH(K, MCQ_FWD)
#undef MCQ_FWD
    MCQ_LAUNCH_CHECK("class_loss_fwd_kernel");
    class_loss_reduce_kernel<<<(N * K + 31) / 32, 1024, 0, st>>>(part_prob, part_lp, ns, N, K, prob_sum, logprob_sum);
    MCQ_LAUNCH_CHECK("class_loss_reduce_kernel");
    return MCQ_OK;
}

int class_loss_bwd_partials() { return 148 * 8 * 8; }

int launch_class_loss_bwd(const float *xw, const float *bias, const int64_t *idx, int64_t B, int N, int K,
                          const float *g_lp, const float *g_prob, float *grad_logits, float *part_gx,
                          cudaStream_t st) {
    if (K > 256) {
        set_error("class loss: codebook_size %d > 256", K);
        return MCQ_EUNSUPPORTED;
    }
    const int spw = K >= 32 ? 1 : 2;
    int64_t blocks = (B * N + 8 * spw - 1) / (8 * spw);
    if (blocks > 148 * 8) blocks = 148 * 8;
    MCQ_CUDA(cudaMemsetAsync(part_gx, 0, sizeof(float) * class_loss_bwd_partials(), st));  // unused warps stay 0
    if ((K == 16 || K == 32) && 256 % N == 0) {
        int64_t nb = (B * N + 255) / 256;
        if (nb > 148 * 8) nb = 148 * 8;
        if (K == 16)
            class_loss_bwd_small_kernel<16><<<(unsigned)nb, 256, 0, st>>>(xw, bias, idx, B, N, g_lp, g_prob, grad_logits, part_gx);
        else
            class_loss_bwd_small_kernel<32><<<(unsigned)nb, 256, 0, st>>>(xw, bias, idx, B, N, g_lp, g_prob, grad_logits, part_gx);
        MCQ_LAUNCH_CHECK("class_loss_bwd_small_kernel");
        return MCQ_OK;
    }
#define MCQ_BWD(W, EPL)                                                                                              \
    class_loss_bwd_kernel<W, EPL><<<(unsigned)blocks, 256, 0, st>>>(xw, bias, idx, B, N, K, g_lp, g_prob, grad_logits, \
                                                                    part_gx)
    MCQ_LOSS_DISPATCH(K, MCQ_BWD)
#undef MCQ_BWD
    MCQ_LAUNCH_CHECK("class_loss_bwd_kernel");
    return MCQ_OK;
}

}  // namespace mcq

namespace mcq {

int launch_index_counts(const int64_t *idx, int64_t B, int N, int K, float *counts, unsigned *scratch, cudaStream_t st) {
    const int NK = N * K;
    if ((size_t)NK * sizeof(unsigned) > 48 * 1024) {
        set_error("index counts: %d bins do not fit the shared-memory histogram", NK);
        return MCQ_EUNSUPPORTED;
    }
    MCQ_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned) * NK, st));
    int64_t blocks = (B * N + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    index_counts_kernel<<<(unsigned)blocks, 256, NK * sizeof(unsigned), st>>>(idx, B, N, K, scratch);
    MCQ_LAUNCH_CHECK("index_counts_kernel");
    counts_to_float_kernel<<<(NK + 255) / 256, 256, 0, st>>>(scratch, NK, counts);
    MCQ_LAUNCH_CHECK("counts_to_float_kernel");
    return MCQ_OK;
}

int column_sum_partials(int C) { return COLSUM_PARTS * C; }

int launch_column_sums(const float *X, int64_t R, int C, float *out, float *part, cudaStream_t st) {
    int parts = COLSUM_PARTS;
    if (parts > R) parts = (int)(R > 0 ? R : 1);
    colsum_part_kernel<<<parts, 256, 0, st>>>(X, R, C, part);
    MCQ_LAUNCH_CHECK("colsum_part_kernel");
    colsum_reduce_kernel<<<(C + 7) / 8, 256, 0, st>>>(part, parts, C, out);
    MCQ_LAUNCH_CHECK("colsum_reduce_kernel");
    return MCQ_OK;
}

}  // namespace mcq
