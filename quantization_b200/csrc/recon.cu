// recon.cu -- the reconstruction term of Quantizer.compute_loss (quantization.py:209-216) without its (B, dim)
// intermediates:   rel = sum (x_hat - x)^2 / (sum (x - mean)^2 + 1e-20),   x_hat = decode(indexes).
// The reference (and a plain PyTorch mirror) materialises x_hat, the difference, both squares and, in backward, the
// gradient of x_hat -- about twenty passes over (B, dim) fp32 tensors per trainer step.  Here:
//   forward : one pass -- gather-sum the N scaled centers of a frame (the decode order, n ascending), subtract x (read
//             in its own dtype), accumulate both sums; fixed work assignment + fixed-order reduction (reproducible);
//   backward: one pass -- recompute x_hat - x, scale by the (device-resident) coefficient and scatter-add straight
//             into the gradient of the scaled centers (128-bit vector atomics when dim is a multiple of 128).
//             (Measured and dropped: a private copy of the gradient per CTA in shared memory for small tables -- float
//             atomics on shared memory compile to compare-and-swap loops, 177 us against 130 us for global atomics
//             at 65,536 frames, N*K = 128.)
#include <stdlib.h>

#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RECON_BLOCKS = 148 * 8;  // 8 CTAs of 8 warps per SM
constexpr int RECON_WARPS = RECON_BLOCKS * 8;

template <typename T>
__device__ __forceinline__ float ldx(const T *p);
template <>
__device__ __forceinline__ float ldx<float>(const float *p) { return __ldcs(p); }
template <>
__device__ __forceinline__ float ldx<__half>(const __half *p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float ldx<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// x_hat[d] - x[d] for d = lane, lane + 32, ... (up to MAXC values per lane), x_hat summed n = 0..N-1 like decode
template <typename T, int MAXC, typename R>
__device__ __forceinline__ void frame_error(const T *__restrict__ xb, const float *__restrict__ cs, const R *rows,
                                            int N, int D, int lane, float (&e)[MAXC], float (&xv)[MAXC]) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int d = c * 32 + lane;
        float acc = 0.0f;
        if (d < D) {
            for (int n = 0; n < N; ++n) acc = acc + __ldg(cs + (size_t)rows[n] * D + d);
            xv[c] = ldx<T>(xb + d);
        } else {
            xv[c] = 0.0f;
        }
        e[c] = acc - xv[c];
    }
}

template <typename T, int MAXC>
__global__ void __launch_bounds__(256) recon_fwd_kernel(const T *__restrict__ x, const int64_t *__restrict__ idx, int64_t B,
                                                        int N, int K, int D, const float *__restrict__ cs,
                                                        const float *__restrict__ mean, float *__restrict__ partials) {
    __shared__ long long rows_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *rows = rows_s[warp];
    float mu[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) mu[c] = (c * 32 + lane < D) ? __ldg(mean + c * 32 + lane) : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        for (int n = lane; n < N; n += 32) {
            long long k = idx[(size_t)b * N + n];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            rows[n] = (long long)n * K + k;
        }
        __syncwarp();
        float e[MAXC], xv[MAXC];
        frame_error<T, MAXC, long long>(x + (size_t)b * D, cs, rows, N, D, lane, e, xv);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            s1 = fmaf(e[c], e[c], s1);
            const float dm = (c * 32 + lane < D) ? xv[c] - mu[c] : 0.0f;
            s2 = fmaf(dm, dm, s2);
        }
        __syncwarp();
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        partials[2 * (blockIdx.x * 8 + warp)] = s1;
        partials[2 * (blockIdx.x * 8 + warp) + 1] = s2;
    }
}

__global__ void __launch_bounds__(1024) recon_reduce_kernel(const float *__restrict__ partials, int nwarps,
                                                            float *__restrict__ sums) {
    __shared__ double sh[2][32];
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < nwarps; i += 1024) {
        a += (double)partials[2 * i];
        c += (double)partials[2 * i + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        c += __shfl_xor_sync(FULL, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = a;
        sh[1][threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = sh[0][threadIdx.x];
        c = sh[1][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(FULL, a, o);
            c += __shfl_xor_sync(FULL, c, o);
        }
        if (threadIdx.x == 0) {
            sums[0] = (float)a;
            sums[1] = (float)c;
        }
    }
}

template <typename T, int MAXC>
__global__ void __launch_bounds__(256) recon_bwd_kernel(const T *__restrict__ x, const int64_t *__restrict__ idx, int64_t B,
                                                        int N, int K, int D, const float *__restrict__ cs,
                                                        const float *__restrict__ coef, float *__restrict__ grad) {
    __shared__ long long rows_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *rows = rows_s[warp];
    const float cf = __ldg(coef);
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        for (int n = lane; n < N; n += 32) {
            long long k = idx[(size_t)b * N + n];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            rows[n] = (long long)n * K + k;
        }
        __syncwarp();
        float e[MAXC], xv[MAXC];
        frame_error<T, MAXC, long long>(x + (size_t)b * D, cs, rows, N, D, lane, e, xv);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int d = c * 32 + lane;
            if (d < D) {
                const float g = cf * e[c];
                for (int n = 0; n < N; ++n) atomicAdd(grad + (size_t)rows[n] * D + d, g);
            }
        }
        __syncwarp();
    }
}

// ---- 128-bit path: dim a multiple of 128; lane owns the float4 at column (c * 32 + lane) * 4, c < VC = dim / 128 ----
template <typename T>
__device__ __forceinline__ float4 ld4(const T *p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
template <>
__device__ __forceinline__ float4 ld4<__half>(const __half *p) {
    const uint2 u = __ldcs(reinterpret_cast<const uint2 *>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 u = __ldcs(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}

template <typename T, int VC>
__device__ __forceinline__ void frame_error4(const T *__restrict__ xb, const float *__restrict__ cs, const int *rows, int N,
                                             int D, int lane, float4 (&e)[VC], float4 (&xv)[VC]) {
#pragma unroll
    for (int c = 0; c < VC; ++c) {
        e[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        xv[c] = ld4<T>(xb + (c * 32 + lane) * 4);
    }
    for (int n = 0; n < N; ++n) {
        const float4 *row = reinterpret_cast<const float4 *>(cs + (size_t)rows[n] * D);
#pragma unroll
        for (int c = 0; c < VC; ++c) {
            const float4 v = __ldg(row + c * 32 + lane);
            e[c].x += v.x;
            e[c].y += v.y;
            e[c].z += v.z;
            e[c].w += v.w;
        }
    }
#pragma unroll
    for (int c = 0; c < VC; ++c) {
        e[c].x -= xv[c].x;
        e[c].y -= xv[c].y;
        e[c].z -= xv[c].z;
        e[c].w -= xv[c].w;
    }
}

template <typename T, int VC>
__global__ void __launch_bounds__(256) recon_fwd4_kernel(const T *__restrict__ x, const int64_t *__restrict__ idx, int64_t B,
                                                         int N, int K, int D, const float *__restrict__ cs,
                                                         const float *__restrict__ mean, float *__restrict__ partials) {
    __shared__ int rows_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *rows = rows_s[warp];
    float4 mu[VC];
#pragma unroll
    for (int c = 0; c < VC; ++c) mu[c] = __ldg(reinterpret_cast<const float4 *>(mean) + c * 32 + lane);
    float s1 = 0.0f, s2 = 0.0f;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        for (int n = lane; n < N; n += 32) {
            long long k = idx[(size_t)b * N + n];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            rows[n] = n * K + (int)k;
        }
        __syncwarp();
        float4 e[VC], xv[VC];
        frame_error4<T, VC>(x + (size_t)b * D, cs, rows, N, D, lane, e, xv);
#pragma unroll
        for (int c = 0; c < VC; ++c) {
            s1 = fmaf(e[c].x, e[c].x, s1);
            s1 = fmaf(e[c].y, e[c].y, s1);
            s1 = fmaf(e[c].z, e[c].z, s1);
            s1 = fmaf(e[c].w, e[c].w, s1);
            const float a0 = xv[c].x - mu[c].x, a1 = xv[c].y - mu[c].y, a2 = xv[c].z - mu[c].z, a3 = xv[c].w - mu[c].w;
            s2 = fmaf(a0, a0, s2);
            s2 = fmaf(a1, a1, s2);
            s2 = fmaf(a2, a2, s2);
            s2 = fmaf(a3, a3, s2);
        }
        __syncwarp();
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        partials[2 * (blockIdx.x * 8 + warp)] = s1;
        partials[2 * (blockIdx.x * 8 + warp) + 1] = s2;
    }
}

template <typename T, int VC>
__global__ void __launch_bounds__(256) recon_bwd4_kernel(const T *__restrict__ x, const int64_t *__restrict__ idx, int64_t B,
                                                         int N, int K, int D, const float *__restrict__ cs,
                                                         const float *__restrict__ coef, float *__restrict__ grad) {
    __shared__ int rows_s[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *rows = rows_s[warp];
    const float cf = __ldg(coef);
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        for (int n = lane; n < N; n += 32) {
            long long k = idx[(size_t)b * N + n];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            rows[n] = n * K + (int)k;
        }
        __syncwarp();
        float4 e[VC], xv[VC];
        frame_error4<T, VC>(x + (size_t)b * D, cs, rows, N, D, lane, e, xv);
#pragma unroll
        for (int c = 0; c < VC; ++c) e[c] = make_float4(cf * e[c].x, cf * e[c].y, cf * e[c].z, cf * e[c].w);
        for (int n = 0; n < N; ++n) {
            float4 *row = reinterpret_cast<float4 *>(grad + (size_t)rows[n] * D);
#pragma unroll
            for (int c = 0; c < VC; ++c) atomicAdd(row + c * 32 + lane, e[c]);
        }
        __syncwarp();
    }
}

// ---- backward for small tables (codebook_size 16: trainer phase 1).  65,536 frames scatter into 128 rows, so atomics --
// global or shared -- serialise on a few addresses (149 us, ncu: issue 16 %).  Here the sums are formed in REGISTERS:
// a CTA stages coef * (x_hat - x) of 32 frames in shared memory; warp (n, h) owns codebook n and the h-th 128 columns,
// keeps the 16 rows of that codebook as 16 float4 accumulators per lane and adds each staged frame into the one its
// (warp-uniform) code selects; one round of vector atomics per CTA at the end.  Measured 119 us at 65,536 frames x 8
// codebooks (the atomic kernel: 142-149 us); what remains is the latency of staging a tile (16 warps per SM at 128
// registers, two barriers per 32 frames) -- double-buffering the tile is the next step.  (Measured and dropped: 32
// warps per CTA with float2 accumulators, i.e. half the register state per warp: 136 us.) ----
template <typename T, int VC>
__global__ void __launch_bounds__(512, 1) recon_bwd_k16_kernel(const T *__restrict__ x, const int64_t *__restrict__ idx,
                                                               int64_t B, int N, int D, const float *__restrict__ cs,
                                                               const float *__restrict__ coef, float *__restrict__ grad) {
    constexpr int K = 16, F = 32;
    extern __shared__ float4 tile[];              // [F][D / 4]
    __shared__ int rows_s[F][64];                 // n * K + code, per staged frame
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;           // == N * VC
    const int my_n = warp / VC, my_h = warp % VC;
    const int D4 = D >> 2;
    const float cf = __ldg(coef);
    float4 a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15;
    a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = a9 = a10 = a11 = a12 = a13 = a14 = a15 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t b0 = (int64_t)blockIdx.x * F; b0 < B; b0 += (int64_t)gridDim.x * F) {
        // codes of the whole tile in one coalesced pass (takes a dependent global round trip out of every frame)
        for (int i = threadIdx.x; i < F * N; i += blockDim.x) {
            const int f = i / N, n = i - f * N;
            const int64_t b = b0 + f;
            int r = -1;
            if (b < B) {
                long long k = idx[(size_t)b * N + n];
                k = k < 0 ? 0 : (k >= K ? K - 1 : k);
                r = n * K + (int)k;
            }
            rows_s[f][n] = r;
        }
        __syncthreads();
        // stage: warp w computes frames w, w + nwarps, ... of this tile
        for (int f = warp; f < F; f += nwarps) {
            const int64_t b = b0 + f;
            if (b < B) {
                float4 e[VC], xv[VC];
                frame_error4<T, VC>(x + (size_t)b * D, cs, rows_s[f], N, D, lane, e, xv);
#pragma unroll
                for (int c = 0; c < VC; ++c)
                    tile[f * D4 + c * 32 + lane] = make_float4(cf * e[c].x, cf * e[c].y, cf * e[c].z, cf * e[c].w);
            }
        }
        __syncthreads();
        // accumulate: my codebook, my 128 columns
        if (my_n < N) {
#pragma unroll 4
            for (int f = 0; f < F; ++f) {
                const int r = rows_s[f][my_n];
                if (r < 0) continue;
                const float4 v = tile[f * D4 + my_h * 32 + lane];
#define MCQ_ACC(i) case i: a##i.x += v.x; a##i.y += v.y; a##i.z += v.z; a##i.w += v.w; break;
                switch (r - my_n * K) {
                    MCQ_ACC(0) MCQ_ACC(1) MCQ_ACC(2) MCQ_ACC(3) MCQ_ACC(4) MCQ_ACC(5) MCQ_ACC(6) MCQ_ACC(7)
                    MCQ_ACC(8) MCQ_ACC(9) MCQ_ACC(10) MCQ_ACC(11) MCQ_ACC(12) MCQ_ACC(13) MCQ_ACC(14) MCQ_ACC(15)
                    default: break;
                }
#undef MCQ_ACC
            }
        }
        __syncthreads();
    }
    if (my_n < N) {
        float4 *g = reinterpret_cast<float4 *>(grad + (size_t)my_n * K * D) + my_h * 32 + lane;
#define MCQ_FLUSH(i) atomicAdd(g + (size_t)i * D4, a##i);
        MCQ_FLUSH(0) MCQ_FLUSH(1) MCQ_FLUSH(2) MCQ_FLUSH(3) MCQ_FLUSH(4) MCQ_FLUSH(5) MCQ_FLUSH(6) MCQ_FLUSH(7)
        MCQ_FLUSH(8) MCQ_FLUSH(9) MCQ_FLUSH(10) MCQ_FLUSH(11) MCQ_FLUSH(12) MCQ_FLUSH(13) MCQ_FLUSH(14) MCQ_FLUSH(15)
#undef MCQ_FLUSH
    }
}

template <typename T>
int recon_fwd_t(const T *x, const int64_t *idx, int64_t B, int N, int K, int D, const float *cs, const float *mean,
                float *sums, float *partials, cudaStream_t st) {
    int64_t blocks = (B + 7) / 8;
    if (blocks > RECON_BLOCKS) blocks = RECON_BLOCKS;
    if (blocks < 1) blocks = 1;
    if (D % 128 == 0) {
#define MCQ_RF4(VC) recon_fwd4_kernel<T, VC><<<(unsigned)blocks, 256, 0, st>>>(x, idx, B, N, K, D, cs, mean, partials)
        switch (D / 128) {
            case 1: MCQ_RF4(1); break;
            case 2: MCQ_RF4(2); break;
            case 3: MCQ_RF4(3); break;
            case 4: MCQ_RF4(4); break;
            case 5: MCQ_RF4(5); break;
            case 6: MCQ_RF4(6); break;
            case 7: MCQ_RF4(7); break;
            default: MCQ_RF4(8); break;
        }
#undef MCQ_RF4
        MCQ_LAUNCH_CHECK("recon_fwd4_kernel");
        recon_reduce_kernel<<<1, 1024, 0, st>>>(partials, (int)blocks * 8, sums);
        MCQ_LAUNCH_CHECK("recon_reduce_kernel");
        return MCQ_OK;
    }
#define MCQ_RF(MAXC) recon_fwd_kernel<T, MAXC><<<(unsigned)blocks, 256, 0, st>>>(x, idx, B, N, K, D, cs, mean, partials)
    const int chunks = (D + 31) / 32;
    if (chunks <= 2) MCQ_RF(2);
    else if (chunks <= 4) MCQ_RF(4);
    else if (chunks <= 8) MCQ_RF(8);
    else if (chunks <= 16) MCQ_RF(16);
    else MCQ_RF(32);
#undef MCQ_RF
    MCQ_LAUNCH_CHECK("recon_fwd_kernel");
    recon_reduce_kernel<<<1, 1024, 0, st>>>(partials, (int)blocks * 8, sums);
    MCQ_LAUNCH_CHECK("recon_reduce_kernel");
    return MCQ_OK;
}

template <typename T, int MAXC>
int recon_bwd_launch(const T *x, const int64_t *idx, int64_t B, int N, int K, int D, const float *cs, const float *coef,
                     float *grad, cudaStream_t st) {
    int64_t blocks = (B + 7) / 8;
    if (blocks > RECON_BLOCKS) blocks = RECON_BLOCKS;
    recon_bwd_kernel<T, MAXC><<<(unsigned)blocks, 256, 0, st>>>(x, idx, B, N, K, D, cs, coef, grad);
    MCQ_LAUNCH_CHECK("recon_bwd_kernel");
    return MCQ_OK;
}

template <typename T>
int recon_bwd_t(const T *x, const int64_t *idx, int64_t B, int N, int K, int D, const float *cs, const float *coef,
                float *grad, cudaStream_t st) {
    if (K == 16 && D % 128 == 0 && (D / 128 <= 4 || D / 128 == 8) && N * (D / 128) <= 16 && B >= 8192 &&
        !getenv("MCQ_RECON_BWD_ATOMIC")) {
        const int vc = D / 128;
        const size_t smem = (size_t)32 * D * sizeof(float);
        int64_t blocks = (B + 31) / 32;
        if (blocks > 148) blocks = 148;
        const unsigned threads = (unsigned)(N * vc * 32);
#define MCQ_RBK(VC)                                                                                       \
    do {                                                                                                  \
        auto kern = recon_bwd_k16_kernel<T, VC>;                                                          \
        MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<(unsigned)blocks, threads, smem, st>>>(x, idx, B, N, D, cs, coef, grad);                   \
    } while (0)
        switch (vc) {
            case 1: MCQ_RBK(1); break;
            case 2: MCQ_RBK(2); break;
            case 3: MCQ_RBK(3); break;
            case 4: MCQ_RBK(4); break;
            default: MCQ_RBK(8); break;
        }
#undef MCQ_RBK
        MCQ_LAUNCH_CHECK("recon_bwd_k16_kernel");
        return MCQ_OK;
    }
    if (D % 128 == 0) {
        int64_t blocks = (B + 7) / 8;
        if (blocks > RECON_BLOCKS) blocks = RECON_BLOCKS;
#define MCQ_RB4(VC) recon_bwd4_kernel<T, VC><<<(unsigned)blocks, 256, 0, st>>>(x, idx, B, N, K, D, cs, coef, grad)
        switch (D / 128) {
            case 1: MCQ_RB4(1); break;
            case 2: MCQ_RB4(2); break;
            case 3: MCQ_RB4(3); break;
            case 4: MCQ_RB4(4); break;
            case 5: MCQ_RB4(5); break;
            case 6: MCQ_RB4(6); break;
            case 7: MCQ_RB4(7); break;
            default: MCQ_RB4(8); break;
        }
#undef MCQ_RB4
        MCQ_LAUNCH_CHECK("recon_bwd4_kernel");
        return MCQ_OK;
    }
    const int chunks = (D + 31) / 32;
    if (chunks <= 2) return recon_bwd_launch<T, 2>(x, idx, B, N, K, D, cs, coef, grad, st);
    if (chunks <= 4) return recon_bwd_launch<T, 4>(x, idx, B, N, K, D, cs, coef, grad, st);
    if (chunks <= 8) return recon_bwd_launch<T, 8>(x, idx, B, N, K, D, cs, coef, grad, st);
    if (chunks <= 16) return recon_bwd_launch<T, 16>(x, idx, B, N, K, D, cs, coef, grad, st);
    return recon_bwd_launch<T, 32>(x, idx, B, N, K, D, cs, coef, grad, st);
}

int check_recon(const char *who, int64_t B, int N, int K, int D) {
    if (B <= 0 || N < 1 || N > 64 || K < 1 || D < 1 || D > 1024) {
        set_error("%s: need num_frames > 0, 1 <= num_codebooks <= 64, 1 <= dim <= 1024 (got B=%lld N=%d K=%d D=%d)", who,
                  (long long)B, N, K, D);
        return D > 1024 ? MCQ_EUNSUPPORTED : MCQ_EINVAL;
    }
    return MCQ_OK;
}

}  // namespace

}  // namespace mcq

using namespace mcq;

extern "C" {

int mcq_recon_loss_partials(void) { return 2 * RECON_WARPS; }

int mcq_recon_loss_forward(const void *x, int x_dtype, const int64_t *idx, int64_t B, int N, int K, int D,
                           const float *scaled_centers, const float *mean, float *sums, float *partials, void *stream) {
    int rc = check_recon("mcq_recon_loss_forward", B, N, K, D);
    if (rc) return rc;
    if (!x || !idx || !scaled_centers || !mean || !sums || !partials) {
        set_error("mcq_recon_loss_forward: null pointer");
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (x_dtype) {
        case MCQ_F32: return recon_fwd_t<float>((const float *)x, idx, B, N, K, D, scaled_centers, mean, sums, partials, st);
        case MCQ_F16: return recon_fwd_t<__half>((const __half *)x, idx, B, N, K, D, scaled_centers, mean, sums, partials, st);
        case MCQ_BF16:
            return recon_fwd_t<__nv_bfloat16>((const __nv_bfloat16 *)x, idx, B, N, K, D, scaled_centers, mean, sums, partials,
                                              st);
        default: break;
    }
    set_error("mcq_recon_loss_forward: unknown dtype %d", x_dtype);
    return MCQ_EINVAL;
}

int mcq_recon_loss_backward(const void *x, int x_dtype, const int64_t *idx, int64_t B, int N, int K, int D,
                            const float *scaled_centers, const float *coef, float *grad_scaled_centers, void *stream) {
    int rc = check_recon("mcq_recon_loss_backward", B, N, K, D);
    if (rc) return rc;
    if (!x || !idx || !scaled_centers || !coef || !grad_scaled_centers) {
        set_error("mcq_recon_loss_backward: null pointer");
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (x_dtype) {
        case MCQ_F32: return recon_bwd_t<float>((const float *)x, idx, B, N, K, D, scaled_centers, coef, grad_scaled_centers, st);
        case MCQ_F16:
            return recon_bwd_t<__half>((const __half *)x, idx, B, N, K, D, scaled_centers, coef, grad_scaled_centers, st);
        case MCQ_BF16:
            return recon_bwd_t<__nv_bfloat16>((const __nv_bfloat16 *)x, idx, B, N, K, D, scaled_centers, coef,
                                              grad_scaled_centers, st);
        default: break;
    }
    set_error("mcq_recon_loss_backward: unknown dtype %d", x_dtype);
    return MCQ_EINVAL;
}

}  // extern "C"
