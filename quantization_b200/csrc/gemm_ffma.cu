// gemm_ffma.cu -- plain fp32 CUDA-core GEMM  C (M, NK) = (a_scale * A) (M, D) . Bm (NK, D)^T.
// Used as the cross-check of the tcgen05 fp16x2 GEMM (gemm_tc.cu) in the tests and for shapes the tensor-core
// kernel does not tile (dim not a multiple of 64); also holds the classifier arg-max (quantization.py:297-301).
#include "common.cuh"

namespace mcq {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_ffma_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                        float *__restrict__ C, int64_t M, int NK, int D,
                                                        const float *__restrict__ a_scale) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * TM;
    const int n0 = blockIdx.x * TN;
    const float s = a_scale ? *a_scale : 1.0f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += TK) {
        for (int e = threadIdx.x; e < TM * TK; e += 256) {
            int row = e >> 4, kk = e & 15, k = k0 + kk;
            float a = 0.f, b = 0.f;
            if (k < D) {
                if (m0 + row < M) {
                    a = A[(size_t)(m0 + row) * D + k];
                    if (a_scale) a = __fmul_rn(s, a);  // quantization.py:278 scales x before the GEMM
                }
                if (n0 + row < NK) b = Bm[(size_t)(n0 + row) * D + k];
            }
            As[kk][row] = a;
            Bs[kk][row] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < NK) C[(size_t)m * NK + n] = acc[i][j];
        }
    }
}

int launch_gemm_ffma(const float *A, const float *Bm, float *C, int64_t M, int NK, int D, const float *a_scale,
                     cudaStream_t st) {
    if (M <= 0) return MCQ_OK;
    int64_t my = (M + TM - 1) / TM;
    // gridDim.y is limited to 65535: loop over slabs of rows
    const int64_t slab = 65535;
    for (int64_t y0 = 0; y0 < my; y0 += slab) {
        int64_t ny = my - y0 < slab ? my - y0 : slab;
        dim3 grid((NK + TN - 1) / TN, (unsigned)ny);
        int64_t row0 = y0 * TM;
        gemm_ffma_kernel<<<grid, 256, 0, st>>>(A + (size_t)row0 * D, Bm, C + (size_t)row0 * NK, M - row0, NK, D,
                                               a_scale);
        MCQ_LAUNCH_CHECK("gemm_ffma_kernel");
    }
    return MCQ_OK;
}

// idx[b, n] = argmax_k (logits[b, n*K + k] + bias[n*K + k]), first maximum on ties (quantization.py:297-301).
// One warp per (frame, codebook) pair.
__global__ void argmax_init_kernel(const float *__restrict__ logits, const float *__restrict__ bias, int64_t B, int N,
                                   int K, int32_t *__restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < B * N; item += nwarps) {
        const int64_t b = item / N;
        const int n = (int)(item - b * N);
        const float *row = logits + (size_t)b * N * K + (size_t)n * K;
        const float *bs = bias + (size_t)n * K;
        float best = 0.f;
        int bk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            float v = row[k] + bs[k];
            if (bk == 0x7fffffff || v > best) {
                best = v;
                bk = k;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, off);
            int ok = __shfl_xor_sync(0xffffffffu, bk, off);
            bool take = (ok != 0x7fffffff) && (bk == 0x7fffffff || ov > best || (ov == best && ok < bk));
            if (take) {
                best = ov;
                bk = ok;
            }
        }
        if (lane == 0) idx[item] = bk == 0x7fffffff ? 0 : bk;
    }
}

// Small codebooks (K = 16 or 32: trainer phase 1): one thread per (frame, codebook) pair, the K logits are K*4
// contiguous bytes read as float4 -- a warp per pair would leave half the lanes idle and spend its time in shuffles.
template <int K>
__global__ void __launch_bounds__(256) argmax_small_kernel(const float *__restrict__ logits, const float *__restrict__ bias,
                                                           int64_t items, int N, int32_t *__restrict__ idx) {
    for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(item % N);
        const float4 *row = reinterpret_cast<const float4 *>(logits + (size_t)item * K);
        const float4 *bs = reinterpret_cast<const float4 *>(bias + (size_t)n * K);
        float best = 0.f;
        int bk = 0;
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            const float4 v = __ldcs(row + q);
            const float4 c = __ldg(bs + q);
            const float e[4] = {v.x + c.x, v.y + c.y, v.z + c.z, v.w + c.w};
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if ((q == 0 && t == 0) || e[t] > best) {  // strict >: first maximum on ties
                    best = e[t];
                    bk = q * 4 + t;
                }
        }
        idx[item] = bk;
    }
}

int launch_argmax_init(const float *logits, const float *bias, int64_t B, int N, int K, int32_t *idx, cudaStream_t st) {
    if (B <= 0) return MCQ_OK;
    int64_t items = B * N;
    if (K == 16 || K == 32) {
        int64_t blocks = (items + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (K == 16)
            argmax_small_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(logits, bias, items, N, idx);
        else
            argmax_small_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(logits, bias, items, N, idx);
        MCQ_LAUNCH_CHECK("argmax_small_kernel");
        return MCQ_OK;
    }
    int64_t blocks = (items + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    argmax_init_kernel<<<(unsigned)blocks, 256, 0, st>>>(logits, bias, B, N, K, idx);
    MCQ_LAUNCH_CHECK("argmax_init_kernel");
    return MCQ_OK;
}

}  // namespace mcq
