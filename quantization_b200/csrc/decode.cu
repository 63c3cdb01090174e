// decode.cu -- Quantizer.decode (quantization.py:117-148) as one fused gather-sum kernel, its gradient w.r.t. the
// scaled centers, and the index (un)packing of Quantizer.encode (:266-272) / _maybe_separate_indexes (:551-573).
//
// decode is HBM bound: per frame it reads ncols code bytes and writes D output elements; the scaled codebooks
// (N*K*D*4 bytes, 1-64 MB) are read through L2.  One warp owns one frame at a time; each lane owns 4 consecutive
// features per 128-feature slab and adds the N selected rows in codebook order n = 0..N-1 in fp32, which is the
// reference's sum(dim=0) order bit for bit (for N <= 16; torch re-associates longer sums, tests allow 1e-5 there).
#include "common.cuh"

namespace mcq {

namespace {

__device__ __forceinline__ int log2i(int v) { return 31 - __clz(v); }

// Reads the index of codebook n for one frame from a row of codes.  r = N / ncols sub-indexes per column,
// sub-index j of column c is (v >> (j * log2 K)) & (K-1)   (quantization.py:566-573 with K a power of two).
template <typename CT>
__device__ __forceinline__ int read_index(const CT *row, int n, int r, int lgK, int K) {
    if (r == 1) return (int)row[n];
    const int c = n / r, j = n - c * r;
    const unsigned v = (unsigned)row[c];
    return (int)((v >> (j * lgK)) & (unsigned)(K - 1));
}

template <typename OT> __device__ __forceinline__ void store4(OT *p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float *p, float4 v) {
    __stcs(reinterpret_cast<float4 *>(p), v);
}
template <> __device__ __forceinline__ void store4<__half>(__half *p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned *>(&a);
    u.y = *reinterpret_cast<unsigned *>(&b);
    __stcs(reinterpret_cast<uint2 *>(p), u);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned *>(&a);
    u.y = *reinterpret_cast<unsigned *>(&b);
    __stcs(reinterpret_cast<uint2 *>(p), u);
}
template <typename OT> __device__ __forceinline__ void store1(OT *p, float v);
template <> __device__ __forceinline__ void store1<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void store1<__half>(__half *p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
    *p = __float2bfloat16_rn(v);
}

constexpr int DEC_MAXN = 64;

// NT > 0: num_codebooks known at compile time (all N <= 8 row loads of a slab are in flight before the first add;
// measured +12 % at N = 8, -9 % at N = 16 where 64 registers of loads cost occupancy, so 16 stays generic);
// NT == 0: generic loop.  The adds stay in codebook order n = 0..N-1 either way.
template <typename CT, typename OT, bool VEC4, int NT>
__global__ void __launch_bounds__(256) decode_kernel(const CT *__restrict__ codes, int64_t B, int ncols, int Nrt, int K,
                                                     int D, const float *__restrict__ cs, OT *__restrict__ out) {
    __shared__ int sidx[8][DEC_MAXN];
    const int N = NT > 0 ? NT : Nrt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = N / ncols, lgK = log2i(K);
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        const CT *row = codes + (size_t)b * ncols;
        for (int n = lane; n < N; n += 32) {
            int k = read_index<CT>(row, n, r, lgK, K);
            if (k < 0 || k >= K) k = 0;  // out-of-range codes are rejected by the host layer before the launch
            sidx[warp][n] = (n * K + k) * D;  // element offset of the selected row
        }
        __syncwarp();
        OT *o = out + (size_t)b * D;
        if (VEC4) {
            if constexpr (NT > 0) {
                int ro[NT];
#pragma unroll
                for (int n = 0; n < NT; ++n) ro[n] = sidx[warp][n];
                for (int d = lane * 4; d < D; d += 128) {
                    float4 c[NT];
#pragma unroll
                    for (int n = 0; n < NT; ++n) c[n] = __ldg(reinterpret_cast<const float4 *>(cs + ro[n] + d));
                    float4 acc = c[0];
#pragma unroll
                    for (int n = 1; n < NT; ++n) {
                        acc.x = acc.x + c[n].x;
                        acc.y = acc.y + c[n].y;
                        acc.z = acc.z + c[n].z;
                        acc.w = acc.w + c[n].w;
                    }
                    store4<OT>(o + d, acc);
                }
            } else {
                for (int d = lane * 4; d < D; d += 128) {
                    float4 acc = __ldg(reinterpret_cast<const float4 *>(cs + sidx[warp][0] + d));
                    for (int n = 1; n < N; ++n) {
                        const float4 c = __ldg(reinterpret_cast<const float4 *>(cs + sidx[warp][n] + d));
                        acc.x = acc.x + c.x;
                        acc.y = acc.y + c.y;
                        acc.z = acc.z + c.z;
                        acc.w = acc.w + c.w;
                    }
                    store4<OT>(o + d, acc);
                }
            }
        } else {
            for (int d = lane; d < D; d += 32) {
                float acc = __ldg(cs + sidx[warp][0] + d);
                for (int n = 1; n < N; ++n) acc = acc + __ldg(cs + sidx[warp][n] + d);
                store1<OT>(o + d, acc);
            }
        }
        __syncwarp();
    }
}

template <typename CT, typename OT>
int launch_decode_t(const CT *codes, int64_t B, int ncols, int N, int K, int D, const float *cs, OT *out,
                    cudaStream_t st) {
    int64_t blocks = (B + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const unsigned g = (unsigned)blocks;
    if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(out) % 16 == 0)) {
        switch (N) {
            case 2: decode_kernel<CT, OT, true, 2><<<g, 256, 0, st>>>(codes, B, ncols, N, K, D, cs, out); break;
            case 4: decode_kernel<CT, OT, true, 4><<<g, 256, 0, st>>>(codes, B, ncols, N, K, D, cs, out); break;
            case 8: decode_kernel<CT, OT, true, 8><<<g, 256, 0, st>>>(codes, B, ncols, N, K, D, cs, out); break;
            default: decode_kernel<CT, OT, true, 0><<<g, 256, 0, st>>>(codes, B, ncols, N, K, D, cs, out); break;
        }
    } else {
        decode_kernel<CT, OT, false, 0><<<g, 256, 0, st>>>(codes, B, ncols, N, K, D, cs, out);
    }
    MCQ_LAUNCH_CHECK("decode_kernel");
    return MCQ_OK;
}

template <typename CT>
int launch_decode_c(const CT *codes, int64_t B, int ncols, int N, int K, int D, const float *cs, void *out,
                    int out_dtype, cudaStream_t st) {
    switch (out_dtype) {
        case MCQ_F32: return launch_decode_t<CT, float>(codes, B, ncols, N, K, D, cs, (float *)out, st);
        case MCQ_F16: return launch_decode_t<CT, __half>(codes, B, ncols, N, K, D, cs, (__half *)out, st);
        case MCQ_BF16:
            return launch_decode_t<CT, __nv_bfloat16>(codes, B, ncols, N, K, D, cs, (__nv_bfloat16 *)out, st);
        default: break;
    }
    set_error("decode: unknown output dtype %d", out_dtype);
    return MCQ_EINVAL;
}

// grad[n, idx[b,n], :] += grad_out[b, :]
__global__ void __launch_bounds__(256) decode_backward_kernel(const float *__restrict__ go,
                                                              const int64_t *__restrict__ idx, int64_t B, int N, int K,
                                                              int D, float *__restrict__ grad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
        const float *g = go + (size_t)b * D;
        for (int n = 0; n < N; ++n) {
            int64_t k = idx[(size_t)b * N + n];
            if (k < 0 || k >= K) continue;
            float *dst = grad + ((size_t)n * K + (size_t)k) * D;
            for (int d = lane; d < D; d += 32) atomicAdd(dst + d, g[d]);
        }
    }
}

// (B, N) int32 -> packed bytes (quantization.py:266-272) or (B, N) int64 / int32
__global__ void pack_kernel(const int32_t *__restrict__ idx, int64_t B, int N, int K, int ncols, void *codes,
                            int codes_dtype) {
    const int r = N / ncols, lgK = log2i(K);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B * ncols;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / ncols;
        const int c = (int)(i - b * ncols);
        if (codes_dtype == MCQ_U8) {
            unsigned v = 0;
            for (int j = 0; j < r; ++j) v |= (unsigned)idx[(size_t)b * N + (size_t)c * r + j] << (j * lgK);
            ((uint8_t *)codes)[i] = (uint8_t)v;
        } else if (codes_dtype == MCQ_I64) {
            ((int64_t *)codes)[i] = idx[i];
        } else {
            ((int32_t *)codes)[i] = idx[i];
        }
    }
}

__global__ void i64_to_i32_kernel(const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n, int K) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = src[i];
        dst[i] = (int32_t)(v < 0 ? 0 : (v >= K ? K - 1 : v));
    }
}

__global__ void i32_to_i64_kernel(const int32_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

unsigned grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace

int launch_decode(const void *codes, int codes_dtype, int64_t B, int ncols, int N, int K, int D, const float *cs,
                  void *out, int out_dtype, cudaStream_t st) {
    if (B <= 0) return MCQ_OK;
    if (N > DEC_MAXN) {
        set_error("decode: num_codebooks %d > %d", N, DEC_MAXN);
        return MCQ_EUNSUPPORTED;
    }
    switch (codes_dtype) {
        case MCQ_U8: return launch_decode_c<uint8_t>((const uint8_t *)codes, B, ncols, N, K, D, cs, out, out_dtype, st);
        case MCQ_I64:
            return launch_decode_c<int64_t>((const int64_t *)codes, B, ncols, N, K, D, cs, out, out_dtype, st);
        case MCQ_I32:
            return launch_decode_c<int32_t>((const int32_t *)codes, B, ncols, N, K, D, cs, out, out_dtype, st);
        default: break;
    }
    set_error("decode: unknown codes dtype %d", codes_dtype);
    return MCQ_EINVAL;
}

int launch_decode_backward(const float *grad_out, const int64_t *idx, int64_t B, int N, int K, int D, float *grad,
                           cudaStream_t st) {
    if (B <= 0) return MCQ_OK;
    int64_t blocks = (B + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    decode_backward_kernel<<<(unsigned)blocks, 256, 0, st>>>(grad_out, idx, B, N, K, D, grad);
    MCQ_LAUNCH_CHECK("decode_backward_kernel");
    return MCQ_OK;
}

int launch_pack(const int32_t *idx, int64_t B, int N, int K, void *codes, int codes_dtype, cudaStream_t st) {
    if (B <= 0) return MCQ_OK;
    const int ncols = codes_dtype == MCQ_U8 ? mcq_packed_cols(N, K) : N;
    pack_kernel<<<grid_for(B * ncols), 256, 0, st>>>(idx, B, N, K, ncols, codes, codes_dtype);
    MCQ_LAUNCH_CHECK("pack_kernel");
    return MCQ_OK;
}

int launch_i64_to_i32(const int64_t *src, int32_t *dst, int64_t n, int K, cudaStream_t st) {
    if (n <= 0) return MCQ_OK;
    i64_to_i32_kernel<<<grid_for(n), 256, 0, st>>>(src, dst, n, K);
    MCQ_LAUNCH_CHECK("i64_to_i32_kernel");
    return MCQ_OK;
}

int launch_i32_to_i64(const int32_t *src, int64_t *dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return MCQ_OK;
    i32_to_i64_kernel<<<grid_for(n), 256, 0, st>>>(src, dst, n);
    MCQ_LAUNCH_CHECK("i32_to_i64_kernel");
    return MCQ_OK;
}

}  // namespace mcq
