// decode.cu -- Quantizer.decode (quantization.py:117-148) as one fused gather-sum kernel, its gradient w.r.t. the
// scaled centers, and the index (un)packing of Quantizer.encode (:266-272) / _maybe_separate_indexes (:551-573).
//
// decode is HBM bound: per frame it reads ncols code bytes and writes D output elements; the scaled codebooks
// (N*K*D*4 bytes, 1-64 MB) are read through L2.  One warp owns one frame at a time; each lane owns 4 consecutive
// features per 128-feature slab and adds the N selected rows in codebook order n = 0..N-1 in fp32, which is the
// reference's sum(dim=0) order bit for bit (for N <= 16; torch re-associates longer sums, tests allow 1e-5 there).
#include <stdlib.h>

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

// ---- slab decode: large batches of byte codes ---------------------------------------------------------------------
// decode_kernel above reads every selected row through L1 from L2: 16 KB of gathers per frame at 8 codebooks x 512
// features for 2 KB written, and at ~1.1 Gvec/s it sits on the L2 -> SM rate (18 TB/s), not on HBM
// (profiles/r01_search2.md).  Here a CTA owns a SLAB of 32 features: the 128-byte slab segments of all rows of the
// first NS codebooks live in its shared memory (K * 128 bytes per codebook: 7 of 8 codebooks of 256 entries fit the
// 227 KB; codebooks NS..N-1 are still read through L1), and the frames stream through it.  A quarter warp owns one
// frame (8 lanes x float4 = the 128 bytes of the slab), so every shared-memory request is one conflict-free wavefront
// per frame-row and every store a full 128-byte line.  Same arithmetic as decode_kernel: rows added in codebook order
// n = 0..N-1 in fp32 -- bit-identical output.  L2 traffic per frame drops from N * D * 4 to (N - NS) * D * 4 bytes;
// what remains is the shared-memory pipe itself (128 B/clk/SM: N loads + 1 store per 128 bytes written).
constexpr int SLAB_W = 32;               // features per slab
constexpr int SLAB_THREADS = 1024;       // 32 warps: 128 frames in flight per CTA
constexpr int SLAB_SMEM_MAX = 227 * 1024 - 1024;

// KT = 256: codebook_size known at compile time (row offsets become immediates and bytes cannot be out of range);
// KT = 0: run-time codebook_size K <= 256.
template <typename OT, int NT, int NS, int KT>
__global__ void __launch_bounds__(SLAB_THREADS, 1)
    decode_slab_kernel(const uint8_t *__restrict__ codes, int64_t B, int Krt, int D, int cps,
                       const float *__restrict__ cs, OT *__restrict__ out) {
    const int K = KT > 0 ? KT : Krt;
    extern __shared__ __align__(16) float slab[];  // [NS * K][32]
    const int nslabs = D / SLAB_W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fq = lane >> 3, sub = lane & 7;  // frame within the quad, float4 within the slab segment
    const int64_t nquads = (B + 3) >> 2;
    const unsigned kmask = (unsigned)(K - 1);
    for (int u = blockIdx.x; u < nslabs * cps; u += gridDim.x) {
        const int sl = u / cps, part = u - sl * cps;
        const int d0 = sl * SLAB_W;
        __syncthreads();  // the previous slab is no longer read
        for (int i = threadIdx.x; i < NS * K * (SLAB_W / 4); i += SLAB_THREADS) {
            const int row = i >> 3, c4 = i & 7;
            reinterpret_cast<float4 *>(slab)[i] = __ldg(reinterpret_cast<const float4 *>(cs + (size_t)row * D + d0) + c4);
        }
        __syncthreads();
        const float *gbase = cs + d0 + sub * 4;
        const float *sbase = slab + sub * 4;
        // quads of this part: q = part * 32 + warp, step cps * 32
        int64_t q = (int64_t)part * 32 + warp;
        const int64_t qstep = (int64_t)cps * 32;
        unsigned c_lo = 0, c_hi = 0;
        auto load_codes = [&](int64_t qq, unsigned &lo, unsigned &hi) {
            int64_t b = qq * 4 + fq;
            if (b >= B) b = B - 1;
            if constexpr (NT == 8) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(codes + (size_t)b * 8));
                lo = v.x;
                hi = v.y;
            } else {
                lo = __ldg(reinterpret_cast<const unsigned *>(codes + (size_t)b * 4));
                hi = 0;
            }
        };
        if (q < nquads) load_codes(q, c_lo, c_hi);
        for (; q < nquads; q += qstep) {
            unsigned n_lo = 0, n_hi = 0;
            if (q + qstep < nquads) load_codes(q + qstep, n_lo, n_hi);  // next quad's codes ahead of this quad's rows
            float4 c[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                unsigned k = __byte_perm(n < 4 ? c_lo : c_hi, 0, 0x4440 | (n & 3));  // byte n of the frame's codes
                if (KT != 256 && k > kmask) k = 0;  // out-of-range codes decode as entry 0, like decode_kernel
                if (n < NS)  // NS is a template parameter: resolved when the loop is unrolled
                    c[n] = *reinterpret_cast<const float4 *>(sbase + k * SLAB_W + n * K * SLAB_W);
                else
                    c[n] = __ldg(reinterpret_cast<const float4 *>(gbase + (size_t)(n * K + (int)k) * D));
            }
            float4 acc = c[0];
#pragma unroll
            for (int n = 1; n < NT; ++n) {
                acc.x = acc.x + c[n].x;
                acc.y = acc.y + c[n].y;
                acc.z = acc.z + c[n].z;
                acc.w = acc.w + c[n].w;
            }
            const int64_t b = q * 4 + fq;
            if (b < B) store4<OT>(out + (size_t)b * D + d0 + sub * 4, acc);
            c_lo = n_lo;
            c_hi = n_hi;
        }
    }
}

// frames from which the slab kernel is used (below it the fill of the slabs and the idle SMs of a small grid cost more
// than the L2 gathers; MCQ_DECODE_SLAB=0 / 1 forces the choice for A/B measurements)
constexpr int64_t SLAB_MIN_FRAMES = 16384;

template <typename OT>
int try_launch_decode_slab(const uint8_t *codes, int64_t B, int ncols, int N, int K, int D, const float *cs, OT *out,
                           cudaStream_t st, bool *launched) {
    *launched = false;
    static int force = -1;
    if (force < 0) {
        const char *e = getenv("MCQ_DECODE_SLAB");
        force = e ? (atoi(e) ? 1 : 0) : 2;
    }
    if (force == 0) return MCQ_OK;
    if (ncols != N || (N != 4 && N != 8) || D % SLAB_W != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(codes) % N != 0 || reinterpret_cast<uintptr_t>(cs) % 16 != 0)
        return MCQ_OK;
    if (force == 2 && B < SLAB_MIN_FRAMES) return MCQ_OK;
    const int per_cb = K * SLAB_W * (int)sizeof(float);
    int NS = SLAB_SMEM_MAX / per_cb;
    if (NS > N) NS = N;
    if (NS < N - 1 || NS < 1) return MCQ_OK;  // at most one codebook left to the L2 path
    const int smem = NS * per_cb;
    int dev = 0, sms = 148;
    MCQ_CUDA(cudaGetDevice(&dev));
    MCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int nslabs = D / SLAB_W;
    const int cps = sms / nslabs > 0 ? sms / nslabs : 1;  // CTAs per slab: 9 x 16 slabs = 144 CTAs at 512 features
    int grid = nslabs * cps;
    if (grid > sms) grid = sms;
    auto launch = [&](auto kern) -> int {
        MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, SLAB_THREADS, smem, st>>>(codes, B, K, D, cps, cs, out);
        MCQ_LAUNCH_CHECK("decode_slab_kernel");
        return MCQ_OK;
    };
    int rc;
    if (N == 8 && K == 256)
        rc = launch(decode_slab_kernel<OT, 8, 7, 256>);  // 7 x 32 KB of slab rows, the last codebook through L1
    else if (N == 4 && K == 256)
        rc = launch(decode_slab_kernel<OT, 4, 4, 256>);
    else if (N == 8 && NS == 8)
        rc = launch(decode_slab_kernel<OT, 8, 8, 0>);
    else if (N == 4 && NS == 4)
        rc = launch(decode_slab_kernel<OT, 4, 4, 0>);
    else
        return MCQ_OK;
    if (rc == MCQ_OK) *launched = true;
    return rc;
}

template <typename CT, typename OT>
int launch_decode_t(const CT *codes, int64_t B, int ncols, int N, int K, int D, const float *cs, OT *out,
                    cudaStream_t st) {
    if constexpr (sizeof(CT) == 1) {
        bool launched = false;
        const int rc = try_launch_decode_slab<OT>(reinterpret_cast<const uint8_t *>(codes), B, ncols, N, K, D, cs, out, st,
                                                  &launched);
        if (rc != MCQ_OK || launched) return rc;
    }
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
