// gemm_tn.cu -- out (C1, C2) = A^T . Bm for A (R, C1) fp32 and Bm (R, C2) fp32 / fp16 / bf16, reduction over the R rows
// (frames): the shape of every weight gradient on the path -- d loss / d to_logits.weight = grad_logits^T . x in
// QuantizerTrainer.step (the reference's autograd forms it with an fp32 SGEMM, quantization.py:279 backward), and the
// linear2b / linear1 weight gradients of JointCodebookLoss.  Few output tiles, a reduction of 10^4..10^5: a split-K job.
//
// fp32-faithful on the tensor cores exactly like the forward products (gemm_tc.cu): each operand is written as a sum of
// two fp16 pieces after scaling by a power of two, three tcgen05 products are accumulated in fp32.  The tcgen05 kernel
// wants both operands with the reduction dimension contiguous, so the pack kernel here TRANSPOSES while it splits
// (tile through shared memory, coalesced on both sides).  The scale is one power of two per matrix (from its largest
// magnitude): entries far below the largest lose relative, not absolute, accuracy -- the error of every output stays
// <= 2^-18 of sum_r |a_r b_r| (measured 2^-19.3 .. 2^-21.8, at or below the library SGEMM on the same inputs:
// tests/test_gpu_parity.py::test_gemm_tn_split_k).  The k_splits partial products are summed in a fixed order.
#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// *out = max |x| as the bit pattern of a non-negative float (uint order == float order); *out must start at 0.
template <typename T>
__global__ void __launch_bounds__(256) absmax_kernel(const T *__restrict__ x, int64_t rows, int cols, int64_t ld,
                                                     unsigned *__restrict__ out) {
    float m = 0.0f;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += nwarps) {  // a warp per row: coalesced, no 64-bit divisions
        const T *row = x + r * ld;
        for (int c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(to_f32(row[c])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

// power of two that brings the largest magnitude into [2^13, 2^14) (fp16 pieces then never overflow)
__device__ __forceinline__ float scale_for(unsigned absmax_bits) {
    const float m = __uint_as_float(absmax_bits);
    if (!(m > 0.0f) || !isfinite(m)) return 1.0f;
    int e;
    frexpf(m, &e);                    // m = f * 2^e, f in [0.5, 1)
    int sh = 14 - e;                  // m * 2^sh in [2^13, 2^14)
    sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
    return ldexpf(1.0f, sh);
}

// dst[i] = 1 / scale for i < n: the per-row factor array the GEMM epilogue multiplies by
__global__ void fill_inv_scale_kernel(const unsigned *__restrict__ absmax_bits, float *__restrict__ dst, int n) {
    const float inv = 1.0f / scale_for(*absmax_bits);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = inv;
}

// planes[p][c][r] (p = 0, 1; c < Cp; r < Rp) = fp16 pieces of scale * X[r][c]; zero outside the matrix.
// One CTA transposes a 64 (rows) x 64 (columns) tile: rows of X are read 256 contiguous bytes at a time, stored
// transposed in shared memory, and written back as __half2 pairs -- a warp writes 128 contiguous bytes of one plane row.
template <typename T>
__global__ void __launch_bounds__(256) pack_transposed_kernel(const T *__restrict__ X, int64_t R, int C, int64_t ld,
                                                              const unsigned *__restrict__ absmax_bits,
                                                              __half *__restrict__ planes, int64_t Rp, int Cp) {
    __shared__ __align__(8) float tile[64][66];  // [column][row]
    const float s = scale_for(*absmax_bits);
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 4 rows of 64 threads
#pragma unroll 4
    for (int i = ty; i < 64; i += 4) {
        const int64_t r = r0 + i;
        const int c = c0 + tx;
        tile[tx][i] = (r < R && c < C) ? to_f32(X[r * ld + c]) * s : 0.0f;
    }
    __syncthreads();
    const size_t plane = (size_t)Cp * (size_t)Rp;
    const int rr = threadIdx.x & 31, cc = threadIdx.x >> 5;  // lane: row pair, warp: column
#pragma unroll 4
    for (int i = cc; i < 64; i += 8) {
        const int c = c0 + i;
        if (c < Cp) {
            const float2 v = *reinterpret_cast<const float2 *>(&tile[i][2 * rr]);
            const __half a0 = __float2half_rn(v.x), b0 = __float2half_rn(v.y);
            const __half a1 = __float2half_rn(v.x - __half2float(a0)), b1 = __float2half_rn(v.y - __half2float(b0));
            const size_t o = (size_t)c * Rp + (size_t)(r0 + 2 * rr);
            *reinterpret_cast<__half2 *>(planes + o) = __halves2half2(a0, b0);
            *reinterpret_cast<__half2 *>(planes + plane + o) = __halves2half2(a1, b1);
        }
    }
}

// out[i][j] = sum_ks part[ks][i][j] (ks ascending: reproducible), i < C1, j < C2; part rows have length ldp
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float *__restrict__ part, int k_splits, size_t stride,
                                                            int ldp, int C1, int C2, float *__restrict__ out) {
    const int64_t n = (int64_t)C1 * C2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C2), c = (int)(i - (int64_t)r * C2);
        float a = 0.0f;
        for (int ks = 0; ks < k_splits; ++ks) a += part[(size_t)ks * stride + (size_t)r * ldp + c];
        out[i] = a;
    }
}

// Non-transposing pack for products with the reduction along the rows' own elements (out = A . B^T):
// planes[p][r][c] (r < Rp, c < Cp) = fp16 pieces of 2^e_r * X[r][c], inv_scale[r] = 2^-e_r with e_r from the row's
// largest magnitude (the same per-row scaling as the encode path, prepare.cu); zero outside the matrix.  Warp per row.
__global__ void __launch_bounds__(256) pack_rows_kernel(const float *__restrict__ X, int64_t R, int C, int64_t ld,
                                                        __half *__restrict__ planes, float *__restrict__ inv_scale,
                                                        int64_t Rp, int Cp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)Rp * (size_t)Cp;
    for (int64_t r = warp; r < Rp; r += nwarps) {
        const float *row = X + (size_t)r * ld;
        float m = 0.0f;
        if (r < R)
            for (int c = lane; c < C; c += 32) m = fmaxf(m, fabsf(row[c]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        const float s = scale_for(__float_as_uint(m));
        if (lane == 0) inv_scale[r] = 1.0f / s;
        __half *h0 = planes + (size_t)r * Cp, *h1 = h0 + plane;
        for (int c = 2 * lane; c < Cp; c += 64) {  // two columns per lane: 4-byte stores
            float v0 = 0.0f, v1 = 0.0f;
            if (r < R) {
                if (c < C) v0 = row[c] * s;
                if (c + 1 < C) v1 = row[c + 1] * s;
            }
            const __half a0 = __float2half_rn(v0), b0 = __float2half_rn(v1);
            const __half a1 = __float2half_rn(v0 - __half2float(a0)), b1 = __float2half_rn(v1 - __half2float(b0));
            *reinterpret_cast<__half2 *>(h0 + c) = __halves2half2(a0, b0);
            *reinterpret_cast<__half2 *>(h1 + c) = __halves2half2(a1, b1);
        }
    }
}

int launch_pack_rows(const float *X, int64_t R, int C, int64_t ld, __half *planes, float *inv_scale, int64_t Rp, int Cp,
                     cudaStream_t st) {
    int64_t blocks = (Rp + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, R, C, ld, planes, inv_scale, Rp, Cp);
    MCQ_LAUNCH_CHECK("pack_rows_kernel");
    return MCQ_OK;
}

struct NtPlan {
    int64_t Mp;
    int Np, Kp;
    size_t off_pa, off_pb, off_sa, off_sb, bytes;
};

NtPlan nt_plan(int64_t m, int n, int k) {
    NtPlan p;
    p.Mp = (int64_t)align_up((size_t)m, 128);
    p.Np = (int)align_up((size_t)n, 128);  // rows of a plane of B (the tile loader reads whole 64/128-row boxes)
    p.Kp = (int)align_up((size_t)k, 64);
    size_t o = 0;
    p.off_pa = o; o += align_up((size_t)2 * p.Mp * p.Kp * sizeof(__half), 1024);
    p.off_pb = o; o += align_up((size_t)2 * p.Np * p.Kp * sizeof(__half), 1024);
    p.off_sa = o; o += align_up((size_t)p.Mp * sizeof(float), 1024);
    p.off_sb = o; o += align_up((size_t)p.Np * sizeof(float), 1024);
    p.bytes = o;
    return p;
}

struct TnPlan {
    int C1p, C2p, splits;
    int64_t Rp;
    size_t off_pa, off_pb, off_part, off_sa, off_sb, off_max, bytes;
};

TnPlan tn_plan(int64_t R, int C1, int C2) {
    TnPlan p;
    p.C1p = (int)align_up((size_t)C1, 128);
    p.C2p = (int)align_up((size_t)C2, 64);
    const int bn = (p.C2p % 128 == 0) ? 128 : 64;
    const int64_t mn_tiles = (int64_t)(p.C1p / 128) * (p.C2p / bn);
    const int64_t kb = (R + 63) / 64;  // 64-row blocks of the reduction
    // Short accumulation chains: the tensor core adds every 16-deep product into the fp32 accumulator with a rounding
    // that is not round-to-nearest, so the error grows with the number of accumulations per output.  At most 16
    // k-blocks (1,024 rows, 64 accumulations of the main product) per split keeps it near 2^-20 of sum |a b|; the
    // partial products cost splits * C1p * C2p * 4 bytes (capped at 512 MB) and one extra pass to sum.
    int64_t splits = (kb + 15) / 16;
    const int64_t min_splits = (2 * 148 + mn_tiles - 1) / mn_tiles;  // and about two waves of tiles on 148 SMs
    if (splits < min_splits) splits = min_splits;
    if (splits > kb) splits = kb;
    const int64_t cap = ((int64_t)512 << 20) / ((int64_t)p.C1p * p.C2p * 4);
    if (splits > cap) splits = cap;
    if (splits < 1) splits = 1;
    if (splits > 1024) splits = 1024;
    const int64_t kps = (kb + splits - 1) / splits;
    p.splits = (int)splits;
    p.Rp = kps * splits * 64;
    size_t o = 0;
    p.off_pa = o;   o += align_up((size_t)2 * p.C1p * p.Rp * sizeof(__half), 1024);
    p.off_pb = o;   o += align_up((size_t)2 * align_up((size_t)p.C2p, 128) * p.Rp * sizeof(__half), 1024);
    p.off_part = o; o += align_up((size_t)p.splits * p.C1p * p.C2p * sizeof(float), 1024);
    p.off_sa = o;   o += align_up((size_t)p.C1p * sizeof(float), 1024);
    p.off_sb = o;   o += align_up(align_up((size_t)p.C2p, 128) * sizeof(float), 1024);
    p.off_max = o;  o += 1024;
    p.bytes = o;
    return p;
}

template <typename T>
int pack_operand(const T *X, int64_t R, int C, int64_t ld, unsigned *absmax, __half *planes, int64_t Rp, int Cp,
                 float *inv_scale, int n_scale, cudaStream_t st) {
    int64_t blocks = (R + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    absmax_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(X, R, C, ld, absmax);
    MCQ_LAUNCH_CHECK("absmax_kernel");
    fill_inv_scale_kernel<<<(n_scale + 255) / 256, 256, 0, st>>>(absmax, inv_scale, n_scale);
    MCQ_LAUNCH_CHECK("fill_inv_scale_kernel");
    dim3 grid((unsigned)(Rp / 64), (unsigned)((Cp + 63) / 64));
    pack_transposed_kernel<T><<<grid, 256, 0, st>>>(X, R, C, ld, absmax, planes, Rp, Cp);
    MCQ_LAUNCH_CHECK("pack_transposed_kernel");
    return MCQ_OK;
}

}  // namespace

}  // namespace mcq

using namespace mcq;

extern "C" {

size_t mcq_gemm_tn_workspace_bytes(int64_t rows, int c1, int c2) {
    if (rows <= 0 || c1 <= 0 || c2 <= 0) return 0;
    return tn_plan(rows, c1, c2).bytes;
}

int mcq_gemm_tn(const float *a, int64_t lda, const void *b, int b_dtype, int64_t ldb, int64_t rows, int c1, int c2,
                float *out, void *workspace, size_t workspace_bytes, void *stream) {
    if (rows <= 0 || c1 <= 0 || c2 <= 0 || lda < c1 || ldb < c2) {
        set_error("mcq_gemm_tn: bad shape rows=%lld c1=%d c2=%d lda=%lld ldb=%lld", (long long)rows, c1, c2,
                  (long long)lda, (long long)ldb);
        return MCQ_EINVAL;
    }
    if (!a || !b || !out || !workspace) {
        set_error("mcq_gemm_tn: null pointer");
        return MCQ_EINVAL;
    }
    const TnPlan p = tn_plan(rows, c1, c2);
    if (workspace_bytes < p.bytes) {
        set_error("mcq_gemm_tn: workspace of %zu bytes, need %zu", workspace_bytes, p.bytes);
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    unsigned *mx = (unsigned *)(ws + p.off_max);
    MCQ_CUDA(cudaMemsetAsync(mx, 0, 8, st));
    const int C2p128 = (int)align_up((size_t)p.C2p, 128);
    int rc;
    if ((rc = pack_operand<float>(a, rows, c1, lda, mx, (__half *)(ws + p.off_pa), p.Rp, p.C1p, (float *)(ws + p.off_sa),
                                  p.C1p, st)))
        return rc;
    __half *pb = (__half *)(ws + p.off_pb);
    float *sb = (float *)(ws + p.off_sb);
    switch (b_dtype) {
        case MCQ_F32: rc = pack_operand<float>((const float *)b, rows, c2, ldb, mx + 1, pb, p.Rp, C2p128, sb, C2p128, st); break;
        case MCQ_F16: rc = pack_operand<__half>((const __half *)b, rows, c2, ldb, mx + 1, pb, p.Rp, C2p128, sb, C2p128, st); break;
        case MCQ_BF16:
            rc = pack_operand<__nv_bfloat16>((const __nv_bfloat16 *)b, rows, c2, ldb, mx + 1, pb, p.Rp, C2p128, sb, C2p128, st);
            break;
        default: set_error("mcq_gemm_tn: unknown dtype %d", b_dtype); return MCQ_EINVAL;
    }
    if (rc) return rc;
    float *part = (float *)(ws + p.off_part);
    if ((rc = launch_gemm_tc_splitk((const __half *)(ws + p.off_pa), (const float *)(ws + p.off_sa), pb, sb, part, p.C1p,
                                    p.C2p, (int)p.Rp, p.splits, st)))
        return rc;
    int64_t blocks = ((int64_t)c1 * c2 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(part, p.splits, (size_t)p.C1p * p.C2p, p.C2p, c1, c2, out);
    MCQ_LAUNCH_CHECK("splitk_reduce_kernel");
    return MCQ_OK;
}

size_t mcq_gemm_nt_workspace_bytes(int64_t m, int n, int k) {
    if (m <= 0 || n <= 0 || k <= 0) return 0;
    return nt_plan(m, n, k).bytes;
}

int mcq_gemm_nt(const float *a, int64_t lda, const float *b, int64_t ldb, int64_t m, int n, int k, float *out,
                int64_t ldc, int accumulate, void *workspace, size_t workspace_bytes, void *stream) {
    if (m <= 0 || n <= 0 || k <= 0 || lda < k || ldb < k || ldc < n) {
        set_error("mcq_gemm_nt: bad shape m=%lld n=%d k=%d lda=%lld ldb=%lld ldc=%lld", (long long)m, n, k, (long long)lda,
                  (long long)ldb, (long long)ldc);
        return MCQ_EINVAL;
    }
    if (n % 64 != 0 || (ldc & 3) || ((uintptr_t)out & 15)) {
        set_error("mcq_gemm_nt: n=%d must be a multiple of 64, ldc=%lld a multiple of 4, out 16-byte aligned", n,
                  (long long)ldc);
        return MCQ_EUNSUPPORTED;
    }
    if (!a || !b || !out || !workspace) {
        set_error("mcq_gemm_nt: null pointer");
        return MCQ_EINVAL;
    }
    const NtPlan p = nt_plan(m, n, k);
    if (workspace_bytes < p.bytes) {
        set_error("mcq_gemm_nt: workspace of %zu bytes, need %zu", workspace_bytes, p.bytes);
        return MCQ_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    int rc;
    if ((rc = launch_pack_rows(a, m, k, lda, (__half *)(ws + p.off_pa), (float *)(ws + p.off_sa), p.Mp, p.Kp, st))) return rc;
    if ((rc = launch_pack_rows(b, n, k, ldb, (__half *)(ws + p.off_pb), (float *)(ws + p.off_sb), p.Np, p.Kp, st))) return rc;
    return launch_gemm_tc_general((const __half *)(ws + p.off_pa), (const float *)(ws + p.off_sa),
                                  (const __half *)(ws + p.off_pb), (const float *)(ws + p.off_sb), out, ldc, m, p.Mp, n,
                                  p.Kp, accumulate, st);
}

}  // extern "C"
