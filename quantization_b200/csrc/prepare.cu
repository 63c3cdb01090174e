// prepare.cu -- once-per-parameter-version work: scaled centers (quantization.py:77-79), the exp() of the logits
// scale (:278), the Gram table G = Cs Cs^T (replaces all_centers_sumsq :411 and every per-frame delta product
// :413-416, :533-535), and the bf16 three-way operand splits the tcgen05 GEMM consumes.
#include "common.cuh"

namespace mcq {

// exp(raw * speed) as the reference evaluates it: fp32 product, then exp.  The exp itself is taken in
// double and rounded once (== correctly rounded fp32 exp).
__device__ __forceinline__ float scale_of(const float *raw, float speed) {
    float prod = __fmul_rn(*raw, speed);
    return (float)exp((double)prod);
}

// a = a1 + a2 + a3 exactly (8 significand bits each), for normal-range a.
__device__ __forceinline__ void split3(float a, __nv_bfloat16 &h1, __nv_bfloat16 &h2, __nv_bfloat16 &h3) {
    h1 = __float2bfloat16_rn(a);
    float r1 = a - __bfloat162float(h1);
    h2 = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(h2);
    h3 = __float2bfloat16_rn(r2);
}

__global__ void prep_params_kernel(const float *__restrict__ centers, const float *__restrict__ centers_scale,
                                   const float *__restrict__ w, const float *__restrict__ bias,
                                   const float *__restrict__ logits_scale, float speed, int NK, int NKp, int D, int Dp,
                                   float *__restrict__ cs, float *__restrict__ wout, float *__restrict__ bout,
                                   float *__restrict__ scal, __nv_bfloat16 *__restrict__ csplit,
                                   __nv_bfloat16 *__restrict__ wsplit) {
    const float s = scale_of(centers_scale, speed);
    const size_t total = (size_t)NK * Dp;
    const size_t plane = (size_t)NKp * Dp;  // each split plane holds NKp (multiple of 128) rows; the tail rows stay zero
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / Dp;
        int d = (int)(i - r * Dp);
        float c = 0.f, ww = 0.f;
        if (d < D) {
            c = __fmul_rn(s, centers[r * D + d]);
            ww = w[r * D + d];
            cs[r * D + d] = c;
            wout[r * D + d] = ww;
        }
        __nv_bfloat16 a, b, e;
        split3(c, a, b, e);
        csplit[i] = a;
        csplit[plane + i] = b;
        csplit[2 * plane + i] = e;
        split3(ww, a, b, e);
        wsplit[i] = a;
        wsplit[plane + i] = b;
        wsplit[2 * plane + i] = e;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < NK; i += blockDim.x) bout[i] = bias[i];
        if (threadIdx.x == 0) {
            scal[0] = s;
            scal[1] = scale_of(logits_scale, speed);
        }
    }
}

// G[r][s] = (float) sum_d (double)cs[r][d] * (double)cs[s][d], d ascending.  64x64 tile per CTA, 4x4 per thread.
// The (r,s) and (s,r) entries see the same products in the same order, so G is bitwise symmetric.
__global__ void __launch_bounds__(256) gram_kernel(const float *__restrict__ cs, int NK, int D, float *__restrict__ G) {
    __shared__ float As[16][65];
    __shared__ float Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r0 = blockIdx.y * 64, s0 = blockIdx.x * 64;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < D; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            int row = e >> 4, kk = e & 15;
            int k = k0 + kk;
            float a = 0.f, b = 0.f;
            if (k < D) {
                if (r0 + row < NK) a = cs[(size_t)(r0 + row) * D + k];
                if (s0 + row < NK) b = cs[(size_t)(s0 + row) * D + k];
            }
            As[kk][row] = a;
            Bs[kk][row] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = (double)As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (double)Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = r0 + ty * 4 + i, s = s0 + tx * 4 + j;
            if (r < NK && s < NK) {
                float g = (float)acc[i][j];
                G[(size_t)r * NK + s] = g;
                if (r == s) G[(size_t)NK * NK + r] = g;  // diagonal copy |c_r|^2 appended after the table
            }
        }
}

int launch_prepare(const float *centers, const float *centers_scale, const float *w, const float *bias,
                   const float *logits_scale, float scale_speed, const Prepared &L, char *blob, cudaStream_t st) {
    float *cs = (float *)(blob + L.off_cs);
    size_t total = (size_t)L.NK * L.Dp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    prep_params_kernel<<<blocks, 256, 0, st>>>(centers, centers_scale, w, bias, logits_scale, scale_speed, L.NK,
                                               (int)align_up((size_t)L.NK, 128), L.D, L.Dp, cs, (float *)(blob + L.off_w), (float *)(blob + L.off_bias),
                                               (float *)(blob + L.off_scal), (__nv_bfloat16 *)(blob + L.off_csplit),
                                               (__nv_bfloat16 *)(blob + L.off_wsplit));
    MCQ_LAUNCH_CHECK("prep_params_kernel");
    dim3 grid((L.NK + 63) / 64, (L.NK + 63) / 64);
    gram_kernel<<<grid, 256, 0, st>>>(cs, L.NK, L.D, (float *)(blob + L.off_gram));
    MCQ_LAUNCH_CHECK("gram_kernel");
    return MCQ_OK;
}

// ---- x staging: any dtype -> fp32 copy + bf16 splits of x and of fl(lscale * x) -----------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void split_x_kernel(const T *__restrict__ x, int64_t B, int64_t Mp, int D, int Dp,
                               const float *__restrict__ scal, float *__restrict__ xf,
                               __nv_bfloat16 *__restrict__ xsplit, __nv_bfloat16 *__restrict__ lsplit,
                               bool want_logits) {
    const float ls = scal[1];
    const size_t total = (size_t)Mp * Dp;
    const size_t plane = total;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i / Dp;
        int d = (int)(i - r * Dp);
        float v = 0.f;
        if (r < (size_t)B && d < D) v = to_f32<T>(x[r * D + d]);
        if (d < D) xf[r * D + d] = v;
        __nv_bfloat16 a, b, e;
        split3(v, a, b, e);
        xsplit[i] = a;
        xsplit[plane + i] = b;
        xsplit[2 * plane + i] = e;
        if (want_logits) {
            split3(__fmul_rn(ls, v), a, b, e);  // quantization.py:278: exp(logits_scale*speed) * x, then the GEMM
            lsplit[i] = a;
            lsplit[plane + i] = b;
            lsplit[2 * plane + i] = e;
        }
    }
}

int launch_split_x(const void *x, int x_dtype, int64_t B, const Prepared &L, const char *blob, const Workspace &W,
                   char *ws, bool want_logits_split, cudaStream_t st) {
    size_t total = (size_t)W.Mp * L.Dp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    const float *scal = (const float *)(blob + L.off_scal);
    float *xf = (float *)(ws + W.off_xf);
    __nv_bfloat16 *xs = (__nv_bfloat16 *)(ws + W.off_xsplit), *lsp = (__nv_bfloat16 *)(ws + W.off_lsplit);
    switch (x_dtype) {
        case MCQ_F32:
            split_x_kernel<float><<<blocks, 256, 0, st>>>((const float *)x, B, W.Mp, L.D, L.Dp, scal, xf, xs, lsp,
                                                          want_logits_split);
            break;
        case MCQ_F16:
            split_x_kernel<__half><<<blocks, 256, 0, st>>>((const __half *)x, B, W.Mp, L.D, L.Dp, scal, xf, xs, lsp,
                                                           want_logits_split);
            break;
        case MCQ_BF16:
            split_x_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16 *)x, B, W.Mp, L.D, L.Dp, scal,
                                                                  xf, xs, lsp, want_logits_split);
            break;
        default:
            set_error("unknown x dtype %d", x_dtype);
            return MCQ_EINVAL;
    }
    MCQ_LAUNCH_CHECK("split_x_kernel");
    return MCQ_OK;
}

}  // namespace mcq
