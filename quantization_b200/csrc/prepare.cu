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

// ---- fp16 two-way split with an exact power-of-two row scale ---------------------------------------------------------
// A row a[] is scaled by 2^e so that its largest magnitude lands in [2^13, 2^14), then every element is written as
// a0 + a1 with a0 = half(a 2^e), a1 = half(a 2^e - a0): |a 2^e - a0 - a1| <= 2^-23 |a 2^e| for elements within 2^-17 of the
// row maximum (smaller ones are limited by fp16's subnormal quantum 2^-24, i.e. 2^-38 of the row maximum in absolute
// terms -- far below the fp32 rounding of the large elements).  The GEMM multiplies its result by 2^-e (exact).
// Products of fp16 pieces are exact in the tensor core's fp32 accumulator; the dropped a1 b1 term is <= 2^-22 |a b|.
__device__ __forceinline__ void row_scale_pow2(float rowmax, float &scale, float &inv) {
    scale = 1.0f;
    inv = 1.0f;
    if (rowmax > 0.0f && rowmax < INFINITY) {
        int q;
        frexpf(rowmax, &q);  // rowmax = m 2^q, m in [0.5, 1)
        int e = 14 - q;
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        scale = ldexpf(1.0f, e);
        inv = ldexpf(1.0f, -e);
    }
}
__device__ __forceinline__ void split2(float a, __half &h0, __half &h1) {
    h0 = __float2half_rn(a);
    h1 = __float2half_rn(a - __half2float(h0));
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One warp per codebook row r: scaled centers (quantization.py:77-79), copies of the classifier parameters, and the
// row-scaled fp16 splits of both.
__global__ void __launch_bounds__(256) prep_params_kernel(const float *__restrict__ centers,
                                                          const float *__restrict__ centers_scale,
                                                          const float *__restrict__ w, const float *__restrict__ bias,
                                                          const float *__restrict__ logits_scale, float speed, int NK,
                                                          int NKp, int D, int Dp, float *__restrict__ cs,
                                                          float *__restrict__ wout, float *__restrict__ bout,
                                                          float *__restrict__ scal, __half *__restrict__ csplit,
                                                          __half *__restrict__ wsplit, float *__restrict__ cscale,
                                                          float *__restrict__ wscale) {
    const float s = scale_of(centers_scale, speed);
    const size_t plane = (size_t)NKp * Dp;  // each split plane holds NKp (multiple of 128) rows; the tail rows stay zero
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int r = gw; r < NK; r += nw) {
        float mc = 0.f, mw = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float c = __fmul_rn(s, centers[(size_t)r * D + d]);
            const float ww = w[(size_t)r * D + d];
            cs[(size_t)r * D + d] = c;
            wout[(size_t)r * D + d] = ww;
            mc = fmaxf(mc, fabsf(c));
            mw = fmaxf(mw, fabsf(ww));
        }
        mc = warp_max_f(mc);
        mw = warp_max_f(mw);
        float sc, ic, sw, iw;
        row_scale_pow2(mc, sc, ic);
        row_scale_pow2(mw, sw, iw);
        for (int d = lane; d < Dp; d += 32) {
            float c = 0.f, ww = 0.f;
            if (d < D) {
                c = __fmul_rn(s, centers[(size_t)r * D + d]) * sc;
                ww = w[(size_t)r * D + d] * sw;
            }
            __half h0, h1;
            split2(c, h0, h1);
            csplit[(size_t)r * Dp + d] = h0;
            csplit[plane + (size_t)r * Dp + d] = h1;
            split2(ww, h0, h1);
            wsplit[(size_t)r * Dp + d] = h0;
            wsplit[plane + (size_t)r * Dp + d] = h1;
        }
        if (lane == 0) {
            cscale[r] = ic;
            wscale[r] = iw;
        }
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
    int blocks = (L.NK + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    prep_params_kernel<<<blocks, 256, 0, st>>>(centers, centers_scale, w, bias, logits_scale, scale_speed, L.NK,
                                               (int)align_up((size_t)L.NK, 128), L.D, L.Dp, cs,
                                               (float *)(blob + L.off_w), (float *)(blob + L.off_bias),
                                               (float *)(blob + L.off_scal), (__half *)(blob + L.off_csplit),
                                               (__half *)(blob + L.off_wsplit), (float *)(blob + L.off_cscale),
                                               (float *)(blob + L.off_wscale));
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

// One warp per frame row (rows >= B of the padded chunk are zero): fp32 copy, row-scaled fp16 splits of x and of
// fl(lscale * x) (quantization.py:278: exp(logits_scale*speed) * x, then the GEMM), and the rows' inverse scales.
template <typename T>
__global__ void __launch_bounds__(256) split_x_kernel(const T *__restrict__ x, int64_t B, int64_t Mp, int D, int Dp,
                                                      const float *__restrict__ scal, float *__restrict__ xf,
                                                      __half *__restrict__ xsplit, __half *__restrict__ lsplit,
                                                      float *__restrict__ xscale, float *__restrict__ lscale,
                                                      bool want_logits, bool want_xf) {
    const float ls = scal[1];
    const size_t plane = (size_t)Mp * Dp;
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const __half z = __float2half_rn(0.f);
    for (int64_t r = gw; r < Mp; r += nw) {
        const bool live = r < B;
        float mx = 0.f;
        if (live)
            for (int d = lane; d < D; d += 32) {
                const float v = to_f32<T>(x[(size_t)r * D + d]);
                if (want_xf) xf[(size_t)r * D + d] = v;
                mx = fmaxf(mx, fabsf(v));
            }
        else if (want_xf)
            for (int d = lane; d < D; d += 32) xf[(size_t)r * D + d] = 0.f;
        mx = warp_max_f(mx);
        float sx, ix, sl = 1.f, il = 1.f;
        row_scale_pow2(mx, sx, ix);
        if (want_logits) row_scale_pow2(fabsf(__fmul_rn(ls, mx)), sl, il);  // rounding is monotone: max |fl(ls v)|
        for (int d = lane; d < Dp; d += 32) {
            __half a0 = z, a1 = z, b0 = z, b1 = z;
            if (live && d < D) {
                const float v = to_f32<T>(x[(size_t)r * D + d]);
                split2(v * sx, a0, a1);
                if (want_logits) split2(__fmul_rn(ls, v) * sl, b0, b1);
            }
            xsplit[(size_t)r * Dp + d] = a0;
            xsplit[plane + (size_t)r * Dp + d] = a1;
            if (want_logits) {
                lsplit[(size_t)r * Dp + d] = b0;
                lsplit[plane + (size_t)r * Dp + d] = b1;
            }
        }
        if (lane == 0) {
            xscale[r] = ix;
            if (want_logits) lscale[r] = il;
        }
    }
}

int launch_split_x(const void *x, int x_dtype, int64_t B, const Prepared &L, const char *blob, const Workspace &W,
                   char *ws, bool want_logits_split, cudaStream_t st) {
    int64_t blocks64 = ((int64_t)W.Mp + 7) / 8;
    int blocks = (int)(blocks64 > 148 * 32 ? 148 * 32 : blocks64);
    if (blocks < 1) blocks = 1;
    const float *scal = (const float *)(blob + L.off_scal);
    float *xf = (float *)(ws + W.off_xf);
    __half *xs = (__half *)(ws + W.off_xsplit), *lsp = (__half *)(ws + W.off_lsplit);
    float *xsc = (float *)(ws + W.off_xscale), *lsc = (float *)(ws + W.off_lscale);
    // the fp32 copy is only read by the CUDA-core GEMM (shapes the tcgen05 kernel does not tile, MCQ_GEMM=ffma)
    const bool want_xf = !(use_tensor_core_gemm() && L.NK % 64 == 0);
    switch (x_dtype) {
        case MCQ_F32:
            split_x_kernel<float><<<blocks, 256, 0, st>>>((const float *)x, B, W.Mp, L.D, L.Dp, scal, xf, xs, lsp, xsc,
                                                          lsc, want_logits_split, want_xf);
            break;
        case MCQ_F16:
            split_x_kernel<__half><<<blocks, 256, 0, st>>>((const __half *)x, B, W.Mp, L.D, L.Dp, scal, xf, xs, lsp, xsc,
                                                           lsc, want_logits_split, want_xf);
            break;
        case MCQ_BF16:
            split_x_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16 *)x, B, W.Mp, L.D, L.Dp, scal,
                                                                  xf, xs, lsp, xsc, lsc, want_logits_split, want_xf);
            break;
        default:
            set_error("unknown x dtype %d", x_dtype);
            return MCQ_EINVAL;
    }
    MCQ_LAUNCH_CHECK("split_x_kernel");
    return MCQ_OK;
}

}  // namespace mcq
