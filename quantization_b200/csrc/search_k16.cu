// search_k16.cu -- refinement search for codebook_size 16, 8 codebooks: the first phase of QuantizerTrainer at
// bytes_per_frame = 4 (quantization.py:616-628), BASELINE config 3.  Same arithmetic contract and tie rules as search.cu /
// oracle/mcq_gram_model.c (tested bit for bit against both); what is different from the K = 256 kernels:
//
//   * the whole Gram table is 128 x 128 floats: it lives in shared memory (one copy per CTA, row stride 132 floats so
//     that reads of one column from several rows spread over the banks); no global gathers at all;
//   * the selections are small (16 -> 8 per codebook, 64 -> 8, 64 -> 16), so several run side by side in sub-warp
//     groups: 8 groups of 4 lanes at level 1, 4 groups of 8 lanes for the first merge, 2 groups of 16 for the second;
//     a group's candidates are blocked over its lanes, so "lowest lane among equals" is "lowest flat index among
//     equals"; the group minimum is a shuffle butterfly;
//   * the last merge (16 x 16) folds E_b[i][q] = sum_a D_ab(i_a, q) per left candidate i and right slot q first (the
//     inner sum of the contract), then dot(i,j) = sum_b E_b[i][j_b].
//
// Level schedule at K = 16 (base cut-off 8, quantization.py:453-463): keep 8 of 16 per codebook; merge pairs (8 x 8)
// keep 8; merge pairs of pairs (8 x 8) keep 16; merge the two quads (16 x 16) keep 1.
#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int KQ = 16, NQ = 8, NKQ = KQ * NQ;  // 128 rows / columns
constexpr int GS = 132;                        // shared-memory row stride of G (floats)
constexpr int WPC = 16;                        // warps per CTA

struct alignas(16) WarpK16 {
    float2 lists[9][32];   // per-lane sorted columns (key, flat) of the running selections + sentinel row
    float kd1[NQ][8];      // level 1: kept deltas and entries of each codebook
    int kk1[NQ][8];
    float kd2[4][8];       // after the first merge: deltas, slot tuples (2 x 4 bits)
    unsigned kt2[4][8];
    float kd3[2][16];      // after the second merge: deltas, slot tuples (4 x 4 bits)
    unsigned kt3[2][16];
    float es[16][8];       // last merge: E_b[i][q] of the current b
    int old[NQ];
};

template <int W>
__device__ __forceinline__ float group_min(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

__device__ __forceinline__ int le_mask16(float a, float b) {
    int r;
    asm("set.le.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}

// R smallest of each group's W * KPL candidates (lane li of a group holds flats li*KPL .. li*KPL + KPL-1), ascending by
// (key, flat).  `emit(r, key, flat)` is called by the lane that pops the r-th smallest of its group.
template <int W, int KPL, int R, class Emit>
__device__ __forceinline__ void group_select(WarpK16 &s, const float (&key)[KPL], int lane, Emit emit) {
    int rank[KPL];
#pragma unroll
    for (int t = 0; t < KPL; ++t) rank[t] = KPL - 1 - t;
#pragma unroll
    for (int t = 1; t < KPL; ++t)
#pragma unroll
        for (int u = 0; u < t; ++u) {
            const int c = le_mask16(key[u], key[t]);  // -1 when key[u] sorts before key[t] (equal keys keep index order)
            rank[t] -= c;
            rank[u] += c;
        }
    const int li = lane & (W - 1);
#pragma unroll
    for (int t = 0; t < KPL; ++t) s.lists[rank[t]][lane] = make_float2(key[t], __int_as_float(li * KPL + t));
    s.lists[KPL][lane] = make_float2(__int_as_float(0x7f800000), __int_as_float(0));
    const float2 *col = &s.lists[0][lane];
    float2 head = col[0];
    const unsigned gmask = (W == 32 ? FULL : ((1u << W) - 1u)) << (lane & ~(W - 1));
    const unsigned lower = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float m = group_min<W>(head.x);
        const bool p = head.x == m;
        const unsigned b = __ballot_sync(FULL, p) & gmask;
        if (p && (b & lower) == 0u) {  // lowest lane of the group among equals = lowest flat index
            emit(r, head.x, __float_as_int(head.y));
            col += 32;
            head = *col;
        }
    }
    __syncwarp();
}

// G entry from the CTA's shared copy
__device__ __forceinline__ float gs(const float *__restrict__ Gs, int row, int col) { return Gs[row * GS + col]; }

__device__ __forceinline__ void pass_k16(WarpK16 &s, const float *__restrict__ Gs, const float *__restrict__ diag,
                                         const float *__restrict__ Pb, int lane) {
    // ---- level 1: group n = lane / 4 handles codebook n; lane c = lane % 4 holds entries 4c .. 4c+3 ----
    {
        const int n = lane >> 2, c = lane & 3;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int mm = 0; mm < NQ - 1; ++mm) {
            const int m = mm + (mm >= n ? 1 : 0);  // ascending m, skipping n
            const float4 g = *reinterpret_cast<const float4 *>(Gs + (m * KQ + s.old[m]) * GS + n * KQ + c * 4);
            acc[0] = acc[0] + g.x;
            acc[1] = acc[1] + g.y;
            acc[2] = acc[2] + g.z;
            acc[3] = acc[3] + g.w;
        }
        const float4 p = __ldg(reinterpret_cast<const float4 *>(Pb + n * KQ + c * 4));
        const float4 d = *reinterpret_cast<const float4 *>(diag + n * KQ + c * 4);
        float v[4];
        v[0] = fmaf(2.0f, acc[0] - p.x, d.x);
        v[1] = fmaf(2.0f, acc[1] - p.y, d.y);
        v[2] = fmaf(2.0f, acc[2] - p.z, d.z);
        v[3] = fmaf(2.0f, acc[3] - p.w, d.w);
        const int on = s.old[n];
        float vs = v[0];
#pragma unroll
        for (int t = 1; t < 4; ++t) vs = ((on & 3) == t) ? v[t] : vs;
        const float vold = __shfl_sync(FULL, vs, (lane & ~3) | (on >> 2));
        float key[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) key[t] = v[t] - vold;
        group_select<4, 4, 8>(s, key, lane, [&](int r, float k, int flat) {
            s.kd1[n][r] = k;
            s.kk1[n][r] = flat;  // flat index == codebook entry
        });
    }
    // ---- first merge: group g = lane / 8 merges codebooks e = 2g, o = 2g+1; lane i = lane % 8 holds (i, j = 0..7) ----
    {
        const int g = lane >> 3, i = lane & 7;
        const int e = 2 * g, o = e + 1;
        const int rowe = e * KQ + s.kk1[e][i], rowe_old = e * KQ + s.old[e], colo_old = o * KQ + s.old[o];
        const float u = gs(Gs, rowe, colo_old), w = gs(Gs, rowe_old, colo_old), kde = s.kd1[e][i];
        float key[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int colo = o * KQ + s.kk1[o][j];
            const float d = ((gs(Gs, rowe, colo) - u) - gs(Gs, rowe_old, colo)) + w;
            key[j] = fmaf(2.0f, d, kde + s.kd1[o][j]);
        }
        group_select<8, 8, 8>(s, key, lane, [&](int r, float k, int flat) {
            s.kd2[g][r] = k;
            s.kt2[g][r] = (unsigned)(flat >> 3) | ((unsigned)(flat & 7) << 4);
        });
    }
    // ---- second merge: group h = lane / 16 merges pair groups e = 2h (codebooks 4h, 4h+1), o = 2h+1 (4h+2, 4h+3);
    //      lane li = lane % 16 holds flats 4 li .. 4 li + 3: i = li / 2, j = 4 (li % 2) + t ----
    {
        const int h = lane >> 4, li = lane & 15;
        const int e = 2 * h, o = e + 1, a0 = 4 * h, a1 = a0 + 1, b0 = a0 + 2, b1 = a0 + 3;
        const int i = li >> 1, jb = (li & 1) * 4;
        const unsigned ti = s.kt2[e][i];
        const int ra0 = a0 * KQ + s.kk1[a0][ti & 15], ra1 = a1 * KQ + s.kk1[a1][ti >> 4];
        const int ra0o = a0 * KQ + s.old[a0], ra1o = a1 * KQ + s.old[a1];
        const int cb0o = b0 * KQ + s.old[b0], cb1o = b1 * KQ + s.old[b1];
        const float u00 = gs(Gs, ra0, cb0o), u10 = gs(Gs, ra1, cb0o), u01 = gs(Gs, ra0, cb1o), u11 = gs(Gs, ra1, cb1o);
        const float w00 = gs(Gs, ra0o, cb0o), w10 = gs(Gs, ra1o, cb0o), w01 = gs(Gs, ra0o, cb1o), w11 = gs(Gs, ra1o, cb1o);
        const float kde = s.kd2[e][i];
        float key[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const unsigned tj = s.kt2[o][jb + t];
            const int c0 = b0 * KQ + s.kk1[b0][tj & 15], c1 = b1 * KQ + s.kk1[b1][tj >> 4];
            const float d00 = ((gs(Gs, ra0, c0) - u00) - gs(Gs, ra0o, c0)) + w00;
            const float d10 = ((gs(Gs, ra1, c0) - u10) - gs(Gs, ra1o, c0)) + w10;
            const float d01 = ((gs(Gs, ra0, c1) - u01) - gs(Gs, ra0o, c1)) + w01;
            const float d11 = ((gs(Gs, ra1, c1) - u11) - gs(Gs, ra1o, c1)) + w11;
            const float wb0 = d00 + d10, wb1 = d01 + d11;  // inner sums over a, then b ascending
            key[t] = fmaf(2.0f, wb0 + wb1, kde + s.kd2[o][jb + t]);
        }
        group_select<16, 4, 16>(s, key, lane, [&](int r, float k, int flat) {
            s.kd3[h][r] = k;
            s.kt3[h][r] = s.kt2[e][flat >> 3] | (s.kt2[o][flat & 7] << 8);
        });
    }
    // ---- last merge: 16 x 16 candidates, flat = i*16 + j; lane holds i = lane / 2, j = 8 (lane % 2) + t ----
    {
        const int i = lane >> 1, qh = (lane & 1) * 4, jb = (lane & 1) * 8;
        const unsigned ti = s.kt3[0][i];
        float dot[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) dot[t] = 0.0f;
#pragma unroll 1
        for (int lb = 0; lb < 4; ++lb) {
            const int b = 4 + lb, cbo = b * KQ + s.old[b];
            // E_b[i][q] for my i and q = qh .. qh+3 (level-1 slots of codebook b)
            int cq[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) cq[t] = b * KQ + s.kk1[b][qh + t];
            float E[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int ra = a * KQ + s.kk1[a][(ti >> (4 * a)) & 15], rao = a * KQ + s.old[a];
                const float u = gs(Gs, ra, cbo), w = gs(Gs, rao, cbo);
#pragma unroll
                for (int t = 0; t < 4; ++t) E[t] = E[t] + (((gs(Gs, ra, cq[t]) - u) - gs(Gs, rao, cq[t])) + w);
            }
            *reinterpret_cast<float4 *>(&s.es[i][qh]) = make_float4(E[0], E[1], E[2], E[3]);
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int jq = (s.kt3[1][jb + t] >> (4 * lb)) & 15;
                dot[t] = dot[t] + s.es[i][jq];
            }
            __syncwarp();
        }
        const float kde = s.kd3[0][i];
        float best = fmaf(2.0f, dot[0], kde + s.kd3[1][jb]);
        int bt = 0;
#pragma unroll
        for (int t = 1; t < 8; ++t) {
            const float k = fmaf(2.0f, dot[t], kde + s.kd3[1][jb + t]);
            if (k < best) {
                best = k;
                bt = t;
            }
        }
        const float m = group_min<32>(best);
        const unsigned bal = __ballot_sync(FULL, best == m);
        const int wl = bal ? (__ffs(bal) - 1) : 0;  // lowest lane among equals = lowest flat index (blocked flats)
        const int flat = __shfl_sync(FULL, lane * 8 + bt, wl);
        const unsigned te = s.kt3[0][flat >> 4], to = s.kt3[1][flat & 15];
        __syncwarp();  // every lane has read old[] before it is overwritten
        if (lane < 8) {
            const unsigned tt = lane < 4 ? te : to;
            s.old[lane] = s.kk1[lane][(tt >> (4 * (lane & 3))) & 15];
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WPC * 32, 1)
    search_k16n8_kernel(const float *__restrict__ P, const float *__restrict__ G, int64_t B, int iters,
                        const int32_t *__restrict__ idx_in, int32_t *__restrict__ idx_out,
                        unsigned *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Gs = reinterpret_cast<float *>(smem_raw);         // [128][GS]
    float *diag = Gs + NKQ * GS;                              // [128]
    WarpK16 *wm = reinterpret_cast<WarpK16 *>(diag + NKQ);
    for (int e = threadIdx.x; e < NKQ * NKQ; e += blockDim.x) Gs[(e >> 7) * GS + (e & 127)] = G[e];
    for (int e = threadIdx.x; e < NKQ; e += blockDim.x) diag[e] = G[(size_t)NKQ * NKQ + e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpK16 &s = wm[warp];
    const int64_t nwarps = (int64_t)gridDim.x * WPC;
    unsigned npass = 0, nframes = 0;
    for (int64_t b = (int64_t)blockIdx.x * WPC + warp; b < B;) {
        if (lane < NQ) s.old[lane] = idx_in[(size_t)b * NQ + lane];
        __syncwarp();
        const float *Pb = P + (size_t)b * NKQ;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int prev = (lane < NQ) ? s.old[lane] : 0;
            pass_k16(s, Gs, diag, Pb, lane);
            const int now = (lane < NQ) ? s.old[lane] : 0;
            ++npass;
            if (__all_sync(FULL, prev == now)) break;  // fixed point: the remaining passes are no-ops
        }
        ++nframes;
        if (lane < NQ) idx_out[(size_t)b * NQ + lane] = s.old[lane];
        if (work_counter != nullptr) {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(work_counter, 1u);
            b = nwarps + (int64_t)__shfl_sync(FULL, t, 0);
        } else {
            b += nwarps;
        }
        __syncwarp();
    }
    search_stats_add(work_counter, npass, nframes, lane);
}

}  // namespace

bool search_k16_supports(int N, int K) { return K == KQ && N == NQ; }

int launch_search_k16(const float *P, const float *gram, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
                      cudaStream_t st, unsigned *work_counter) {
    if (B <= 0) return MCQ_OK;
    const size_t smem = sizeof(float) * (NKQ * GS + NKQ) + sizeof(WarpK16) * WPC;
    MCQ_CUDA(cudaFuncSetAttribute(search_k16n8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    MCQ_CUDA(cudaGetDevice(&dev));
    MCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t need = (B + WPC - 1) / WPC;
    int64_t grid = sms;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    search_k16n8_kernel<<<(unsigned)grid, WPC * 32, smem, st>>>(P, gram, B, iters, idx_in, idx_out, work_counter);
    MCQ_LAUNCH_CHECK("search_k16n8_kernel");
    return MCQ_OK;
}

}  // namespace mcq
