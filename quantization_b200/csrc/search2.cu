// search2.cu -- second version of the refinement search (all passes of Quantizer._refine_indexes,
// quantization.py:308-547, for a batch of frames in ONE launch) for codebook_size 256 and 2, 4 or 8 codebooks: the
// inference configurations.  Same tables (P = x Cs^T per frame, G = Cs Cs^T per parameter version), same
// arithmetic contract and tie rules as search.cu / oracle/mcq_gram_model.c -- the tests compare both kernels with
// that model bit for bit -- but organised around what the first version's profile showed (profiles/r01_ncu_summary.md:
// issue bound, 55 % of the instructions in the sorted top-k extraction, 25 % in building difference tables):
//
//   * one warp per frame; a lane owns 8 CONSECUTIVE candidates (flat = lane*8 + t), so level-1 rows are float4 loads
//     and "lowest lane among equals" is "lowest flat index among equals" (the contract's tie rule);
//   * sorted top-R: each lane rank-sorts its 8 keys into its own shared-memory column, then R steps of
//     redux.sync.min.f32 (CREDUX) + ballot pop the global minimum from the column heads;
//   * every u/v term of a difference D_ab(p,q) = ((G[ap,bq] - G[ap,b_old]) - G[a_old,bq]) + G[a_old,b_old] comes from one
//     cached gather uv[a][m][p] = G[(m,old_m),(a,kk_a[p])] (G is bitwise symmetric);
//   * merge of single codebooks (16x16): one G gather per joint candidate; merge of codebook pairs (16x16): four;
//     merge of codebook quads (32x32): 16x16 tables T_ab per codebook pair, folded per candidate row into
//     E_b[i][q] = sum_a T_ab[i_a][q] -- exactly the contract's inner sum -- then dot(i,j) = sum_b E_b[i][j_b].
#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int K2 = 256;

__device__ __forceinline__ float credux_min(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

template <int N>
struct alignas(16) WarpMem2 {
    static constexpr int NG2 = (N >= 2) ? N / 2 : 1;  // groups after the first merge
    static constexpr int NG3 = (N >= 4) ? N / 4 : 1;  // groups after the second merge
    float2 lists[9][32];     // per-lane sorted columns (key, flat) + one row of +inf sentinels; rows 0..7 double as E
    float tab[16][16];       // level-1 scratch (v of the 256 candidates); T_ab of the quad merge
    float2 out1[N][16];      // level-1 kept candidates of each codebook: (delta, k)
    float uv[N][N][16];      // uv[a][m][p] = G[(m,old_m), (a, kk_a[p])]
    float2 sel[32];          // result of the current selection: (key, flat), ascending
    float kd2[NG2][16];      // kept deltas / slot tuples after the first merge
    unsigned kt2[NG2][16];
    float kd3[NG3][32];      // ... after the second merge
    unsigned kt3[NG3][32];
    int old[N];              // indexes at the start of the pass (and its result)
    unsigned used[8];        // quad merge: which level-1 slots of each codebook the 32+32 candidates still use
};

template <int N>
__device__ __forceinline__ int kk_of(const WarpMem2<N> &s, int a, int slot) {
    return __float_as_int(s.out1[a][slot].y);
}

// The R smallest of the warp's 256 candidates (8 per lane, flat index lane*8 + t), ascending by (key, flat),
// written to s.sel[0..R).  quantization.py:474-487 (sort + keep the first K_cutoff).
template <int N, int R>
__device__ __forceinline__ void select_sorted(WarpMem2<N> &s, const float (&key)[8], int lane) {
    int rank[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) rank[t] = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t)
#pragma unroll
        for (int u = 0; u < t; ++u) {
            const bool le = key[u] <= key[t];  // equal keys keep index order
            rank[t] += le ? 1 : 0;
            rank[u] += le ? 0 : 1;
        }
#pragma unroll
    for (int t = 0; t < 8; ++t) s.lists[rank[t]][lane] = make_float2(key[t], __int_as_float(lane * 8 + t));
    // a lane reads back only its own column: no warp synchronisation needed here
    const float2 *col = &s.lists[0][lane];
    const unsigned lt = (1u << lane) - 1u;
    int pos = 0;
    float2 head = col[0];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float m = credux_min(head.x);
        const bool p = head.x == m;
        const unsigned b = __ballot_sync(FULL, p);
        const bool mine = p && ((b & lt) == 0u);  // lowest lane among equals = lowest flat index
        if (mine) {
            s.sel[r] = head;
            pos += 32;
        }
        head = col[pos];
    }
    __syncwarp();
}

// Flat index of the smallest of the warp's 256 candidates (lowest flat index among equals).
__device__ __forceinline__ int select_best(const float (&key)[8], int lane) {
    float best = key[0];
    int bt = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t)
        if (key[t] < best) {
            best = key[t];
            bt = t;
        }
    const float m = credux_min(best);
    const unsigned b = __ballot_sync(FULL, best == m);
    const int w = b ? (__ffs(b) - 1) : 0;
    return __shfl_sync(FULL, lane * 8 + bt, w);
}

// Level 1 (quantization.py:401-418 with the per-codebook constants dropped) + top-16 per codebook.
template <int N>
__device__ __forceinline__ void level1(WarpMem2<N> &s, const float *__restrict__ Pb, const float *__restrict__ G,
                                       int lane) {
    constexpr int NK = N * K2;
    const float *diag = G + (size_t)NK * NK;
    float *scratch = &s.tab[0][0];
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
        float acc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.0f;
        const float *colbase = G + n * K2 + lane * 8;
#pragma unroll
        for (int mm = 0; mm < N - 1; ++mm) {
            const int m = mm + (mm >= n ? 1 : 0);  // ascending m, skipping n
            const float4 *row = reinterpret_cast<const float4 *>(colbase + (size_t)(m * K2 + s.old[m]) * NK);
            const float4 a = __ldg(row), b = __ldg(row + 1);
            acc[0] = acc[0] + a.x;
            acc[1] = acc[1] + a.y;
            acc[2] = acc[2] + a.z;
            acc[3] = acc[3] + a.w;
            acc[4] = acc[4] + b.x;
            acc[5] = acc[5] + b.y;
            acc[6] = acc[6] + b.z;
            acc[7] = acc[7] + b.w;
        }
        const float4 *pp = reinterpret_cast<const float4 *>(Pb + n * K2 + lane * 8);
        const float4 *dp = reinterpret_cast<const float4 *>(diag + n * K2 + lane * 8);
        const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1), d0 = __ldg(dp), d1 = __ldg(dp + 1);
        float v[8];
        v[0] = fmaf(2.0f, acc[0] - p0.x, d0.x);
        v[1] = fmaf(2.0f, acc[1] - p0.y, d0.y);
        v[2] = fmaf(2.0f, acc[2] - p0.z, d0.z);
        v[3] = fmaf(2.0f, acc[3] - p0.w, d0.w);
        v[4] = fmaf(2.0f, acc[4] - p1.x, d1.x);
        v[5] = fmaf(2.0f, acc[5] - p1.y, d1.y);
        v[6] = fmaf(2.0f, acc[6] - p1.z, d1.z);
        v[7] = fmaf(2.0f, acc[7] - p1.w, d1.w);
        reinterpret_cast<float4 *>(scratch)[lane * 2] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(scratch)[lane * 2 + 1] = make_float4(v[4], v[5], v[6], v[7]);
        __syncwarp();
        const float vold = scratch[s.old[n]];
        float key[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) key[t] = v[t] - vold;
        select_sorted<N, 16>(s, key, lane);
        if (lane < 16) s.out1[n][lane] = s.sel[lane];  // flat index == codebook entry k
        __syncwarp();
    }
}

template <int N>
__device__ __forceinline__ void gather_uv(WarpMem2<N> &s, const float *__restrict__ G, int lane) {
    constexpr int NK = N * K2;
    const int p = lane & 15, mh = lane >> 4;
#pragma unroll 1
    for (int a = 0; a < N; ++a) {
        const float *col = G + a * K2 + kk_of<N>(s, a, p);
#pragma unroll
        for (int r = 0; r < N / 2; ++r) {
            const int m = 2 * r + mh;
            if (m != a) s.uv[a][m][p] = __ldg(col + (size_t)(m * K2 + s.old[m]) * NK);
        }
    }
    __syncwarp();
}

// Merge of two single codebooks e = 2g, o = 2g+1: 16 x 16 joint candidates (quantization.py:504-547 at L = 1).
template <int N, bool FINAL>
__device__ __forceinline__ void merge1(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    constexpr int NK = N * K2;
    const int e = 2 * g, o = e + 1;
    const int i = lane >> 1, jb = (lane & 1) * 8;
    const float2 ke = s.out1[e][i];
    const float *rowp = G + (size_t)(e * K2 + __float_as_int(ke.y)) * NK + o * K2;
    const float u = s.uv[e][o][i];
    const float w = __ldg(G + (size_t)(e * K2 + s.old[e]) * NK + o * K2 + s.old[o]);
    float gv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) gv[t] = __ldg(rowp + kk_of<N>(s, o, jb + t));
    float key[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float v = s.uv[o][e][jb + t];
        const float d = ((gv[t] - u) - v) + w;
        key[t] = fmaf(2.0f, d, ke.x + s.out1[o][jb + t].x);
    }
    if constexpr (FINAL) {
        const int flat = select_best(key, lane);
        if (lane == 0) {
            const int ne = kk_of<N>(s, e, flat >> 4), no = kk_of<N>(s, o, flat & 15);
            s.old[e] = ne;
            s.old[o] = no;
        }
        __syncwarp();
    } else {
        select_sorted<N, 16>(s, key, lane);
        if (lane < 16) {
            const float2 r = s.sel[lane];
            const int flat = __float_as_int(r.y);
            s.kd2[g][lane] = r.x;
            s.kt2[g][lane] = (unsigned)(flat >> 4) | ((unsigned)(flat & 15) << 4);
        }
        __syncwarp();
    }
}

// Merge of two codebook pairs: groups e = 2g (codebooks 4g, 4g+1) and o = 2g+1 (4g+2, 4g+3), 16 x 16 candidates.
template <int N, bool FINAL>
__device__ __forceinline__ void merge2(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    constexpr int NK = N * K2;
    const int e = 2 * g, o = e + 1;
    const int a0 = 4 * g, a1 = a0 + 1, b0 = a0 + 2, b1 = a0 + 3;
    const int i = lane >> 1, jb = (lane & 1) * 8;
    const unsigned ti = s.kt2[e][i];
    const int ia0 = ti & 15, ia1 = ti >> 4;
    const float kde = s.kd2[e][i];
    const float *rp0 = G + (size_t)(a0 * K2 + kk_of<N>(s, a0, ia0)) * NK;
    const float *rp1 = G + (size_t)(a1 * K2 + kk_of<N>(s, a1, ia1)) * NK;
    const float u00 = s.uv[a0][b0][ia0], u10 = s.uv[a1][b0][ia1], u01 = s.uv[a0][b1][ia0], u11 = s.uv[a1][b1][ia1];
    const float *ro0 = G + (size_t)(a0 * K2 + s.old[a0]) * NK;
    const float *ro1 = G + (size_t)(a1 * K2 + s.old[a1]) * NK;
    const int cb0 = b0 * K2 + s.old[b0], cb1 = b1 * K2 + s.old[b1];
    const float w00 = __ldg(ro0 + cb0), w10 = __ldg(ro1 + cb0), w01 = __ldg(ro0 + cb1), w11 = __ldg(ro1 + cb1);
    float g00[8], g10[8], g01[8], g11[8];
    unsigned tjs[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const unsigned tj = s.kt2[o][jb + t];
        tjs[t] = tj;
        const int c0 = b0 * K2 + kk_of<N>(s, b0, tj & 15), c1 = b1 * K2 + kk_of<N>(s, b1, tj >> 4);
        g00[t] = __ldg(rp0 + c0);
        g10[t] = __ldg(rp1 + c0);
        g01[t] = __ldg(rp0 + c1);
        g11[t] = __ldg(rp1 + c1);
    }
    float key[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int q0 = tjs[t] & 15, q1 = tjs[t] >> 4;
        const float d00 = ((g00[t] - u00) - s.uv[b0][a0][q0]) + w00;
        const float d10 = ((g10[t] - u10) - s.uv[b0][a1][q0]) + w10;
        const float d01 = ((g01[t] - u01) - s.uv[b1][a0][q1]) + w01;
        const float d11 = ((g11[t] - u11) - s.uv[b1][a1][q1]) + w11;
        const float wb0 = d00 + d10, wb1 = d01 + d11;  // inner sums over a, then b ascending
        const float dot = wb0 + wb1;
        key[t] = fmaf(2.0f, dot, kde + s.kd2[o][jb + t]);
    }
    if constexpr (FINAL) {
        const int flat = select_best(key, lane);
        if (lane == 0) {
            const unsigned te = s.kt2[e][flat >> 4], to = s.kt2[o][flat & 15];
            const int n0 = kk_of<N>(s, a0, te & 15), n1 = kk_of<N>(s, a1, te >> 4);
            const int n2 = kk_of<N>(s, b0, to & 15), n3 = kk_of<N>(s, b1, to >> 4);
            s.old[a0] = n0;
            s.old[a1] = n1;
            s.old[b0] = n2;
            s.old[b1] = n3;
        }
        __syncwarp();
    } else {
        select_sorted<N, 32>(s, key, lane);
        {
            const float2 r = s.sel[lane];
            const int flat = __float_as_int(r.y);
            s.kd3[g][lane] = r.x;
            s.kt3[g][lane] = s.kt2[e][flat >> 4] | (s.kt2[o][flat & 15] << 8);
        }
        __syncwarp();
    }
}

// Final merge of two codebook quads (N = 8): 32 x 32 joint candidates, candidate flat = i*32 + j.
template <int N>
__device__ __forceinline__ void merge4_final(WarpMem2<N> &s, const float *__restrict__ G, int lane) {
    constexpr int NK = N * K2;
    const unsigned ti = s.kt3[0][lane];  // as row i = lane: my slots of codebooks 0..3
    const unsigned tj = s.kt3[1][lane];  // as column j = lane: my slots of codebooks 4..7
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned ua = __reduce_or_sync(FULL, 1u << ((ti >> (4 * c)) & 15));
        const unsigned ub = __reduce_or_sync(FULL, 1u << ((tj >> (4 * c)) & 15));
        if (lane == 0) {
            s.used[c] = ua;
            s.used[4 + c] = ub;
        }
    }
    __syncwarp();
    float dot[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dot[i] = 0.0f;
    float(*es)[16] = reinterpret_cast<float(*)[16]>(&s.lists[0][0]);
    const int p = lane >> 1, qb = (lane & 1) * 8;
#pragma unroll 1
    for (int lb = 0; lb < 4; ++lb) {
        const int b = 4 + lb;
        const unsigned ub = s.used[4 + lb];
        float E[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) E[q] = 0.0f;
#pragma unroll 1
        for (int a = 0; a < 4; ++a) {
            const bool rowu = (s.used[a] >> p) & 1u;
            const float *rowp = G + (size_t)(a * K2 + kk_of<N>(s, a, p)) * NK + b * K2;
            const float u = s.uv[a][b][p];
            const float w = __ldg(G + (size_t)(a * K2 + s.old[a]) * NK + b * K2 + s.old[b]);
            float gv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                gv[t] = 0.0f;
                if (rowu && ((ub >> (qb + t)) & 1u)) gv[t] = __ldg(rowp + kk_of<N>(s, b, qb + t));
            }
            float d[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) d[t] = ((gv[t] - u) - s.uv[b][a][qb + t]) + w;
            float4 *trow = reinterpret_cast<float4 *>(&s.tab[p][qb]);
            trow[0] = make_float4(d[0], d[1], d[2], d[3]);
            trow[1] = make_float4(d[4], d[5], d[6], d[7]);
            __syncwarp();
            const float4 *mine = reinterpret_cast<const float4 *>(&s.tab[(ti >> (4 * a)) & 15][0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 r = mine[c];
                E[4 * c + 0] = E[4 * c + 0] + r.x;
                E[4 * c + 1] = E[4 * c + 1] + r.y;
                E[4 * c + 2] = E[4 * c + 2] + r.z;
                E[4 * c + 3] = E[4 * c + 3] + r.w;
            }
            __syncwarp();
        }
        float4 *erow = reinterpret_cast<float4 *>(&es[lane][0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) erow[c] = make_float4(E[4 * c], E[4 * c + 1], E[4 * c + 2], E[4 * c + 3]);
        __syncwarp();
        const int jq = (tj >> (4 * lb)) & 15;
#pragma unroll
        for (int i = 0; i < 32; ++i) dot[i] = dot[i] + es[i][jq];
        __syncwarp();
    }
    const float kdo = s.kd3[1][lane];
    float best = fmaf(2.0f, dot[0], s.kd3[0][0] + kdo);
    int bi = 0;
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        const float key = fmaf(2.0f, dot[i], s.kd3[0][i] + kdo);
        if (key < best) {
            best = key;
            bi = i;
        }
    }
    const float m = credux_min(best);
    const unsigned c = (best == m) ? (unsigned)(bi * 32 + lane) : 0x7fffffffu;
    unsigned flat = __reduce_min_sync(FULL, c);
    if (flat == 0x7fffffffu) flat = 0;  // only reachable with NaN scores
    const unsigned te = s.kt3[0][flat >> 5], to = s.kt3[1][flat & 31];
    if (lane < 8) {
        const unsigned tt = lane < 4 ? te : to;
        s.old[lane] = kk_of<N>(s, lane, (tt >> (4 * (lane & 3))) & 15);
    }
    __syncwarp();
}

template <int N>
__device__ __forceinline__ void refine_pass2(WarpMem2<N> &s, const float *__restrict__ Pb,
                                             const float *__restrict__ G, int lane) {
    level1<N>(s, Pb, G, lane);
    gather_uv<N>(s, G, lane);
    if constexpr (N == 2) {
        merge1<N, true>(s, G, 0, lane);
    } else {
#pragma unroll 1
        for (int g = 0; g < N / 2; ++g) merge1<N, false>(s, G, g, lane);
        if constexpr (N == 4) {
            merge2<N, true>(s, G, 0, lane);
        } else {
#pragma unroll 1
            for (int g = 0; g < N / 4; ++g) merge2<N, false>(s, G, g, lane);
            merge4_final<N>(s, G, lane);
        }
    }
}

template <int N>
struct Launch2 {
    static constexpr int WPC = (N == 8) ? 5 : 8;  // warps per CTA
};

template <int N>
__global__ void __launch_bounds__(Launch2<N>::WPC * 32, (N == 8 ? 4 : 3))
    search2_kernel(const float *__restrict__ P, const float *__restrict__ G, int64_t B, int iters,
                   const int32_t *__restrict__ idx_in, int32_t *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int wpc = Launch2<N>::WPC;
    WarpMem2<N> &s = reinterpret_cast<WarpMem2<N> *>(smem_raw)[warp];
    s.lists[8][lane] = make_float2(__int_as_float(0x7f800000), __int_as_float(0));
    s.sel[lane] = make_float2(0.0f, __int_as_float(0));
    __syncwarp();
    for (int64_t b = (int64_t)blockIdx.x * wpc + warp; b < B; b += (int64_t)gridDim.x * wpc) {
        if (lane < N) s.old[lane] = idx_in[(size_t)b * N + lane];
        __syncwarp();
        const float *Pb = P + (size_t)b * (N * K2);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int prev = (lane < N) ? s.old[lane] : 0;
            refine_pass2<N>(s, Pb, G, lane);
            const int now = (lane < N) ? s.old[lane] : 0;
            // a pass that returns its input is a fixed point of a deterministic map: the remaining passes are no-ops
            if (__all_sync(FULL, prev == now)) break;
        }
        if (lane < N) idx_out[(size_t)b * N + lane] = s.old[lane];
        __syncwarp();
    }
}

template <int N>
int launch2(const float *P, const float *G, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
            cudaStream_t st) {
    constexpr int wpc = Launch2<N>::WPC;
    const size_t smem = sizeof(WarpMem2<N>) * wpc;
    auto kern = search2_kernel<N>;
    MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 148;
    MCQ_CUDA(cudaGetDevice(&dev));
    MCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t need = (B + wpc - 1) / wpc;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, wpc * 32, smem, st>>>(P, G, B, iters, idx_in, idx_out);
    MCQ_LAUNCH_CHECK("search2_kernel");
    return MCQ_OK;
}

}  // namespace

bool search2_supports(int N, int K) { return K == 256 && (N == 2 || N == 4 || N == 8); }

int launch_search2(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
                   int32_t *idx_out, cudaStream_t st) {
    if (B <= 0) return MCQ_OK;
    if (K == 256) {
        switch (N) {
            case 2: return launch2<2>(P, gram, B, iters, idx_in, idx_out, st);
            case 4: return launch2<4>(P, gram, B, iters, idx_in, idx_out, st);
            case 8: return launch2<8>(P, gram, B, iters, idx_in, idx_out, st);
            default: break;
        }
    }
    set_error("search2: (K=%d, N=%d) is not supported", K, N);
    return MCQ_EUNSUPPORTED;
}

}  // namespace mcq
