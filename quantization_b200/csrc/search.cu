// search.cu -- all passes of Quantizer._refine_indexes (quantization.py:308-547) for a batch of frames in ONE launch,
// driven by two tables instead of per-frame vectors:
//     P[b, r] = <x_b, c_r>  (gemm_tc.cu, once per frame)        G[r, s] = <c_r, c_s>  (prepare.cu, L2 resident)
// One warp owns one frame.  Nothing of the reference's (B, N, K, dim) "deltas" is ever materialised: every inner
// product of two deltas the reference forms with a per-frame GEMM (:533-535) is four entries of G.
//
// The arithmetic contract (what is rounded where, and every tie rule) is written down once, in
// oracle/mcq_gram_model.c, and the kernel is tested bit-for-bit against that model.
//
// Level schedule (quantization.py:453-547), all compile-time here: keep `cut1` = 16 (8 when K <= 16) candidates per
// codebook, then repeatedly merge neighbouring groups (Kc x Kc joint candidates) and keep cutoff(L) of them, until
// one group is left and its best joint candidate becomes the new indexes.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned KEY_REMOVED = 0xffffffffu;

__host__ __device__ constexpr int cutoff_of(int base, int L) {  // quantization.py:455-463
    int c = base;
    while (L >= 4) {
        L /= 4;
        c *= 2;
    }
    return c < 128 ? c : 128;
}

// Monotone map float -> uint (ascending), with -0 folded onto +0 so that the order is exactly the `<` order
// on floats that the model sorts by.
__device__ __forceinline__ unsigned fkey(float f) {
    f = f + 0.0f;
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

template <int K, int N>
struct Cfg {
    static constexpr int BASE = (K <= 16) ? 8 : 16;
    static constexpr int CUT1 = (N == 1) ? 1 : BASE;  // kept per codebook after level 1
    static constexpr int CPL1 = (K + 31) / 32;        // level-1 candidates per lane
    static constexpr int NK = N * K;
    static constexpr int HALF = (N / 2 > 0) ? N / 2 : 1;
    static constexpr int TW = (HALF + 7) / 8;         // 32-bit words per stored slot tuple (4 bits per codebook)
    static constexpr int LIST = N * CUT1;             // entries of the level-1 lists (the largest)
    static constexpr int LIST2 = HALF * CUT1;         // entries of any later list
    static constexpr int TABN = CUT1 * CUT1;
    // K >= 32 with N >= 32 reaches cutoff 64 (quantization.py:455-463): one or two merges of 64 x 64 joint candidates,
    // too many to keep in registers -- they are accumulated in shared memory (Level::run_big)
    static constexpr int BIGN = (BASE == 16 && N >= 32) ? 4096 : 1;
};

// Per-warp shared memory.
template <int K, int N>
struct WarpSmem {
    using C = Cfg<K, N>;
    int old[N];                              // indexes at the start of the pass
    unsigned short um[N];                    // per codebook: which level-1 slots the current lists still use
    unsigned char kk[N][16];                 // level-1 kept candidates: slot -> codebook entry
    float kd[2][C::LIST];                    // kept deltas (ping-pong between levels)
    unsigned kt[2][C::LIST2 > 0 ? C::LIST2 : 1][C::TW];  // kept slot tuples of levels >= 2
    float tab[C::HALF][C::TABN];             // D_ab tables of the codebook pairs being merged
    unsigned big[C::BIGN];                   // dots, then sortable keys, of a 4096-candidate merge
};

// Removes and returns the smallest (key, candidate) of the warp's CPL*32 candidates, candidate c = t*32 + lane.
template <int CPL>
__device__ __forceinline__ void extract_min(unsigned (&key)[CPL], int lane, unsigned &mkey, int &mc) {
    unsigned lk = key[0];
    int lt = 0;
#pragma unroll
    for (int t = 1; t < CPL; ++t)
        if (key[t] < lk) {
            lk = key[t];
            lt = t;
        }
    const unsigned m = __reduce_min_sync(FULL, lk);
    const unsigned c = (lk == m) ? (unsigned)(lt * 32 + lane) : 0xffffffffu;
    const unsigned cw = __reduce_min_sync(FULL, c);
    const bool mine = (c == cw);
#pragma unroll
    for (int t = 0; t < CPL; ++t)
        if (mine && t == lt) key[t] = KEY_REMOVED;
    mkey = m;
    mc = (int)cw;
}

template <int TW>
__device__ __forceinline__ int nib(const unsigned (&t)[TW], int la) {
    return (t[la >> 3] >> ((la & 7) * 4)) & 15;
}

// One merge level: Ncur groups with Kc kept candidates covering L codebooks each.
template <int K, int N, int Ncur, int Kc, int L>
struct Level {
    using C = Cfg<K, N>;
    static constexpr int CPL = (Kc * Kc) / 32;  // joint candidates per lane (Kc >= 8)
    static constexpr int NEWN = Ncur / 2;
    static constexpr int NEWK = (NEWN == 1) ? 1 : cutoff_of(C::BASE, 2 * L);
    static constexpr int CUT1 = C::CUT1;
    static constexpr int TPL = (CUT1 * CUT1) / 32;  // table entries per lane

    // The L tables D_ab = ((g - u) - v) + w of codebook b against every codebook a of the even group (only the slots
    // the kept candidates still use).
    __device__ static __forceinline__ void build_tables(WarpSmem<K, N> &s, const float *__restrict__ G, int e, int b,
                                                        int lane) {
        const size_t NK = C::NK;
        const size_t cb_old = (size_t)b * K + s.old[b];
        const unsigned umb = s.um[b];
#pragma unroll 1
        for (int la = 0; la < L; ++la) {
            const int a = e * L + la;
            const size_t ra_old = ((size_t)a * K + s.old[a]) * NK;
            const unsigned uma = s.um[a];
            const float w = __ldg(G + ra_old + cb_old);
#pragma unroll
            for (int tt = 0; tt < (TPL > 0 ? TPL : 1); ++tt) {
                const int ent = tt * 32 + lane;
                const int q = ent % CUT1, p = ent / CUT1;
                if (ent < CUT1 * CUT1 && ((uma >> p) & 1u) && ((umb >> q) & 1u)) {
                    const size_t ra = ((size_t)a * K + s.kk[a][p]) * NK;
                    const size_t cb = (size_t)b * K + s.kk[b][q];
                    const float g = __ldg(G + ra + cb);
                    const float u = __ldg(G + ra + cb_old);
                    const float v = __ldg(G + ra_old + cb);
                    s.tab[la][ent] = ((g - u) - v) + w;
                }
            }
        }
    }

    // Stores the r-th kept candidate (flat index c = i*Kc + j) of merged group m, or the final indexes.
    __device__ static __forceinline__ void emit(WarpSmem<K, N> &s, int cur, int m, int r, unsigned mk, int c) {
        const int nxt = cur ^ 1;
        const int e = 2 * m, o = 2 * m + 1;
        const int ci = c / Kc, cj = c % Kc;
        unsigned te[C::TW], to[C::TW];
        if constexpr (L == 1) {
            te[0] = (unsigned)ci;
            to[0] = (unsigned)cj;
        } else {
#pragma unroll
            for (int w = 0; w < C::TW; ++w) {
                te[w] = s.kt[cur][e * Kc + ci][w];
                to[w] = s.kt[cur][o * Kc + cj][w];
            }
        }
        if constexpr (NEWN == 1) {
            // final winner: tuple covers all N codebooks; decode slots to codebook entries
#pragma unroll 1
            for (int la = 0; la < L; ++la) {
                s.old[la] = s.kk[la][nib<C::TW>(te, la)];
                s.old[L + la] = s.kk[L + la][nib<C::TW>(to, la)];
            }
        } else {
            unsigned tn[C::TW];
#pragma unroll
            for (int w = 0; w < C::TW; ++w) tn[w] = 0u;
            if constexpr (L < 8) {
                tn[0] = te[0] | (to[0] << (4 * L));
            } else {
#pragma unroll
                for (int w = 0; w < L / 8; ++w) {
                    tn[w] = te[w];
                    tn[L / 8 + w] = to[w];
                }
            }
            s.kd[nxt][m * NEWK + r] = fkey_inv(mk);
#pragma unroll
            for (int w = 0; w < C::TW; ++w) s.kt[nxt][m * NEWK + r][w] = tn[w];
        }
    }

    // Merge of group pair m with Kc = 64: the 4096 dots / keys live in shared memory, candidate c = i*64 + j is owned by
    // lane c % 32 (so j = lane or lane + 32).  Same arithmetic and tie rules as the register path below.
    __device__ static __forceinline__ void merge_big(WarpSmem<K, N> &s, const float *__restrict__ G, int cur, int m,
                                                     int lane) {
        constexpr int NC = Kc * Kc, PER = NC / 32;
        const int e = 2 * m, o = 2 * m + 1;
        unsigned tj[2][C::TW];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int w = 0; w < C::TW; ++w) tj[h][w] = s.kt[cur][o * Kc + h * 32 + lane][w];
#pragma unroll 4
        for (int t = 0; t < PER; ++t) s.big[t * 32 + lane] = __float_as_uint(0.0f);
#pragma unroll 1
        for (int lb = 0; lb < L; ++lb) {
            const int b = o * L + lb;
            const int sj0 = nib<C::TW>(tj[0], lb), sj1 = nib<C::TW>(tj[1], lb);
            build_tables(s, G, e, b, lane);
            __syncwarp();
#pragma unroll 2
            for (int t = 0; t < PER; ++t) {
                const int i = t >> 1;
                const int sjb = (t & 1) ? sj1 : sj0;
                unsigned ti[C::TW];
#pragma unroll
                for (int w = 0; w < C::TW; ++w) ti[w] = s.kt[cur][e * Kc + i][w];
                float w = 0.0f;
#pragma unroll
                for (int la = 0; la < L; ++la) w = w + s.tab[la][nib<C::TW>(ti, la) * CUT1 + sjb];
                s.big[t * 32 + lane] = __float_as_uint(__uint_as_float(s.big[t * 32 + lane]) + w);
            }
            __syncwarp();
        }
        // keys; each lane keeps the smallest (key, t) of its own 128
        const float dj0 = s.kd[cur][o * Kc + lane], dj1 = s.kd[cur][o * Kc + 32 + lane];
        unsigned lk = KEY_REMOVED;
        int lt = 0;
#pragma unroll 2
        for (int t = 0; t < PER; ++t) {
            const float base = s.kd[cur][e * Kc + (t >> 1)] + ((t & 1) ? dj1 : dj0);
            const unsigned k = fkey(fmaf(2.0f, __uint_as_float(s.big[t * 32 + lane]), base));
            s.big[t * 32 + lane] = k;
            if (k < lk) {
                lk = k;
                lt = t;
            }
        }
#pragma unroll 1
        for (int r = 0; r < NEWK; ++r) {
            const unsigned mk = __reduce_min_sync(FULL, lk);
            const unsigned c = (lk == mk) ? (unsigned)(lt * 32 + lane) : 0xffffffffu;
            const unsigned cw = __reduce_min_sync(FULL, c);
            if (lane == 0) emit(s, cur, m, r, mk, (int)cw);
            if (NEWK > 1 && c == cw) {  // the owner removes it and rescans its column
                s.big[cw] = KEY_REMOVED;
                lk = KEY_REMOVED;
                lt = 0;
#pragma unroll 4
                for (int t = 0; t < PER; ++t) {
                    const unsigned k = s.big[t * 32 + lane];
                    if (k < lk) {
                        lk = k;
                        lt = t;
                    }
                }
            }
        }
        __syncwarp();
    }

    // Merge of group pair m with at most 1024 joint candidates: dots and keys in registers.
    __device__ static __forceinline__ void merge_regs(WarpSmem<K, N> &s, const float *__restrict__ G, int cur, int m,
                                                      int lane) {
        {
            const int e = 2 * m, o = 2 * m + 1;
            // this lane's fixed right-hand candidate j and its slot tuple
            const int j = lane % Kc;
            unsigned tj[C::TW];
            if constexpr (L == 1) {
                tj[0] = (unsigned)j;
            } else {
#pragma unroll
                for (int w = 0; w < C::TW; ++w) tj[w] = s.kt[cur][o * Kc + j][w];
            }
            float dot[CPL];
#pragma unroll
            for (int t = 0; t < CPL; ++t) dot[t] = 0.0f;

#pragma unroll 1
            for (int lb = 0; lb < L; ++lb) {
                const int b = o * L + lb;
                const int sjb = nib<C::TW>(tj, lb);
                build_tables(s, G, e, b, lane);
                __syncwarp();
                // ---- accumulate: dot(i,j) += sum_a D_ab[slot_i(a)][slot_j(b)] ----
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const int i = (t * 32 + lane) / Kc;
                    unsigned ti[C::TW];
                    if constexpr (L == 1) {
                        ti[0] = (unsigned)i;
                    } else {
#pragma unroll
                        for (int w = 0; w < C::TW; ++w) ti[w] = s.kt[cur][e * Kc + i][w];
                    }
                    float w = 0.0f;
#pragma unroll
                    for (int la = 0; la < L; ++la) w = w + s.tab[la][nib<C::TW>(ti, la) * CUT1 + sjb];
                    dot[t] = dot[t] + w;
                }
                __syncwarp();
            }
            // ---- joint deltas (quantization.py:533-535 minus the common |x_err|^2) ----
            unsigned key[CPL];
            const float dj = s.kd[cur][o * Kc + j];
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int i = (t * 32 + lane) / Kc;
                const float base = s.kd[cur][e * Kc + i] + dj;
                key[t] = fkey(fmaf(2.0f, dot[t], base));
            }
            // ---- keep the NEWK best, ascending, ties by flat index i*Kc + j (quantization.py:474-487) ----
#pragma unroll 1
            for (int r = 0; r < NEWK; ++r) {
                unsigned mk;
                int c;
                extract_min<CPL>(key, lane, mk, c);
                // candidate c = t*32 + lane  ->  flat = i*Kc + j with i = c / Kc, j = c % Kc: identical numbering
                if (lane == 0) emit(s, cur, m, r, mk, c);
            }
            __syncwarp();
        }
    }

    __device__ static __forceinline__ void run(WarpSmem<K, N> &s, const float *__restrict__ G, int cur, int lane) {
        static_assert(Kc * Kc >= 32 && Kc <= 64, "joint candidate count per merge must be 64..4096");
        const int nxt = cur ^ 1;
#pragma unroll 1
        for (int m = 0; m < NEWN; ++m) {
            if constexpr (Kc == 64)
                merge_big(s, G, cur, m, lane);
            else
                merge_regs(s, G, cur, m, lane);
        }
        if constexpr (NEWN > 1) {
            // used-slot masks of the new lists
#pragma unroll 1
            for (int m = 0; m < NEWN; ++m) {
                unsigned t[(NEWK + 31) / 32][C::TW];
#pragma unroll
                for (int h = 0; h < (NEWK + 31) / 32; ++h)
#pragma unroll
                    for (int w = 0; w < C::TW; ++w)
                        t[h][w] = (h * 32 + lane < NEWK) ? s.kt[nxt][m * NEWK + h * 32 + lane][w] : 0u;
#pragma unroll 1
                for (int la = 0; la < 2 * L; ++la) {
                    unsigned bit = 0u;
#pragma unroll
                    for (int h = 0; h < (NEWK + 31) / 32; ++h)
                        if (h * 32 + lane < NEWK) bit |= 1u << nib<C::TW>(t[h], la);
                    bit = __reduce_or_sync(FULL, bit);
                    if (lane == 0) s.um[m * 2 * L + la] = (unsigned short)bit;
                }
            }
            __syncwarp();
            Level<K, N, NEWN, NEWK, 2 * L>::run(s, G, nxt, lane);
        }
    }
};

// One pass for the frame owned by this warp.  s.old holds the indexes on entry and on exit.
template <int K, int N>
__device__ __forceinline__ void refine_pass(WarpSmem<K, N> &s, const float *__restrict__ Pb,
                                            const float *__restrict__ G, int lane) {
    using C = Cfg<K, N>;
    const size_t NK = C::NK;
    const float *diag = G + NK * NK;
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
        float acc[C::CPL1];
#pragma unroll
        for (int t = 0; t < C::CPL1; ++t) acc[t] = 0.0f;
#pragma unroll 1
        for (int m = 0; m < N; ++m) {
            if (m == n) continue;
            const float *row = G + ((size_t)m * K + s.old[m]) * NK + (size_t)n * K;
#pragma unroll
            for (int t = 0; t < C::CPL1; ++t) {
                const int k = t * 32 + lane;
                if (k < K) acc[t] = acc[t] + __ldg(row + k);
            }
        }
        const int on = s.old[n];
        float v[C::CPL1];
        float vo = 0.0f;
#pragma unroll
        for (int t = 0; t < C::CPL1; ++t) {
            const int k = t * 32 + lane;
            v[t] = 0.0f;
            if (k < K) {
                const float cross = acc[t] - __ldg(Pb + (size_t)n * K + k);
                v[t] = fmaf(2.0f, cross, __ldg(diag + (size_t)n * K + k));
            }
            if (k == on) vo = v[t];
        }
        const float vold = __shfl_sync(FULL, vo, on & 31);
        unsigned key[C::CPL1];
#pragma unroll
        for (int t = 0; t < C::CPL1; ++t) {
            const int k = t * 32 + lane;
            key[t] = (k < K) ? fkey(v[t] - vold) : KEY_REMOVED;
        }
#pragma unroll 1
        for (int r = 0; r < C::CUT1; ++r) {
            unsigned mk;
            int c;
            extract_min<C::CPL1>(key, lane, mk, c);
            if (lane == 0) {
                if constexpr (N == 1) {
                    s.old[0] = c;
                } else {
                    s.kd[0][n * C::CUT1 + r] = fkey_inv(mk);
                    s.kk[n][r] = (unsigned char)c;
                }
            }
        }
        if constexpr (N > 1) {
            if (lane == 0) s.um[n] = (unsigned short)((1u << C::CUT1) - 1u);
        }
    }
    __syncwarp();
    if constexpr (N > 1) Level<K, N, N, C::CUT1, 1>::run(s, G, 0, lane);
    __syncwarp();
}

template <int K, int N>
__global__ void __launch_bounds__(256) search_kernel(const float *__restrict__ P, const float *__restrict__ G,
                                                     int64_t B, int iters, const int32_t *__restrict__ idx_in,
                                                     int32_t *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    WarpSmem<K, N> &s = reinterpret_cast<WarpSmem<K, N> *>(smem_raw)[warp];
    for (int64_t b = (int64_t)blockIdx.x * wpc + warp; b < B; b += (int64_t)gridDim.x * wpc) {
        if (lane < N) s.old[lane] = idx_in[(size_t)b * N + lane];
        if (N > 32 && lane + 32 < N) s.old[lane + 32] = idx_in[(size_t)b * N + lane + 32];
        __syncwarp();
        const float *Pb = P + (size_t)b * Cfg<K, N>::NK;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            int prev0 = (lane < N) ? s.old[lane] : 0;
            int prev1 = (N > 32 && lane + 32 < N) ? s.old[lane + 32] : 0;
            refine_pass<K, N>(s, Pb, G, lane);
            int now0 = (lane < N) ? s.old[lane] : 0;
            int now1 = (N > 32 && lane + 32 < N) ? s.old[lane + 32] : 0;
            // a pass that returns its input is a fixed point of a deterministic map: the remaining passes are no-ops
            if (__all_sync(FULL, prev0 == now0 && prev1 == now1)) break;
        }
        if (lane < N) idx_out[(size_t)b * N + lane] = s.old[lane];
        if (N > 32 && lane + 32 < N) idx_out[(size_t)b * N + lane + 32] = s.old[lane + 32];
        __syncwarp();
    }
}

template <int K, int N>
int launch_one(const float *P, const float *G, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
               cudaStream_t st) {
    constexpr size_t per_warp = sizeof(WarpSmem<K, N>);
    int wpc = 8;
    while (wpc > 1 && per_warp * wpc > 200 * 1024) wpc /= 2;
    const size_t smem = per_warp * wpc;
    auto kern = search_kernel<K, N>;
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
    MCQ_LAUNCH_CHECK("search_kernel");
    return MCQ_OK;
}

template <int K>
int dispatch_n(int N, const float *P, const float *G, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
               cudaStream_t st) {
    switch (N) {
        case 1: return launch_one<K, 1>(P, G, B, iters, idx_in, idx_out, st);
        case 2: if constexpr (K >= 16) return launch_one<K, 2>(P, G, B, iters, idx_in, idx_out, st); break;
        case 4: if constexpr (K >= 16) return launch_one<K, 4>(P, G, B, iters, idx_in, idx_out, st); break;
        case 8: if constexpr (K >= 16) return launch_one<K, 8>(P, G, B, iters, idx_in, idx_out, st); break;
        case 16: if constexpr (K >= 16) return launch_one<K, 16>(P, G, B, iters, idx_in, idx_out, st); break;
        case 32: if constexpr (K >= 16) return launch_one<K, 32>(P, G, B, iters, idx_in, idx_out, st); break;
        case 64: if constexpr (K >= 16) return launch_one<K, 64>(P, G, B, iters, idx_in, idx_out, st); break;
        default: break;
    }
    set_error("search: (K=%d, N=%d) is not supported by this build", K, N);
    return MCQ_EUNSUPPORTED;
}

}  // namespace

int launch_search(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
                  int32_t *idx_out, cudaStream_t st, unsigned *work_counter) {
    if (B <= 0) return MCQ_OK;
    {
        // MCQ_SEARCH=v1 keeps the generic first version for every shape (used by the tests to cross-check)
        const char *e = getenv("MCQ_SEARCH");
        const bool force_v1 = e && strcmp(e, "v1") == 0;
        if (!force_v1 && search2_supports(N, K)) return launch_search2(P, gram, B, N, K, iters, idx_in, idx_out, st, work_counter);
        if (!force_v1 && search_k16_supports(N, K)) return launch_search_k16(P, gram, B, iters, idx_in, idx_out, st, work_counter);
    }
    switch (K) {
        case 2: return dispatch_n<2>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 4: return dispatch_n<4>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 8: return dispatch_n<8>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 16: return dispatch_n<16>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 32: return dispatch_n<32>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 64: return dispatch_n<64>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 128: return dispatch_n<128>(N, P, gram, B, iters, idx_in, idx_out, st);
        case 256: return dispatch_n<256>(N, P, gram, B, iters, idx_in, idx_out, st);
        default: break;
    }
    set_error("search: codebook_size %d not supported", K);
    return MCQ_EUNSUPPORTED;
}

}  // namespace mcq
