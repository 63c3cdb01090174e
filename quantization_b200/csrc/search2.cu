// search2.cu -- second version of the refinement search (all passes of Quantizer._refine_indexes,
// quantization.py:308-547, for a batch of frames in ONE launch) for codebook_size 256 and 2, 4, 8 or 16 codebooks: the
// inference configurations.  Same tables (P = x Cs^T per frame, G = Cs Cs^T per parameter version), same
// arithmetic contract and tie rules as search.cu / oracle/mcq_gram_model.c -- the tests compare both kernels with
// that model bit for bit -- but organised around what the first version's profile showed (profiles/r01_ncu_summary.md:
// issue bound, 55 % of the instructions in the sorted top-k extraction, 25 % in building difference tables):
//
//   * one warp per frame; a lane owns 8 CONSECUTIVE candidates (flat = lane*8 + t), so level-1 rows are float4 loads
//     and "lowest lane among equals" is "lowest flat index among equals" (the contract's tie rule);
//   * sorted top-R: each lane rank-sorts its 8 keys into its own shared-memory column, then R steps of
//     redux.sync.min.f32 (CREDUX) pop the global minimum from the column heads -- one reduction per pop (predicated
//     PTX, pop_step), with an exact two-reduction loop as the fallback when two lanes hold bit-equal keys;
//   * every u/v term of a difference D_ab(p,q) = ((G[ap,bq] - G[ap,b_old]) - G[a_old,bq]) + G[a_old,b_old] comes from one
//     cached gather uv[a][m][p] = G[(m,old_m),(a,kk_a[p])] (G is bitwise symmetric);
//   * merge of single codebooks (16x16): one G gather per joint candidate; merge of codebook pairs (16x16): four;
//     merge of codebook quads (32x32): 16x16 tables T_ab per codebook pair, folded per candidate row into
//     E_b[i][q] = sum_a T_ab[i_a][q] -- exactly the contract's inner sum -- then dot(i,j) = sum_b E_b[i][j_b]; the
//     tables hold only the slots the surviving candidates still use (columns compacted, unused rows skipped).
// Round-2 measurements, ablations and the variants that lost are in profiles/r02_search.md.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mcq {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int K2 = 256;
#ifndef MCQ_POP_UNROLL
#define MCQ_POP_UNROLL 4  // the pop loop is the bulk of the code; full unrolling costs instruction-cache misses
#endif
constexpr int POP_UNROLL = MCQ_POP_UNROLL;
// Interleave the pop chains of two merges of the same level (needs a second list region: +2.8 KB of shared memory per
// warp).  0 = one merge at a time.  Measured at C2 (75,776 frames, static striding): 5.17 ms paired vs 4.78 ms one at a
// time -- the extra shared memory costs more (L1 left for the gathers) than the second chain hides; default off.
#ifndef MCQ_MERGE_PAIR
#define MCQ_MERGE_PAIR 0
#endif
// Launch shape: ONE CTA of 16 warps per SM (128 registers per thread).  Measured at C2 (75,776 frames, static striding):
// 4 CTAs x 5 warps (96 regs) 5.59 ms, 2 x 8 (128 regs) 5.12 ms, 1 x 20 (96 regs) 5.40 ms, 1 x 12 (152 regs) 5.23 ms,
// 1 x 16 (128 regs) 4.90 ms: fewer, fatter warps win (no spills, and the shared memory one CTA does not take stays L1).
// N = 2, 4: one CTA of 24 warps (0.92 vs 0.97 ms for 3 x 8 at N = 4); N = 16: two CTAs of 8 warps (10.9 vs 11.2 ms for 4 x 4).
#ifndef MCQ_S2_MINB
#define MCQ_S2_MINB 1
#endif
#ifndef MCQ_S2_WPC
#define MCQ_S2_WPC 16
#endif
#ifndef MCQ_S2_WPC4
#define MCQ_S2_WPC4 24
#endif
#ifndef MCQ_S2_MINB4
#define MCQ_S2_MINB4 1
#endif
#ifndef MCQ_S16_PAIR
#define MCQ_S16_PAIR 1  // interleave the level-1 selections of two codebooks in the 16-codebook kernel as well
#endif
#ifndef MCQ_S16_WPC
#define MCQ_S16_WPC 8
#endif
#ifndef MCQ_S16_MINB
#define MCQ_S16_MINB 2
#endif

__device__ __forceinline__ float credux_min(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

// Scattered 4-byte read of G (element offset).  A texture-object variant of these gathers (TEX data pipe instead of
// the LSU pipe) was measured and is slower (9.3 ms vs 7.2 ms per 75,776 frames, profiles/r01_search2.md).
__device__ __forceinline__ float gat(const float *__restrict__ G, unsigned idx) { return __ldg(G + idx); }

// (a0, a1) += (b0, b1): one FADD2 (packed fp32 add, each half an IEEE round-to-nearest add)
__device__ __forceinline__ void fadd2(float &a0, float &a1, float b0, float b1) {
    asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\t"
        "mov.b64 {%0, %1}, x;\n\t}"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}

// -1 if a <= b else 0 (one FSET)
__device__ __forceinline__ int le_mask(float a, float b) {
    int r;
    asm("set.le.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}

constexpr int TSTR = 20;  // row stride (floats) of the 16-column tables: 80 B keeps float4 rows bank-spread
// float offset of row p of a 16-column table: rows 8..15 are shifted by 16 floats so that a warp writing rows
// (t, t+8) x 16 columns hits 32 distinct banks
__device__ __forceinline__ int trow_off(int p) { return p * TSTR + ((p >> 3) << 4); }
constexpr int TAB_FLOATS = 16 * TSTR + 16;

template <int N>
struct alignas(16) WarpMem2 {
    static constexpr int NG2 = (N >= 2) ? N / 2 : 1;  // groups after the first merge
    static constexpr int NG3 = (N >= 4) ? N / 4 : 1;  // groups after the second merge
    union {
        float2 lists[9][32];  // per-lane sorted columns (key, flat) + one row of sentinels (list_sentinel)
        float es[32][TSTR];   // quad merge: E_b[i][q]  (overwrites the sentinel row; restored after the merge)
    };
#if MCQ_MERGE_PAIR
    union {
        float tab[TAB_FLOATS];   // T_ab of the quad merge (trow_off)
        float2 lists2[10][32];   // second column set of the interleaved merge selections (merge1_pair, merge2_pair)
    };
    float2 sel[32];          // result of the current selection: (key, flat), ascending
    float2 sel2[32];         // result of the second of two interleaved merge selections
    float kd2[NG2][16];      // kept deltas / slot tuples after the first merge
    unsigned kt2[NG2][16];
#else
    // the quad merge's table never lives at the same time as the selection results and the first merge's survivors
    // (they are dead once the second merge has stored kd3 / kt3): sharing the bytes keeps 16 warps under 164 KB, the
    // next smaller shared-memory carve-out, which leaves L1 92 KB instead of 60
    union {
        float tab[TAB_FLOATS];   // T_ab of the quad merge (trow_off)
        struct {
            float2 sel[32];      // result of the current selection: (key, flat), ascending
            float kd2[NG2][16];  // kept deltas / slot tuples after the first merge
            unsigned kt2[NG2][16];
        };
    };
#endif
    float kd1[N][16];        // level-1 kept candidates of each codebook: delta ...
    int kk[N][16];           // ... and codebook entry k
    unsigned rowk[N][16];    // (n*K + kk[n][p]) * NK: element offset of the G row of each kept candidate
    float uv[N][N][16];      // uv[a][m][p] = G[(m,old_m), (a, kk_a[p])]
    float kd3[NG3][32];      // ... after the second merge
    unsigned kt3[NG3][32];
    int old[N];              // indexes at the start of the pass (and its result)
    unsigned rowoff[N];      // (m*K + old[m]) * NK: element offset of the G row of each current entry
    unsigned used[8];        // quad merge: which level-1 slots of each codebook the 32+32 candidates still use
};

// Scratch of the second of two interleaved level-1 selections (level1<.., PAIR>): its column lists live in the (not yet
// gathered) uv region, its result in the (not yet written) kd2 region.
template <int N>
__device__ __forceinline__ float2 (*pair_lists2(WarpMem2<N> &s))[32] {
    static_assert(sizeof(s.uv) >= 10 * 32 * sizeof(float2), "pair memory");
    return reinterpret_cast<float2(*)[32]>(&s.uv[0][0][0]);
}
template <int N>
__device__ __forceinline__ float2 *pair_sel2(WarpMem2<N> &s) {
    static_assert(sizeof(s.kd2) >= 16 * sizeof(float2), "pair memory");
    return reinterpret_cast<float2 *>(&s.kd2[0][0]);
}
struct WarpMem16;
__device__ __forceinline__ float2 (*pair_lists2(WarpMem16 &s))[32];
__device__ __forceinline__ float2 *pair_sel2(WarpMem16 &s);

// Sentinel below the 8 entries of a lane's column: NaN.  redux.sync.min.f32 ignores NaN inputs (the result is NaN only
// when every lane's head is the sentinel) and `head == min` is false for it, so an exhausted lane never pops again and
// its column pointer never leaves the list (row 9, the prefetched successor of the sentinel, is still inside the union).
__device__ __forceinline__ float2 list_sentinel() { return make_float2(__int_as_float(0x7fc00000), __int_as_float(0)); }

// ---- sorted top-R of the warp's 256 candidates ------------------------------------------------------------------------
// Every lane rank-sorts its 8 keys into its own shared-memory column (s.lists, row 8 = NaN sentinels); R "pops" then
// take the global minimum from the column heads.  The contract orders candidates by (key, flat index).
//
// Fast path (one warp reduction per pop): the lane whose head EQUALS the minimum pops.  That is the contract's order
// unless two lanes hold bit-equal keys at the same time (then both pop in one step).  Such a step is not looked for
// while popping -- after the loop the number of popped candidates is summed over the warp, and if it is not exactly R
// (a tie somewhere, or NaN keys, which never pop) the selection is redone by the exact loop below, which breaks ties
// with a second reduction over the flat indexes.  Measured: exact ties occur on degenerate inputs only (all-zero
// frames, duplicated codebook rows); the profile of the two-reduction loop had 39 % of the kernel's stall samples.
template <class Mem, int R>
__device__ __noinline__ void pops_exact(Mem &s, float2 (*lists)[32], float2 *sel, int lane) {
    __syncwarp();  // sel[] holds the fast path's stores of other lanes: order them before this loop's
    const float2 *col = &lists[0][lane];
    float2 head = col[0], nxt = col[32];
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
        const float m = credux_min(head.x);
        // among the lanes holding the minimum the lowest flat index wins (contract: ascending (key, flat))
        const unsigned f = (head.x == m) ? (unsigned)__float_as_int(head.y) : 0x7fffffffu;
        const bool mine = (f == __reduce_min_sync(FULL, f)) && f != 0x7fffffffu;
        if (mine) {
            sel[r] = head;
            head = nxt;
            col += 32;
            nxt = col[32];  // at most row 9: inside the union (es) even after the sentinel row
        }
    }
    __syncwarp();
}

// One pop step of the fast path in predicated PTX with 32-bit shared-memory addresses: 8 instructions.  (The C++ form
// of the same statements compiled to 16: 64-bit pointer arithmetic on the generic column pointer and register copies;
// the pop loops are 256 steps per frame-pass -- profiles/r02_search.md.)
//   m = min over the warp of the column heads (NaN sentinels ignored); the lane whose head equals m stores it to
//   sel[r], takes its prefetched successor as the new head and prefetches the entry after that.
//   hk/hf: head (key, flat), nk/nf: its successor, a: shared address of the head's row in this lane's column (rows are
//   256 bytes apart), sel_a: shared address of sel[first step of the group], U: step within the group.
// When two lanes hold bit-equal minima both pop in the same step and both store to sel[r]: a write-after-write that
// compute-sanitizer's racecheck reports.  It is benign -- the popped count is then not R and pops_exact redoes the whole
// selection, overwriting sel -- and -DMCQ_POP_RACEFREE=1 proves it is the only one: the store is then issued only by a
// lane that pops alone (one vote + three integer instructions more per pop: 4.76 vs 4.50 ms per launch), racecheck
// reports no error and the
// results are bit-identical (profiles/r02_racecheck_search.log).
#ifndef MCQ_POP_RACEFREE
#define MCQ_POP_RACEFREE 0
#endif
template <int U>
__device__ __forceinline__ void pop_step(float &hk, float &hf, float &nk, float &nf, unsigned &a, unsigned sel_a) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .f32 m;\n\t"
        "redux.sync.min.f32 m, %0, 0xffffffff;\n\t"
        "setp.eq.f32 p, %0, m;\n\t"
#if MCQ_POP_RACEFREE
        ".reg .pred s;\n\t"
        ".reg .b32 b, c;\n\t"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
        "add.u32 c, b, -1;\n\t"
        "and.b32 c, c, b;\n\t"
        "setp.eq.and.u32 s, c, 0, p;\n\t"
        "@s st.shared.v2.f32 [%5+%6], {%0, %1};\n\t"
#else
        "@p st.shared.v2.f32 [%5+%6], {%0, %1};\n\t"
#endif
        "@p mov.f32 %0, %2;\n\t"
        "@p mov.f32 %1, %3;\n\t"
        "@p ld.shared.v2.f32 {%2, %3}, [%4+512];\n\t"
        "@p add.u32 %4, %4, 256;\n\t"
        "}"
        : "+f"(hk), "+f"(hf), "+f"(nk), "+f"(nf), "+r"(a)
        : "r"(sel_a), "n"(U * 8)
        : "memory");
}

// R pops of one column set; returns the number of entries this lane popped
template <int R>
__device__ __forceinline__ int pop_loop(float2 (*lists)[32], float2 *sel, int lane) {
    static_assert(R % 4 == 0, "pop groups of four");
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(&lists[0][lane]);
    unsigned a = a0, sa = (unsigned)__cvta_generic_to_shared(sel);
    float2 head = lists[0][lane], nxt = lists[1][lane];  // the successor is fetched ahead of the pop that needs it
#pragma unroll(POP_UNROLL / 4 > 0 ? POP_UNROLL / 4 : 1)
    for (int r0 = 0; r0 < R; r0 += 4) {
        pop_step<0>(head.x, head.y, nxt.x, nxt.y, a, sa);
        pop_step<1>(head.x, head.y, nxt.x, nxt.y, a, sa);
        pop_step<2>(head.x, head.y, nxt.x, nxt.y, a, sa);
        pop_step<3>(head.x, head.y, nxt.x, nxt.y, a, sa);
        sa += 32;
    }
    return (int)((a - a0) >> 8);
}

// two interleaved chains; returns popped(A) + (popped(B) << 16)
template <int R>
__device__ __forceinline__ int pop_loop2(float2 (*listsA)[32], float2 *selA, float2 (*listsB)[32], float2 *selB, int lane) {
    static_assert(R % 4 == 0, "pop groups of four");
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(&listsA[0][lane]);
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(&listsB[0][lane]);
    unsigned a = a0, b = b0;
    unsigned sa = (unsigned)__cvta_generic_to_shared(selA), sb = (unsigned)__cvta_generic_to_shared(selB);
    float2 hA = listsA[0][lane], nA = listsA[1][lane], hB = listsB[0][lane], nB = listsB[1][lane];
#pragma unroll(POP_UNROLL / 4 > 0 ? POP_UNROLL / 4 : 1)
    for (int r0 = 0; r0 < R; r0 += 4) {
        pop_step<0>(hA.x, hA.y, nA.x, nA.y, a, sa);
        pop_step<0>(hB.x, hB.y, nB.x, nB.y, b, sb);
        pop_step<1>(hA.x, hA.y, nA.x, nA.y, a, sa);
        pop_step<1>(hB.x, hB.y, nB.x, nB.y, b, sb);
        pop_step<2>(hA.x, hA.y, nA.x, nA.y, a, sa);
        pop_step<2>(hB.x, hB.y, nB.x, nB.y, b, sb);
        pop_step<3>(hA.x, hA.y, nA.x, nA.y, a, sa);
        pop_step<3>(hB.x, hB.y, nB.x, nB.y, b, sb);
        sa += 32;
        sb += 32;
    }
    return (int)((a - a0) >> 8) + ((int)((b - b0) >> 8) << 16);
}

// (key[t], flat[t]) of a lane -> its column of `lists`, ascending by (key, t):
//   rank[t] = #{u < t: key[u] <= key[t]} + #{u > t: key[u] < key[t]}
__device__ __forceinline__ void rank_sort_store(float2 (*lists)[32], const float (&key)[8], const int (&flat)[8], int lane) {
    int rank[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) rank[t] = 7 - t;
#pragma unroll
    for (int t = 1; t < 8; ++t)
#pragma unroll
        for (int u = 0; u < t; ++u) {
            const int c = le_mask(key[u], key[t]);  // -1 when key[u] sorts before key[t]
            rank[t] -= c;
            rank[u] += c;
        }
#pragma unroll
    for (int t = 0; t < 8; ++t) lists[rank[t]][lane] = make_float2(key[t], __int_as_float(flat[t]));
}

// The R smallest of the warp's 256 candidates (8 per lane; candidate t of a lane has flat index flat[t], ascending in
// t), ascending by (key, flat), written to s.sel[0..R).  quantization.py:474-487 (sort + keep the first K_cutoff).
template <class Mem, int R>
__device__ __forceinline__ void select_sorted(Mem &s, const float (&key)[8], const int (&flat)[8], int lane) {
    rank_sort_store(s.lists, key, flat, lane);
    // a lane reads back only its own column: no warp synchronisation needed here.  The prefetch reaches at most row
    // 9: inside the union (es) even after the sentinel row.
    const int popped = pop_loop<R>(s.lists, s.sel, lane);
    if (__reduce_add_sync(FULL, popped) != R) pops_exact<Mem, R>(s, s.lists, s.sel, lane);
    __syncwarp();
}

// Two independent selections at once (same flat numbering): the two pop chains are interleaved, so the latency of one
// chain's warp reduction is covered by the other's.  The second selection uses caller-supplied memory: `lists2` (10 rows
// of 32 float2; row 8 gets the sentinels here, row 9 is only ever prefetched) and `sel2` (R entries).
template <class Mem, int R>
__device__ __forceinline__ void select_sorted2(Mem &s, const float (&keyA)[8], const float (&keyB)[8],
                                               const int (&flat)[8], int lane, float2 (*lists2)[32], float2 *sel2) {
    rank_sort_store(s.lists, keyA, flat, lane);
    rank_sort_store(lists2, keyB, flat, lane);
    lists2[8][lane] = list_sentinel();
    // both chains in one reduction: A's count in the low half, B's in the high half
    const int popped = pop_loop2<R>(s.lists, s.sel, lists2, sel2, lane);
    const int tot = __reduce_add_sync(FULL, popped);
    if ((tot & 0xffff) != R) pops_exact<Mem, R>(s, s.lists, s.sel, lane);
    if ((tot >> 16) != R) pops_exact<Mem, R>(s, lists2, sel2, lane);
    __syncwarp();
}

// Flat index of the smallest of the warp's 256 candidates (lowest flat index among equals).
__device__ __forceinline__ int select_best(const float (&key)[8], const int (&flat)[8], int lane) {
    float best = key[0];
    int bf = flat[0];
#pragma unroll
    for (int t = 1; t < 8; ++t)
        if (key[t] < best) {
            best = key[t];
            bf = flat[t];
        }
    const float m = credux_min(best);
    const unsigned f = (best == m) ? (unsigned)bf : 0x7fffffffu;
    const unsigned w = __reduce_min_sync(FULL, f);
    return w == 0x7fffffffu ? 0 : (int)w;  // the guard is only reachable with NaN scores
}

// Level-1 deltas of codebook n (quantization.py:401-418 with the per-codebook constants dropped):
//   key[k] = v[k] - v[old_n],  v[k] = fmaf(2, sum_{m != n} G[(m,old_m),(n,k)] - P[n,k], G[(n,k),(n,k)])
// Lane L owns entries 4L..4L+3 and 128+4L..128+4L+3, so every float4 request of the warp is 512 contiguous bytes.
template <int N, class Mem>
__device__ __forceinline__ void level1_keys(Mem &s, const float *__restrict__ Pb, const float *__restrict__ Gp, int n,
                                            int lane, float (&key)[8]) {
    constexpr unsigned NK = N * K2;
    const float *diag = Gp + (size_t)NK * NK;
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.0f;
    const unsigned colbase = n * K2 + lane * 4;
#pragma unroll(N <= 8 ? N - 1 : 5)  // rows in flight per batch: all N-1 up to 8 codebooks, 5 of the 15 at N = 16
    for (int mm = 0; mm < N - 1; ++mm) {
        const int m = mm + (mm >= n ? 1 : 0);  // ascending m, skipping n
        const float4 *row = reinterpret_cast<const float4 *>(Gp + (s.rowoff[m] + colbase));
        const float4 a = __ldg(row), b = __ldg(row + 32);
        fadd2(acc[0], acc[1], a.x, a.y);
        fadd2(acc[2], acc[3], a.z, a.w);
        fadd2(acc[4], acc[5], b.x, b.y);
        fadd2(acc[6], acc[7], b.z, b.w);
    }
    const float4 *pp = reinterpret_cast<const float4 *>(Pb + colbase);
    const float4 *dp = reinterpret_cast<const float4 *>(diag + colbase);
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 32), d0 = __ldg(dp), d1 = __ldg(dp + 32);
    float v[8];
    v[0] = fmaf(2.0f, acc[0] - p0.x, d0.x);
    v[1] = fmaf(2.0f, acc[1] - p0.y, d0.y);
    v[2] = fmaf(2.0f, acc[2] - p0.z, d0.z);
    v[3] = fmaf(2.0f, acc[3] - p0.w, d0.w);
    v[4] = fmaf(2.0f, acc[4] - p1.x, d1.x);
    v[5] = fmaf(2.0f, acc[5] - p1.y, d1.y);
    v[6] = fmaf(2.0f, acc[6] - p1.z, d1.z);
    v[7] = fmaf(2.0f, acc[7] - p1.w, d1.w);
    // v of the current entry old[n]: it lives in lane (old >> 2) & 31 as element (old & 3) + 4 * (old >= 128)
    const int on = s.old[n];
    const int tsel = (on & 3) | ((on >> 5) & 4);
    float vs = v[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) vs = (tsel == t) ? v[t] : vs;
    const float vold = __shfl_sync(FULL, vs, (on >> 2) & 31);
#pragma unroll
    for (int t = 0; t < 8; ++t) key[t] = v[t] - vold;
}

template <int N, class Mem>
__device__ __forceinline__ void level1_store(Mem &s, int n, int lane, const float2 *sel) {
    constexpr unsigned NK = N * K2;
    if (lane < 16) {
        const float2 r = sel[lane];
        s.kd1[n][lane] = r.x;
        s.kk[n][lane] = __float_as_int(r.y);  // flat index == codebook entry k
        s.rowk[n][lane] = (unsigned)(n * K2 + __float_as_int(r.y)) * NK;
    }
}

// Level 1 + top-16 per codebook.  PAIR: two codebooks per step with interleaved pop chains; the second selection's
// lists live in the (not yet gathered) uv region and its result in the (not yet written) kd2 region.
template <int N, class Mem, bool PAIR>
__device__ __forceinline__ void level1(Mem &s, const float *__restrict__ Pb, const float *__restrict__ Gp, int lane) {
    int flat[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) flat[t] = (t < 4 ? 0 : 128 - 4) + lane * 4 + t;
    if constexpr (PAIR) {
        float2(*lists2)[32] = pair_lists2(s);  // scratch of the second chain: regions that are dead during level 1
        float2 *sel2 = pair_sel2(s);
#pragma unroll 1
        for (int n = 0; n < N; n += 2) {
            float keyA[8], keyB[8];
            level1_keys<N, Mem>(s, Pb, Gp, n, lane, keyA);
            level1_keys<N, Mem>(s, Pb, Gp, n + 1, lane, keyB);
            select_sorted2<Mem, 16>(s, keyA, keyB, flat, lane, lists2, sel2);
            level1_store<N, Mem>(s, n, lane, s.sel);
            level1_store<N, Mem>(s, n + 1, lane, sel2);
            __syncwarp();
        }
    } else {
#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            float key[8];
            level1_keys<N, Mem>(s, Pb, Gp, n, lane, key);
            select_sorted<Mem, 16>(s, key, flat, lane);
            level1_store<N, Mem>(s, n, lane, s.sel);
            __syncwarp();
        }
    }
}

template <int N>
__device__ __forceinline__ void gather_uv(WarpMem2<N> &s, const float *__restrict__ G, int lane) {
    const int p = lane & 15, mh = lane >> 4;
    unsigned ro[N / 2];
#pragma unroll
    for (int r = 0; r < N / 2; ++r) ro[r] = s.rowoff[2 * r + mh];
    constexpr int AB = (N >= 4) ? 4 : N;  // codebooks per batch: AB * N/2 loads in flight per lane
#pragma unroll 1
    for (int a0 = 0; a0 < N; a0 += AB) {
        float val[AB][N / 2];
#pragma unroll
        for (int aa = 0; aa < AB; ++aa) {
            const unsigned col = (a0 + aa) * K2 + s.kk[a0 + aa][p];
#pragma unroll
            for (int r = 0; r < N / 2; ++r) val[aa][r] = gat(G, ro[r] + col);  // (m == a is read but not used)
        }
#pragma unroll
        for (int aa = 0; aa < AB; ++aa)
#pragma unroll
            for (int r = 0; r < N / 2; ++r) s.uv[a0 + aa][2 * r + mh][p] = val[aa][r];
    }
    __syncwarp();
}

// Merge of two single codebooks e = 2g, o = 2g+1: 16 x 16 joint candidates (quantization.py:504-547 at L = 1).
// Lane (hi, j) scores candidates (i = 8*hi + t, j), t = 0..7: a request reads two G rows x 16 columns.
template <int N>
__device__ __forceinline__ void merge1_keys(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane, float (&key)[8]) {
    const int e = 2 * g, o = e + 1;
    const int j = lane & 15, ib = (lane >> 4) * 8;
    const unsigned ko = o * K2 + s.kk[o][j];
    const float v = s.uv[o][e][j];
    const float kdo = s.kd1[o][j];
    const float w = gat(G, s.rowoff[e] + o * K2 + s.old[o]);
    unsigned rowp[8];
    float u[8], kde[8];
    {
        const uint4 *rp = reinterpret_cast<const uint4 *>(&s.rowk[e][ib]);
        const float4 *up = reinterpret_cast<const float4 *>(&s.uv[e][o][ib]);
        const float4 *dp = reinterpret_cast<const float4 *>(&s.kd1[e][ib]);
        const uint4 r0 = rp[0], r1 = rp[1];
        const float4 u0 = up[0], u1 = up[1], d0 = dp[0], d1 = dp[1];
        rowp[0] = r0.x; rowp[1] = r0.y; rowp[2] = r0.z; rowp[3] = r0.w;
        rowp[4] = r1.x; rowp[5] = r1.y; rowp[6] = r1.z; rowp[7] = r1.w;
        u[0] = u0.x; u[1] = u0.y; u[2] = u0.z; u[3] = u0.w; u[4] = u1.x; u[5] = u1.y; u[6] = u1.z; u[7] = u1.w;
        kde[0] = d0.x; kde[1] = d0.y; kde[2] = d0.z; kde[3] = d0.w;
        kde[4] = d1.x; kde[5] = d1.y; kde[6] = d1.z; kde[7] = d1.w;
    }
    float gv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) gv[t] = gat(G, rowp[t] + ko);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float d = ((gv[t] - u[t]) - v) + w;
        key[t] = fmaf(2.0f, d, kde[t] + kdo);
    }
}

// flat index of the joint candidate (i = 8*hi + t, j) a lane scores in merge1 / merge2: i * 16 + j
__device__ __forceinline__ void merge_flat(int lane, int (&flat)[8]) {
    const int j = lane & 15, ib = (lane >> 4) * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) flat[t] = (ib + t) * 16 + j;
}

template <int N>
__device__ __forceinline__ void merge1_store(WarpMem2<N> &s, int g, int lane, const float2 *sel) {
    if (lane < 16) {
        const float2 r = sel[lane];
        const int fl = __float_as_int(r.y);
        s.kd2[g][lane] = r.x;
        s.kt2[g][lane] = (unsigned)(fl >> 4) | ((unsigned)(fl & 15) << 4);
    }
}

template <int N, bool FINAL>
__device__ __forceinline__ void merge1(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    const int e = 2 * g, o = e + 1;
    float key[8];
    int flat[8];
    merge1_keys<N>(s, G, g, lane, key);
    merge_flat(lane, flat);
    if constexpr (FINAL) {
        const int fl = select_best(key, flat, lane);
        __syncwarp();  // every lane has read old[] (above) before lane 0 overwrites it
        if (lane == 0) {
            const int ne = s.kk[e][fl >> 4], no = s.kk[o][fl & 15];
            s.old[e] = ne;
            s.old[o] = no;
        }
        __syncwarp();
    } else {
        select_sorted<WarpMem2<N>, 16>(s, key, flat, lane);
        merge1_store<N>(s, g, lane, s.sel);
        __syncwarp();
    }
}

// Two merges of single codebooks (g, g + 1) with interleaved pop chains (like the level-1 pairs).
template <int N>
__device__ __forceinline__ void merge1_pair(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    float keyA[8], keyB[8];
    int flat[8];
    merge1_keys<N>(s, G, g, lane, keyA);
    merge1_keys<N>(s, G, g + 1, lane, keyB);
    merge_flat(lane, flat);
    select_sorted2<WarpMem2<N>, 16>(s, keyA, keyB, flat, lane, s.lists2, s.sel2);
    merge1_store<N>(s, g, lane, s.sel);
    merge1_store<N>(s, g + 1, lane, s.sel2);
    __syncwarp();
}

// Merge of two codebook pairs: groups e = 2g (codebooks 4g, 4g+1) and o = 2g+1 (4g+2, 4g+3), 16 x 16 candidates.
// Same lane mapping as merge1: lane (hi, j) scores (i = 8*hi + t, j).
template <int N>
__device__ __forceinline__ void merge2_keys(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane, float (&key)[8]) {
    const int e = 2 * g, o = e + 1;
    const int a0 = 4 * g, a1 = a0 + 1, b0 = a0 + 2, b1 = a0 + 3;
    const int j = lane & 15, ib = (lane >> 4) * 8;
    const unsigned tj = s.kt2[o][j];
    const int q0 = tj & 15, q1 = tj >> 4;
    const unsigned c0 = b0 * K2 + s.kk[b0][q0], c1 = b1 * K2 + s.kk[b1][q1];
    const float v00 = s.uv[b0][a0][q0], v10 = s.uv[b0][a1][q0], v01 = s.uv[b1][a0][q1], v11 = s.uv[b1][a1][q1];
    const float kdo = s.kd2[o][j];
    const unsigned cb0 = b0 * K2 + s.old[b0], cb1 = b1 * K2 + s.old[b1];
    const float w00 = gat(G, s.rowoff[a0] + cb0), w10 = gat(G, s.rowoff[a1] + cb0);
    const float w01 = gat(G, s.rowoff[a0] + cb1), w11 = gat(G, s.rowoff[a1] + cb1);
    unsigned tis[8];
    float kde[8];
    {
        const uint4 *tp = reinterpret_cast<const uint4 *>(&s.kt2[e][ib]);
        const float4 *dp = reinterpret_cast<const float4 *>(&s.kd2[e][ib]);
        const uint4 t0 = tp[0], t1 = tp[1];
        const float4 d0 = dp[0], d1 = dp[1];
        tis[0] = t0.x; tis[1] = t0.y; tis[2] = t0.z; tis[3] = t0.w; tis[4] = t1.x; tis[5] = t1.y; tis[6] = t1.z; tis[7] = t1.w;
        kde[0] = d0.x; kde[1] = d0.y; kde[2] = d0.z; kde[3] = d0.w;
        kde[4] = d1.x; kde[5] = d1.y; kde[6] = d1.z; kde[7] = d1.w;
    }
    float g00[8], g10[8], g01[8], g11[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const unsigned rp0 = s.rowk[a0][tis[t] & 15], rp1 = s.rowk[a1][tis[t] >> 4];
        g00[t] = gat(G, rp0 + c0);
        g10[t] = gat(G, rp1 + c0);
        g01[t] = gat(G, rp0 + c1);
        g11[t] = gat(G, rp1 + c1);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int ia0 = tis[t] & 15, ia1 = tis[t] >> 4;
        const float d00 = ((g00[t] - s.uv[a0][b0][ia0]) - v00) + w00;
        const float d10 = ((g10[t] - s.uv[a1][b0][ia1]) - v10) + w10;
        const float d01 = ((g01[t] - s.uv[a0][b1][ia0]) - v01) + w01;
        const float d11 = ((g11[t] - s.uv[a1][b1][ia1]) - v11) + w11;
        const float wb0 = d00 + d10, wb1 = d01 + d11;  // inner sums over a, then b ascending
        const float dot = wb0 + wb1;
        key[t] = fmaf(2.0f, dot, kde[t] + kdo);
    }
}

template <int N>
__device__ __forceinline__ void merge2_store(WarpMem2<N> &s, int g, int lane, const float2 *sel) {
    const int e = 2 * g, o = e + 1;
    const float2 r = sel[lane];
    const int fl = __float_as_int(r.y);
    s.kd3[g][lane] = r.x;
    s.kt3[g][lane] = s.kt2[e][fl >> 4] | (s.kt2[o][fl & 15] << 8);
}

template <int N, bool FINAL>
__device__ __forceinline__ void merge2(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    const int e = 2 * g, o = e + 1;
    const int a0 = 4 * g, a1 = a0 + 1, b0 = a0 + 2, b1 = a0 + 3;
    float key[8];
    int flat[8];
    merge2_keys<N>(s, G, g, lane, key);
    merge_flat(lane, flat);
    if constexpr (FINAL) {
        const int fl = select_best(key, flat, lane);
        __syncwarp();  // every lane has read old[] (above) before lane 0 overwrites it
        if (lane == 0) {
            const unsigned te = s.kt2[e][fl >> 4], to = s.kt2[o][fl & 15];
            const int n0 = s.kk[a0][te & 15], n1 = s.kk[a1][te >> 4];
            const int n2 = s.kk[b0][to & 15], n3 = s.kk[b1][to >> 4];
            s.old[a0] = n0;
            s.old[a1] = n1;
            s.old[b0] = n2;
            s.old[b1] = n3;
        }
        __syncwarp();
    } else {
        select_sorted<WarpMem2<N>, 32>(s, key, flat, lane);
        merge2_store<N>(s, g, lane, s.sel);
        __syncwarp();
    }
}

// The two merges of codebook pairs of N = 8 (g = 0, 1) with interleaved pop chains.
template <int N>
__device__ __forceinline__ void merge2_pair(WarpMem2<N> &s, const float *__restrict__ G, int g, int lane) {
    float keyA[8], keyB[8];
    int flat[8];
    merge2_keys<N>(s, G, g, lane, keyA);
    merge2_keys<N>(s, G, g + 1, lane, keyB);
    merge_flat(lane, flat);
    select_sorted2<WarpMem2<N>, 32>(s, keyA, keyB, flat, lane, s.lists2, s.sel2);
    merge2_store<N>(s, g, lane, s.sel);
    merge2_store<N>(s, g + 1, lane, s.sel2);
    __syncwarp();
}

// Final merge of two codebook quads (N = 8): 32 x 32 joint candidates, candidate flat = i*32 + j.
template <int N>
__device__ __forceinline__ void merge4_final(WarpMem2<N> &s, const float *__restrict__ G, int lane) {
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
    // table T_ab: lane (hi, q) computes rows p = 8*hi + t of column q -- a request reads two G rows x 16 columns.
    // Only the slots the 32 + 32 candidates still use matter (5.8 of 16 per codebook on average, tools note in
    // profiles/r02_search.md): the columns of a table are COMPACTED to the used slots of codebook b (column q goes to
    // position popc(used_b below q)), so that a candidate row folds ceil(nq / 4) float4 instead of 4, and rows no
    // candidate uses are neither gathered nor stored.  Same values, same order of additions per (i, j).
    const int q = lane & 15, pb = (lane >> 4) * 8;
    const unsigned below_q = (1u << q) - 1u;
#pragma unroll 1
    for (int lb = 0; lb < 4; ++lb) {
        const int b = 4 + lb;
        const unsigned ub = s.used[4 + lb];
        const int nq4 = (__popc(ub) + 3) >> 2;                 // float4 groups of a compacted table row (warp-uniform)
        const int tbase = trow_off(pb) + __popc(ub & below_q);  // my compacted column
        const unsigned cq = b * K2 + s.kk[b][q];
        const bool colu = (ub >> q) & 1u;  // does any candidate still use slot q of codebook b?
        const unsigned cbo = b * K2 + s.old[b];
        float E[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) E[c] = 0.0f;
        // the tables of two codebooks a are gathered together (16 loads in flight per lane) and folded in a order;
        // the second table lives in the lists / es union, which is idle until E_b is written below
        float *const tabs[2] = {&s.tab[0], &s.es[0][0]};
#pragma unroll 1
        for (int ap = 0; ap < 4; ap += 2) {
            float gv[2][8], w[2];
            unsigned msk[2];
#pragma unroll
            for (int aa = 0; aa < 2; ++aa) {
                const int a = ap + aa;
                msk[aa] = colu ? (s.used[a] >> pb) & 0xffu : 0u;  // rows of my half that are still in use
                const uint4 *rp = reinterpret_cast<const uint4 *>(&s.rowk[a][pb]);
                const uint4 r0 = rp[0], r1 = rp[1];
                const unsigned ra[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    gv[aa][t] = 0.0f;
                    if ((msk[aa] >> t) & 1u) gv[aa][t] = gat(G, ra[t] + cq);
                }
                w[aa] = gat(G, s.rowoff[a] + cbo);
            }
#pragma unroll
            for (int aa = 0; aa < 2; ++aa) {
                const int a = ap + aa;
                const float4 *up = reinterpret_cast<const float4 *>(&s.uv[a][b][pb]);
                const float4 u0 = up[0], u1 = up[1];
                const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                const float v = s.uv[b][a][q];
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if ((msk[aa] >> t) & 1u)  // tbase + t*TSTR = trow_off(pb + t) + compacted column
                        tabs[aa][tbase + t * TSTR] = ((gv[aa][t] - u[t]) - v) + w[aa];
            }
            __syncwarp();
#pragma unroll
            for (int aa = 0; aa < 2; ++aa) {
                const float4 *mine =
                    reinterpret_cast<const float4 *>(&tabs[aa][trow_off((ti >> (4 * (ap + aa))) & 15)]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nq4) {
                        const float4 r = mine[c];
                        fadd2(E[4 * c + 0], E[4 * c + 1], r.x, r.y);
                        fadd2(E[4 * c + 2], E[4 * c + 3], r.z, r.w);
                    }
                }
            }
            __syncwarp();
        }
        float4 *erow = reinterpret_cast<float4 *>(&s.es[lane][0]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < nq4) erow[c] = make_float4(E[4 * c], E[4 * c + 1], E[4 * c + 2], E[4 * c + 3]);
        __syncwarp();
        const int jq = __popc(ub & ((1u << ((tj >> (4 * lb)) & 15)) - 1u));  // compacted column of my slot of codebook b
#pragma unroll
        for (int i = 0; i < 32; ++i) dot[i] = dot[i] + s.es[i][jq];
        __syncwarp();
    }
    s.lists[8][lane] = list_sentinel();  // es overwrote the sentinels
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
        s.old[lane] = s.kk[lane][(tt >> (4 * (lane & 3))) & 15];
    }
    __syncwarp();
}

template <int N>
__device__ __forceinline__ void refine_pass2(WarpMem2<N> &s, const float *__restrict__ Pb, const float *__restrict__ G, int lane) {
    if (lane < N) s.rowoff[lane] = (unsigned)(lane * K2 + s.old[lane]) * (unsigned)(N * K2);
    __syncwarp();
    level1<N, WarpMem2<N>, (N == 8)>(s, Pb, G, lane);
    gather_uv<N>(s, G, lane);
    if constexpr (N == 2) {
        merge1<N, true>(s, G, 0, lane);
    } else {
#if MCQ_MERGE_PAIR
#pragma unroll 1
        for (int g = 0; g < N / 2; g += 2) merge1_pair<N>(s, G, g, lane);
#else
#pragma unroll 1
        for (int g = 0; g < N / 2; ++g) merge1<N, false>(s, G, g, lane);
#endif
        if constexpr (N == 4) {
            merge2<N, true>(s, G, 0, lane);
        } else {
#if MCQ_MERGE_PAIR
            merge2_pair<N>(s, G, 0, lane);
#else
#pragma unroll 1
            for (int g = 0; g < N / 4; ++g) merge2<N, false>(s, G, g, lane);
#endif
            merge4_final<N>(s, G, lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 16 codebooks (BASELINE config 4).  Same building blocks; the differences to the N <= 8 path:
//   * the u/v terms do not fit shared memory for all 240 codebook pairs (16 KB), so each merge gathers the ones of ITS
//     pairs into a 2 KB buffer first (ul / vl);
//   * one more level: the quad merges (32 x 32) keep 32 of 1024 candidates (four 256-candidate selections + one over
//     their 128 survivors), and the final merge joins two octets: 64 tables, eight per right-hand codebook.
struct alignas(16) WarpMem16 {
    union {
        float2 lists[9][32];
        float es[32][TSTR];
    };
    union {
        float tab[TAB_FLOATS];  // T_ab of the wide merges
        float2 cand[4][32];     // survivors of the four 256-candidate selections of a quad merge
    };
    float kd1[16][16];
    int kk[16][16];
    unsigned rowk[16][16];
    float ul[16][16];        // u terms of the current merge: ul[c][p] = G[(b,old_b), (a, kk_a[p])], c = pair index
    float vl[16][16];        // v terms: vl[c][q] = G[(a,old_a), (b, kk_b[q])]
    float2 sel[32];
    float kd2[8][16];
    unsigned kt2[8][16];     // 2 slots (8 bits)
    float kd3[4][32];
    unsigned kt3[4][32];     // 4 slots (16 bits)
    float kd4[2][32];
    unsigned kt4[2][32];     // 8 slots (32 bits)
    int old[16];
    unsigned rowoff[16];
    unsigned used[16];
};

// second level-1 chain of the 16-codebook kernel: lists in kd2 .. kt4 (exactly 10 rows of 32 float2, all written only by
// the merges), result in the u-term buffer
__device__ __forceinline__ float2 (*pair_lists2(WarpMem16 &s))[32] {
    static_assert(offsetof(WarpMem16, kt4) + sizeof(WarpMem16::kt4) - offsetof(WarpMem16, kd2) == 10 * 32 * sizeof(float2),
                  "pair memory");
    return reinterpret_cast<float2(*)[32]>(&s.kd2[0][0]);
}
__device__ __forceinline__ float2 *pair_sel2(WarpMem16 &s) { return reinterpret_cast<float2 *>(&s.ul[0][0]); }

// ul/vl of the NA x NB codebook pairs (a_base + la, b_base + lb), pair index c = la * NB + lb.
// Lanes 0..15 fetch u (slot p = lane), lanes 16..31 fetch v (slot q = lane - 16).
template <int NA, int NB>
__device__ __forceinline__ void gather_uv_local(WarpMem16 &s, const float *__restrict__ G, int a_base, int b_base, int lane) {
    const int sl = lane & 15;
    const bool isv = lane >= 16;
    constexpr int NC = NA * NB;
    constexpr int BATCH = NC < 4 ? NC : 4;
#pragma unroll 1
    for (int c0 = 0; c0 < NC; c0 += BATCH) {
        float val[BATCH];
#pragma unroll
        for (int cc = 0; cc < BATCH; ++cc) {
            const int c = c0 + cc;
            const int a = a_base + c / NB, b = b_base + c % NB;
            // u: row of (b, old_b), column (a, kk_a[p]);  v: row of (a, old_a), column (b, kk_b[q])
            const unsigned idx = isv ? s.rowoff[a] + b * K2 + s.kk[b][sl] : s.rowoff[b] + a * K2 + s.kk[a][sl];
            val[cc] = gat(G, idx);
        }
#pragma unroll
        for (int cc = 0; cc < BATCH; ++cc) {
            if (isv)
                s.vl[c0 + cc][sl] = val[cc];
            else
                s.ul[c0 + cc][sl] = val[cc];
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void merge1_16(WarpMem16 &s, const float *__restrict__ G, int g, int lane) {
    const int e = 2 * g, o = e + 1;
    gather_uv_local<1, 1>(s, G, e, o, lane);
    const int j = lane & 15, ib = (lane >> 4) * 8;
    const unsigned ko = o * K2 + s.kk[o][j];
    const float v = s.vl[0][j];
    const float kdo = s.kd1[o][j];
    const float w = gat(G, s.rowoff[e] + o * K2 + s.old[o]);
    float key[8];
    int flat[8];
    float gv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) gv[t] = gat(G, s.rowk[e][ib + t] + ko);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float d = ((gv[t] - s.ul[0][ib + t]) - v) + w;
        key[t] = fmaf(2.0f, d, s.kd1[e][ib + t] + kdo);
        flat[t] = (ib + t) * 16 + j;
    }
    select_sorted<WarpMem16, 16>(s, key, flat, lane);
    if (lane < 16) {
        const float2 r = s.sel[lane];
        const int fl = __float_as_int(r.y);
        s.kd2[g][lane] = r.x;
        s.kt2[g][lane] = (unsigned)(fl >> 4) | ((unsigned)(fl & 15) << 4);
    }
    __syncwarp();
}

__device__ __forceinline__ void merge2_16(WarpMem16 &s, const float *__restrict__ G, int g, int lane) {
    const int e = 2 * g, o = e + 1;
    const int a0 = 4 * g, a1 = a0 + 1, b0 = a0 + 2, b1 = a0 + 3;
    gather_uv_local<2, 2>(s, G, a0, b0, lane);  // pair index c = la * 2 + lb
    const int j = lane & 15, ib = (lane >> 4) * 8;
    const unsigned tj = s.kt2[o][j];
    const int q0 = tj & 15, q1 = tj >> 4;
    const unsigned c0 = b0 * K2 + s.kk[b0][q0], c1 = b1 * K2 + s.kk[b1][q1];
    const float v00 = s.vl[0][q0], v10 = s.vl[2][q0], v01 = s.vl[1][q1], v11 = s.vl[3][q1];
    const float kdo = s.kd2[o][j];
    const unsigned cb0 = b0 * K2 + s.old[b0], cb1 = b1 * K2 + s.old[b1];
    const float w00 = gat(G, s.rowoff[a0] + cb0), w10 = gat(G, s.rowoff[a1] + cb0);
    const float w01 = gat(G, s.rowoff[a0] + cb1), w11 = gat(G, s.rowoff[a1] + cb1);
    float g00[8], g10[8], g01[8], g11[8];
    unsigned tis[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        tis[t] = s.kt2[e][ib + t];
        const unsigned rp0 = s.rowk[a0][tis[t] & 15], rp1 = s.rowk[a1][tis[t] >> 4];
        g00[t] = gat(G, rp0 + c0);
        g10[t] = gat(G, rp1 + c0);
        g01[t] = gat(G, rp0 + c1);
        g11[t] = gat(G, rp1 + c1);
    }
    float key[8];
    int flat[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int ia0 = tis[t] & 15, ia1 = tis[t] >> 4;
        const float d00 = ((g00[t] - s.ul[0][ia0]) - v00) + w00;
        const float d10 = ((g10[t] - s.ul[2][ia1]) - v10) + w10;
        const float d01 = ((g01[t] - s.ul[1][ia0]) - v01) + w01;
        const float d11 = ((g11[t] - s.ul[3][ia1]) - v11) + w11;
        const float wb0 = d00 + d10, wb1 = d01 + d11;
        const float dot = wb0 + wb1;
        key[t] = fmaf(2.0f, dot, s.kd2[e][ib + t] + kdo);
        flat[t] = (ib + t) * 16 + j;
    }
    select_sorted<WarpMem16, 32>(s, key, flat, lane);
    {
        const float2 r = s.sel[lane];
        const int fl = __float_as_int(r.y);
        s.kd3[g][lane] = r.x;
        s.kt3[g][lane] = s.kt2[e][fl >> 4] | (s.kt2[o][fl & 15] << 8);
    }
    __syncwarp();
}

// dot(i, j) of the 32 x 32 joint candidates of two groups of NA codebooks each (a_base.., b_base..), lane = column j,
// dot[i] over the rows i.  ti / tj: this lane's slot tuple as row i = lane / as column j = lane (4 bits per codebook).
template <int NA>
__device__ __forceinline__ void wide_dots(WarpMem16 &s, const float *__restrict__ G, int a_base, int b_base, unsigned ti, unsigned tj,
                                          int lane, float (&dot)[32]) {
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        const unsigned ua = __reduce_or_sync(FULL, 1u << ((ti >> (4 * c)) & 15));
        const unsigned ub = __reduce_or_sync(FULL, 1u << ((tj >> (4 * c)) & 15));
        if (lane == 0) {
            s.used[c] = ua;
            s.used[8 + c] = ub;
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) dot[i] = 0.0f;
    if constexpr (NA == 4) gather_uv_local<4, 4>(s, G, a_base, b_base, lane);  // all 16 pairs fit: c = la * 4 + lb
    // as in merge4_final: table columns compacted to the slots of codebook b that are still in use, unused rows skipped
    const int q = lane & 15, pb = (lane >> 4) * 8;
    const unsigned below_q = (1u << q) - 1u;
#pragma unroll 1
    for (int lb = 0; lb < NA; ++lb) {
        const int b = b_base + lb;
        if constexpr (NA == 8) gather_uv_local<8, 1>(s, G, a_base, b, lane);   // the 8 pairs of this b: c = la
        const unsigned ub = s.used[8 + lb];
        const int nq4 = (__popc(ub) + 3) >> 2;
        const int tbase = trow_off(pb) + __popc(ub & below_q);
        const unsigned cq = b * K2 + s.kk[b][q];
        const bool colu = (ub >> q) & 1u;
        const unsigned cbo = b * K2 + s.old[b];
        float E[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) E[c] = 0.0f;
#pragma unroll 1
        for (int la = 0; la < NA; ++la) {
            const int a = a_base + la;
            const int c = (NA == 4) ? la * 4 + lb : la;
            const unsigned msk = colu ? (s.used[la] >> pb) & 0xffu : 0u;
            unsigned ra[8];
            float u[8];
            {
                const uint4 *rp = reinterpret_cast<const uint4 *>(&s.rowk[a][pb]);
                const float4 *up = reinterpret_cast<const float4 *>(&s.ul[c][pb]);
                const uint4 r0 = rp[0], r1 = rp[1];
                const float4 u0 = up[0], u1 = up[1];
                ra[0] = r0.x; ra[1] = r0.y; ra[2] = r0.z; ra[3] = r0.w; ra[4] = r1.x; ra[5] = r1.y; ra[6] = r1.z; ra[7] = r1.w;
                u[0] = u0.x; u[1] = u0.y; u[2] = u0.z; u[3] = u0.w; u[4] = u1.x; u[5] = u1.y; u[6] = u1.z; u[7] = u1.w;
            }
            float gv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                gv[t] = 0.0f;
                if ((msk >> t) & 1u) gv[t] = gat(G, ra[t] + cq);
            }
            const float v = s.vl[c][q];
            const float w = gat(G, s.rowoff[a] + cbo);
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if ((msk >> t) & 1u) s.tab[tbase + t * TSTR] = ((gv[t] - u[t]) - v) + w;  // trow_off(pb + t) + compacted column
            __syncwarp();
            const float4 *mine = reinterpret_cast<const float4 *>(&s.tab[trow_off((ti >> (4 * la)) & 15)]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                if (cc < nq4) {
                    const float4 r = mine[cc];
                    fadd2(E[4 * cc + 0], E[4 * cc + 1], r.x, r.y);
                    fadd2(E[4 * cc + 2], E[4 * cc + 3], r.z, r.w);
                }
            }
            __syncwarp();
        }
        float4 *erow = reinterpret_cast<float4 *>(&s.es[lane][0]);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (cc < nq4) erow[cc] = make_float4(E[4 * cc], E[4 * cc + 1], E[4 * cc + 2], E[4 * cc + 3]);
        __syncwarp();
        const int jq = __popc(ub & ((1u << ((tj >> (4 * lb)) & 15)) - 1u));
#pragma unroll
        for (int i = 0; i < 32; ++i) dot[i] = dot[i] + s.es[i][jq];
        __syncwarp();
    }
    s.lists[8][lane] = list_sentinel();  // es overwrote the sentinels
    __syncwarp();
}

// Quad merge of N = 16 (not final): groups e = 2g (codebooks 8g..8g+3) and o = 2g+1 (8g+4..8g+7); keeps the 32 best
// of the 1024 joint candidates, ascending by (score, flat = i*32 + j).
__device__ __forceinline__ void merge4_16(WarpMem16 &s, const float *__restrict__ G, int g, int lane) {
    const int e = 2 * g, o = e + 1;
    const unsigned ti = s.kt3[e][lane], tj = s.kt3[o][lane];
    float dot[32];
    wide_dots<4>(s, G, 8 * g, 8 * g + 4, ti, tj, lane, dot);
    const float kdo = s.kd3[o][lane];
#pragma unroll
    for (int i = 0; i < 32; ++i) dot[i] = fmaf(2.0f, dot[i], s.kd3[e][i] + kdo);  // dot[] now holds the scores
    // four selections over the row blocks i in [8r, 8r+8), then (r == 4) one over their 4 x 32 survivors.  One copy of
    // the selection code: the loop is not unrolled, the block's scores are picked out of the register array by a switch.
#pragma unroll 1
    for (int r = 0; r < 5; ++r) {
        float key[8];
        int flat[8];
        if (r < 4) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                key[t] = r == 0 ? dot[t] : (r == 1 ? dot[8 + t] : (r == 2 ? dot[16 + t] : dot[24 + t]));
                flat[t] = (8 * r + t) * 32 + lane;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 c = s.cand[t][lane];  // block t holds flats in [256 t, 256 t + 256): ascending in t
                key[t] = c.x;
                flat[t] = __float_as_int(c.y);
            }
#pragma unroll
            for (int t = 4; t < 8; ++t) {
                key[t] = __int_as_float(0x7f800000);
                flat[t] = 0x7ffffff0 + t;
            }
        }
        select_sorted<WarpMem16, 32>(s, key, flat, lane);
        if (r < 4) {
            s.cand[r][lane] = s.sel[lane];
            __syncwarp();
        }
    }
    {
        const float2 r = s.sel[lane];
        const int fl = __float_as_int(r.y) & 1023;
        s.kd4[g][lane] = r.x;
        s.kt4[g][lane] = s.kt3[e][fl >> 5] | (s.kt3[o][fl & 31] << 16);
    }
    __syncwarp();
}

// Final merge of N = 16: the two octets, 32 x 32 candidates, best one wins.
__device__ __forceinline__ void merge8_final_16(WarpMem16 &s, const float *__restrict__ G, int lane) {
    const unsigned ti = s.kt4[0][lane], tj = s.kt4[1][lane];
    float dot[32];
    wide_dots<8>(s, G, 0, 8, ti, tj, lane, dot);
    const float kdo = s.kd4[1][lane];
    float best = fmaf(2.0f, dot[0], s.kd4[0][0] + kdo);
    int bi = 0;
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        const float key = fmaf(2.0f, dot[i], s.kd4[0][i] + kdo);
        if (key < best) {
            best = key;
            bi = i;
        }
    }
    const float m = credux_min(best);
    const unsigned c = (best == m) ? (unsigned)(bi * 32 + lane) : 0x7fffffffu;
    unsigned flat = __reduce_min_sync(FULL, c);
    if (flat == 0x7fffffffu) flat = 0;  // only reachable with NaN scores
    const unsigned te = s.kt4[0][flat >> 5], to = s.kt4[1][flat & 31];
    if (lane < 16) {
        const unsigned tt = lane < 8 ? te : to;
        s.old[lane] = s.kk[lane][(tt >> (4 * (lane & 7))) & 15];
    }
    __syncwarp();
}

__device__ __forceinline__ void refine_pass16(WarpMem16 &s, const float *__restrict__ Pb, const float *__restrict__ G, int lane) {
    if (lane < 16) s.rowoff[lane] = (unsigned)(lane * K2 + s.old[lane]) * (unsigned)(16 * K2);
    __syncwarp();
    level1<16, WarpMem16, MCQ_S16_PAIR != 0>(s, Pb, G, lane);
#pragma unroll 1
    for (int g = 0; g < 8; ++g) merge1_16(s, G, g, lane);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) merge2_16(s, G, g, lane);
#pragma unroll 1
    for (int g = 0; g < 2; ++g) merge4_16(s, G, g, lane);
    merge8_final_16(s, G, lane);
}

constexpr int WPC16 = MCQ_S16_WPC;

__global__ void __launch_bounds__(WPC16 * 32, MCQ_S16_MINB)
    search2_kernel16(const float *__restrict__ P, const float *__restrict__ G, int64_t B, int iters, const int32_t *__restrict__ idx_in,
                     int32_t *__restrict__ idx_out, unsigned *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpMem16 &s = reinterpret_cast<WarpMem16 *>(smem_raw)[warp];
    s.lists[8][lane] = list_sentinel();
    s.sel[lane] = make_float2(0.0f, __int_as_float(0));
    __syncwarp();
    const int64_t nwarps = (int64_t)gridDim.x * WPC16;
    unsigned npass = 0, nframes = 0;
    for (int64_t b = (int64_t)blockIdx.x * WPC16 + warp; b < B;) {
        if (lane < 16) s.old[lane] = idx_in[(size_t)b * 16 + lane];
        __syncwarp();
        const float *Pb = P + (size_t)b * (16 * K2);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int prev = (lane < 16) ? s.old[lane] : 0;
            refine_pass16(s, Pb, G, lane);
            const int now = (lane < 16) ? s.old[lane] : 0;
            ++npass;
            if (__all_sync(FULL, prev == now)) break;  // fixed point: the remaining passes are no-ops
        }
        ++nframes;
        if (lane < 16) idx_out[(size_t)b * 16 + lane] = s.old[lane];
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

int launch16(const float *P, const float *Gp, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
             cudaStream_t st, unsigned *work_counter) {
    const float *G = Gp;
    const size_t smem = sizeof(WarpMem16) * WPC16;
    auto kern = search2_kernel16;
    MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MCQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPC16 * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 148;
    MCQ_CUDA(cudaGetDevice(&dev));
    MCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t need = (B + WPC16 - 1) / WPC16;
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, WPC16 * 32, smem, st>>>(P, G, B, iters, idx_in, idx_out, work_counter);
    MCQ_LAUNCH_CHECK("search2_kernel16");
    return MCQ_OK;
}


template <int N>
struct Launch2 {
    static constexpr int WPC = (N == 8) ? MCQ_S2_WPC : MCQ_S2_WPC4;  // warps per CTA
    static constexpr int MINB = (N == 8) ? MCQ_S2_MINB : MCQ_S2_MINB4;
};

template <int N>
__global__ void __launch_bounds__(Launch2<N>::WPC * 32, Launch2<N>::MINB)
    search2_kernel(const float *__restrict__ P, const float *__restrict__ G, int64_t B, int iters, const int32_t *__restrict__ idx_in,
                   int32_t *__restrict__ idx_out, unsigned *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int wpc = Launch2<N>::WPC;
    WarpMem2<N> &s = reinterpret_cast<WarpMem2<N> *>(smem_raw)[warp];
    s.lists[8][lane] = list_sentinel();
    s.sel[lane] = make_float2(0.0f, __int_as_float(0));
    __syncwarp();
    // the first frame of a warp is its global warp index; further frames come from the work counter when there is
    // one (frames take 2..iters passes, so static striding would leave a tail), else by striding over the batch
    const int64_t nwarps = (int64_t)gridDim.x * wpc;
    unsigned npass = 0, nframes = 0;
    for (int64_t b = (int64_t)blockIdx.x * wpc + warp; b < B;) {
        if (lane < N) s.old[lane] = idx_in[(size_t)b * N + lane];
        __syncwarp();
        const float *Pb = P + (size_t)b * (N * K2);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const int prev = (lane < N) ? s.old[lane] : 0;
            refine_pass2<N>(s, Pb, G, lane);
            const int now = (lane < N) ? s.old[lane] : 0;
            ++npass;
            // a pass that returns its input is a fixed point of a deterministic map: the remaining passes are no-ops
            if (__all_sync(FULL, prev == now)) break;
        }
        ++nframes;
        if (lane < N) idx_out[(size_t)b * N + lane] = s.old[lane];
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

template <int N>
int launch2t(const float *P, const float *G, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
             cudaStream_t st, unsigned *work_counter) {
    constexpr int wpc = Launch2<N>::WPC;
    const size_t smem = sizeof(WarpMem2<N>) * wpc;
    auto kern = search2_kernel<N>;
    MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef MCQ_S2_CARVEOUT
    // percent of the 228 KB the SM may give to shared memory; the rest of the 256 KB is L1
    MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, MCQ_S2_CARVEOUT));
#endif
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
    kern<<<(unsigned)grid, wpc * 32, smem, st>>>(P, G, B, iters, idx_in, idx_out, work_counter);
    MCQ_LAUNCH_CHECK("search2_kernel");
    return MCQ_OK;
}

template <int N>
int launch2(const float *P, const float *Gp, int64_t B, int iters, const int32_t *idx_in, int32_t *idx_out,
            cudaStream_t st, unsigned *work_counter) {
    return launch2t<N>(P, Gp, B, iters, idx_in, idx_out, st, work_counter);
}

}  // namespace

bool search2_supports(int N, int K) { return K == 256 && (N == 2 || N == 4 || N == 8 || N == 16); }

int launch_search2(const float *P, const float *gram, int64_t B, int N, int K, int iters, const int32_t *idx_in,
                   int32_t *idx_out, cudaStream_t st, unsigned *work_counter) {
    if (B <= 0) return MCQ_OK;
    if (K == 256) {
        switch (N) {
            case 2: return launch2<2>(P, gram, B, iters, idx_in, idx_out, st, work_counter);
            case 4: return launch2<4>(P, gram, B, iters, idx_in, idx_out, st, work_counter);
            case 8: return launch2<8>(P, gram, B, iters, idx_in, idx_out, st, work_counter);
            case 16: return launch16(P, gram, B, iters, idx_in, idx_out, st, work_counter);
            default: break;
        }
    }
    set_error("search2: (K=%d, N=%d) is not supported", K, N);
    return MCQ_EUNSUPPORTED;
}

}  // namespace mcq
