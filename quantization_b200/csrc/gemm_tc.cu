// gemm_tc.cu -- fp32-faithful GEMM on the 5th-generation tensor cores:  C (Mp, NK) fp32 = A (Mp, D) . B (NK, D)^T
// with both operands given as row-scaled two-way fp16 splits a 2^e = a0 + a1 (11 significand bits each, prepare.cu).
// The three leading cross products  a0b0 + a0b1 + a1b0  are issued as tcgen05.mma kind::f16 (fp16 in, fp32 accumulate
// in TMEM) -- products of fp16 pieces are exact in fp32; the dropped a1b1 term is <= 2^-22 |a b| per element, under the
// rounding noise of the fp32 accumulation itself -- and the epilogue multiplies by the rows' exact 2^-e factors.
// (Round 1 started with a three-way bf16 split and six products: same accuracy class, twice the MMA work.)
// A single-pass bf16 or tf32 GEMM flips 0.1-7 % of the codes, which is why the split is
// there.
//
// This replaces the reference's `to_logits(x)` addmm (quantization.py:279) and the per-pass scoring matmul
// (quantization.py:413-416; here done once per frame as P = x Cs^T, see search.cu).
//
// Structure: persistent CTAs (one per SM), 128 x BN output tile, K blocks of 64 fp16 (one 128-byte swizzle atom).
//   warp 0      TMA producer: per K block 2 A-plane boxes + 2 B-plane boxes -> 128B-swizzled smem, 3-stage ring
//   warp 1      TMEM allocator + single-thread MMA issuer (12 tcgen05.mma per K block), tcgen05.commit -> mbarriers
//   warps 2..5  epilogue: tcgen05.ld 32x32b.x32 -> registers -> 16-byte global stores, overlapped with the next
//               tile's MMAs through a double-buffered TMEM accumulator
#include <cuda.h>

#include "common.cuh"

namespace mcq {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements = 128 bytes = one swizzle atom row
constexpr int STAGES = 3;
constexpr int PLANES = 2;
constexpr int NUM_THREADS = 192;

template <int BN>
struct TcCfg {
    static constexpr uint32_t A_PLANE = BM * 128;  // bytes of one plane of the A stage
    static constexpr uint32_t B_PLANE = BN * 128;
    static constexpr uint32_t STAGE = PLANES * A_PLANE + PLANES * B_PLANE;
    // epilogue staging: per epilogue warp 32 rows x (32 + 4) floats, so that the global stores are full 128-byte rows
    static constexpr uint32_t EPI_STAGE = 4 * 32 * 36 * 4;
    static constexpr uint32_t SMEM = STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_STAGE;
    // Two fp32 accumulators per tile -- `main` takes only the a0*b0 products, `corr` the two small cross terms --
    // double buffered.  The tensor core truncates when it adds into the accumulator, so every MMA costs up to an
    // ulp of |acc|; keeping the MMAs that carry < 2^-11 of the magnitude out of the main accumulator keeps that
    // bias to one MMA per K step.
    static constexpr uint32_t TMEM_COLS = 4 * BN;
    // instruction descriptor: D=f32 (bit 4), A=B=f16 (format fields 0), both K-major, N, M
    // (cute::UMMA::InstrDescriptor bit layout)
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// K-major, 128-byte swizzle, rows of exactly 128 bytes: SBO = 8 rows * 128 B, LBO unused (1), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// the three products kept (plane of A, plane of B): the two corrections, then the main product
constexpr int NPROD = 3;
__device__ constexpr int kProdA[NPROD] = {1, 0, 0};
__device__ constexpr int kProdB[NPROD] = {0, 1, 0};

// ARGMAX: instead of storing the tile, the epilogue adds `bias` and keeps, per row, the first maximum of the tile's BN
// columns: part_val / part_idx [row][n_tile] (the classifier arg-max of quantization.py:297-301 fused into the GEMM;
// argmax_merge_kernel folds the K / BN tiles of a codebook).
template <int BN, bool ARGMAX>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fp16x2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   float *__restrict__ C, int m_tiles, int n_tiles, int k_blocks, int ldc, int a_plane_rows,
                   int b_plane_rows, const float *__restrict__ a_scale, const float *__restrict__ b_scale,
                   const float *__restrict__ bias, float *__restrict__ part_val, int *__restrict__ part_idx,
                   int k_splits, size_t c_split_stride, int m_valid, int accumulate) {
    // m_valid: rows >= m_valid of the (padded) product are not stored; accumulate != 0: C += product.
    // k_splits > 1 (split-K, for products with few output tiles and a long reduction): tile t covers k-blocks
    // [ks * k_blocks, (ks + 1) * k_blocks) with ks = t / (m_tiles * n_tiles) and stores into C + ks * c_split_stride;
    // the caller sums the k_splits partial results in a fixed order.
    using Cfg = TcCfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * Cfg::STAGE;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base address word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mn_tiles = m_tiles * n_tiles;
    const int num_tiles = mn_tiles * k_splits;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"(Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mn = tile % mn_tiles, kb0 = (tile / mn_tiles) * k_blocks;
                const int m0 = (mn / n_tiles) * BM, n0 = (mn % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t st = base + s * Cfg::STAGE;
                    mbar_expect_tx(full_bar(s), Cfg::STAGE);
#pragma unroll
                    for (int p = 0; p < PLANES; ++p) {
                        tma_load_2d(st + p * Cfg::A_PLANE, &tma_a, (kb0 + kb) * BK, p * a_plane_rows + m0, full_bar(s));
                        tma_load_2d(st + PLANES * Cfg::A_PLANE + p * Cfg::B_PLANE, &tma_b, (kb0 + kb) * BK,
                                    p * b_plane_rows + n0, full_bar(s));
                    }
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), aph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_main = tmem_base + (uint32_t)(acc * 2 * BN);
                const uint32_t tmem_corr = tmem_main + (uint32_t)BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * Cfg::STAGE;
#pragma unroll
                    for (int pr = 0; pr < NPROD; ++pr) {
                        const uint32_t a_addr = st + kProdA[pr] * Cfg::A_PLANE;
                        const uint32_t b_addr = st + PLANES * Cfg::A_PLANE + kProdB[pr] * Cfg::B_PLANE;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if (pr == NPROD - 1)
                                umma_f16(tmem_main, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32),
                                          Cfg::IDESC, (kb | k) != 0 ? 1u : 0u);
                            else
                                umma_f16(tmem_corr, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32),
                                          Cfg::IDESC, (kb | pr | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
                    if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    aph ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mn = tile % mn_tiles;
            const int m0 = (mn / n_tiles) * BM, n0 = (mn % n_tiles) * BN;
            float *Ct = C + (size_t)(tile / mn_tiles) * c_split_stride;
            mbar_wait(tfull_bar(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float sa = __ldg(a_scale + m0 + q * 32 + lane);  // 2^-e of my row of A
            float best = 0.0f;
            int bk = -1;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32], rc[32];
                const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + c * 32);
                tmem_ld32(lane_base, r);
                tmem_ld32(lane_base + (uint32_t)BN, rc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float *sbp = b_scale + n0 + c * 32;  // 2^-e of the rows of B = my columns
                if constexpr (ARGMAX) {
                    const float *bs = bias + n0 + c * 32;
#pragma unroll
                    for (int v = 0; v < 32; ++v) {
                        const float val =
                            (__uint_as_float(r[v]) + __uint_as_float(rc[v])) * (sa * __ldg(sbp + v)) + __ldg(bs + v);
                        if (bk < 0 || val > best) {  // first maximum wins, like torch.argmax
                            best = val;
                            bk = n0 + c * 32 + v;
                        }
                    }
                } else {
                    // scale, transpose through shared memory, store full rows: lane l writes 16 bytes of row
                    // (l / 8 + 4 k), so one instruction covers four complete 128-byte row segments
                    float *stg = reinterpret_cast<float *>(smem_raw + (bar_base + 256u - smem_u32(smem_raw))) + q * (32 * 36);
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float4 sb = __ldg(reinterpret_cast<const float4 *>(sbp) + v);
                        float4 o;
                        o.x = (__uint_as_float(r[4 * v + 0]) + __uint_as_float(rc[4 * v + 0])) * (sa * sb.x);
                        o.y = (__uint_as_float(r[4 * v + 1]) + __uint_as_float(rc[4 * v + 1])) * (sa * sb.y);
                        o.z = (__uint_as_float(r[4 * v + 2]) + __uint_as_float(rc[4 * v + 2])) * (sa * sb.z);
                        o.w = (__uint_as_float(r[4 * v + 3]) + __uint_as_float(rc[4 * v + 3])) * (sa * sb.w);
                        *reinterpret_cast<float4 *>(stg + lane * 36 + 4 * v) = o;
                    }
                    __syncwarp();
                    float *cbase = Ct + (size_t)(m0 + q * 32) * ldc + n0 + c * 32 + (lane & 7) * 4;
                    const int rows_left = m_valid - (m0 + q * 32);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int row = k * 4 + (lane >> 3);
                        float4 o = *reinterpret_cast<const float4 *>(stg + row * 36 + (lane & 7) * 4);
                        if (row < rows_left) {
                            float4 *dst = reinterpret_cast<float4 *>(cbase + (size_t)row * ldc);
                            if (accumulate) {
                                const float4 old = *dst;
                                o.x += old.x;
                                o.y += old.y;
                                o.z += old.z;
                                o.w += old.w;
                            }
                            *dst = o;
                        }
                    }
                    __syncwarp();
                }
            }
            if constexpr (ARGMAX) {
                const size_t slot = (size_t)(m0 + q * 32 + lane) * n_tiles + (n0 / BN);
                part_val[slot] = best;
                part_idx[slot] = bk;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                aph ^= 1u;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map(CUtensorMap *m, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return MCQ_ECUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (rows=%llu cols=%llu box=%u)", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols, box_rows);
        return MCQ_ECUDA;
    }
    return MCQ_OK;
}

template <int BN, bool ARGMAX>
int launch_bn(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale, float *C,
              int64_t Mp, int NK, int Dp, cudaStream_t st, const float *bias = nullptr, float *part_val = nullptr,
              int *part_idx = nullptr, int k_splits = 1, int64_t ldc = 0, int64_t m_valid = -1, int accumulate = 0,
              int64_t a_plane_rows = 0) {
    // a_plane_rows: row distance between the two fp16 planes of A (default Mp; larger when only the first Mp rows of
    // a bigger split buffer take part, e.g. the tail chunk of a batch)
    using Cfg = TcCfg<BN>;
    const uint64_t NKp = align_up((size_t)NK, 128);
    if (a_plane_rows <= 0) a_plane_rows = Mp;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map(&ma, a_split, (uint64_t)PLANES * (uint64_t)a_plane_rows, (uint64_t)Dp, BM))) return rc;
    if ((rc = make_map(&mb, b_split, (uint64_t)PLANES * NKp, (uint64_t)Dp, BN))) return rc;
    auto kern = gemm_fp16x2_kernel<BN, ARGMAX>;
    MCQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    int dev = 0, sms = 148;
    MCQ_CUDA(cudaGetDevice(&dev));
    MCQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int m_tiles = (int)(Mp / BM), n_tiles = NK / BN;
    int64_t tiles = (int64_t)m_tiles * n_tiles * k_splits;
    int grid = (int)(tiles < sms ? tiles : sms);
    kern<<<grid, NUM_THREADS, Cfg::SMEM, st>>>(ma, mb, C, m_tiles, n_tiles, Dp / BK / k_splits, ldc > 0 ? (int)ldc : NK,
                                               (int)a_plane_rows, (int)NKp, a_scale, b_scale, bias, part_val, part_idx, k_splits,
                                               (size_t)Mp * (size_t)NK, m_valid >= 0 ? (int)m_valid : (int)Mp, accumulate);
    MCQ_LAUNCH_CHECK("gemm_fp16x2_kernel");
    return MCQ_OK;
}

}  // namespace

int launch_gemm_tc(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale, float *C,
                   int64_t Mp, int NK, int Dp, cudaStream_t st) {
    if (Mp <= 0) return MCQ_OK;
    if (Mp % BM != 0 || Dp % BK != 0 || NK % 64 != 0) {
        set_error("gemm_tc: Mp=%lld Dp=%d NK=%d not tile aligned", (long long)Mp, Dp, NK);
        return MCQ_EINVAL;
    }
    if (NK % 128 == 0) return launch_bn<128, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st);
    return launch_bn<64, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st);
}

// Split-K variant: `k_splits` partial products C[ks] (each Mp x NK, contiguous) over Dp / k_splits columns each.
int launch_gemm_tc_splitk(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                          float *C, int64_t Mp, int NK, int Dp, int k_splits, cudaStream_t st) {
    if (Mp <= 0) return MCQ_OK;
    if (Mp % BM != 0 || k_splits < 1 || Dp % (BK * k_splits) != 0 || NK % 64 != 0) {
        set_error("gemm_tc_splitk: Mp=%lld Dp=%d NK=%d splits=%d not tile aligned", (long long)Mp, Dp, NK, k_splits);
        return MCQ_EINVAL;
    }
    if (NK % 128 == 0)
        return launch_bn<128, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st, nullptr, nullptr, nullptr,
                                     k_splits);
    return launch_bn<64, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st, nullptr, nullptr, nullptr, k_splits);
}

// General form: C (m_valid x NK, row stride ldc) = or += A . B^T from packed operands (Mp = m_valid rounded up to 128).
int launch_gemm_tc_general(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                           float *C, int64_t ldc, int64_t m_valid, int64_t Mp, int NK, int Dp, int accumulate,
                           cudaStream_t st, int64_t a_plane_rows) {
    if (Mp <= 0) return MCQ_OK;
    if (Mp % BM != 0 || Dp % BK != 0 || NK % 64 != 0 || ldc < NK || (ldc & 3) || m_valid > Mp ||
        (a_plane_rows > 0 && a_plane_rows < Mp)) {
        set_error("gemm_tc_general: Mp=%lld Dp=%d NK=%d ldc=%lld not supported", (long long)Mp, Dp, NK, (long long)ldc);
        return MCQ_EINVAL;
    }
    if (NK % 128 == 0)
        return launch_bn<128, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st, nullptr, nullptr, nullptr, 1,
                                     ldc, m_valid, accumulate, a_plane_rows);
    return launch_bn<64, false>(a_split, a_scale, b_split, b_scale, C, Mp, NK, Dp, st, nullptr, nullptr, nullptr, 1, ldc,
                                m_valid, accumulate, a_plane_rows);
}

// idx[b][n] = first maximum over the K / 128 tile maxima of codebook n (ascending tile = ascending column order)
__global__ void argmax_merge_kernel(const float *__restrict__ part_val, const int *__restrict__ part_idx, int64_t B,
                                    int N, int K, int n_tiles, int32_t *__restrict__ idx) {
    const int tpc = K / 128;  // tiles per codebook
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B * N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / N;
        const int n = (int)(i - b * N);
        const size_t base = (size_t)b * n_tiles + (size_t)n * tpc;
        float best = part_val[base];
        int bk = part_idx[base];
        for (int t = 1; t < tpc; ++t) {
            const float v = part_val[base + t];
            if (v > best) {
                best = v;
                bk = part_idx[base + t];
            }
        }
        idx[i] = bk - n * K;
    }
}

bool gemm_tc_argmax_supported(int NK, int K) { return NK % 128 == 0 && K % 128 == 0; }

// logits GEMM with the classifier arg-max fused into its epilogue: idx (B, N) int32.  `scratch` holds
// Mp * (NK / 128) * 8 bytes of per-tile maxima.
int launch_gemm_tc_argmax(const __half *a_split, const float *a_scale, const __half *b_split, const float *b_scale,
                          int64_t Mp, int NK, int Dp, const float *bias, int64_t B, int N, int K, void *scratch,
                          int32_t *idx, cudaStream_t st) {
    if (Mp <= 0) return MCQ_OK;
    if (Mp % BM != 0 || Dp % BK != 0 || !gemm_tc_argmax_supported(NK, K)) {
        set_error("gemm_tc_argmax: Mp=%lld Dp=%d NK=%d K=%d not supported", (long long)Mp, Dp, NK, K);
        return MCQ_EINVAL;
    }
    const int n_tiles = NK / 128;
    float *part_val = (float *)scratch;
    int *part_idx = (int *)(part_val + (size_t)Mp * n_tiles);
    int rc = launch_bn<128, true>(a_split, a_scale, b_split, b_scale, nullptr, Mp, NK, Dp, st, bias, part_val, part_idx);
    if (rc) return rc;
    int64_t blocks = (B * N + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    argmax_merge_kernel<<<(unsigned)blocks, 256, 0, st>>>(part_val, part_idx, B, N, K, n_tiles, idx);
    MCQ_LAUNCH_CHECK("argmax_merge_kernel");
    return MCQ_OK;
}

}  // namespace mcq
