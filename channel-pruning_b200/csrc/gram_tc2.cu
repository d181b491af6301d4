// cp_gram, tensor-core mode (CP_GRAM_3XTF32 in the header), second generation: split-fp16 operands prepared once,
// a persistent TMA -> tcgen05 pipeline with no conversion work on the critical path.
//
//   G = X'X (K x K),  Bxy = X'(Y - b) (K x n)      X: N x K fp32 row-major, N ~ 5e3..1e5, K = c*k*k
//
// The reference does this arithmetic in float64 on the CPU (numpy matmul / LAPACK inside LinearRegression.fit,
// lib/decompose.py:665-666, and the LASSO design products lib/decompose.py:428-457).  tcgen05 has no fp32/fp64
// MMA; the first-generation kernel (gram_tc.cu) used three kind::tf32 products of a hi/lo split made by converter
// warps inside the GEMM -- 46 % tensor-pipe utilisation, bound by the per-k-block hand-shake of those warps
// (profiles/r1c_summary.md).  This version removes them:
//
//   prep   xs = fl32(x - s_col)             s = fp32 column mean (removes the rank-one mean component)
//          v  = xs * 2^e_col                power-of-two column scale (exact): max|v| in [2^9, 2^10)
//          v  = hi + lo                     hi = fp16_rn(v), lo = fp16_rn(v - hi): 22 mantissa bits kept -- the
//                                           same split precision as tf32 (10 explicit bits each), but kind::f16
//                                           runs at twice the tf32 rate and the operands are half as wide
//          written TRANSPOSED (operand row = column of X, reduction index contiguous, zero padded) so that the
//          GEMM reads plain K-major SWIZZLE_128B tiles with TMA; the same pass produces the fp64 column sums
//          and sums of squares of xs (fixed summation order)
//   gemm   P += hi'hi + hi'lo + lo'hi       three kind::f16 MMAs (128 x 256 x 16) per k-step, fp32 accumulation
//          in TMEM, both operands from shared memory; the tensor core truncates when it adds into its fp32
//          accumulator, so an accumulator takes 128 rows (24 additions), then the drain warps add it into fp32
//          registers with round-to-nearest while the MMAs continue on the second accumulator
//   reduce fp64 sum of the row splits, exact rescale by 2^-(e_i + e_j), shift undone exactly
//          (G = P + s T' + T s' + N s s',  T = column sums of xs in fp64), diagonal from the fp64 squares,
//          lower triangle written in the same pass.
//
// GEMM anatomy: persistent grid (one CTA per SM), static round-robin over (output tile, row split) items.
//   warp 8   TMA producer: per 64-row stage four boxes (A hi/lo 64 x 128, B hi/lo 64 x 256; diagonal tiles take
//            the A operand out of the B tile), 2 stages of 96 KB
//   warp 9   TMEM allocator + single-thread MMA issuer; tcgen05.commit frees the stage / publishes the accumulator
//   warps 0-7  drain: thread = accumulator row x 128 columns, tcgen05.ld, fp32 adds, one fp32 partial tile per item
// Bound: L2 -> SM operand traffic (96 KB per 1536 tensor-pipe cycles = 62 B/clk/SM against ~42 B/clk/SM of L2
// slice throughput), then the tensor pipe.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

int cp_gram_fp64_products(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype,
                          int n, int64_t ldy, const float *y_bias, const int32_t *rows, int nrows, double *G,
                          double *Bxy, double *sx, double *sy, double *yy, cudaStream_t stream);
bool cp_gram_tc_eligible(const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype, int n, int64_t ldy,
                         const int32_t *rows, bool wantB);

namespace {

using namespace cptc;

constexpr int TM = 128, TN = 256;  // output tile (rows of A x rows of B), reduction rows per stage
constexpr int A_TILE = TM * 128;            // bytes of one A operand tile (hi or lo): 128 rows x 128 B
constexpr int B_TILE = TN * 128;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;  // 96 KB
constexpr int STAGES = 2;
constexpr int SUB_STAGES = 2;               // stages per accumulator run (128 rows)
constexpr int NDRAIN_WARPS = 8;
constexpr int NTHREADS = 32 * (NDRAIN_WARPS + 2);
constexpr int W_TMA = NDRAIN_WARPS, W_MMA = NDRAIN_WARPS + 1;  // single-thread roles on the highest warp ids
constexpr int T_TMA = 32 * W_TMA;
constexpr int OFF_BAR = STAGES * STAGE_BYTES;
constexpr int NBAR = 2 * STAGES + 4;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;  // + alignment slack
constexpr int RS = 32;                      // row splits of the statistics passes (partials combined in fixed order)
constexpr int PT = 64;                      // prep tile: 64 rows x 64 columns

struct Tc2Params {
    float *partial;      // [nsplit][ntiles][128][256]
    int64_t Np;          // padded row count (multiple of 64) = inner extent of the operand matrix
    int rows_per_split;  // multiple of 128
    int nsplit;
    int tiles_sym;       // number of 128 x 256 tiles covering the upper triangle of G (0 when G is not requested)
    int tk;              // ceil(K / 128): A tiles
    int tjx;             // ceil(K / 256): B tiles inside X
    int tjy0;            // first B tile of the Y region (= Kp / 256)
    int tnb;             // B tiles of the Y region
    int ntiles;
    int mtot;            // operand rows of one half (hi); the lo half starts at row mtot
};

// item -> (tile, split) -> (ti, tJ); the tiles of the upper triangle first (row ti: tJ = ti/2 .. tjx-1), then X'Y
struct Item {
    int tile, split, ti, tj, nst;
    int64_t r_begin;
    bool diag;  // the A rows are part of the B tile (ti / 2 == tJ inside X)
};
template <bool PAIR>
__device__ __forceinline__ Item decode_item(const Tc2Params &P, int w) {
    Item it;
    it.split = w / P.ntiles;
    it.tile = w - it.split * P.ntiles;
    int l = it.tile;
    if (l < P.tiles_sym) {
        int ti = 0;
        if (PAIR) {  // 256 x 256 tiles: row ti holds tJ = ti .. tjx-1
            while (l >= P.tjx - ti) { l -= P.tjx - ti; ++ti; }
            it.tj = ti + l;
        } else {     // 128 x 256 tiles: row ti holds tJ = ti/2 .. tjx-1
            while (l >= P.tjx - (ti >> 1)) { l -= P.tjx - (ti >> 1); ++ti; }
            it.tj = (ti >> 1) + l;
        }
        it.ti = ti;
        it.diag = (l == 0);
    } else {
        l -= P.tiles_sym;
        it.ti = l / P.tnb;
        it.tj = P.tjy0 + (l - it.ti * P.tnb);
        it.diag = false;
    }
    it.r_begin = (int64_t)it.split * P.rows_per_split;
    int64_t r_end = it.r_begin + P.rows_per_split;
    if (r_end > P.Np) r_end = P.Np;
    it.nst = (int)((r_end - it.r_begin) / KS);
    return it;
}

// ------------------------------------------------------------------ the GEMM
__global__ void __launch_bounds__(NTHREADS, 1)
gram_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Tc2Params P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = P.ntiles * P.nsplit;

    auto bar = [&](int i) { return sbase + OFF_BAR + 8 * i; };
    constexpr int FULL = 0, EMPTY = STAGES, ACC_FULL = 2 * STAGES, ACC_EMPTY = 2 * STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);

    if (threadIdx.x == T_TMA) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(FULL + s), 1);
            mbar_init(bar(EMPTY + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar(ACC_FULL + a), 1);
            mbar_init(bar(ACC_EMPTY + a), NDRAIN_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {  // all 512 TMEM columns: two fp32 accumulators of 256 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t g = 0;  // stages issued so far
            for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
                const Item it = decode_item<false>(P, w);
                const int rowA = it.ti * TM, rowB = it.tj * TN;
                for (int st = 0; st < it.nst; ++st, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    mbar_wait(bar(EMPTY + s), ph ^ 1);
                    const uint32_t dst = sbase + s * STAGE_BYTES;
                    const int r0 = (int)(it.r_begin + (int64_t)st * KS);
                    mbar_arrive_expect_tx(bar(FULL + s), it.diag ? 2 * B_TILE : STAGE_BYTES);
                    if (!it.diag) {
                        tma_load_2d(dst, &mapA, bar(FULL + s), r0, rowA);
                        tma_load_2d(dst + A_TILE, &mapA, bar(FULL + s), r0, P.mtot + rowA);
                    }
                    tma_load_2d(dst + 2 * A_TILE, &mapB, bar(FULL + s), r0, rowB);
                    tma_load_2d(dst + 2 * A_TILE + B_TILE, &mapB, bar(FULL + s), r0, P.mtot + rowB);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D fp32 (1 << 4), A and B fp16 (format 0), both K-major, N >> 3 at 17, M >> 4 at 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t g = 0, gc = 0;  // stages / accumulator runs so far
            for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
                const Item it = decode_item<false>(P, w);
                for (int st = 0; st < it.nst; ++st, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    const int kk = st % SUB_STAGES;
                    const uint32_t ab = gc & 1;
                    if (kk == 0) mbar_wait(bar(ACC_EMPTY + ab), ((gc >> 1) & 1) ^ 1);  // accumulator drained
                    mbar_wait(bar(FULL + s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t stage = sbase + s * STAGE_BYTES;
                    const uint32_t b_hi = stage + 2 * A_TILE, b_lo = b_hi + B_TILE;
                    const uint32_t a_hi = it.diag ? b_hi + (uint32_t)(it.ti & 1) * A_TILE : stage;
                    const uint32_t a_lo = it.diag ? b_lo + (uint32_t)(it.ti & 1) * A_TILE : stage + A_TILE;
                    const uint32_t acc = tmem_base + ab * TN;
#pragma unroll
                    for (int ks = 0; ks < KS / 16; ++ks) {
                        const uint32_t off = ks * 32;  // 16 fp16 = 32 bytes along K inside the 128-byte swizzled row
                        const uint32_t first = (kk == 0 && ks == 0) ? 0u : 1u;
                        umma_f16_ss(acc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, first);
                        umma_f16_ss(acc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1u);
                        umma_f16_ss(acc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, 1u);
                    }
                    umma_commit(bar(EMPTY + s));  // stage free once these MMAs have read it
                    if (kk == SUB_STAGES - 1 || st == it.nst - 1) {
                        umma_commit(bar(ACC_FULL + ab));
                        ++gc;
                    }
                }
            }
        }
    } else {
        // ===================== drain warps =====================
        // thread -> accumulator row m (TMEM lane; a warp reaches the lanes of its quadrant, warp % 4) x 128 columns
        const int quad = warp & 3, half = warp >> 2;
        const int m = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
        uint32_t gc = 0;
        for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
            const Item it = decode_item<false>(P, w);
            const int nsub = (it.nst + SUB_STAGES - 1) / SUB_STAGES;
            float accv[128];
#pragma unroll
            for (int e = 0; e < 128; ++e) accv[e] = 0.f;
            for (int c = 0; c < nsub; ++c, ++gc) {
                const uint32_t ab = gc & 1;
                mbar_wait(bar(ACC_FULL + ab), (gc >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int gq = 0; gq < 2; ++gq) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(lane_addr + ab * TN + (uint32_t)(gq * 64), r0);
                    tmem_ld32(lane_addr + ab * TN + (uint32_t)(gq * 64 + 32), r1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        accv[gq * 64 + e] = __fadd_rn(accv[gq * 64 + e], __uint_as_float(r0[e]));
                        accv[gq * 64 + 32 + e] = __fadd_rn(accv[gq * 64 + 32 + e], __uint_as_float(r1[e]));
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(ACC_EMPTY + ab));
            }
            float *dst = P.partial + ((size_t)it.split * P.ntiles + it.tile) * (size_t)(TM * TN) + (size_t)m * TN + half * 128;
#pragma unroll
            for (int e = 0; e < 128; e += 4)
                *reinterpret_cast<float4 *>(dst + e) = make_float4(accv[e], accv[e + 1], accv[e + 2], accv[e + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------ the GEMM, CTA-pair version (cta_group::2)
// Two CTAs of a cluster (one TPC) compute one 256 x 256 tile: UMMA M = 256 (128 accumulator rows in the tensor memory
// of each CTA), N = 256 with each CTA holding half of the B rows in ITS shared memory.  Per CTA and 64-row stage:
// A hi/lo (own 128 rows) + B hi/lo (own 128 of the 256 B rows) = 64 KB for the same 1536 tensor-pipe cycles -- two
// thirds of the operand traffic of the single-CTA tile -- which also makes room for a third stage.
//   both CTAs   TMA producer (cp.async.bulk.tensor ... cta_group::2: complete_tx lands on the LEADER's full barrier),
//               8 drain warps (own accumulator rows), arrive on the leader's accumulator-empty barrier
//   leader      expect_tx for both CTAs' bytes, single-thread MMA issuer, commits multicast to both CTAs
constexpr int PS_TILE = 128 * 128;                 // bytes of one 128-row operand tile (hi or lo)
constexpr int PS_STAGE_BYTES = 4 * PS_TILE;        // A hi, A lo, B-half hi, B-half lo
constexpr int PS_STAGES = 3;
constexpr int PS_OFF_BAR = PS_STAGES * PS_STAGE_BYTES;
constexpr int PS_NBAR = 2 * PS_STAGES + 4;
constexpr int PS_OFF_TMEM = PS_OFF_BAR + PS_NBAR * 8;
constexpr int PS_SMEM_BYTES = PS_OFF_TMEM + 16 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gram_tc2_pair_kernel(const __grid_constant__ CUtensorMap mapA, const Tc2Params P) {
    extern __shared__ unsigned char smem_dyn[];
    // both CTAs of the pair must use the same shared-memory offsets: the dynamic segment starts at the same offset in
    // every CTA of a launch, so the aligned base is the same too
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const int nitems = P.ntiles * P.nsplit;

    auto bar = [&](int i) { return sbase + PS_OFF_BAR + 8 * i; };
    constexpr int FULL = 0, EMPTY = PS_STAGES, ACC_FULL = 2 * PS_STAGES, ACC_EMPTY = 2 * PS_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + PS_OFF_TMEM);

    if (threadIdx.x == T_TMA) {
        for (int s = 0; s < PS_STAGES; ++s) {
            mbar_init(bar(FULL + s), 1);
            mbar_init(bar(EMPTY + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar(ACC_FULL + a), 1);
            mbar_init(bar(ACC_EMPTY + a), 2 * NDRAIN_WARPS);  // the drain warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t g = 0;
            for (int w = cid; w < nitems; w += ncl) {
                const Item it = decode_item<true>(P, w);
                const int rowA = it.ti * 256 + (int)rank * 128, rowB = it.tj * 256 + (int)rank * 128;
                for (int st = 0; st < it.nst; ++st, ++g) {
                    const int s = g % PS_STAGES;
                    const uint32_t ph = (g / PS_STAGES) & 1;
                    mbar_wait(bar(EMPTY + s), ph ^ 1);
                    const uint32_t dst = sbase + s * PS_STAGE_BYTES;
                    const uint32_t lbar = mapa_rank(bar(FULL + s), 0);
                    const int r0 = (int)(it.r_begin + (int64_t)st * KS);
                    if (leader) mbar_arrive_expect_tx(bar(FULL + s), 2 * (it.diag ? 2 * PS_TILE : PS_STAGE_BYTES));
                    if (!it.diag) {
                        tma_load_2d_pair(dst, &mapA, lbar, r0, rowA);
                        tma_load_2d_pair(dst + PS_TILE, &mapA, lbar, r0, P.mtot + rowA);
                    }
                    tma_load_2d_pair(dst + 2 * PS_TILE, &mapA, lbar, r0, rowB);
                    tma_load_2d_pair(dst + 3 * PS_TILE, &mapA, lbar, r0, P.mtot + rowB);
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            uint32_t g = 0, gc = 0;
            for (int w = cid; w < nitems; w += ncl) {
                const Item it = decode_item<true>(P, w);
                for (int st = 0; st < it.nst; ++st, ++g) {
                    const int s = g % PS_STAGES;
                    const uint32_t ph = (g / PS_STAGES) & 1;
                    const int kk = st % SUB_STAGES;
                    const uint32_t ab = gc & 1;
                    if (kk == 0) mbar_wait(bar(ACC_EMPTY + ab), ((gc >> 1) & 1) ^ 1);
                    mbar_wait(bar(FULL + s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t stage = sbase + s * PS_STAGE_BYTES;
                    const uint32_t b_hi = stage + 2 * PS_TILE, b_lo = stage + 3 * PS_TILE;
                    // diagonal tile: the A rows of each CTA are exactly its half of the B rows
                    const uint32_t a_hi = it.diag ? b_hi : stage, a_lo = it.diag ? b_lo : stage + PS_TILE;
                    const uint32_t acc = tmem_base + ab * 256;
#pragma unroll
                    for (int ks = 0; ks < KS / 16; ++ks) {
                        const uint32_t off = ks * 32;
                        const uint32_t first = (kk == 0 && ks == 0) ? 0u : 1u;
                        umma_f16_ss_pair(acc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_hi + off), idesc, first);
                        umma_f16_ss_pair(acc, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(b_lo + off), idesc, 1u);
                        umma_f16_ss_pair(acc, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(b_hi + off), idesc, 1u);
                    }
                    umma_commit_pair(bar(EMPTY + s));
                    if (kk == SUB_STAGES - 1 || st == it.nst - 1) {
                        umma_commit_pair(bar(ACC_FULL + ab));
                        ++gc;
                    }
                }
            }
        }
    } else {
        // ===================== drain warps (both CTAs, own 128 accumulator rows) =====================
        const int quad = warp & 3, half = warp >> 2;
        const int m = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
        uint32_t gc = 0;
        for (int w = cid; w < nitems; w += ncl) {
            const Item it = decode_item<true>(P, w);
            const int nsub = (it.nst + SUB_STAGES - 1) / SUB_STAGES;
            float accv[128];
#pragma unroll
            for (int e = 0; e < 128; ++e) accv[e] = 0.f;
            for (int c = 0; c < nsub; ++c, ++gc) {
                const uint32_t ab = gc & 1;
                mbar_wait(bar(ACC_FULL + ab), (gc >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int gq = 0; gq < 2; ++gq) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(lane_addr + ab * 256 + (uint32_t)(gq * 64), r0);
                    tmem_ld32(lane_addr + ab * 256 + (uint32_t)(gq * 64 + 32), r1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        accv[gq * 64 + e] = __fadd_rn(accv[gq * 64 + e], __uint_as_float(r0[e]));
                        accv[gq * 64 + 32 + e] = __fadd_rn(accv[gq * 64 + 32 + e], __uint_as_float(r1[e]));
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_rank(bar(ACC_EMPTY + ab), 0));
            }
            float *dst = P.partial + ((size_t)it.split * P.ntiles + it.tile) * (size_t)(256 * 256) +
                         (size_t)(rank * 128 + m) * 256 + half * 128;
#pragma unroll
            for (int e = 0; e < 128; e += 4)
                *reinterpret_cast<float4 *>(dst + e) = make_float4(accv[e], accv[e + 1], accv[e + 2], accv[e + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still touch it
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------ operand space
// Operand row o of the GEMM: o < Kp -> column o of X (zero row when o >= K); o >= Kp -> column o - Kp of Y.
// Kp and the Y extent are multiples of 256, so no strip of the kernels below straddles the two matrices.
struct Cols {
    const float *X, *Y;
    int64_t ldx, ldy;
    int K, n, Kp;
};
struct ColRef {
    const float *base;
    int64_t ld;
    int col, ncols;
};
__device__ __forceinline__ ColRef col_ref(const Cols &C, int o) {
    ColRef r;
    if (o < C.Kp) { r.base = C.X; r.ld = C.ldx; r.col = o; r.ncols = C.K; }
    else { r.base = C.Y; r.ld = C.ldy; r.col = o - C.Kp; r.ncols = C.n; }
    return r;
}

// ------------------------------------------------------------------ statistics pass 1: column sums and max |x|
// CTA = 128 operand rows (32 float4 lanes) x 8 row lanes over one row split; fixed summation order.
__global__ void __launch_bounds__(256)
colstat_part(const Cols C, int64_t nrows, int mtot, double *__restrict__ part, float *__restrict__ part_max) {
    __shared__ double s1[8][132];
    __shared__ float s2[8][132];
    const int cq = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int o0 = blockIdx.x * 128 + cq * 4;
    const ColRef cr = col_ref(C, o0);
    const int64_t per = (nrows + RS - 1) / RS;
    const int64_t r0 = (int64_t)blockIdx.y * per;
    const int64_t r1 = r0 + per < nrows ? r0 + per : nrows;
    double a[4] = {0, 0, 0, 0};
    float mx[4] = {0, 0, 0, 0};
    if (cr.col + 3 < cr.ncols) {
#pragma unroll 4
        for (int64_t r = r0 + rg; r < r1; r += 8) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(cr.base + r * cr.ld + cr.col));
            a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w;
            mx[0] = fmaxf(mx[0], fabsf(v.x)); mx[1] = fmaxf(mx[1], fabsf(v.y));
            mx[2] = fmaxf(mx[2], fabsf(v.z)); mx[3] = fmaxf(mx[3], fabsf(v.w));
        }
    } else if (cr.col < cr.ncols) {
        for (int64_t r = r0 + rg; r < r1; r += 8)
            for (int c = 0; c < 4 && cr.col + c < cr.ncols; ++c) {
                const float v = __ldg(cr.base + r * cr.ld + cr.col + c);
                a[c] += (double)v;
                mx[c] = fmaxf(mx[c], fabsf(v));
            }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { s1[rg][cq * 4 + c] = a[c]; s2[rg][cq * 4 + c] = mx[c]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        double t = 0.0;
        float m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t += s1[k][threadIdx.x]; m2 = fmaxf(m2, s2[k][threadIdx.x]); }
        part[(size_t)blockIdx.y * mtot + blockIdx.x * 128 + threadIdx.x] = t;
        part_max[(size_t)blockIdx.y * mtot + blockIdx.x * 128 + threadIdx.x] = m2;
    }
}

// shift[o] = fl32(mean), scale[o] = 2^e with 2 * max|x| * 2^e in [2^9, 2^10), inv[o] = 2^-e (fp64)
__global__ void __launch_bounds__(128)
colstat_finish(const double *__restrict__ part, const float *__restrict__ part_max, int mtot, double invN,
               float *__restrict__ shift, float *__restrict__ scale, double *__restrict__ inv) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= mtot) return;
    double v[RS];
    float m[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) { v[r] = part[(size_t)r * mtot + o]; m[r] = part_max[(size_t)r * mtot + o]; }
    double t = 0.0;
    float mx = 0.f;
#pragma unroll
    for (int r = 0; r < RS; ++r) { t += v[r]; mx = fmaxf(mx, m[r]); }
    shift[o] = (float)(t * invN);
    int e = 0;
    const float b = 2.f * mx;  // |x - mean| <= 2 max|x|
    if (b > 0.f && isfinite(b)) {
        e = 9 - ilogbf(b);
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    scale[o] = ldexpf(1.f, e);
    inv[o] = ldexp(1.0, -e);
}

// ------------------------------------------------------------------ operand preparation (+ statistics pass 2)
// One CTA = a strip of 64 operand rows x the 64-row tiles of one row split.  Reads the data once (coalesced along the
// columns), writes the hi and lo operand rows (coalesced along the reduction index) through a swizzled shared-memory
// tile, and accumulates the fp64 column sums / sums of squares of xs in a fixed order.
__global__ void __launch_bounds__(256)
tc2_prep(const Cols C, int64_t nrows, int64_t Np, int tiles_per_split, int mtot, const float *__restrict__ shift,
         const float *__restrict__ scale, __half *__restrict__ Ohi, __half *__restrict__ Olo, double *__restrict__ part,
         double *__restrict__ part_sq) {
    __shared__ __align__(16) unsigned char tile_hi[PT * 128], tile_lo[PT * 128];
    __shared__ double red[16][PT + 1];
    const int t = threadIdx.x, jq = t & 15, rg = t >> 4;
    const int o0 = blockIdx.x * PT;       // first operand row of the strip
    const ColRef cr = col_ref(C, o0);
    const float *__restrict__ X = cr.base;
    const int64_t ld = cr.ld;
    const int ncols = cr.ncols;
    const int cj = cr.col + jq * 4;
    const bool vec = (cj + 3 < ncols);
    float sh[4], sc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sh[c] = cj + c < ncols ? shift[o0 + jq * 4 + c] : 0.f;
        sc[c] = cj + c < ncols ? scale[o0 + jq * 4 + c] : 0.f;
    }
    double a[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const int64_t ntile = Np / PT;
    const int64_t tb = (int64_t)blockIdx.y * tiles_per_split;
    int64_t te = tb + tiles_per_split;
    if (te > ntile) te = ntile;

    float4 cur[4], nxt[4];
    auto load = [&](int64_t tile, float4(&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = tile * PT + rg * 4 + i;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) {
                const float *p = X + r * ld + cj;
                if (vec) x = __ldg(reinterpret_cast<const float4 *>(p));
                else {
                    if (cj + 0 < ncols) x.x = __ldg(p + 0);
                    if (cj + 1 < ncols) x.y = __ldg(p + 1);
                    if (cj + 2 < ncols) x.z = __ldg(p + 2);
                }
            }
            v[i] = x;
        }
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tb < te) load(tb, cur);
    for (int64_t tile = tb; tile < te; ++tile) {
        if (tile + 1 < te) load(tile + 1, nxt);
        // split: thread owns rows rg*4 .. rg*4+3 of the tile and 4 columns
        __half hi[4][4], lo[4][4];  // [column][row]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool live = tile * PT + rg * 4 + i < nrows;
            const float xv[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float xs = (live && cj + c < ncols) ? __fsub_rn(xv[c], sh[c]) : 0.f;
                const double x64 = (double)xs;
                a[c] += x64;
                q[c] = fma(x64, x64, q[c]);
                const float v = __fmul_rn(xs, sc[c]);
                const __half h = __float2half_rn(v);
                hi[c][i] = h;
                lo[c][i] = __float2half_rn(__fsub_rn(v, __half2float(h)));
            }
        }
        // operand row j of the strip: 64 reduction values = 8 chunks of 16 B; chunk index swizzled by j >> 2
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = jq * 4 + c;
            const uint32_t off = (uint32_t)j * 128u + (uint32_t)(((rg >> 1) ^ (jq & 7)) << 4) + (uint32_t)((rg & 1) << 3);
            uint2 ph, pl;
            ph.x = (uint32_t)__half_as_ushort(hi[c][0]) | ((uint32_t)__half_as_ushort(hi[c][1]) << 16);
            ph.y = (uint32_t)__half_as_ushort(hi[c][2]) | ((uint32_t)__half_as_ushort(hi[c][3]) << 16);
            pl.x = (uint32_t)__half_as_ushort(lo[c][0]) | ((uint32_t)__half_as_ushort(lo[c][1]) << 16);
            pl.y = (uint32_t)__half_as_ushort(lo[c][2]) | ((uint32_t)__half_as_ushort(lo[c][3]) << 16);
            *reinterpret_cast<uint2 *>(tile_hi + off) = ph;
            *reinterpret_cast<uint2 *>(tile_lo + off) = pl;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int j = (t >> 3) + 32 * it, ch = t & 7;
            const uint32_t off = (uint32_t)j * 128u + (uint32_t)((ch ^ ((j >> 2) & 7)) << 4);
            const size_t o = (size_t)(o0 + j) * (size_t)Np + (size_t)tile * PT + (size_t)ch * 8;
            *reinterpret_cast<uint4 *>(Ohi + o) = *reinterpret_cast<const uint4 *>(tile_hi + off);
            *reinterpret_cast<uint4 *>(Olo + o) = *reinterpret_cast<const uint4 *>(tile_lo + off);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
    }
    // column statistics of the strip: 16 row groups summed in a fixed order
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[rg][jq * 4 + c] = pass ? q[c] : a[c];
        __syncthreads();
        if (t < PT) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) s += red[k][t];
            (pass ? part_sq : part)[(size_t)blockIdx.y * mtot + o0 + t] = s;
        }
        __syncthreads();
    }
}

// T[o] = sum of the split partials (fixed order), SQ likewise; the caller's column sums:
// out[col] = T + N * ((double)shift - bias)   (sx for the X part, sy for the Y part)
__global__ void __launch_bounds__(128)
tc2_stat_finish(const double *__restrict__ part, const double *__restrict__ part_sq, int nsplit_used, int mtot, Cols C,
                const float *__restrict__ shift, const float *__restrict__ y_bias, double Nd, double *__restrict__ T,
                double *__restrict__ SQ, double *__restrict__ sx, double *__restrict__ sy) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= mtot) return;
    double v[RS], w[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        v[r] = r < nsplit_used ? part[(size_t)r * mtot + o] : 0.0;
        w[r] = r < nsplit_used ? part_sq[(size_t)r * mtot + o] : 0.0;
    }
    double t = 0.0, t2 = 0.0;
#pragma unroll
    for (int r = 0; r < RS; ++r) { t += v[r]; t2 += w[r]; }
    T[o] = t;
    SQ[o] = t2;
    if (o < C.Kp) {
        if (sx && o < C.K) sx[o] = t + Nd * (double)shift[o];
    } else if (sy && o - C.Kp < C.n) {
        const int j = o - C.Kp;
        sy[j] = t + Nd * ((double)shift[o] - (y_bias ? (double)y_bias[j] : 0.0));
    }
}

// ------------------------------------------------------------------ reduction of the row splits
// CTA = 32 rows x 128 columns (one 128-column block) of a tile; thread = 4 consecutive columns x 4 rows (float4 loads of
// every split's partial, all issued before the first use).
//   C[i,j] = inv_i inv_j sum_s P_s[i,j] + uA_i TB_j + TA_i uB_j + N uA_i uB_j,   u = (double)shift32 - bias.
// G tiles: 128-blocks below the diagonal block are skipped, the diagonal 128-block reads the upper element for both
// (i,j) and (j,i) (the tensor core produced them with different rounding; G must be bitwise symmetric), blocks above
// it are also written transposed (the lower triangle of G) through shared memory.  TR = rows of a tile (128, or 256
// for the pair kernel).
template <int TR>
__global__ void __launch_bounds__(256)
reduce_tc2(const float *__restrict__ partial, int nsplit, int ntiles, int tiles_sym, int tjx, int tjy0, int tnb, int Kp,
           const float *__restrict__ shift, const double *__restrict__ inv, const double *__restrict__ T,
           const double *__restrict__ SQ, const float *__restrict__ y_bias, double Nd, int K, int n,
           double *__restrict__ G, double *__restrict__ Bxy) {
    __shared__ double tr[32][129];
    int l = blockIdx.x, ti, tj;
    const bool sym = l < tiles_sym;
    if (sym) {
        ti = 0;
        if (TR == 128) {
            while (l >= tjx - (ti >> 1)) { l -= tjx - (ti >> 1); ++ti; }
            tj = (ti >> 1) + l;
        } else {
            while (l >= tjx - ti) { l -= tjx - ti; ++ti; }
            tj = ti + l;
        }
    } else {
        l -= tiles_sym;
        ti = l / tnb;
        tj = tjy0 + (l - ti * tnb);
    }
    const int rowbase = ti * TR, colbase = tj * TN;  // colbase in operand space
    const int sr = blockIdx.y, i0 = rowbase + sr * 32;
    if (i0 >= K) return;
    const int Nn = sym ? K : n;
    const int coff = sym ? 0 : Kp;  // operand index of column 0 of C
    const int oj0 = colbase + blockIdx.z * 128, j0 = oj0 - coff;
    if (j0 >= Nn) return;
    const int rb128 = i0 >> 7, cb128 = j0 >> 7;
    if (sym && cb128 < rb128) return;
    const bool dblock = sym && cb128 == rb128, upper = sym && cb128 > rb128;
    const size_t tile_elems = (size_t)TR * TN, split_stride = (size_t)ntiles * tile_elems;
    const float *p0 = partial + (size_t)blockIdx.x * tile_elems;
    double *__restrict__ Cm = sym ? G : Bxy;
    const int64_t ldc = sym ? K : n;
    const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int jc = j0 + cq * 4, ojc = oj0 + cq * 4;  // first of this thread's 4 columns
    const int ec0 = blockIdx.z * 128 + cq * 4;        // ... inside the tile

    double sv[4][4];
    if (!dblock) {
        double d[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) d[p][e] = 0.0;
        for (int c = 0; c < nsplit; ++c) {
            float4 v[4];
#pragma unroll
            for (int p = 0; p < 4; ++p)
                v[p] = __ldg(reinterpret_cast<const float4 *>(p0 + (size_t)c * split_stride +
                                                              (size_t)(sr * 32 + rl + 8 * p) * TN + ec0));
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                d[p][0] += (double)v[p].x; d[p][1] += (double)v[p].y; d[p][2] += (double)v[p].z; d[p][3] += (double)v[p].w;
            }
        }
        double invj[4], ubj[4], Tj[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool ok = jc + e < Nn;
            invj[e] = ok ? inv[ojc + e] : 0.0;
            Tj[e] = ok ? T[ojc + e] : 0.0;
            ubj[e] = ok ? (double)shift[ojc + e] - ((!sym && y_bias) ? (double)y_bias[jc + e] : 0.0) : 0.0;
        }
        const bool vec2 = ((ldc & 1) == 0) && (jc + 3 < Nn) && ((reinterpret_cast<uintptr_t>(Cm) & 15) == 0);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int i = i0 + rl + 8 * p;
            if (i < K) {
                const double invi = inv[i], ua = (double)shift[i], Ti = T[i];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    sv[p][e] = d[p][e] * (invi * invj[e]) + (ua * Tj[e] + Ti * ubj[e] + Nd * ua * ubj[e]);
                double *dst = Cm + (int64_t)i * ldc + jc;
                if (vec2) {
                    *reinterpret_cast<double2 *>(dst) = make_double2(sv[p][0], sv[p][1]);
                    *reinterpret_cast<double2 *>(dst + 2) = make_double2(sv[p][2], sv[p][3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (jc + e < Nn) dst[e] = sv[p][e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) sv[p][e] = 0.0;
            }
        }
    } else {
        // diagonal 128-block: element-wise, (i > j) reads the mirrored element
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int lr = rl + 8 * p, i = i0 + lr;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = jc + e;
                double s = 0.0;
                if (i < K && j < K) {
                    int er = sr * 32 + lr, ec = ec0 + e;
                    if (i > j) {
                        const int nr = (colbase + ec) - rowbase, nc = (rowbase + er) - colbase;
                        er = nr;
                        ec = nc;
                    }
                    for (int c = 0; c < nsplit; ++c) s += (double)p0[(size_t)c * split_stride + (size_t)er * TN + ec];
                    s *= inv[i] * inv[j];
                    if (i == j) s = SQ[i];
                    const double ua = (double)shift[i], ub = (double)shift[j];
                    s += ua * T[j] + T[i] * ub + Nd * ua * ub;
                    G[(int64_t)i * K + j] = s;
                }
                sv[p][e] = s;
            }
        }
    }
    if (upper) {  // the transposed copy: G[j][i] for the 32 x 128 strip (CTA-uniform condition)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) tr[rl + 8 * p][cq * 4 + e] = sv[p][e];
        __syncthreads();
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const int col = w * 16 + k;
            const int j = j0 + col, i = i0 + lane;
            if (i < K && j < K) G[(int64_t)j * K + i] = tr[lane][col];
        }
    }
}

}  // namespace

// CPB200_GRAM_PAIR=0 selects the single-CTA 128 x 256 tiles (A/B measurements); default: CTA pairs, 256 x 256 tiles
static bool tc2_use_pair() {
    static const bool pair = [] { const char *e = getenv("CPB200_GRAM_PAIR"); return !(e && e[0] == '0'); }();
    return pair;
}

int cp_gram_tc2(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype, int n,
                int64_t ldy, const float *y_bias, const int32_t *rows, int nrows, double *G, double *Bxy, double *sx,
                double *sy, double *yy, cudaStream_t stream) {
    const bool wantB = Bxy != nullptr;
    if (yy != nullptr || !cp_gram_tc_eligible(X, N, K, ldx, Yraw, y_dtype, n, ldy, rows, wantB || sy != nullptr))
        return cp_gram_fp64_products(h, X, N, K, ldx, Yraw, y_dtype, n, ldy, y_bias, rows, nrows, G, Bxy, sx, sy, yy, stream);
    if (sy != nullptr && !wantB)  // column sums of Y alone: nothing for the tensor cores to do
        return cp_gram_fp64_products(h, X, N, K, ldx, Yraw, y_dtype, n, ldy, y_bias, rows, nrows, G, Bxy, sx, sy, yy, stream);
    const bool pair = tc2_use_pair();
    const int TR = pair ? 256 : TM;
    const int tjx = cp_cdiv(K, TN), tk = cp_cdiv(K, TR);
    const int Kp = tjx * TN;
    const int tnb = wantB ? cp_cdiv(n, TN) : 0;
    const int np_ = tnb * TN;
    const int mtot = Kp + np_;
    int tiles_sym = 0;
    if (G)
        for (int ti = 0; ti < tk; ++ti) tiles_sym += pair ? tjx - ti : tjx - (ti >> 1);
    const int ntiles = tiles_sym + tk * tnb;
    const int64_t Np = cp_cdiv(N, KS) * (int64_t)KS;
    const int units = pair ? h->num_sms / 2 : h->num_sms;  // CTAs or CTA pairs working concurrently

    // Row splits.  An item (tile x split) costs its stages (~1.2 us each) plus a small hand-over; items run
    // round-robin on the persistent grid; every split adds one fp32 partial per tile (written, then read by the
    // reduction).  A split is a multiple of the 128-row accumulator run and at most 32 runs (the fp32 register
    // sums stay below 4e-7 of the partial sum whatever N is).
    constexpr int64_t SUB = SUB_STAGES * KS, MAX_SPLIT_ROWS = 32 * SUB;
    const int ns_min = (int)cp_cdiv(N, MAX_SPLIT_ROWS);
    int nsplit = ns_min, rps = (int)(cp_cdiv(cp_cdiv(N, ns_min), SUB) * SUB);
    if (ntiles > 0) {
        double best = 1e300;
        const int max_ns = (int)cp_cdiv(N, SUB);
        for (int ns = ns_min; ns <= max_ns && ns < ns_min + 32; ++ns) {
            const int64_t r = cp_cdiv(cp_cdiv(N, ns), SUB) * SUB;
            const int ns_eff = (int)cp_cdiv(N, r);
            const double rounds = (double)cp_cdiv((int64_t)ntiles * ns_eff, units);
            const double cost = rounds * (1.2 * (double)(r / KS) + 1.0) +
                                (double)ns_eff * (2.0 * ntiles * TR * TN * 4.0 / 5.0e6);
            if (cost < best * 0.98) {
                best = cost;
                nsplit = ns_eff;
                rps = (int)r;
            }
        }
    }

    const size_t part_elems = (size_t)nsplit * ntiles * TR * TN;
    const size_t op_elems = 2 * (size_t)mtot * (size_t)Np;  // hi rows, then lo rows
    const size_t need = cp_carver::need(part_elems, 4) + cp_carver::need(op_elems, 2) + 3 * cp_carver::need(mtot, 8) +
                        2 * cp_carver::need((size_t)RS * mtot, 8) + cp_carver::need((size_t)RS * mtot, 4) +
                        2 * cp_carver::need(mtot, 4);
    void *ws = nullptr;
    int rc = cp_ws_reserve(h, need, &ws);
    if (rc) return rc;
    cp_carver cv(ws);
    float *partial = cv.take<float>(part_elems);
    __half *ops = cv.take<__half>(op_elems);
    double *T = cv.take<double>(mtot), *SQ = cv.take<double>(mtot), *inv = cv.take<double>(mtot);
    double *cpart = cv.take<double>((size_t)RS * mtot), *cpart_sq = cv.take<double>((size_t)RS * mtot);
    float *cpart_max = cv.take<float>((size_t)RS * mtot);
    float *shift = cv.take<float>(mtot), *scale = cv.take<float>(mtot);
    const double invN = 1.0 / (double)N, Nd = (double)N;
    __half *Ohi = ops, *Olo = ops + (size_t)mtot * (size_t)Np;
    const int64_t ntile_rows = Np / PT;
    const int tiles_per_split = cp_cdiv(ntile_rows, RS);
    const int splits_used = cp_cdiv(ntile_rows, tiles_per_split);
    Cols C{X, (const float *)Yraw, ldx, ldy, K, wantB ? n : 0, Kp};

    colstat_part<<<dim3(mtot / 128, RS), 256, 0, stream>>>(C, N, mtot, cpart, cpart_max);
    CP_CHECK_LAUNCH();
    colstat_finish<<<cp_cdiv(mtot, 128), 128, 0, stream>>>(cpart, cpart_max, mtot, invN, shift, scale, inv);
    CP_CHECK_LAUNCH();
    tc2_prep<<<dim3(mtot / PT, splits_used), 256, 0, stream>>>(C, N, Np, tiles_per_split, mtot, shift, scale, Ohi, Olo, cpart,
                                                               cpart_sq);
    CP_CHECK_LAUNCH();
    tc2_stat_finish<<<cp_cdiv(mtot, 128), 128, 0, stream>>>(cpart, cpart_sq, splits_used, mtot, C, shift, y_bias, Nd, T, SQ, sx,
                                                            wantB ? sy : nullptr);
    CP_CHECK_LAUNCH();
    if (ntiles > 0) {
        CUtensorMap mapA, mapB;
        rc = make_map16(h, &mapA, ops, Np, 2 * (int64_t)mtot, TM);
        if (rc) return rc;
        Tc2Params P{};
        P.partial = partial; P.Np = Np; P.rows_per_split = rps; P.nsplit = nsplit; P.tiles_sym = tiles_sym; P.tk = tk;
        P.tjx = tjx; P.tjy0 = Kp / TN; P.tnb = tnb; P.ntiles = ntiles; P.mtot = mtot;
        const int nitems = ntiles * nsplit;
        const bool prof = h->gram_profile;
        if (prof) CP_CUDA(cudaEventRecord(h->ev_gram0, stream));
        if (pair) {
            static cp_per_device_flag configured;
            if (bool *done = configured.slot(); !*done) {
                CP_CUDA(cudaFuncSetAttribute(gram_tc2_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_BYTES));
                *done = true;
            }
            const int ncl = nitems < units ? nitems : units;
            gram_tc2_pair_kernel<<<2 * ncl, NTHREADS, PS_SMEM_BYTES, stream>>>(mapA, P);
        } else {
            rc = make_map16(h, &mapB, ops, Np, 2 * (int64_t)mtot, TN);
            if (rc) return rc;
            static cp_per_device_flag configured;
            if (bool *done = configured.slot(); !*done) {
                CP_CUDA(cudaFuncSetAttribute(gram_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
                *done = true;
            }
            const int grid = nitems < units ? nitems : units;
            gram_tc2_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(mapA, mapB, P);
        }
        CP_CHECK_LAUNCH();
        if (prof) CP_CUDA(cudaEventRecord(h->ev_gram1, stream));
        if (pair)
            reduce_tc2<256><<<dim3(ntiles, 8, 2), 256, 0, stream>>>(partial, nsplit, ntiles, tiles_sym, tjx, P.tjy0, tnb, Kp, shift,
                                                                   inv, T, SQ, y_bias, Nd, K, n, G, Bxy);
        else
            reduce_tc2<128><<<dim3(ntiles, 4, 2), 256, 0, stream>>>(partial, nsplit, ntiles, tiles_sym, tjx, P.tjy0, tnb, Kp, shift,
                                                                   inv, T, SQ, y_bias, Nd, K, n, G, Bxy);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}
