// cp_gram, mode CP_GRAM_3XTF32: the tall-skinny Gram / cross product on the 5th-gen tensor cores.
//
//   G = X'X (K x K),  Bxy = X'(Y - b) (K x n)      X: N x K fp32 row-major, N ~ 5e3..1e5, K = c*k*k
//
// The reference does this arithmetic in float64 on the CPU (numpy matmul / LAPACK inside
// LinearRegression.fit, lib/decompose.py:665-666).  tcgen05 has no fp32/fp64 MMA, so the kernel
// evaluates every product with the error-compensated 3xTF32 scheme on SHIFTED data:
//
//   xs = fl32(x - s_col)            s = fp32 column mean (one extra pass): removes the rank-one
//                                   mean component that would otherwise dominate the fp32 sums
//   xs = hi + lo                    hi = tf32_rn(xs), lo = tf32_rn(xs - hi)  (22 mantissa bits kept)
//   P += hi'hi + hi'lo + lo'hi      three kind::tf32 MMAs per k-step, fp32 accumulation in TMEM
//
// and keeps each fp32 accumulation short (the tensor core truncates when it adds into its fp32 accumulator):
// the N rows are cut into row splits (one CTA per 128x128 output tile and split) and, inside a CTA, into
// 128-row sub-chunks that alternate between two TMEM accumulators; finished sub-chunks are added into fp32
// registers with round-to-nearest, every CTA writes ONE fp32 partial tile, and a second kernel sums the
// partials of the splits in fp64 and undoes the shift exactly
// (G = P + s T' + T s' + N s s',  T = column sums of xs in fp64).  The diagonal comes from an fp64 pass.
//
// CTA anatomy (8 converter warps + 2 single-thread roles on the highest warp ids, one CTA per SM):
//   TMA warp     cp.async.bulk.tensor 2D boxes of raw fp32 X (32 rows x 128 cols for A and for B), 4-stage ring,
//                mbarrier complete_tx; the first boxes are issued before the rest of the set-up
//   MMA warp     TMEM allocator + single-thread tcgen05.mma issuer: TS mode (A from tensor memory, B from shared
//                memory), 12 MMAs of 128x128x8 per 32-row k-block, tcgen05.commit frees the operand stage and,
//                at the end of a sub-chunk, publishes the accumulator
//   converters   thread = one operand row (= its TMEM lane) x 16 k-values: shift, split hi/lo, A -> tcgen05.st
//                (columns = k), B -> K-major 128B-swizzled shared memory (X is "MN-major" in memory; the
//                transposition is free because the data passes through registers for the split anyway);
//                between k-blocks they drain the accumulator of the PREVIOUS sub-chunk (tcgen05.ld) into
//                registers, so the epilogue overlaps the MMAs; finally one fp32 partial tile per CTA.
// Bound: tensor pipe (3 passes -> at most 1/3 of the dense TF32 rate in algorithmic flops); measured state and
// the ablation study are in profiles/r1c_summary.md.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

int cp_gram_fp64_products(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype,
                          int n, int64_t ldy, const float *y_bias, const int32_t *rows, int nrows, double *G,
                          double *Bxy, double *sx, double *sy, double *yy, cudaStream_t stream);

namespace {

constexpr int TM = 128, TN = 128, KB = 32;           // output tile, rows per k-block
constexpr int RAW_TILE = KB * TM * 4;                // 16 KB raw fp32 box
constexpr int OP_TILE = TM * 128;                    // 16 KB operand tile (128 rows x 128 B)
constexpr int STAGES = 3;      // operand stages: B hi/lo in shared memory (K-major swizzled), A hi/lo in TMEM
constexpr int RSTAGES = 4;     // raw fp32 stages in flight
#ifndef CP_TC_CONV_WARPS
#define CP_TC_CONV_WARPS 8
#endif
constexpr int NCONV = 32 * CP_TC_CONV_WARPS;   // converter / drain threads (8 or 16 warps)
constexpr int NTHREADS = NCONV + 64;           // converters first, then the TMA warp, then the MMA warp
// Warp roles.  The two single-thread roles take the HIGHEST warp ids: the warp scheduler favours high warp ids, and
// with the roles at warps 0/1 the polling converter warps starved the TMA / MMA issuers of issue slots.
constexpr int W_TMA = CP_TC_CONV_WARPS, W_MMA = CP_TC_CONV_WARPS + 1;
constexpr int T_TMA = 32 * W_TMA, T_MMA = 32 * W_MMA;
constexpr int WPQ = CP_TC_CONV_WARPS / 4;      // converter warps per TMEM lane quadrant
constexpr int KPT = KB / WPQ;                  // k-values (rows of the k-block) converted per thread and operand
constexpr int CPT = TN / WPQ;                  // accumulator columns drained per thread
static_assert(KPT == 8 || KPT == 16, "converter layout");
constexpr int NCONV_WARPS = NCONV / 32;  // barrier arrivals are per warp (one elected lane after __syncwarp)
// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_OPS = 0;                                    // STAGES x {Bhi, Blo}
constexpr int OFF_RAW = OFF_OPS + STAGES * 2 * OP_TILE;       // RSTAGES x {rawA, rawB}
constexpr int OFF_SHIFT = OFF_RAW + RSTAGES * 2 * RAW_TILE;    // shiftA[128], shiftB[128] fp32
constexpr int OFF_BAR = OFF_SHIFT + 2 * TM * 4;               // mbarriers
constexpr int NBAR = 2 * RSTAGES + 2 * STAGES + 2;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;              // + alignment slack

struct TcParams {
    const float *shiftA;  // K   fp32 column shifts of X
    const float *shiftB;  // n   fp32 column shifts of Y (B operand of the X'Y tiles)
    float *partial;       // [nsplit][ntiles][128][128]
    int K, nB;            // columns of X, columns of Y
    int64_t N;
    int rows_per_chunk;   // rows per split (one CTA per tile and split)
    int tiles_sym;        // number of upper-triangular G tiles (0 when G is not requested)
    int tk;               // ceil(K / 128)
    int tnb;              // ceil(nB / 128) for the X'Y tiles
    int ntiles;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
template <bool TEST_WAIT>
__device__ __forceinline__ void mbar_wait_impl(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        if (TEST_WAIT) {  // non-blocking probe: the thread keeps polling instead of being suspended
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        }
        if (spin > (1u << 26)) __trap();
    }
}
// ablation builds (make ablate): CP_TC_ABLATE_MASK is a compile-time constant, so the product build carries none of it
// bit0: converters skip their work, bit1: no MMAs, bit2: no TMA loads, bit3: no fences, bit4: no drain, bit6: test_wait
#ifndef CP_TC_ABLATE_MASK
#define CP_TC_ABLATE_MASK 0
#endif
#define TC_ABLATE(bit) ((CP_TC_ABLATE_MASK) & (bit))
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_impl<(TC_ABLATE(64) != 0)>(bar, parity); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    // cute::UMMA::SmemDescriptor: start[0,14) (>>4) | LBO[16,30) | SBO[32,46) | version[46,48) = 1 | layout[61,64) = 2
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// A operand from tensor memory (lanes = rows of A, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ __attribute__((unused)) void tmem_st(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ __attribute__((unused)) void tmem_st(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// round-to-nearest (ties away) to the 10-bit tf32 mantissa with two full-rate integer ops; identical to
// cvt.rna.tf32.f32 for finite inputs, but not issued on the quarter-rate conversion pipe
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// ------------------------------------------------------------------ optional in-kernel timeline (profiling builds only)
#ifdef CP_TC_TIMING
__device__ long long cp_tc_times[64][16];
#define TC_T(slot) do { if (timed) cp_tc_times[blockIdx.x][slot] = clock64(); } while (0)
#define TC_ACC(slot, expr) do { long long _a = 0; if (timed) _a = clock64(); expr; if (timed) cp_tc_times[blockIdx.x][slot] += clock64() - _a; } while (0)
#else
#define TC_T(slot) do { } while (0)
#define TC_ACC(slot, expr) do { expr; } while (0)
#endif

// ------------------------------------------------------------------ main kernel
// One CTA = one 128x128 output tile x one row split.  The rows of the split are consumed in sub-chunks of
// CHUNK_KB k-blocks (256 rows); sub-chunk c accumulates into TMEM accumulator pair (c & 1) -- two accumulators
// alternating by k-block, so no fp32 accumulator takes more than 48 truncating additions -- while the
// converter warps drain pair ((c-1) & 1) into fp32 registers (round-to-nearest adds).  One fp32 partial
// tile per CTA leaves the kernel.
constexpr int CHUNK_KB = 4;
constexpr uint32_t TM_ACC = 0, TM_A = 2 * TN;  // TMEM columns: two accumulators, then STAGES x {A hi (32), A lo (32)}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__global__ void __launch_bounds__(NTHREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, const TcParams P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef CP_TC_TIMING
    const bool timed = blockIdx.x < 64 && blockIdx.y == 0 && (threadIdx.x == T_TMA || threadIdx.x == T_MMA || threadIdx.x == 0);
    if (timed && threadIdx.x == T_TMA) { for (int i = 8; i < 16; ++i) cp_tc_times[blockIdx.x][i] = 0; }
    if (threadIdx.x == T_TMA) TC_T(0);
#endif

    // ---- work item
    int l = blockIdx.x, ti, tj;
    bool is_xy = false;
    if (l < P.tiles_sym) {
        ti = 0;
        while (l >= P.tk - ti) { l -= P.tk - ti; ++ti; }
        tj = ti + l;
    } else {
        l -= P.tiles_sym;
        is_xy = true;
        ti = l / P.tnb;
        tj = l - ti * P.tnb;
    }
    const bool diag = !is_xy && ti == tj;  // A and B operands are the same tile
    const CUtensorMap *mapA = &mapX, *mapB = is_xy ? &mapY : &mapX;
    const int64_t r_begin = (int64_t)blockIdx.y * P.rows_per_chunk;
    int64_t r_end = r_begin + P.rows_per_chunk;
    if (r_end > P.N) r_end = P.N;
    const int nkb = (int)((r_end - r_begin + KB - 1) / KB);
    const int nchunk = (nkb + CHUNK_KB - 1) / CHUNK_KB;

    auto bar = [&](int i) { return sbase + OFF_BAR + 8 * i; };
    // barrier indices
    constexpr int RAW_FULL = 0, RAW_EMPTY = RSTAGES, OPS_FULL = 2 * RSTAGES, OPS_EMPTY = 2 * RSTAGES + STAGES;
    constexpr int ACC_FULL = 2 * RSTAGES + 2 * STAGES;  // two: one per accumulator pair
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);

    if (threadIdx.x == T_TMA) {
        for (int s = 0; s < RSTAGES; ++s) {
            mbar_init(bar(RAW_FULL + s), 1);
            mbar_init(bar(RAW_EMPTY + s), NCONV_WARPS);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(OPS_FULL + s), NCONV_WARPS);
            mbar_init(bar(OPS_EMPTY + s), 1);
        }
        mbar_init(bar(ACC_FULL + 0), 1);
        mbar_init(bar(ACC_FULL + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first RSTAGES boxes need nothing but their barriers: get them in flight before the rest of the set-up
        for (int kb = 0; kb < nkb && kb < RSTAGES; ++kb) {
            if (TC_ABLATE(4)) { mbar_arrive(bar(RAW_FULL + kb)); continue; }
            mbar_arrive_expect_tx(bar(RAW_FULL + kb), diag ? RAW_TILE : 2 * RAW_TILE);
            const int row = (int)(r_begin + (int64_t)kb * KB);
            tma_load_2d(sbase + OFF_RAW + (kb * 2 + 0) * RAW_TILE, mapA, bar(RAW_FULL + kb), ti * TM, row);
            if (!diag) tma_load_2d(sbase + OFF_RAW + (kb * 2 + 1) * RAW_TILE, mapB, bar(RAW_FULL + kb), tj * TN, row);
        }
    }
    if (warp == W_MMA) {  // TMEM: all 512 columns = 2 pairs of fp32 accumulators of 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // column shifts of the two operand tiles
    float *shA = reinterpret_cast<float *>(smem + OFF_SHIFT), *shB = shA + TM;
    if (threadIdx.x < TM) {
        const int ca = ti * TM + threadIdx.x;
        shA[threadIdx.x] = ca < P.K ? P.shiftA[ca] : 0.f;
        const int cb = tj * TN + threadIdx.x;
        const int nB = is_xy ? P.nB : P.K;
        const float *shiftB = is_xy ? P.shiftB : P.shiftA;
        shB[threadIdx.x] = cb < nB ? shiftB[cb] : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == T_TMA) TC_T(1);

    if (warp == W_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = RSTAGES; kb < nkb; ++kb) {
                const int s = kb % RSTAGES;
                const uint32_t ph = (kb / RSTAGES) & 1;
                TC_ACC(8, mbar_wait(bar(RAW_EMPTY + s), ph ^ 1));  // slot free
                if (TC_ABLATE(4)) { mbar_arrive(bar(RAW_FULL + s)); continue; }
                mbar_arrive_expect_tx(bar(RAW_FULL + s), diag ? RAW_TILE : 2 * RAW_TILE);
                const int row = (int)(r_begin + (int64_t)kb * KB);
                tma_load_2d(sbase + OFF_RAW + (s * 2 + 0) * RAW_TILE, mapA, bar(RAW_FULL + s), ti * TM, row);
                if (!diag) tma_load_2d(sbase + OFF_RAW + (s * 2 + 1) * RAW_TILE, mapB, bar(RAW_FULL + s), tj * TN, row);
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // InstrDescriptor: c_format F32 (1<<4) | a,b TF32 (2<<7, 2<<10) | K-major | N>>3 at 17 | M>>4 at 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                const int c = kb / CHUNK_KB, kk = kb - c * CHUNK_KB;
                TC_ACC(9, mbar_wait(bar(OPS_FULL + s), ph));
                if (kb == 0) TC_T(3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = sbase + OFF_OPS + s * 2 * OP_TILE, b_lo = b_hi + OP_TILE;
                const uint32_t a_hi = tmem_base + TM_A + (uint32_t)(s * 2 * KB), a_lo = a_hi + KB;
                // The tensor core truncates when it adds into the fp32 accumulator, so long positive sums drift
                // low in proportion to the number of additions: a sub-chunk (4 k-blocks, 48 additions) per accumulator.
                // Accumulator (c & 1) was drained by the converters before they delivered this sub-chunk's operands.
                const uint32_t acc = tmem_base + TM_ACC + (uint32_t)((c & 1) * TN);
#pragma unroll
                for (int ks = 0; ks < KB / 8; ++ks) {
                    if (TC_ABLATE(2)) break;
                    const uint32_t off = ks * 32;  // 8 tf32 = 32 bytes along K inside the 128-byte swizzled row
                    const uint32_t first = (kk == 0 && ks == 0) ? 0u : 1u;
                    umma_tf32_ts(acc, a_hi + ks * 8, umma_desc_k_sw128(b_hi + off), idesc, first);
                    umma_tf32_ts(acc, a_hi + ks * 8, umma_desc_k_sw128(b_lo + off), idesc, 1u);
                    umma_tf32_ts(acc, a_lo + ks * 8, umma_desc_k_sw128(b_hi + off), idesc, 1u);
                }
                umma_commit(bar(OPS_EMPTY + s));  // operand stage free once these MMAs have read it
                if (kk == CHUNK_KB - 1 || kb == nkb - 1) umma_commit(bar(ACC_FULL + (c & 1)));  // pair complete
            }
            TC_T(4);
        }
    } else {
        // ===================== converters + drain =====================
        // thread -> (m, kh): m = row of the operands (column of the raw boxes) = TMEM lane, which a warp can only
        // reach inside its own quadrant (warp % 4); kh = which KPT-row slice of the 32-row k-block this thread converts
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int kh = warp >> 2;
        // Shared-space (32-bit) addresses and explicit ld/st.shared: through the aligned generic pointer the
        // compiler emits generic LD/ST and, fearing aliasing, serialises load -> split -> store per 4 values.
        // All raw values of the thread are loaded first so that their latency overlaps.
        const float shva = shA[m], shvb = shB[m];
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        auto convert = [&](int rs, int s, bool same, int nvalid, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            float va[KPT], vb[KPT];
            const uint32_t rawa = sbase + OFF_RAW + (rs * 2 + 0) * RAW_TILE + (uint32_t)(kh * KPT * TM + m) * 4u;
#pragma unroll
            for (int e = 0; e < KPT; ++e) va[e] = lds_f32(rawa + (uint32_t)e * TM * 4u);
            if (!same) {
#pragma unroll
                for (int e = 0; e < KPT; ++e) vb[e] = lds_f32(rawa + RAW_TILE + (uint32_t)e * TM * 4u);
            }
            // A: shift, split, straight into tensor memory (columns = k)
            float ah[KPT], al[KPT];
#pragma unroll
            for (int e = 0; e < KPT; ++e) {
                const float x = (FULL || kh * KPT + e < nvalid) ? __fsub_rn(va[e], shva) : 0.f;
                ah[e] = tf32_rn(x);
                al[e] = tf32_rn(__fsub_rn(x, ah[e]));
            }
            const uint32_t ta = lane_addr + TM_A + (uint32_t)(s * 2 * KB + kh * KPT);
            tmem_st(ta, ah);
            tmem_st(ta + KB, al);
            // B: K-major 128B-swizzled shared memory (the same values when the tile is on the diagonal)
            const uint32_t hi = sbase + OFF_OPS + s * 2 * OP_TILE + (uint32_t)m * 128u;
#pragma unroll
            for (int q4 = 0; q4 < KPT / 4; ++q4) {
                const int q = kh * (KPT / 4) + q4;
                float h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (same) {
                        h[e] = ah[4 * q4 + e];
                        l[e] = al[4 * q4 + e];
                    } else {
                        const float x = (FULL || 4 * q + e < nvalid) ? __fsub_rn(vb[4 * q4 + e], shvb) : 0.f;
                        h[e] = tf32_rn(x);
                        l[e] = tf32_rn(__fsub_rn(x, h[e]));
                    }
                }
                const uint32_t off = (uint32_t)((q ^ (m & 7)) << 4);  // Swizzle<3,4,3>: 16B chunk ^= row & 7
                sts_v4(hi + off, h[0], h[1], h[2], h[3]);
                sts_v4(hi + OP_TILE + off, l[0], l[1], l[2], l[3]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        // drain: this thread owns row (quad*32 + lane) x 64 columns of the tile
        // drain: this thread owns row m x CPT columns (slice kh) of the tile
        float accv[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) accv[e] = 0.f;
        auto drain = [&](int c) {
            TC_ACC(11, mbar_wait(bar(ACC_FULL + (c & 1)), (uint32_t)((c >> 1) & 1)));  // sub-chunk c is the (c>>1)-th use
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < CPT / 32; ++g) {
                uint32_t r[32];
                tmem_ld32(lane_addr + TM_ACC + (uint32_t)((c & 1) * TN + kh * CPT + g * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 32; ++e) accv[g * 32 + e] = __fadd_rn(accv[g * 32 + e], __uint_as_float(r[e]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        };
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES, rs = kb % RSTAGES;
            const uint32_t ph = (kb / STAGES) & 1, rph = (kb / RSTAGES) & 1;
            const int c = kb / CHUNK_KB, kk = kb - c * CHUNK_KB;
            TC_ACC(10, mbar_wait(bar(RAW_FULL + rs), rph));     // raw boxes landed
            if (kb == 0) TC_T(2);
            TC_ACC(12, mbar_wait(bar(OPS_EMPTY + s), ph ^ 1));  // operand stage no longer read by the tensor core
            const int64_t row0 = r_begin + (int64_t)kb * KB;
            const int nvalid = (int)((r_end - row0) < KB ? (r_end - row0) : KB);
            if (TC_ABLATE(1)) { }
            else if (nvalid == KB) convert(rs, s, diag, KB, std::true_type{});
            else convert(rs, s, diag, nvalid, std::false_type{});
            if (!TC_ABLATE(8)) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy (UMMA)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // tcgen05.st (A) ordered before the arrive
            }
            // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise in shared memory
            // (~500 cycles per k-block, measured with profiles/tc_timeline.py)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(OPS_FULL + s));    // operands ready
                mbar_arrive(bar(RAW_EMPTY + rs));  // raw stage free
            }
            // the previous sub-chunk's pair is complete by now (the tensor core is at most two k-blocks behind):
            // drain it while the MMAs of this sub-chunk run on the other pair
            if (c > 0) {
                const int kbc = nkb - c * CHUNK_KB;
                const int dpoint = kbc > 2 ? 2 : kbc - 1;
                if (kk == dpoint && !TC_ABLATE(16)) drain(c - 1);
            }
        }
        TC_T(5);
        drain(nchunk - 1);
        // ---- fp32 partial tile
        float *dst = P.partial + ((size_t)blockIdx.y * P.ntiles + blockIdx.x) * (size_t)(TM * TN) + (size_t)m * TN + kh * CPT;
#pragma unroll
        for (int e = 0; e < CPT; e += 4)
            *reinterpret_cast<float4 *>(dst + e) = make_float4(accv[e], accv[e + 1], accv[e + 2], accv[e + 3]);
    }
    if (threadIdx.x == 0) TC_T(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
    if (threadIdx.x == T_MMA) TC_T(7);
}

// ------------------------------------------------------------------ small kernels around it
constexpr int RS = 32;  // row splits of the column-sum passes (parallelism; partials combined in fixed order)

// partial column sums over a slice of the rows: raw values (shift == nullptr) or fl32(x - shift[col]);
// optionally sums of squares; fp64, fixed order
__global__ void __launch_bounds__(256)
colsum_part(const float *__restrict__ X, int64_t ld, int ncols, int64_t nrows, const float *__restrict__ shift,
            double *__restrict__ part, double *__restrict__ part_sq) {
    __shared__ double s1[8][33], s2[8][33];
    const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const int64_t per = (nrows + RS - 1) / RS;
    const int64_t r0 = (int64_t)blockIdx.y * per;
    const int64_t r1 = r0 + per < nrows ? r0 + per : nrows;
    double a = 0.0, q = 0.0;
    if (col < ncols) {
        const float sh = shift ? shift[col] : 0.f;
        for (int64_t r = r0 + rg; r < r1; r += 8) {
            const double v = (double)__fsub_rn(__ldg(X + r * ld + col), sh);
            a += v;
            q = fma(v, v, q);
        }
    }
    s1[rg][cx] = a;
    s2[rg][cx] = q;
    __syncthreads();
    if (rg == 0 && col < ncols) {
        double t = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t += s1[k][cx]; t2 += s2[k][cx]; }
        part[(size_t)blockIdx.y * ncols + col] = t;
        if (part_sq) part_sq[(size_t)blockIdx.y * ncols + col] = t2;
    }
}

// out[j] = sum_rs part[rs][j]; optionally shift32[j] = fl32(out[j] / N)
__global__ void colsum_finish(const double *__restrict__ part, const double *__restrict__ part_sq, int ncols, double invN,
                              double *__restrict__ out, double *__restrict__ out_sq, float *__restrict__ shift_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double t = 0.0, t2 = 0.0;
    for (int r = 0; r < RS; ++r) {
        t += part[(size_t)r * ncols + j];
        if (part_sq) t2 += part_sq[(size_t)r * ncols + j];
    }
    if (out) out[j] = t;
    if (out_sq) out_sq[j] = t2;
    if (shift_out) shift_out[j] = (float)(t * invN);
}

// Sums the fp32 partial tiles over the row chunks in fp64 and undoes the shift:
//   C[i,j] = sum_c P_c[i,j] + uA_i TB_j + TA_i uB_j + N uA_i uB_j ,  u = (double)shift32 - bias
__global__ void __launch_bounds__(256)
reduce_tc(const float *__restrict__ partial, int nchunks, int ntiles, int tile0, int tiles_cols, int sym, int tk,
          const float *__restrict__ shiftA, const double *__restrict__ TA, const float *__restrict__ shiftB,
          const float *__restrict__ biasB, const double *__restrict__ TB, const double *__restrict__ diag_sq, double Nd,
          int M, int Nn, double *__restrict__ C, int64_t ldc) {
    int l = blockIdx.x, ti, tj;
    if (sym) {
        ti = 0;
        while (l >= tk - ti) { l -= tk - ti; ++ti; }
        tj = ti + l;
    } else {
        ti = l / tiles_cols;
        tj = l - ti * tiles_cols;
    }
    const size_t tile_elems = (size_t)TM * TN;
    const float *p0 = partial + (size_t)(tile0 + blockIdx.x) * tile_elems;
    const int e0 = blockIdx.y * (TM * TN / 4);  // four CTAs per tile
    for (int e = e0 + threadIdx.x; e < e0 + TM * TN / 4; e += 256) {
        const int i = ti * TM + e / TN, j = tj * TN + e % TN;
        if (i >= M || j >= Nn) continue;
        double s = 0.0;
        // inside a diagonal tile the tensor core produced both (i,j) and (j,i) with different rounding:
        // read the upper one for both so that G is bitwise symmetric
        const int er = (sym && ti == tj && (e / TN) > (e % TN)) ? (e % TN) * TN + (e / TN) : e;
        for (int c = 0; c < nchunks; ++c) s += (double)p0[(size_t)c * ntiles * tile_elems + er];
        // the diagonal is a sum of squares (no cancellation: worst case for the truncating fp32
        // accumulation) and costs one fused pass to get exactly -- use the fp64 value
        if (sym && i == j) s = diag_sq[i];
        const double ua = (double)shiftA[i];
        const double ub = (double)shiftB[j] - (biasB ? (double)biasB[j] : 0.0);
        s += ua * TB[j] + TA[i] * ub + Nd * ua * ub;
        C[(int64_t)i * ldc + j] = s;
    }
}

// out[j] = T[j] + N * ((double)shift[j] - bias[j])
__global__ void finish_sums(const double *__restrict__ T, const float *__restrict__ shift, const float *__restrict__ bias,
                            double Nd, int ncols, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncols) out[j] = T[j] + Nd * ((double)shift[j] - (bias ? (double)bias[j] : 0.0));
}

__global__ void __launch_bounds__(256)
mirror_upper_tiles_tc(double *__restrict__ C, int M, int64_t ldc) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x, by = blockIdx.y;
    if ((by * 32) / TM >= (bx * 32) / TN) return;  // strictly-upper 128-tiles only
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = by * 32 + r, j = bx * 32 + tx;
        if (i < M && j < M) t[r][tx] = C[(int64_t)i * ldc + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bx * 32 + r, i = by * 32 + tx;
        if (i < M && j < M) C[(int64_t)j * ldc + i] = t[tx][r];
    }
}

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(cp_handle_t h, CUtensorMap *map, const float *base, int64_t rows, int cols, int64_t ld) {
    if (!h->tmap_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) CP_FAIL(CP_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        h->tmap_encode = fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TM, (cuuint32_t)KB};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((encode_fn_t)h->tmap_encode)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box,
                                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CP_FAIL(CP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CP_OK;
}

}  // namespace

#ifdef CP_TC_TIMING
extern "C" int cp_debug_tc_times(long long *host_out) {  // 64 x 16 clock64 stamps of the last gram_tc_kernel launch
    return (int)cudaMemcpyFromSymbol(host_out, cp_tc_times, sizeof(long long) * 64 * 16);
}
#endif

bool cp_gram_tc_eligible(const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype, int n, int64_t ldy,
                         const int32_t *rows, bool wantB) {
    if (rows != nullptr || N < 64 || K < 64) return false;
    if ((((uintptr_t)X) & 15) || (ldx % 4)) return false;  // TMA: 16-byte aligned base and row pitch
    if (wantB && (y_dtype != CP_F32 || (((uintptr_t)Yraw) & 15) || (ldy % 4) || n < 1)) return false;
    return true;
}

int cp_gram_tc(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype, int n,
               int64_t ldy, const float *y_bias, const int32_t *rows, int nrows, double *G, double *Bxy, double *sx,
               double *sy, double *yy, cudaStream_t stream) {
    const bool wantB = Bxy != nullptr;
    if (yy != nullptr || !cp_gram_tc_eligible(X, N, K, ldx, Yraw, y_dtype, n, ldy, rows, wantB || sy != nullptr))
        return cp_gram_fp64_products(h, X, N, K, ldx, Yraw, y_dtype, n, ldy, y_bias, rows, nrows, G, Bxy, sx, sy, yy, stream);
    const float *Y = (const float *)Yraw;
    const bool haveY = (Bxy != nullptr || sy != nullptr);
    const int tk = cp_cdiv(K, TM);
    const int tiles_sym = G ? tk * (tk + 1) / 2 : 0;
    const int tnb = wantB ? cp_cdiv(n, TN) : 0;
    const int ntiles = tiles_sym + tk * tnb;
    // Row splits: every CTA pays a fixed prologue/epilogue (~10k cycles) and ~1.4k cycles per 32-row k-block, and the
    // grid runs in waves of num_sms CTAs (one per SM: 224 KB of shared memory each).  Pick the number of splits
    // that minimises waves x CTA time; splits are multiples of the 256-row sub-chunk.
    // A CTA adds its sub-chunks in fp32 registers (round to nearest): at most MAX_SPLIT_ROWS rows (32 sub-chunks)
    // per split keeps that error below 4e-7 of the partial sum whatever N is.
    constexpr int64_t SUB = CHUNK_KB * KB, MAX_SPLIT_ROWS = 32 * SUB;
    const int ns_min = (int)cp_cdiv(N, MAX_SPLIT_ROWS);
    int nchunks = ns_min, rpc = (int)(cp_cdiv(cp_cdiv(N, ns_min), SUB) * SUB);
    if (ntiles > 0) {
        double best = 1e300;
        const int max_ns = (int)cp_cdiv(N, SUB);
        for (int ns = ns_min; ns <= max_ns && ns < ns_min + 64; ++ns) {
            const int64_t r = cp_cdiv(cp_cdiv(N, ns), SUB) * SUB;
            const int ns_eff = (int)cp_cdiv(N, r);
            const double waves = (double)cp_cdiv((int64_t)ntiles * ns_eff, h->num_sms);
            const double cost = waves * (10000.0 + 1400.0 * (double)(r / KB));
            if (cost < best * 0.97) {  // prefer fewer splits (less partial traffic) unless clearly better
                best = cost;
                nchunks = ns_eff;
                rpc = (int)r;
            }
        }
    }

    const size_t part_bytes = (size_t)nchunks * ntiles * TM * TN * sizeof(float);
    const size_t need = cp_carver::need(part_bytes, 1) + 2 * cp_carver::need(K, 8) + cp_carver::need(n > 0 ? n : 1, 8) +
                        2 * cp_carver::need((size_t)RS * (K > n ? K : n), 8) +
                        cp_carver::need(K, 4) + cp_carver::need(n > 0 ? n : 1, 4);
    void *ws = nullptr;
    int rc = cp_ws_reserve(h, need, &ws);
    if (rc) return rc;
    cp_carver cv(ws);
    float *partial = cv.take<float>((size_t)nchunks * ntiles * TM * TN);
    double *TX = cv.take<double>(K), *SQX = cv.take<double>(K);
    double *TY = cv.take<double>(n > 0 ? n : 1);
    const int wmax = K > n ? K : n;
    double *cpart = cv.take<double>((size_t)RS * wmax), *cpart_sq = cv.take<double>((size_t)RS * wmax);
    float *shX = cv.take<float>(K), *shY = cv.take<float>(n > 0 ? n : 1);
    const double invN = 1.0 / (double)N, Nd = (double)N;

    colsum_part<<<dim3(cp_cdiv(K, 32), RS), 256, 0, stream>>>(X, ldx, K, N, nullptr, cpart, nullptr);
    CP_CHECK_LAUNCH();
    colsum_finish<<<cp_cdiv(K, 256), 256, 0, stream>>>(cpart, nullptr, K, invN, nullptr, nullptr, shX);
    CP_CHECK_LAUNCH();
    colsum_part<<<dim3(cp_cdiv(K, 32), RS), 256, 0, stream>>>(X, ldx, K, N, shX, cpart, cpart_sq);
    CP_CHECK_LAUNCH();
    colsum_finish<<<cp_cdiv(K, 256), 256, 0, stream>>>(cpart, cpart_sq, K, invN, TX, SQX, nullptr);
    CP_CHECK_LAUNCH();
    if (haveY) {
        colsum_part<<<dim3(cp_cdiv(n, 32), RS), 256, 0, stream>>>(Y, ldy, n, N, nullptr, cpart, nullptr);
        CP_CHECK_LAUNCH();
        colsum_finish<<<cp_cdiv(n, 256), 256, 0, stream>>>(cpart, nullptr, n, invN, nullptr, nullptr, shY);
        CP_CHECK_LAUNCH();
        colsum_part<<<dim3(cp_cdiv(n, 32), RS), 256, 0, stream>>>(Y, ldy, n, N, shY, cpart, nullptr);
        CP_CHECK_LAUNCH();
        colsum_finish<<<cp_cdiv(n, 256), 256, 0, stream>>>(cpart, nullptr, n, invN, TY, nullptr, nullptr);
        CP_CHECK_LAUNCH();
    }
    if (ntiles > 0) {
        CUtensorMap mapA, mapB;
        rc = make_map(h, &mapA, X, N, K, ldx);
        if (rc) return rc;
        if (wantB) {
            rc = make_map(h, &mapB, Y, N, n, ldy);
            if (rc) return rc;
        } else {
            mapB = mapA;
        }
        TcParams P{};
        P.shiftA = shX; P.shiftB = shY; P.partial = partial; P.K = K; P.nB = n; P.N = N; P.rows_per_chunk = rpc;
        P.tiles_sym = tiles_sym; P.tk = tk; P.tnb = tnb; P.ntiles = ntiles;
        static cp_per_device_flag configured;
        if (bool *done = configured.slot(); !*done) {
            CP_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            *done = true;
        }
        // one launch: X'X upper tiles first, then the X'Y tiles (B operand from the Y map)
        gram_tc_kernel<<<dim3(ntiles, nchunks), NTHREADS, SMEM_BYTES, stream>>>(mapA, mapB, P);
        CP_CHECK_LAUNCH();
        if (tiles_sym > 0) {
            reduce_tc<<<dim3(tiles_sym, 4), 256, 0, stream>>>(partial, nchunks, ntiles, 0, 0, 1, tk, shX, TX, shX, nullptr, TX, SQX, Nd,
                                                    K, K, G, K);
            CP_CHECK_LAUNCH();
            if (K > TM) {
                const int nb32 = cp_cdiv(K, 32);
                mirror_upper_tiles_tc<<<dim3(nb32, nb32), 256, 0, stream>>>(G, K, K);
                CP_CHECK_LAUNCH();
            }
        }
        if (wantB) {
            reduce_tc<<<dim3(tk * tnb, 4), 256, 0, stream>>>(partial, nchunks, ntiles, tiles_sym, tnb, 0, tk, shX, TX, shY, y_bias, TY,
                                                   nullptr, Nd, K, n, Bxy, n);
            CP_CHECK_LAUNCH();
        }
    }
    if (sx) {
        finish_sums<<<cp_cdiv(K, 256), 256, 0, stream>>>(TX, shX, nullptr, Nd, K, sx);
        CP_CHECK_LAUNCH();
    }
    if (sy) {
        finish_sums<<<cp_cdiv(n, 256), 256, 0, stream>>>(TY, shY, y_bias, Nd, n, sy);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}
