// fp64 GEMM with asynchronous operand staging -- the workhorse of the least-squares solver (ls.cu): Cholesky
// trailing updates, panel solves through inverted blocks, forward / backward substitutions.
//
//   C[m, nn] = alpha * sum_r a(m, r) * b(nn, r) + beta * C[m, nn]                    (all fp64)
//   a(m, r)  = A[m * lda + r]                                  (reduction index contiguous)
//   b(nn, r) = B_NC ? B[r * ldb + nn] : B[nn * ldb + r]
//
// gemm_f64.cuh stages its operands global -> registers -> shared memory (it has to: it widens fp32 data and
// gathers rows); for fp64 operands that detour costs 16 conflicting STS per thread and stage and exposes the
// global latency once per 16-deep stage.  Here 16-byte cp.async copies land the tiles in shared memory directly,
// NS stages deep, in the layout the MMA fragments want:
//   r-contiguous operand  ->  [tile row][k]   leading dimension BK + 4 doubles  (4 mod 16: the m8n8k4 fragment
//   x-contiguous operand  ->  [k][tile col]   leading dimension T + 4 doubles    loads of a half-warp hit 16 banks)
// Tails (rows beyond the matrix, reduction not a multiple of the stage) are zero-filled by the copy itself
// (src-size operand).  Inner loop: mma.sync.m8n8k4.f64, 2 x 4 warps, warp tile T/2 x T/4.  Tile T = 128
// (throughput) or 64 (latency: everything on a dependency chain).  Requires 16-byte aligned operands and even
// leading dimensions (the callers fall back to gemm_f64.cuh / gemm_small.cuh otherwise).  Bound: FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpasync {

constexpr int BK = 16, NT = 256;
constexpr int LDK = BK + 4;  // [row][k] tiles

enum TileMode { TILES_ALL = 0, TILES_LOWER = 2 };

struct Args {
    const double *A;
    int64_t lda;
    const double *B;
    int64_t ldb;
    double *C;
    int64_t ldc;
    int M, Nn, R;
    double alpha, beta;
    int tile_mode;
    int max_ctas;  // > 0: at most that many CTAs walk the tiles
};

template <int T>
struct Cfg {
    static constexpr int NS = T == 128 ? 4 : 3;                 // stages
    static constexpr int LDX = T + 4;                           // [k][col] tiles
    static constexpr int A_ELEMS = T * LDK;                     // doubles per A stage
    static constexpr int B_ELEMS_RC = T * LDK, B_ELEMS_XC = BK * LDX;
    static constexpr int WM = T / 2, WN = T / 4;                // warp tile
    static constexpr int MI = WM / 8, NJ = WN / 8;              // MMA tiles per warp
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// r-contiguous operand: T rows x BK doubles = T * 8 chunks of 16 bytes
template <int T>
__device__ __forceinline__ void load_rc(double *S, const double *P, int64_t ld, int x0, int xlim, int r0, int rlim) {
#pragma unroll
    for (int q = threadIdx.x; q < T * 8; q += NT) {
        const int row = q >> 3, ch = q & 7;
        const int x = x0 + row, r = r0 + ch * 2;
        int bytes = 0;
        if (x < xlim && r < rlim) bytes = (rlim - r >= 2) ? 16 : 8;
        const double *src = bytes ? P + (int64_t)x * ld + r : P;
        cp_async16(S + row * LDK + ch * 2, src, bytes);
    }
}
// x-contiguous operand: BK rows (k) x T doubles = BK * T / 2 chunks
template <int T>
__device__ __forceinline__ void load_xc(double *S, const double *P, int64_t ld, int x0, int xlim, int r0, int rlim) {
    constexpr int CPR = T / 2;  // chunks per k-row
#pragma unroll
    for (int q = threadIdx.x; q < BK * CPR; q += NT) {
        const int kr = q / CPR, ch = q - kr * CPR;
        const int r = r0 + kr, x = x0 + ch * 2;
        int bytes = 0;
        if (r < rlim && x < xlim) bytes = (xlim - x >= 2) ? 16 : 8;
        const double *src = bytes ? P + (int64_t)r * ld + x : P;
        cp_async16(S + kr * Cfg<T>::LDX + ch * 2, src, bytes);
    }
}

__device__ __forceinline__ int num_tiles_dev(int tm, int tn, int mode) {
    return mode == TILES_LOWER ? tn * tm - tn * (tn - 1) / 2 : tm * tn;
}

template <int T, bool B_NC>
__global__ void __launch_bounds__(NT, T == 128 ? 1 : 2) gemm_async_kernel(const Args g) {
    using C_ = Cfg<T>;
    extern __shared__ __align__(16) double sm_async[];
    constexpr int B_ELEMS = B_NC ? C_::B_ELEMS_XC : C_::B_ELEMS_RC;
    constexpr int STAGE = C_::A_ELEMS + B_ELEMS;
    const int lane = threadIdx.x & 31, wm = threadIdx.x >> 7, wn = (threadIdx.x >> 5) & 3;
    const int tiles_m = (g.M + T - 1) / T, tiles_n = (g.Nn + T - 1) / T;
    const int ntiles = num_tiles_dev(tiles_m, tiles_n, g.tile_mode);
    const int nk = (g.R + BK - 1) / BK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int l = tile, ti, tj;
        if (g.tile_mode == TILES_LOWER) {  // column-tile major: for tj, row tiles ti = tj .. tiles_m-1
            tj = 0;
            while (l >= tiles_m - tj) { l -= tiles_m - tj; ++tj; }
            ti = tj + l;
        } else {
            ti = l / tiles_n;
            tj = l - ti * tiles_n;
        }
        const int m0 = ti * T, n0 = tj * T;
        auto issue = [&](int kb) {
            if (kb < nk) {
                double *S = sm_async + (size_t)(kb % C_::NS) * STAGE;
                load_rc<T>(S, g.A, g.lda, m0, g.M, kb * BK, g.R);
                if constexpr (B_NC) load_xc<T>(S + C_::A_ELEMS, g.B, g.ldb, n0, g.Nn, kb * BK, g.R);
                else load_rc<T>(S + C_::A_ELEMS, g.B, g.ldb, n0, g.Nn, kb * BK, g.R);
            }
            cp_commit();  // one group per slot, empty or not: the wait below counts groups
        };
        double acc[C_::MI][2 * C_::NJ];
#pragma unroll
        for (int i = 0; i < C_::MI; ++i)
#pragma unroll
            for (int j = 0; j < 2 * C_::NJ; ++j) acc[i][j] = 0.0;
#pragma unroll
        for (int s = 0; s < C_::NS - 1; ++s) issue(s);
        for (int kb = 0; kb < nk; ++kb) {
            cp_wait<C_::NS - 2>();   // this thread's copies of stage kb have landed ...
            __syncthreads();         // ... and everybody's; everybody is also done with the slot refilled next
            issue(kb + C_::NS - 1);
            const double *a_s = sm_async + (size_t)(kb % C_::NS) * STAGE;
            const double *b_s = a_s + C_::A_ELEMS;
            const double *ap = a_s + (wm * C_::WM + (lane >> 2)) * LDK + (lane & 3);
#pragma unroll
            for (int k4 = 0; k4 < BK; k4 += 4) {
                double af[C_::MI], bf[C_::NJ];
#pragma unroll
                for (int i = 0; i < C_::MI; ++i) af[i] = ap[8 * i * LDK + k4];
#pragma unroll
                for (int j = 0; j < C_::NJ; ++j) {
                    if constexpr (B_NC) bf[j] = b_s[(k4 + (lane & 3)) * C_::LDX + wn * C_::WN + 8 * j + (lane >> 2)];
                    else bf[j] = b_s[(wn * C_::WN + 8 * j + (lane >> 2)) * LDK + k4 + (lane & 3)];
                }
#pragma unroll
                for (int i = 0; i < C_::MI; ++i)
#pragma unroll
                    for (int j = 0; j < C_::NJ; ++j) dmma884(acc[i][2 * j], acc[i][2 * j + 1], af[i], bf[j]);
            }
        }
        cp_wait<0>();
        __syncthreads();  // the next tile's prologue refills the slots
        // ---- epilogue: acc[i][2j + e] -> row m0 + wm*WM + 8i + (lane >> 2), column n0 + wn*WN + 8j + 2 (lane & 3) + e;
        // every old C value of a group of rows is loaded before any is stored (one memory latency per group)
        const bool rmw = g.beta != 0.0;
        const bool cvec = ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 2 == 0);
        constexpr int GRP = C_::MI >= 4 ? 4 : C_::MI;
#pragma unroll
        for (int ig = 0; ig < C_::MI; ig += GRP) {
            double old[GRP][2 * C_::NJ];
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int m = m0 + wm * C_::WM + 8 * (ig + u) + (lane >> 2);
#pragma unroll
                for (int j = 0; j < C_::NJ; ++j) {
                    const int nn = n0 + wn * C_::WN + 8 * j + 2 * (lane & 3);
                    old[u][2 * j] = old[u][2 * j + 1] = 0.0;
                    if (rmw && m < g.M) {
                        const double *p = g.C + (int64_t)m * g.ldc + nn;
                        if (cvec && nn + 1 < g.Nn) {
                            const double2 v = *reinterpret_cast<const double2 *>(p);
                            old[u][2 * j] = v.x;
                            old[u][2 * j + 1] = v.y;
                        } else {
                            if (nn < g.Nn) old[u][2 * j] = p[0];
                            if (nn + 1 < g.Nn) old[u][2 * j + 1] = p[1];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int m = m0 + wm * C_::WM + 8 * (ig + u) + (lane >> 2);
                if (m >= g.M) continue;
#pragma unroll
                for (int j = 0; j < C_::NJ; ++j) {
                    const int nn = n0 + wn * C_::WN + 8 * j + 2 * (lane & 3);
                    double v0 = acc[ig + u][2 * j] * g.alpha, v1 = acc[ig + u][2 * j + 1] * g.alpha;
                    if (rmw) {
                        v0 = fma(g.beta, old[u][2 * j], v0);
                        v1 = fma(g.beta, old[u][2 * j + 1], v1);
                    }
                    double *p = g.C + (int64_t)m * g.ldc + nn;
                    if (cvec && nn + 1 < g.Nn) {
                        *reinterpret_cast<double2 *>(p) = make_double2(v0, v1);
                    } else {
                        if (nn < g.Nn) p[0] = v0;
                        if (nn + 1 < g.Nn) p[1] = v1;
                    }
                }
            }
        }
    }
}

template <int T>
inline int num_tiles(int M, int Nn, int mode) {
    const int tm = (M + T - 1) / T, tn = (Nn + T - 1) / T;
    if (mode == TILES_LOWER) return tn * tm - tn * (tn - 1) / 2;  // requires tm >= tn
    return tm * tn;
}

inline bool eligible(const Args &g) {
    return ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) &&
           (g.lda % 2 == 0) && (g.ldb % 2 == 0);
}

template <int T, bool B_NC>
inline cudaError_t launch(const Args &g, cudaStream_t stream) {
    using C_ = Cfg<T>;
    if (g.M <= 0 || g.Nn <= 0) return cudaSuccess;
    constexpr size_t smem = (size_t)C_::NS * (C_::A_ELEMS + (B_NC ? C_::B_ELEMS_XC : C_::B_ELEMS_RC)) * sizeof(double);
    auto kern = gemm_async_kernel<T, B_NC>;
    static bool configured[64] = {};  // per instantiation and per device (the attribute is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    bool &done = configured[dev >= 0 && dev < 64 ? dev : 0];
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        done = true;
    }
    unsigned grid = (unsigned)num_tiles<T>(g.M, g.Nn, g.tile_mode);
    if (g.max_ctas > 0 && grid > (unsigned)g.max_ctas) grid = (unsigned)g.max_ctas;
    kern<<<grid, NT, smem, stream>>>(g);
    return cudaGetLastError();
}

}  // namespace cpasync
