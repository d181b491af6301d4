// Small-tile fp64 GEMM for the latency-critical steps of the blocked Cholesky (ls.cu):
// panel triangular solves through inverted diagonal blocks, the update of the NEXT panel's
// columns, and the block steps of the forward / backward substitutions.  Those products have a
// short inner dimension (128) and few output elements, so the 128 x 128 tiles of gemm_f64.cuh
// leave most SMs idle; 64 x 64 tiles put 4x as many CTAs on the same work.
//
//   C[m, nn] = alpha * sum_r a(m, r) * b(nn, r)  (+ beta * C[m, nn])
//
//   a(m, r)  = A[m * lda + r]                               (reduction index contiguous)
//   b(nn, r) = B_NC ? B[r * ldb + nn] : B[nn * ldb + r]
//
// 256 threads, reduction staged 16 deep through double-buffered shared memory with register prefetch; inner loop
// on mma.sync.m8n8k4.f64 (warp tile 32 x 16: 4 x 2 MMA tiles, 6 conflict-free LDS.64 per 8 MMAs) or, with
// CPB200_GEMM=dfma, a 4 x 4 register micro-tile per thread.  Bound: FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace cpsmall {

constexpr int TM = 64, TN = 64, BK = 16, NT = 256;
constexpr int LDS_ = TM + 4;  // padded leading dimension of a staged tile (doubles): 4 mod 16 (MMA fragment loads)

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
    int a_vec, b_vec;  // 16-byte vector loads allowed (alignment checked by the host)
};

// r-contiguous operand: thread t fetches row x = t / 4, r = (t % 4) * 4 .. +4
__device__ __forceinline__ void fetch_rc(const double *P, int64_t ld, int x0, int xlim, int r0, int rlim, bool vec,
                                         double out[4]) {
    const int t = threadIdx.x;
    const int x = x0 + (t >> 2);
    const int r = r0 + (t & 3) * 4;
    if (x >= xlim || r >= rlim) {
        out[0] = out[1] = out[2] = out[3] = 0.0;
        return;
    }
    const double *p = P + (int64_t)x * ld + r;
    if (vec && r + 4 <= rlim) {
        const double2 v0 = __ldg(reinterpret_cast<const double2 *>(p));
        const double2 v1 = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        out[0] = v0.x; out[1] = v0.y; out[2] = v1.x; out[3] = v1.y;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = (r + i < rlim) ? __ldg(p + i) : 0.0;
    }
}
__device__ __forceinline__ void stage_rc(double *S, const double v[4]) {
    const int t = threadIdx.x;
    double *d = S + ((t & 3) * 4) * LDS_ + (t >> 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i * LDS_] = v[i];
}
// x-contiguous operand (element (x, r) = P[r * ld + x]): thread t fetches r = t / 16, x = (t % 16) * 4 .. +4
__device__ __forceinline__ void fetch_xc(const double *P, int64_t ld, int x0, int xlim, int r0, int rlim, bool vec,
                                         double out[4]) {
    const int t = threadIdx.x;
    const int r = r0 + (t >> 4);
    const int x = x0 + (t & 15) * 4;
    if (r >= rlim || x >= xlim) {
        out[0] = out[1] = out[2] = out[3] = 0.0;
        return;
    }
    const double *p = P + (int64_t)r * ld + x;
    if (vec && x + 4 <= xlim) {
        const double2 v0 = __ldg(reinterpret_cast<const double2 *>(p));
        const double2 v1 = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        out[0] = v0.x; out[1] = v0.y; out[2] = v1.x; out[3] = v1.y;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = (x + i < xlim) ? __ldg(p + i) : 0.0;
    }
}
__device__ __forceinline__ void stage_xc(double *S, const double v[4]) {
    const int t = threadIdx.x;
    double *d = S + (t >> 4) * LDS_ + (t & 15) * 4;
    *reinterpret_cast<double2 *>(d) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(d + 2) = make_double2(v[2], v[3]);
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool B_NC, bool DMMA>
__global__ void __launch_bounds__(NT) gemm_small_kernel(const Args g) {
    __shared__ __align__(16) double As[2][BK * LDS_];
    __shared__ __align__(16) double Bs[2][BK * LDS_];
    const int tiles_m = (g.M + TM - 1) / TM, tiles_n = (g.Nn + TN - 1) / TN;
    int l = blockIdx.x, ti, tj;
    if (g.tile_mode == TILES_LOWER) {  // column-tile major: for tj, row tiles ti = tj .. tiles_m-1
        tj = 0;
        while (l >= tiles_m - tj) { l -= tiles_m - tj; ++tj; }
        ti = tj + l;
    } else {
        ti = l / tiles_n;
        tj = l - ti * tiles_n;
    }
    const int m0 = ti * TM, n0 = tj * TN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, wm = threadIdx.x >> 7, wn = (threadIdx.x >> 5) & 3;  // MMA: 2 x 4 warps
    double ra[4], rb[4];
    auto fetch = [&](int r0) {
        fetch_rc(g.A, g.lda, m0, g.M, r0, g.R, g.a_vec, ra);
        if constexpr (B_NC) fetch_xc(g.B, g.ldb, n0, g.Nn, r0, g.R, g.b_vec, rb);
        else fetch_rc(g.B, g.ldb, n0, g.Nn, r0, g.R, g.b_vec, rb);
    };
    auto stage = [&](int buf) {
        stage_rc(As[buf], ra);
        if constexpr (B_NC) stage_xc(Bs[buf], rb); else stage_rc(Bs[buf], rb);
    };
    int buf = 0;
    if (g.R > 0) {
        fetch(0);
        stage(0);
    }
    __syncthreads();
    for (int r0 = 0; r0 < g.R; r0 += BK) {
        const bool has_next = r0 + BK < g.R;
        if (has_next) fetch(r0 + BK);
        const double *a_s = As[buf], *b_s = Bs[buf];
        if constexpr (DMMA) {
            // acc[i][2j + e]: rows wm*32 + 8i + (lane >> 2), columns wn*16 + 8j + 2 (lane & 3) + e   (j < 2)
            const double *ap = a_s + (lane & 3) * LDS_ + wm * 32 + (lane >> 2);
            const double *bp = b_s + (lane & 3) * LDS_ + wn * 16 + (lane >> 2);
#pragma unroll
            for (int k4 = 0; k4 < BK; k4 += 4) {
                double af[4], bf[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = ap[k4 * LDS_ + 8 * i];
#pragma unroll
                for (int j = 0; j < 2; ++j) bf[j] = bp[k4 * LDS_ + 8 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma884(acc[i][2 * j], acc[i][2 * j + 1], af[i], bf[j]);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                double a[4], b[4];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double2 av = *reinterpret_cast<const double2 *>(a_s + kk * LDS_ + q * 32 + ty * 2);
                    const double2 bv = *reinterpret_cast<const double2 *>(b_s + kk * LDS_ + q * 32 + tx * 2);
                    a[2 * q] = av.x; a[2 * q + 1] = av.y;
                    b[2 * q] = bv.x; b[2 * q + 1] = bv.y;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
        if (has_next) stage(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    // epilogue: every C value of the thread is loaded before any is stored (one memory latency, not sixteen)
    const bool rmw = g.beta != 0.0;
    const bool cvec = ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (g.ldc % 2 == 0);
    auto erow = [&](int i) { return DMMA ? wm * 32 + 8 * i + (lane >> 2) : (i >> 1) * 32 + ty * 2 + (i & 1); };
    auto ecol = [&](int q) { return DMMA ? wn * 16 + 8 * q + 2 * (lane & 3) : q * 32 + tx * 2; };
    double old[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + erow(i);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int nn = n0 + ecol(q);
            old[i][2 * q] = old[i][2 * q + 1] = 0.0;
            if (rmw && m < g.M) {
                const double *p = g.C + (int64_t)m * g.ldc + nn;
                if (cvec && nn + 1 < g.Nn) {
                    const double2 v = *reinterpret_cast<const double2 *>(p);
                    old[i][2 * q] = v.x;
                    old[i][2 * q + 1] = v.y;
                } else {
                    if (nn < g.Nn) old[i][2 * q] = p[0];
                    if (nn + 1 < g.Nn) old[i][2 * q + 1] = p[1];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + erow(i);
        if (m >= g.M) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int nn = n0 + ecol(q);
            double v0 = acc[i][2 * q] * g.alpha, v1 = acc[i][2 * q + 1] * g.alpha;
            if (rmw) {
                v0 = fma(g.beta, old[i][2 * q], v0);
                v1 = fma(g.beta, old[i][2 * q + 1], v1);
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

inline int num_tiles(int M, int Nn, int mode) {
    const int tm = (M + TM - 1) / TM, tn = (Nn + TN - 1) / TN;
    if (mode == TILES_LOWER) return tn * tm - tn * (tn - 1) / 2;  // requires tm >= tn
    return tm * tn;
}

template <bool B_NC>
inline cudaError_t launch(const Args &g, cudaStream_t stream) {
    if (g.M <= 0 || g.Nn <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)num_tiles(g.M, g.Nn, g.tile_mode);
    static const bool dmma = [] {
        const char *e = getenv("CPB200_GEMM");
        return !(e && (e[0] == 'd' || e[0] == 'D') && (e[1] == 'f' || e[1] == 'F'));
    }();
    if (dmma) gemm_small_kernel<B_NC, true><<<grid, NT, 0, stream>>>(g);
    else gemm_small_kernel<B_NC, false><<<grid, NT, 0, stream>>>(g);
    return cudaGetLastError();
}

}  // namespace cpsmall
