// Least-squares reconstruction of the surviving weights.
//
//   cp_ls_solve       <- fc_kernel / LinearRegression(fit_intercept=True).fit
//                        (reference lib/decompose.py:622-623, 636-669): centred normal
//                        equations on the principal sub-block of the Gram matrix.
//   cp_ls_solve_dual  <- the same call when N-1 < K' (gelsd's minimum-norm answer),
//                        through the dual (row) normal equations.
//
// Both reduce to one routine: blocked right-looking Cholesky (panel 64) of an SPD matrix
// with the right-hand sides appended as extra ROWS of the same array, so that the forward
// substitution happens for free inside the panel TRSM / trailing update; the diagonal
// blocks are inverted once (64x64, one CTA) and every other operation is an fp64 tile
// GEMM (cpgemm::gemm_kernel).  The backward substitution reuses the inverted blocks.
// Bound: FP64 pipe for the trailing updates (K'^3/3 flop), latency for the 64x64 panels.
#include "common.cuh"
#include "gemm_f64.cuh"

namespace {

constexpr int NB = 64;

// ---------------------------------------------------------------- assemble
// M rows 0..Ks-1    : G[sel_i, sel_j] - sx_i sx_j / N           (Ks x Ks)
// M rows Ks..Ks+n-1 : Bxy[sel_j, t]   - sx_j sy_t / N           (n  x Ks)   (right-hand sides, transposed)
__global__ void __launch_bounds__(256)
ls_assemble(const double *__restrict__ G, const double *__restrict__ Bxy, const double *__restrict__ sx,
            const double *__restrict__ sy, double invN, int K, int n, const int32_t *__restrict__ sel, int Ks,
            double *__restrict__ M, int64_t ld, double *__restrict__ diag0) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= Ks) return;
    const int sj = sel[j];
    if (i < Ks) {
        const int si = sel[i];
        const double v = G[(int64_t)si * K + sj] - sx[si] * sx[sj] * invN;
        M[(int64_t)i * ld + j] = v;
        if (i == j) diag0[i] = v;
    } else {
        const int t = i - Ks;
        M[(int64_t)i * ld + j] = Bxy[(int64_t)sj * n + t] - sx[sj] * sy[t] * invN;
    }
}

// ---------------------------------------------------------------- 64x64 diagonal block: L and L^-1
// One CTA of 1024 threads; thread (j = tid % 64, ig = tid / 64) keeps rows i = ig + 16 m (m = 0..3) of
// column j in registers.  Right-looking factorisation: at step k the 16 threads of column k publish the
// raw column (pivot included) to shared memory, ONE barrier, then every thread applies the
// rank-1 update to its registers.  The inverse is a forward substitution with the same ownership (one
// barrier per row).  1/sqrt(pivot) comes from the fp32 MUFU seed + two fp64 Newton steps.
constexpr int PT = 1024;
constexpr int PR = 4;   // rows per thread
constexpr int PG = 64 / PR;  // row groups
constexpr size_t POTRF_SMEM = (size_t)(NB * (NB + 1) + 5 * NB) * sizeof(double);

__device__ __forceinline__ double rsqrt_newton(double d) {
    double y = (double)rsqrtf((float)d);
    y = y * (1.5 - 0.5 * d * y * y);
    y = y * (1.5 - 0.5 * d * y * y);
    return y;
}

__global__ void __launch_bounds__(PT, 1)
potrf_diag(double *__restrict__ A, int64_t ld, int nb, double *__restrict__ Linv, int32_t *__restrict__ info,
           int j0, const double *__restrict__ diag0) {
    extern __shared__ __align__(16) double psm[];
    double *Ls = psm;                      // [64][65] finished factor (row-major)
    double *col = psm + NB * (NB + 1);     // [2][64] raw column k (double buffered)
    double *xrow = col + 2 * NB;           // [2][64] finished row k of the inverse
    double *rsd = xrow + 2 * NB;           // [64] 1 / L[k][k]
    const int tid = threadIdx.x, j = tid & 63, ig = tid >> 6;
    double a[PR], x[PR];
#pragma unroll
    for (int m = 0; m < PR; ++m) {
        const int i = ig + PG * m;
        double v = 0.0;
        if (i < nb && j < nb && j <= i) v = A[(int64_t)i * ld + j];
        if (i >= nb && i == j) v = 1.0;  // identity padding keeps the arithmetic finite
        a[m] = v;
        x[m] = (i == j) ? 1.0 : 0.0;
    }
    for (int k = 0; k < NB; ++k) {
        double *ck = col + (k & 1) * NB;
        if (j == k) {
#pragma unroll
            for (int m = 0; m < PR; ++m) ck[ig + PG * m] = a[m];
        }
        __syncthreads();
        double d = ck[k];
        // pivot must stay above 1e-12 of the original diagonal entry: the squared form of the
        // sigma < 1e-6 sigma_max cut-off LinearRegression applies (sklearn _base.py:752-753)
        if (!(d > (k < nb ? 1e-12 * diag0[j0 + k] : 0.0))) {
            if (tid == 0 && k < nb) atomicCAS(info, 0, j0 + k + 1);
            d = 1.0;
        }
        const double rs = rsqrt_newton(d);
        if (j == k) {
#pragma unroll
            for (int m = 0; m < PR; ++m) {
                const int i = ig + PG * m;
                Ls[i * (NB + 1) + k] = (i > k) ? ck[i] * rs : (i == k ? d * rs : 0.0);
            }
            if (ig == 0) rsd[k] = rs;
        }
        if (j > k) {
            const double ljk = -ck[j] * rs * rs;  // -L[j][k] / sqrt(d)
#pragma unroll
            for (int m = 0; m < PR; ++m) {
                const int i = ig + PG * m;
                if (i >= j) a[m] = fma(ck[i], ljk, a[m]);
            }
        }
    }
    __syncthreads();
    // X = L^-1: row k of X is final once rows < k have been eliminated
    for (int k = 0; k < NB; ++k) {
        double *xr = xrow + (k & 1) * NB;
        if (ig == (k & (PG - 1))) {
            const int mk = k / PG;
            double xv = 0.0;
#pragma unroll
            for (int m = 0; m < PR; ++m) xv = (m == mk) ? x[m] : xv;
            xv *= rsd[k];
            xr[j] = xv;
            Linv[k * NB + j] = (j <= k) ? xv : 0.0;
        }
        __syncthreads();
        if (j <= k) {  // columns right of the diagonal stay zero
            const double xk = xr[j];
#pragma unroll
            for (int m = 0; m < PR; ++m) {
                const int i = ig + PG * m;
                if (i > k) x[m] = fma(-Ls[i * (NB + 1) + k], xk, x[m]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < PR; ++m) {
        const int i = ig + PG * m;
        if (i < nb && j < nb && j <= i) A[(int64_t)i * ld + j] = Ls[i * (NB + 1) + j];
    }
}

__global__ void __launch_bounds__(256)
ls_output(const double *__restrict__ Wt, int64_t ld, const double *__restrict__ sx, const double *__restrict__ sy,
          const int32_t *__restrict__ sel, int Ks, double invN, double *__restrict__ W_out,
          double *__restrict__ b_out) {
    __shared__ double red[256];
    const int t = blockIdx.x;
    const double *src = Wt + (int64_t)t * ld;
    double s = 0.0;
    for (int i = threadIdx.x; i < Ks; i += 256) {
        const double w = src[i];
        W_out[(int64_t)t * Ks + i] = w;
        s = fma(sx[sel[i]], w, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) b_out[t] = (sy[t] - red[0]) * invN;
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

template <bool B_NC>
int dgemm(const double *A, int64_t lda, const double *B, int64_t ldb, double *C, int64_t ldc, int M, int Nn,
          int64_t R, double alpha, double beta, int tile_mode, cudaStream_t stream) {
    using namespace cpgemm;
    Args g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.M = M; g.Nn = Nn; g.R = R;
    g.nsplit = 1; g.r_per_split = R;
    g.alpha = alpha; g.beta = beta; g.tile_mode = tile_mode;
    g.a_vec = al16(A) && (lda % 2 == 0);
    g.b_vec = al16(B) && (ldb % 2 == 0);
    if (M <= 0 || Nn <= 0) return CP_OK;
    CP_GEMM_LAUNCH((launch<double, double, false, B_NC>(g, stream)));
    return CP_OK;
}

}  // namespace

// In-place: M is (Kd + n) x Kd (leading dimension ld), rows 0..Kd-1 an SPD matrix (lower part
// used), rows Kd.. the transposed right-hand sides.  On return rows Kd.. hold the transposed
// solution  (SPD^-1 Rhs)'.  Linv: scratch of ceil(Kd/64) * 64*64 doubles.
static int chol_solve_inplace(cp_handle_t h, double *M, int64_t ld, int Kd, int n, double *Linv, const double *diag0,
                              int32_t *info, cudaStream_t stream) {
    using namespace cpgemm;
    const int Ktot = Kd + n;
    const int npanel = (Kd + NB - 1) / NB;
    static bool configured = false;
    if (!configured) {
        CP_CUDA(cudaFuncSetAttribute(potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
        configured = true;
    }
    // Two-level blocking: 256-wide outer panels are factored with 64-wide inner blocks whose updates
    // stay inside the panel; the trailing matrix is then updated ONCE per outer panel with inner
    // dimension 256 (4x fewer, 4x deeper tile GEMMs than a plain 64-wide right-looking sweep).
    constexpr int NBO = 4 * NB;
    (void)npanel;
    // Look-ahead: the trailing update of an outer panel is split into the next panel's columns (needed at once, stays
    // on the caller's stream) and the rest, which runs on a low-priority side stream concurrently with the next
    // panel's (latency-bound, one-CTA-at-a-time) factorisation.
    if (!h->side) {
        int lo = 0, hi = 0;
        CP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CP_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, lo));
        CP_CUDA(cudaEventCreateWithFlags(&h->ev_panel, cudaEventDisableTiming));
        CP_CUDA(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
    }
    bool side_pending = false;
    for (int J0 = 0; J0 < Kd; J0 += NBO) {
        const int w = Kd - J0 < NBO ? Kd - J0 : NBO;
        const int J1 = J0 + w;
        for (int jb = J0; jb < J1; jb += NB) {
            const int nb = J1 - jb < NB ? J1 - jb : NB;
            const int j1 = jb + nb;
            double *Lp = Linv + (size_t)(jb / NB) * NB * NB;
            potrf_diag<<<1, PT, POTRF_SMEM, stream>>>(M + (int64_t)jb * ld + jb, ld, nb, Lp, info, jb, diag0);
            CP_CHECK_LAUNCH();
            const int below = Ktot - j1;
            if (below > 0) {
                // block column <- block column * L_d^-T (all rows below, right-hand-side rows included), in place
                double *Pn = M + (int64_t)j1 * ld + jb;
                int rc = dgemm<false>(Pn, ld, Lp, NB, Pn, ld, below, nb, nb, 1.0, 0.0, TILES_ALL, stream);
                if (rc) return rc;
                const int nin = J1 - j1;  // columns of this outer panel still to be factored
                if (nin > 0) {
                    rc = dgemm<false>(Pn, ld, Pn, ld, M + (int64_t)j1 * ld + j1, ld, below, nin, nb, -1.0, 1.0, TILES_ALL,
                                      stream);
                    if (rc) return rc;
                }
            }
        }
        const int ncols = Kd - J1;
        if (ncols > 0) {  // trailing (lower) -= panel * panel', inner dimension w
            const int w2 = ncols < NBO ? ncols : NBO;  // columns of the next outer panel
            const bool fork = ncols > w2;
            if (fork) CP_CUDA(cudaEventRecord(h->ev_panel, stream));  // panel J0..J1 is final
            if (side_pending) {  // the previous panel's far update also touched the columns updated next
                CP_CUDA(cudaStreamWaitEvent(stream, h->ev_side, 0));
                side_pending = false;
            }
            double *Pn = M + (int64_t)J1 * ld + J0;
            int rc = dgemm<false>(Pn, ld, Pn, ld, M + (int64_t)J1 * ld + J1, ld, Ktot - J1, w2, w, -1.0, 1.0, TILES_LOWER, stream);
            if (rc) return rc;
            if (fork) {
                const int J2 = J1 + w2;
                double *Pf = M + (int64_t)J2 * ld + J0;
                CP_CUDA(cudaStreamWaitEvent(h->side, h->ev_panel, 0));
                rc = dgemm<false>(Pf, ld, Pf, ld, M + (int64_t)J2 * ld + J2, ld, Ktot - J2, ncols - w2, w, -1.0, 1.0, TILES_LOWER,
                                  h->side);
                if (rc) return rc;
                CP_CUDA(cudaEventRecord(h->ev_side, h->side));
                side_pending = true;
            }
        }
    }
    if (side_pending) CP_CUDA(cudaStreamWaitEvent(stream, h->ev_side, 0));
    // backward: Wt * L = Zt, outer panels last to first, inner blocks last to first
    double *Zt = M + (int64_t)Kd * ld;
    const int nouter = (Kd + NBO - 1) / NBO;
    for (int P = nouter - 1; P >= 0; --P) {
        const int J0 = P * NBO;
        const int w = Kd - J0 < NBO ? Kd - J0 : NBO;
        const int nin = (w + NB - 1) / NB;
        for (int b = nin - 1; b >= 0; --b) {
            const int jb = J0 + b * NB;
            const int nb = J0 + w - jb < NB ? J0 + w - jb : NB;
            double *Lp = Linv + (size_t)(jb / NB) * NB * NB;
            // Wt_b = Zt_b * Linv_b   (C[t, i] = sum_r Zt[t, jb + r] * Linv[r, i]), in place
            int rc = dgemm<true>(Zt + jb, ld, Lp, NB, Zt + jb, ld, n, nb, nb, 1.0, 0.0, TILES_ALL, stream);
            if (rc) return rc;
            if (jb > J0) {  // remaining columns of this outer panel
                rc = dgemm<true>(Zt + jb, ld, M + (int64_t)jb * ld + J0, ld, Zt + J0, ld, n, jb - J0, nb, -1.0, 1.0, TILES_ALL,
                                 stream);
                if (rc) return rc;
            }
        }
        if (J0 > 0) {  // Zt[:, 0:J0] -= Wt_P * L[J0:J0+w, 0:J0], inner dimension w
            int rc = dgemm<true>(Zt + J0, ld, M + (int64_t)J0 * ld, ld, Zt, ld, n, J0, w, -1.0, 1.0, TILES_ALL, stream);
            if (rc) return rc;
        }
    }
    return CP_OK;
}

extern "C" int cp_ls_solve(cp_handle_t h, const double *G, const double *Bxy, const double *sx, const double *sy,
                           int64_t N, int K, int n, const int32_t *sel_cols, int Ksel, double *W_out, double *b_out,
                           int32_t *info_out, cp_stream_t stream_) {
    CP_REQUIRE(h && G && Bxy && sx && sy && sel_cols && W_out && b_out && info_out, "cp_ls_solve: NULL argument");
    CP_REQUIRE(K > 0 && n > 0 && Ksel > 0 && Ksel <= K && N > 0, "cp_ls_solve: bad shape");
    CP_REQUIRE(N - 1 >= Ksel, "cp_ls_solve: N-1=%lld < K'=%d: centred Gram is singular, use cp_ls_solve_dual",
               (long long)(N - 1), Ksel);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t ld = (Ksel + 7) / 8 * 8;
    const int npanel = (Ksel + NB - 1) / NB;
    const size_t need = cp_carver::need((size_t)(Ksel + n) * ld, 8) + cp_carver::need((size_t)npanel * NB * NB, 8) +
                        cp_carver::need(Ksel, 8);
    void *ws = nullptr;
    int rc = cp_ws_reserve(h, need, &ws);
    if (rc) return rc;
    cp_carver cv(ws);
    double *M = cv.take<double>((size_t)(Ksel + n) * ld);
    double *Linv = cv.take<double>((size_t)npanel * NB * NB);
    double *diag0 = cv.take<double>(Ksel);
    CP_CUDA(cudaMemsetAsync(info_out, 0, sizeof(int32_t), stream));
    const double invN = 1.0 / (double)N;
    dim3 grid(cp_cdiv(Ksel, 256), Ksel + n);
    ls_assemble<<<grid, 256, 0, stream>>>(G, Bxy, sx, sy, invN, K, n, sel_cols, Ksel, M, ld, diag0);
    CP_CHECK_LAUNCH();
    rc = chol_solve_inplace(h, M, ld, Ksel, n, Linv, diag0, info_out, stream);
    if (rc) return rc;
    ls_output<<<n, 256, 0, stream>>>(M + (int64_t)Ksel * ld, ld, sx, sy, sel_cols, Ksel, invN, W_out, b_out);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// ---------------------------------------------------------------- dual (minimum-norm) path
namespace {

// column means of the selected columns of X (fp64, fixed order) and of Y - bias
template <typename T>
__global__ void __launch_bounds__(256)
colmean_sel(const T *__restrict__ X, int64_t ld, const int32_t *__restrict__ sel, int ncols, int64_t N,
            const float *__restrict__ bias, double *__restrict__ mean_out) {
    __shared__ double s1[8][33];
    const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + cx;
    double a = 0.0;
    if (j < ncols) {
        const int col = sel ? sel[j] : j;
        const double b = bias ? (double)bias[col] : 0.0;
        for (int64_t r = rg; r < N; r += 8) a += (double)__ldg(X + r * ld + col) - b;
    }
    s1[rg][cx] = a;
    __syncthreads();
    if (rg == 0 && j < ncols) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s1[k][cx];
        mean_out[j] = t / (double)N;
    }
}

// Xc[r, j] = X[r, sel_j] - mean_j    (N x Ks fp64, ld)
__global__ void __launch_bounds__(256)
center_sel(const float *__restrict__ X, int64_t ldx, const int32_t *__restrict__ sel, int Ks,
           const double *__restrict__ mean, double *__restrict__ Xc, int64_t ld) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (j < Ks) Xc[r * ld + j] = (double)__ldg(X + r * ldx + sel[j]) - mean[j];
}

// rows N..N+n-1 of the augmented matrix: Yc' (n x N):  M[N + t, r] = Y[r, t] - bias_t - ymean_t
template <typename T>
__global__ void __launch_bounds__(256)
dual_rhs(const T *__restrict__ Y, int64_t ldy, const float *__restrict__ bias, const double *__restrict__ ymean,
         int64_t N, int n, double *__restrict__ M, int64_t ld) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int t = blockIdx.y;
    if (r < N) M[(N + t) * ld + r] = (double)__ldg(Y + r * ldy + t) - (bias ? (double)bias[t] : 0.0) - ymean[t];
}

__global__ void add_const_lower(double *__restrict__ M, int64_t ld, int N, double v, double *__restrict__ diag0) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j < N && j <= i) {
        const double x = M[(int64_t)i * ld + j] + v;
        M[(int64_t)i * ld + j] = x;
        if (i == j) diag0[i] = x;
    }
}

__global__ void __launch_bounds__(256)
dual_output(const double *__restrict__ Wt, int64_t ld, const double *__restrict__ xmean,
            const double *__restrict__ ymean, int Ks, double *__restrict__ W_out, double *__restrict__ b_out) {
    __shared__ double red[256];
    const int t = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < Ks; i += 256) {
        const double w = Wt[(int64_t)t * ld + i];
        W_out[(int64_t)t * Ks + i] = w;
        s = fma(xmean[i], w, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) b_out[t] = ymean[t] - red[0];
}

}  // namespace

extern "C" int cp_ls_solve_dual(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw,
                                int y_dtype, int n, int64_t ldy, const float *y_bias, const int32_t *sel_cols, int Ksel,
                                double *W_out, double *b_out, int32_t *info_out, cp_stream_t stream_) {
    using namespace cpgemm;
    CP_REQUIRE(h && X && Yraw && sel_cols && W_out && b_out && info_out, "cp_ls_solve_dual: NULL argument");
    CP_REQUIRE(N > 1 && N < (1 << 15) && K > 0 && n > 0 && Ksel > 0 && Ksel <= K && ldx >= K && ldy >= n,
               "cp_ls_solve_dual: bad shape (N must be < 32768)");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Ni = (int)N;
    const int64_t ldc = (Ksel + 7) / 8 * 8;  // Xc
    const int64_t ldm = (Ni + 7) / 8 * 8;    // dual system
    const int npanel = (Ni + NB - 1) / NB;
    const size_t need = cp_carver::need((size_t)Ni * ldc, 8) + cp_carver::need((size_t)(Ni + n) * ldm, 8) +
                        cp_carver::need((size_t)npanel * NB * NB, 8) + cp_carver::need((size_t)n * ldc, 8) +
                        cp_carver::need(Ksel, 8) + cp_carver::need(n, 8) + cp_carver::need(Ni, 8);
    void *ws = nullptr;
    int rc = cp_ws_reserve(h, need, &ws);
    if (rc) return rc;
    cp_carver cv(ws);
    double *Xc = cv.take<double>((size_t)Ni * ldc);
    double *M = cv.take<double>((size_t)(Ni + n) * ldm);
    double *Linv = cv.take<double>((size_t)npanel * NB * NB);
    double *Wt = cv.take<double>((size_t)n * ldc);
    double *xmean = cv.take<double>(Ksel);
    double *ymean = cv.take<double>(n);
    double *diag0 = cv.take<double>(Ni);
    CP_CUDA(cudaMemsetAsync(info_out, 0, sizeof(int32_t), stream));
    CP_REQUIRE(y_dtype == CP_F32 || y_dtype == CP_F64, "cp_ls_solve_dual: unknown y_dtype %d", y_dtype);
    colmean_sel<float><<<cp_cdiv(Ksel, 32), 256, 0, stream>>>(X, ldx, sel_cols, Ksel, N, nullptr, xmean);
    CP_CHECK_LAUNCH();
    if (y_dtype == CP_F32)
        colmean_sel<float><<<cp_cdiv(n, 32), 256, 0, stream>>>((const float *)Yraw, ldy, nullptr, n, N, y_bias, ymean);
    else
        colmean_sel<double><<<cp_cdiv(n, 32), 256, 0, stream>>>((const double *)Yraw, ldy, nullptr, n, N, y_bias, ymean);
    CP_CHECK_LAUNCH();
    center_sel<<<dim3(cp_cdiv(Ksel, 256), Ni), 256, 0, stream>>>(X, ldx, sel_cols, Ksel, xmean, Xc, ldc);
    CP_CHECK_LAUNCH();
    // H = Xc Xc' (lower tiles) + 1/N
    rc = dgemm<false>(Xc, ldc, Xc, ldc, M, ldm, Ni, Ni, Ksel, 1.0, 0.0, TILES_LOWER, stream);
    if (rc) return rc;
    add_const_lower<<<dim3(cp_cdiv(Ni, 256), Ni), 256, 0, stream>>>(M, ldm, Ni, 1.0 / (double)N, diag0);
    CP_CHECK_LAUNCH();
    if (y_dtype == CP_F32)
        dual_rhs<float><<<dim3(cp_cdiv(Ni, 256), n), 256, 0, stream>>>((const float *)Yraw, ldy, y_bias, ymean, N, n, M, ldm);
    else
        dual_rhs<double><<<dim3(cp_cdiv(Ni, 256), n), 256, 0, stream>>>((const double *)Yraw, ldy, y_bias, ymean, N, n, M, ldm);
    CP_CHECK_LAUNCH();
    rc = chol_solve_inplace(h, M, ldm, Ni, n, Linv, diag0, info_out, stream);
    if (rc) return rc;
    // Wt = At * Xc   (C[t, i] = sum_r At[t, r] * Xc[r, i])
    rc = dgemm<true>(M + (int64_t)Ni * ldm, ldm, Xc, ldc, Wt, ldc, n, Ksel, Ni, 1.0, 0.0, TILES_ALL, stream);
    if (rc) return rc;
    dual_output<<<n, 256, 0, stream>>>(Wt, ldc, xmean, ymean, Ksel, W_out, b_out);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
