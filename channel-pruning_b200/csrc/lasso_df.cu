// Data-form coordinate descent -- the benchmark kernel of SURVEY.md 8(b)/(d) ("LASSO data-form CD: HBM; 4 m c bytes
// per sweep"): the algorithm scikit-learn actually runs for the reference's Lasso.fit(Z, reY)
// (lib/decompose.py:428-457; sklearn 1.9.0 _cd_fast.pyx enet_coordinate_descent, selection='random', no screening),
// on the MATERIALISED design matrix Z (m = S n rows, c columns, fp32, column major) instead of its Gram matrix.
// The product path does not use it (cp_lasso_build + cp_lasso_select work in channel space and never form Z);
// it exists so that the data-form algorithm has a measured bandwidth figure next to the Gram form's latency figure.
//
// Coordinate descent is sequential: coordinate j needs the residual after coordinate j-1.  A grid cannot afford one
// grid-wide reduction per coordinate, so the sweep is processed in blocks of DF_B = 8 coordinates with EXACT
// sequential semantics: one pass over the rows produces, per CTA, the 8 dot products Z_k.R against the residual
// before the block and the 36 inner products Z_k.Z_l of the block's columns; after a grid-wide reduction every CTA
// resolves the 8 soft-threshold updates in order (dot_k - sum_{i<k} delta_i Z_k.Z_i) and then applies them to its
// rows of R while it already accumulates the next block's sums.  Z is streamed from HBM once per sweep (the second
// touch of a block's columns hits L2).  Centring (sklearn's _pre_fit) is implicit: column means are subtracted on
// the fly, Z itself stays raw fp32.  All reductions run in a fixed order: results are reproducible.
// Bound: HBM stream of Z (4 m c bytes per sweep) + two grid barriers per 8 coordinates.
#include "common.cuh"

namespace {

constexpr int DF_T = 256;
constexpr int DF_B = 8;
constexpr int DF_NV = DF_B + DF_B * (DF_B + 1) / 2;  // 44
constexpr unsigned DF_RAND_R_MAX = 2147483647u;

struct DfParams {
    const float *Z;
    int64_t ldz;
    int m, c, G, rows_per_cta;
    const double *zmean, *norm2, *yc;
    double *R, *w;
    double l1, tol_scaled, d_w_tol;
    int max_iter;
    uint32_t seed;
    double *part;              // [2][DF_NV][G]
    double *total;             // [2][DF_NV]
    double *xta_part;          // [G][c + 2]
    double *xta;               // [c + 2]   (Z_j.R for all j, then R.R, R.y)
    unsigned long long *bar;
    double *out;               // n_iter, gap, tol_scaled, sweeps, gap checks
};

__device__ __forceinline__ uint32_t df_rand(uint32_t &s) {  // sklearn/utils/_random.pxd:20-34
    if (s == 0) s = 1;
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s % (DF_RAND_R_MAX + 1u);
}

__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &epoch, int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1ull);
        const unsigned long long target = (++epoch) * (unsigned long long)G;
        unsigned long long v;
        unsigned spin = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
            if (++spin > (1u << 28)) __trap();  // a protocol bug (or CTAs that are not co-resident) traps instead of hanging
        } while (v < target);
    } else {
        ++epoch;
    }
    __syncthreads();
}

// sum over the block of NV per-thread values, result in out[0..NV) (shared); fixed order
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *scratch /* [8][NV] */, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if (lane == 0) scratch[warp * NV + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < DF_T / 32; ++w) t += scratch[w * NV + threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DF_T, 1) df_kernel(const DfParams P) {
    extern __shared__ __align__(16) double dsm[];
    double *Rs = dsm;                          // [rows_per_cta]
    double *ws = Rs + P.rows_per_cta;          // [c] this CTA's copy of the coefficients
    double *scratch = ws + P.c;                // [8][DF_NV]
    double *red = scratch + 8 * DF_NV;         // [DF_NV]
    double *tot = red + DF_NV;                 // [DF_NV]
    double *dl = tot + DF_NV;                  // [DF_B] deltas of the block
    double *zm_s = dl + DF_B;                  // [c] column means   (shared copies: the per-block scalar work of thread 0
    double *n2_s = zm_s + P.c;                 // [c] centred norms   must not wait on global loads)
    __shared__ int js[2][DF_B];
    __shared__ double ctl_gap, ctl_wmax, ctl_dwmax;
    const int tid = threadIdx.x, cta = blockIdx.x, G = P.G;
    const int r0 = cta * P.rows_per_cta;
    const int nrows = max(0, min(P.rows_per_cta, P.m - r0));
    unsigned long long epoch = 0;
    for (int i = tid; i < nrows; i += DF_T) Rs[i] = P.R[r0 + i];
    for (int j = tid; j < P.c; j += DF_T) {
        ws[j] = P.w[j];
        zm_s[j] = P.zmean[j];
        n2_s[j] = P.norm2[j];
    }
    __syncthreads();

    // ---- duality gap (gap_enet + dual_gap_formulation_A, beta = 0): one pass over all columns
    auto gap_check = [&]() {
        double *mine = P.xta_part + (size_t)cta * (P.c + 2);
        for (int j0 = 0; j0 < P.c; j0 += DF_B) {
            double acc[DF_B];
#pragma unroll
            for (int k = 0; k < DF_B; ++k) acc[k] = 0.0;
            for (int i = tid; i < nrows; i += DF_T) {
                const double r = Rs[i];
#pragma unroll
                for (int k = 0; k < DF_B; ++k)
                    if (j0 + k < P.c) acc[k] = fma((double)P.Z[(int64_t)(j0 + k) * P.ldz + r0 + i] - zm_s[j0 + k], r, acc[k]);
            }
            block_reduce<DF_B>(acc, scratch, red);
            if (tid < DF_B && j0 + tid < P.c) mine[j0 + tid] = red[tid];
            __syncthreads();
        }
        double a2[2] = {0.0, 0.0};
        for (int i = tid; i < nrows; i += DF_T) {
            const double r = Rs[i];
            a2[0] = fma(r, r, a2[0]);
            a2[1] = fma(r, P.yc[r0 + i], a2[1]);
        }
        block_reduce<2>(a2, scratch, red);
        if (tid < 2) mine[P.c + tid] = red[tid];
        grid_barrier(P.bar, epoch, G);
        for (int j = cta; j < P.c + 2; j += G) {  // fixed-order sum over CTAs, one warp per value
            if (tid < 32) {
                double t = 0.0;
                for (int q = tid; q < G; q += 32) t += P.xta_part[(size_t)q * (P.c + 2) + j];
#pragma unroll
                for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (tid == 0) P.xta[j] = t;
            }
        }
        grid_barrier(P.bar, epoch, G);
        double dn = 0.0, l1n = 0.0;
        for (int j = tid; j < P.c; j += DF_T) {
            dn = fmax(dn, fabs(P.xta[j]));
            l1n += fabs(ws[j]);
        }
        double v2[1] = {l1n};
        block_reduce<1>(v2, scratch, red);
        const double l1norm = red[0];
        __syncthreads();
        // max over the block
#pragma unroll
        for (int off = 16; off; off >>= 1) dn = fmax(dn, __shfl_xor_sync(0xffffffffu, dn, off));
        if ((tid & 31) == 0) scratch[tid >> 5] = dn;
        __syncthreads();
        if (tid == 0) {
            double d = 0.0;
            for (int w8 = 0; w8 < DF_T / 32; ++w8) d = fmax(d, scratch[w8]);
            const double R_norm2 = P.xta[P.c], Ry = P.xta[P.c + 1];
            const double primal = 0.5 * R_norm2 + P.l1 * l1norm;
            const double scale = d > P.l1 ? P.l1 / d : 1.0;
            const double dual = -0.5 * (scale * scale) * R_norm2 + scale * Ry;
            ctl_gap = primal - dual;
        }
        __syncthreads();
        return ctl_gap;
    };

    int n_iter_ret = 0, sweeps = 0, checks = 0;
    double gap = gap_check();
    ++checks;
    bool broke = false;
    if (gap > P.tol_scaled) {
        uint32_t state = P.seed;
        for (int n_iter = 0; n_iter < P.max_iter; ++n_iter) {
            double w_max = 0.0, d_w_max = 0.0;
            const int nblk = (P.c + DF_B - 1) / DF_B;
            // coordinates of block 0
            if (tid == 0) {
                for (int k = 0; k < DF_B; ++k) js[0][k] = (k < P.c) ? (int)(df_rand(state) % (uint32_t)P.c) : -1;
            }
            __syncthreads();
            // no pending deltas before the first block
            if (tid < DF_B) dl[tid] = 0.0;
            __syncthreads();
            for (int b = 0; b <= nblk; ++b) {
                const int par = b & 1;
                // one pass over the rows: apply the previous block's updates, accumulate this block's sums
                const bool have_prev = b > 0, have_cur = b < nblk;
                int jc[DF_B], jp[DF_B];
                double mc[DF_B], mp[DF_B], dp[DF_B];
#pragma unroll
                for (int k = 0; k < DF_B; ++k) {
                    jc[k] = have_cur ? js[par][k] : -1;
                    jp[k] = have_prev ? js[par ^ 1][k] : -1;
                    mc[k] = jc[k] >= 0 ? zm_s[jc[k]] : 0.0;
                    mp[k] = jp[k] >= 0 ? zm_s[jp[k]] : 0.0;
                    dp[k] = have_prev ? dl[k] : 0.0;
                }
                double acc[DF_NV];
#pragma unroll
                for (int v = 0; v < DF_NV; ++v) acc[v] = 0.0;
                for (int i = tid; i < nrows; i += DF_T) {
                    double r = Rs[i];
                    if (have_prev) {
#pragma unroll
                        for (int k = 0; k < DF_B; ++k)
                            if (jp[k] >= 0 && dp[k] != 0.0)
                                r = fma(-dp[k], (double)P.Z[(int64_t)jp[k] * P.ldz + r0 + i] - mp[k], r);
                        Rs[i] = r;
                    }
                    if (have_cur) {
                        double z[DF_B];
#pragma unroll
                        for (int k = 0; k < DF_B; ++k)
                            z[k] = jc[k] >= 0 ? (double)P.Z[(int64_t)jc[k] * P.ldz + r0 + i] - mc[k] : 0.0;
#pragma unroll
                        for (int k = 0; k < DF_B; ++k) acc[k] = fma(z[k], r, acc[k]);
#pragma unroll
                        for (int k = 0; k < DF_B; ++k)
#pragma unroll
                            for (int l = 0; l <= k; ++l) {
                                const int v = DF_B + k * (k + 1) / 2 + l;  // Z_k . Z_l
                                acc[v] = fma(z[k], z[l], acc[v]);
                            }
                    }
                }
                if (!have_cur) break;
                block_reduce<DF_NV>(acc, scratch, red);
                if (tid < DF_NV) P.part[((size_t)par * DF_NV + tid) * G + cta] = red[tid];
                grid_barrier(P.bar, epoch, G);
                for (int v = cta; v < DF_NV; v += G) {
                    if (tid < 32) {
                        double t = 0.0;
                        for (int q = tid; q < G; q += 32) t += P.part[((size_t)par * DF_NV + v) * G + q];
#pragma unroll
                        for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                        if (tid == 0) P.total[par * DF_NV + v] = t;
                    }
                }
                grid_barrier(P.bar, epoch, G);
                if (tid < DF_NV) tot[tid] = P.total[par * DF_NV + tid];
                __syncthreads();
                if (tid == 0) {  // the 8 updates of the block, in order (every CTA computes the same numbers)
                    double d[DF_B];
                    for (int k = 0; k < DF_B; ++k) {
                        d[k] = 0.0;
                        const int j = js[par][k];
                        if (j < 0) continue;
                        const double nj = n2_s[j];
                        if (nj == 0.0) continue;
                        double dot = tot[k];
                        for (int i2 = 0; i2 < k; ++i2) dot -= d[i2] * tot[DF_B + k * (k + 1) / 2 + i2];
                        const double w_j = ws[j];
                        const double tmp = dot + w_j * nj;
                        const double mag = fabs(tmp) - P.l1;
                        const double w_new = mag > 0.0 ? copysign(mag, tmp) / nj : 0.0;
                        ws[j] = w_new;
                        d[k] = w_new - w_j;
                        w_max = fmax(w_max, fabs(w_new));
                        d_w_max = fmax(d_w_max, fabs(d[k]));
                    }
                    for (int k = 0; k < DF_B; ++k) dl[k] = d[k];
                    // coordinates of the next block
                    for (int k = 0; k < DF_B; ++k) {
                        const int f = (b + 1) * DF_B + k;
                        js[par ^ 1][k] = (b + 1 < nblk && f < P.c) ? (int)(df_rand(state) % (uint32_t)P.c) : -1;
                    }
                    ctl_wmax = w_max;
                    ctl_dwmax = d_w_max;
                }
                __syncthreads();
            }
            __syncthreads();
            w_max = ctl_wmax;
            d_w_max = ctl_dwmax;
            ++sweeps;
            if (w_max == 0.0 || d_w_max / w_max <= P.d_w_tol || n_iter == P.max_iter - 1) {
                gap = gap_check();
                ++checks;
                if (gap <= P.tol_scaled) {
                    broke = true;
                    n_iter_ret = n_iter + 1;
                    break;
                }
            }
        }
        if (!broke) n_iter_ret = P.max_iter;
    }
    for (int i = tid; i < nrows; i += DF_T) P.R[r0 + i] = Rs[i];
    if (cta == 0) {
        for (int j = tid; j < P.c; j += DF_T) P.w[j] = ws[j];
        if (tid == 0) {
            P.out[0] = (double)n_iter_ret;
            P.out[1] = gap;
            P.out[2] = P.tol_scaled;
            P.out[3] = (double)sweeps;
            P.out[4] = (double)checks;
        }
    }
}

// ---- set-up kernels
// Z[a * ldz + s * n + t] = sum_p X[samples[s], a k2 + p] W2[t, a k2 + p]   (fp64 accumulate, fp32 out);  y[s n + t]
__global__ void __launch_bounds__(256)
df_build(const float *__restrict__ X, int64_t ldx, const float *__restrict__ W2, int K, int n, int k2,
         const int32_t *__restrict__ samples, int S, const void *__restrict__ Yraw, int y_dtype, int64_t ldy,
         const float *__restrict__ y_bias, float *__restrict__ Z, int64_t ldz, double *__restrict__ y) {
    const int a = blockIdx.x, s = blockIdx.y;
    const int64_t row = samples[s];
    __shared__ float xs[32];
    if (threadIdx.x < k2) xs[threadIdx.x] = X[row * ldx + (int64_t)a * k2 + threadIdx.x];
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += 256) {
        double acc = 0.0;
        for (int p = 0; p < k2; ++p) acc = fma((double)xs[p], (double)W2[(int64_t)t * K + a * k2 + p], acc);
        Z[(int64_t)a * ldz + (int64_t)s * n + t] = (float)acc;
        if (a == 0) {
            const double yv = y_dtype == CP_F32 ? (double)((const float *)Yraw)[row * ldy + t] : ((const double *)Yraw)[row * ldy + t];
            y[(int64_t)s * n + t] = yv - (y_bias ? (double)y_bias[t] : 0.0);
        }
    }
}

// column mean and centred squared norm of Z (one CTA per column, fixed order)
__global__ void __launch_bounds__(256)
df_colstats(const float *__restrict__ Z, int64_t ldz, int m, double *__restrict__ zmean, double *__restrict__ norm2) {
    __shared__ double red[8];
    __shared__ double mean_s;
    const float *z = Z + (int64_t)blockIdx.x * ldz;
    double a = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) a += (double)z[i];
#pragma unroll
    for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        mean_s = t / (double)m;
    }
    __syncthreads();
    const double mu = mean_s;
    double q = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        const double d = (double)z[i] - mu;
        q = fma(d, d, q);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        zmean[blockIdx.x] = mu;
        norm2[blockIdx.x] = t;
    }
}

// yc = y - mean(y);  R = yc - sum_j w_j (Z_j - mean_j);  out[0] = |yc|^2     (single CTA passes; m <= ~1e6)
__global__ void __launch_bounds__(256)
df_center_y(const double *__restrict__ y, int m, double *__restrict__ yc, double *__restrict__ yn2) {
    __shared__ double red[8];
    __shared__ double mean_s;
    double a = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) a += y[i];
#pragma unroll
    for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        mean_s = t / (double)m;
    }
    __syncthreads();
    double q = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        const double d = y[i] - mean_s;
        yc[i] = d;
        q = fma(d, d, q);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        *yn2 = t;
    }
}
__global__ void __launch_bounds__(256)
df_residual(const float *__restrict__ Z, int64_t ldz, int m, int c, const double *__restrict__ zmean,
            const double *__restrict__ w, const double *__restrict__ yc, double *__restrict__ R) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    double r = yc[i];
    for (int j = 0; j < c; ++j) {
        const double wj = w[j];
        if (wj != 0.0) r = fma(-wj, (double)Z[(int64_t)j * ldz + i] - zmean[j], r);
    }
    R[i] = r;
}

}  // namespace

extern "C" int cp_lasso_dataform_build(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const float *W2, int n,
                                       int c, int k2, const int32_t *samples, int S, const void *Yraw, int y_dtype,
                                       int64_t ldy, const float *y_bias, float *Z_out, int64_t ldz, double *y_out,
                                       cp_stream_t stream_) {
    CP_REQUIRE(h && X && W2 && samples && Yraw && Z_out && y_out, "cp_lasso_dataform_build: NULL argument");
    CP_REQUIRE(N > 0 && c > 0 && k2 > 0 && k2 <= 32 && K == c * k2 && n > 0 && S > 0 && ldx >= K && ldy >= n &&
                   ldz >= (int64_t)S * n,
               "cp_lasso_dataform_build: bad shape");
    CP_REQUIRE(y_dtype == CP_F32 || y_dtype == CP_F64, "cp_lasso_dataform_build: unknown y_dtype %d", y_dtype);
    CP_DEVICE_GUARD(h);
    df_build<<<dim3(c, S), 256, 0, (cudaStream_t)stream_>>>(X, ldx, W2, K, n, k2, samples, S, Yraw, y_dtype, ldy, y_bias, Z_out,
                                                            ldz, y_out);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_lasso_cd_dataform(cp_handle_t h, const float *Z, int64_t ldz, const double *y, int m, int c, double alpha,
                                    double tol, int max_iter, uint32_t seed, double *w, double *out_scalars,
                                    cp_stream_t stream_) {
    CP_REQUIRE(h && Z && y && w && out_scalars, "cp_lasso_cd_dataform: NULL argument");
    CP_REQUIRE(m > 0 && c > 0 && ldz >= m && alpha > 0 && max_iter > 0, "cp_lasso_cd_dataform: bad shape / parameters");
    CP_DEVICE_GUARD(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    // grid: one CTA per SM, all co-resident (cooperative launch): rows split evenly
    int G = h->num_sms;
    int rows = (m + G - 1) / G;
    rows = (rows + 3) / 4 * 4;
    G = (m + rows - 1) / rows;
    const size_t smem = (size_t)(rows + 3 * c + 8 * DF_NV + 2 * DF_NV + DF_B) * sizeof(double);
    CP_REQUIRE(smem <= 200 * 1024, "cp_lasso_cd_dataform: m = %d rows / c = %d columns do not fit the per-CTA residual slice", m, c);
    static cp_per_device_flag configured;
    if (bool *done = configured.slot(); !*done) {
        CP_CUDA(cudaFuncSetAttribute(df_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        *done = true;
    }
    const size_t need = cp_carver::need(c, 8) * 2 + cp_carver::need(m, 8) * 2 + cp_carver::need(1, 8) +
                        cp_carver::need((size_t)2 * DF_NV * G, 8) + cp_carver::need(2 * DF_NV, 8) +
                        cp_carver::need((size_t)G * (c + 2), 8) + cp_carver::need(c + 2, 8) + cp_carver::need(1, 8);
    void *ws = nullptr;
    int rc = cp_ws_reserve(h, need, &ws);
    if (rc) return rc;
    cp_carver cv(ws);
    double *zmean = cv.take<double>(c), *norm2 = cv.take<double>(c);
    double *yc = cv.take<double>(m), *R = cv.take<double>(m);
    double *yn2 = cv.take<double>(1);
    double *part = cv.take<double>((size_t)2 * DF_NV * G), *total = cv.take<double>(2 * DF_NV);
    double *xta_part = cv.take<double>((size_t)G * (c + 2)), *xta = cv.take<double>(c + 2);
    unsigned long long *bar = (unsigned long long *)cv.take<double>(1);
    df_colstats<<<c, 256, 0, stream>>>(Z, ldz, m, zmean, norm2);
    CP_CHECK_LAUNCH();
    df_center_y<<<1, 256, 0, stream>>>(y, m, yc, yn2);
    CP_CHECK_LAUNCH();
    df_residual<<<cp_cdiv(m, 256), 256, 0, stream>>>(Z, ldz, m, c, zmean, w, yc, R);
    CP_CHECK_LAUNCH();
    CP_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned long long), stream));
    double yn2_h = 0.0;  // tol is scaled by |yc|^2 (sklearn: tol *= dot(y, y)): one small read-back
    CP_CUDA(cudaMemcpyAsync(&yn2_h, yn2, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CP_CUDA(cudaStreamSynchronize(stream));
    DfParams P{};
    P.Z = Z; P.ldz = ldz; P.m = m; P.c = c; P.G = G; P.rows_per_cta = rows;
    P.zmean = zmean; P.norm2 = norm2; P.yc = yc; P.R = R; P.w = w;
    P.l1 = alpha * (double)m;  // l1_reg = alpha * n_samples (_coordinate_descent.py:781)
    P.tol_scaled = tol * yn2_h; P.d_w_tol = tol; P.max_iter = max_iter; P.seed = seed;
    P.part = part; P.total = total; P.xta_part = xta_part; P.xta = xta; P.bar = bar; P.out = out_scalars;
    void *args[] = {(void *)&P};
    cp_launch_counter.fetch_add(1, std::memory_order_relaxed);
    CP_CUDA(cudaLaunchCooperativeKernel((const void *)df_kernel, dim3(G), dim3(DF_T), args, smem, stream));
    return CP_OK;
}
