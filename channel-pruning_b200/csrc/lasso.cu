// LASSO channel selection on the device.
//
//   cp_lasso_build   <- construction of Z and sklearn's centring (reference
//                       lib/decompose.py:428-437, sklearn _pre_fit), in channel space
//   cp_lasso_select  <- the alpha search of decompose.dictionary (lib/decompose.py:489-525)
//                       with every Lasso.fit inside it (sklearn 1.9.0
//                       _cd_fast.pyx enet_coordinate_descent, selection='random',
//                       warm_start=True, tol=1e-4, gap-safe screening)
//
// The executable specification of cp_lasso_select is oracle/cd_oracle.c:cp_enet_cd_gram
// (same control flow as sklearn's data-form solver, evaluated on Q = Zc'Zc, q = Zc'yc,
// |yc|^2).  The kernel reproduces that model BIT FOR BIT: every floating-point
// operation that the model rounds separately is issued with an explicit *_rn intrinsic
// (no FMA contraction), reductions run in the model's serial order, and the random
// coordinate order comes from the same 32-bit xorshift.
//
// Coordinate descent is a serial chain (one coordinate update needs the previous one),
// so the search runs in ONE persistent CTA: it is latency bound, not HBM/tensor bound.
// Per coordinate step: the owner thread of coordinate j forms the soft-threshold update
// and publishes delta; after one CTA barrier all threads apply  Qw += delta * Q[j,:]  to the
// elements they own.  Rows of Q are streamed from L2 through an 8-deep cp.async ring that
// runs ahead along the (data-independent) random coordinate sequence.
#include "common.cuh"

namespace {

constexpr int LT = 256;     // threads of the persistent CTA
constexpr int RING = 8;     // prefetch depth (rows of Q in flight)
constexpr int MAXC = 2048;  // largest channel count (shared-memory bound)

// ------------------------------------------------------------------ build
__global__ void __launch_bounds__(256)
lasso_build_Q(const double *__restrict__ Gs, const double *__restrict__ WW, const double *__restrict__ sxs,
              const double *__restrict__ sw, int c, int k2, double m, double *__restrict__ Q) {
    const int b = blockIdx.x * 16 + (threadIdx.x & 15);
    const int a = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (a >= c || b >= c) return;
    const int64_t K = (int64_t)c * k2;
    double s = 0.0, za = 0.0, zb = 0.0;
    for (int p = 0; p < k2; ++p) {
        const double *g = Gs + ((int64_t)a * k2 + p) * K + (int64_t)b * k2;
        const double *w = WW + ((int64_t)a * k2 + p) * K + (int64_t)b * k2;
        for (int q = 0; q < k2; ++q) s = fma(g[q], w[q], s);
        za = fma(sxs[a * k2 + p], sw[a * k2 + p], za);
        zb = fma(sxs[b * k2 + p], sw[b * k2 + p], zb);
    }
    Q[(int64_t)a * c + b] = s - za * zb / m;  // - m * zbar_a * zbar_b
}

// qv[a] = sum_{p,j} W2[j,(a,p)] * Bs[(a,p), j] - m zbar_a ybar ;  block per channel, fixed-shape tree.
__global__ void __launch_bounds__(128)
lasso_build_q(const float *__restrict__ W2, const double *__restrict__ Bs, const double *__restrict__ sxs,
              const double *__restrict__ sw, const double *__restrict__ sys, const double *__restrict__ yys, int c,
              int k2, int n, double m, double *__restrict__ qv, double *__restrict__ yn2) {
    __shared__ double red[128];
    const int a = blockIdx.x;
    const int64_t K = (int64_t)c * k2;
    double s = 0.0;
    for (int e = threadIdx.x; e < k2 * n; e += 128) {
        const int p = e / n, j = e - p * n;
        s = fma((double)W2[(int64_t)j * K + a * k2 + p], Bs[((int64_t)a * k2 + p) * n + j], s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double za = 0.0;
        for (int p = 0; p < k2; ++p) za = fma(sxs[a * k2 + p], sw[a * k2 + p], za);
        double ysum = 0.0;
        for (int j = 0; j < n; ++j) ysum += sys[j];
        qv[a] = red[0] - za * ysum / m;  // - m * (za/m) * (ysum/m)
        if (a == 0) *yn2 = *yys - ysum * ysum / m;
    }
}

// ------------------------------------------------------------------ select
struct SelectParams {
    const double *Q, *qv, *yn2;
    int c;
    double m;
    int rank;
    double lbound, rbound, right0, tol;
    int max_iter;
    const uint32_t *seeds;
    int max_probes;
    uint8_t *out_idxs;
    double *out_coef, *out_scalars, *out_probe_log;
};

__device__ __forceinline__ uint32_t our_rand_r(uint32_t &s) {  // sklearn/utils/_random.pxd:20-34
    if (s == 0) s = 1;
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s & 0x7fffffffu;  // % (RAND_R_MAX + 1)
}
__device__ __forceinline__ uint32_t rand_int(uint32_t end, uint32_t &s) { return our_rand_r(s) % end; }

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct Ctl {        // CTA-wide scalars published by thread 0
    double gap, dual_norm;
    double bc[2][2];  // per-step broadcast: {delta, |w_new|}
    int n_active, nnz, n_zero;
};

__global__ void __launch_bounds__(LT, 1) lasso_select_kernel(const SelectParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = P.c, tid = threadIdx.x;
    double *w = reinterpret_cast<double *>(smem_raw);
    double *Qw = w + c;
    double *qv = Qw + c;
    double *dg = qv + c;
    double *XtA = dg + c;
    double *ring = XtA + c;                                       // RING * c
    uint32_t *active = reinterpret_cast<uint32_t *>(ring + (size_t)RING * c);
    uint32_t *zlist = active + c;                                  // features zeroed by screening
    uint8_t *excluded = reinterpret_cast<uint8_t *>(zlist + c);
    __shared__ Ctl ctl;

    const double *__restrict__ Q = P.Q;
    const double yn2 = *P.yn2;
    for (int e = tid; e < c; e += LT) {
        w[e] = 0.0;
        qv[e] = P.qv[e];
        dg[e] = Q[(int64_t)e * c + e];
    }
    __syncthreads();

    const double tolS = __dmul_rn(P.tol, yn2);
    int probe = 0;
    int status = 0;

    // ---- one Lasso.fit (warm start) at l1 = alpha*m; returns nnz (uniform across threads)
    auto solve = [&](double alpha_user) -> int {
        const double l1 = __dmul_rn(alpha_user, P.m);
        uint32_t state = P.seeds[probe];
        int n_active = c;
        int n_iter_ret = 0;
        double gap = 0.0;

        // Qw = sum_j w[j] * Q[j,:]  (model: serial daxpy per nonzero j)
        for (int e = tid; e < c; e += LT) Qw[e] = 0.0;
        for (int j = 0; j < c; ++j) {
            const double wj = w[j];
            if (wj != 0.0)
                for (int e = tid; e < c; e += LT) Qw[e] = __dadd_rn(Qw[e], __dmul_rn(wj, Q[(int64_t)j * c + e]));
        }
        __syncthreads();

        // gap_enet_gram + dual_gap_formulation_A (beta = 0), serial like the model
        auto gap_check = [&]() {
            if (tid == 0) {
                double q_dot_w = 0.0, wQw = 0.0, dn = 0.0, l1n = 0.0;
                for (int j = 0; j < c; ++j) q_dot_w = __dadd_rn(q_dot_w, __dmul_rn(w[j], qv[j]));
                for (int j = 0; j < c; ++j) wQw = __dadd_rn(wQw, __dmul_rn(w[j], Qw[j]));
                const double R_norm2 = __dadd_rn(__dadd_rn(yn2, wQw), -__dmul_rn(2.0, q_dot_w));
                const double Ry = __dadd_rn(yn2, -q_dot_w);
                for (int j = 0; j < c; ++j) {
                    const double x = __dadd_rn(qv[j], -Qw[j]);
                    XtA[j] = x;
                    const double ax = fabs(x);
                    if (j == 0 || ax > dn) dn = ax;
                    l1n = __dadd_rn(l1n, fabs(w[j]));
                }
                const double primal = __dadd_rn(__dmul_rn(0.5, R_norm2), __dmul_rn(l1, l1n));
                const double scale = dn > l1 ? __ddiv_rn(l1, dn) : 1.0;
                const double dual = __dadd_rn(__dmul_rn(__dmul_rn(-0.5, __dmul_rn(scale, scale)), R_norm2),
                                              __dmul_rn(scale, Ry));
                ctl.gap = __dadd_rn(primal, -dual);
                ctl.dual_norm = dn;
            }
            __syncthreads();
        };
        // gap-safe screening (model: radius = sqrt(2|gap|)/alpha; d_j = (1-|XtA_j/max(alpha,dn)|)/sqrt(Q_jj))
        auto screen = [&](bool first) {
            const double radius = __ddiv_rn(sqrt(__dmul_rn(2.0, fabs(ctl.gap))), l1);
            const double den = l1 > ctl.dual_norm ? l1 : ctl.dual_norm;
            for (int j = tid; j < c; j += LT) {
                uint8_t ex;
                if (first) {
                    if (dg[j] == 0.0) ex = 2;  // zero column: w[j] = 0, excluded (no Qw change needed)
                    else {
                        const double th = __ddiv_rn(XtA[j], den);
                        const double d_j = __ddiv_rn(__dadd_rn(1.0, -fabs(th)), sqrt(dg[j]));
                        ex = d_j <= radius ? 0 : 1;
                    }
                } else if (excluded[j]) ex = 3;  // stays excluded
                else {
                    const double th = __ddiv_rn(XtA[j], den);
                    const double d_j = __ddiv_rn(__dadd_rn(1.0, -fabs(th)), sqrt(dg[j]));
                    ex = d_j <= radius ? 0 : 1;
                }
                excluded[j] = ex;
            }
            __syncthreads();
            if (tid == 0) {
                int na = 0, nz = 0;
                for (int j = 0; j < c; ++j) {
                    const uint8_t ex = excluded[j];
                    if (ex == 0) active[na++] = j;
                    else {
                        if (ex == 1 && w[j] != 0.0) zlist[nz++] = j;
                        if (ex == 2) w[j] = 0.0;
                        excluded[j] = 1;
                    }
                }
                ctl.n_active = na;
                ctl.n_zero = nz;
            }
            __syncthreads();
            const int nz = ctl.n_zero;
            for (int z = 0; z < nz; ++z) {  // Qw -= w[j] * Q[j,:], in ascending j like the model
                const int j = zlist[z];
                const double a = -w[j];
                for (int e = tid; e < c; e += LT) Qw[e] = __dadd_rn(Qw[e], __dmul_rn(a, Q[(int64_t)j * c + e]));
            }
            __syncthreads();
            if (tid == 0)
                for (int z = 0; z < nz; ++z) w[zlist[z]] = 0.0;
            __syncthreads();
            n_active = ctl.n_active;
        };

        gap_check();
        gap = ctl.gap;
        bool converged0 = gap <= tolS;
        if (!converged0) {
            for (int j = tid; j < c; j += LT) excluded[j] = 0;
            __syncthreads();
            screen(true);
            bool broke = false;
            int n_iter = 0;
            for (n_iter = 0; n_iter < P.max_iter; ++n_iter) {
                double w_max = 0.0, d_w_max = 0.0;
                // prime the ring along the coordinate sequence of this sweep
                uint32_t la = state;
                const int nprime = n_active < RING ? n_active : RING;
                for (int d = 0; d < RING; ++d) {
                    if (d < nprime) {
                        const uint32_t jd = active[rand_int((uint32_t)n_active, la)];
                        const double *src = Q + (int64_t)jd * c;
                        double *dst = ring + (size_t)d * c;
                        for (int e = tid; e < c; e += LT) cp_async8(dst + e, src + e);
                    }
                    cp_async_commit();
                }
                for (int f = 0; f < n_active; ++f) {
                    const uint32_t j = active[rand_int((uint32_t)n_active, state)];
                    const int slot = f % RING;
                    if ((int)(j % LT) == tid) {  // owner of coordinate j
                        const double Qjj = dg[j];
                        double delta = 0.0, aw = -1.0;  // aw < 0 flags the "Qjj == 0: continue" case
                        if (Qjj != 0.0) {
                            const double w_j = w[j];
                            const double tmp = __dadd_rn(__dadd_rn(qv[j], -Qw[j]), __dmul_rn(w_j, Qjj));
                            const double sgn = tmp == 0.0 ? 0.0 : (tmp > 0.0 ? 1.0 : -1.0);
                            const double mag = __dadd_rn(fabs(tmp), -l1);
                            const double w_new = __ddiv_rn(__dmul_rn(sgn, mag > 0.0 ? mag : 0.0), Qjj);
                            w[j] = w_new;
                            delta = __dadd_rn(w_new, -w_j);
                            aw = fabs(w_new);
                        }
                        ctl.bc[f & 1][0] = delta;
                        ctl.bc[f & 1][1] = aw;
                    }
                    cp_async_wait<RING - 1>();  // this thread's copies of the row for step f have landed
                    __syncthreads();
                    const double delta = ctl.bc[f & 1][0], aw = ctl.bc[f & 1][1];
                    if (aw >= 0.0) {
                        if (delta != 0.0) {
                            const double *row = ring + (size_t)slot * c;
                            for (int e = tid; e < c; e += LT) Qw[e] = __dadd_rn(Qw[e], __dmul_rn(delta, row[e]));
                        }
                        const double d = fabs(delta);
                        if (d > d_w_max) d_w_max = d;
                        if (aw > w_max) w_max = aw;
                    }
                    if (f + RING < n_active) {
                        const uint32_t jn = active[rand_int((uint32_t)n_active, la)];
                        const double *src = Q + (int64_t)jn * c;
                        double *dst = ring + (size_t)slot * c;
                        for (int e = tid; e < c; e += LT) cp_async8(dst + e, src + e);
                    }
                    cp_async_commit();
                }
                cp_async_wait<0>();
                __syncthreads();
                if (w_max == 0.0 || __ddiv_rn(d_w_max, w_max) <= P.tol || n_iter == P.max_iter - 1) {
                    gap_check();
                    gap = ctl.gap;
                    if (gap <= tolS) { broke = true; break; }
                    screen(false);
                }
            }
            n_iter_ret = broke ? n_iter + 1 : P.max_iter;
        }
        // nnz
        if (tid == 0) {
            int nnz = 0;
            for (int j = 0; j < c; ++j) nnz += (w[j] != 0.0);
            ctl.nnz = nnz;
            double *lg = P.out_probe_log + (size_t)probe * 4;
            lg[0] = alpha_user; lg[1] = (double)nnz; lg[2] = (double)n_iter_ret; lg[3] = gap;
        }
        __syncthreads();
        ++probe;
        return ctl.nnz;
    };

    // ---- alpha search, reference lib/decompose.py:489-525
    double left = 0.0, right = P.right0, alpha = P.right0;
    int nnz = 0;
    bool done = false;
    while (true) {  // :502-515  relax right until fewer than `rank` survive
        if (probe >= P.max_probes) { status = 1; done = true; break; }
        alpha = right;
        nnz = solve(right);
        if (nnz < P.rank) break;
        right = __dmul_rn(right, 2.0);
    }
    while (!done) {  // :516-525  bisection into [lbound, rbound]
        if (probe >= P.max_probes) { status = 1; break; }
        alpha = __dmul_rn(__dadd_rn(left, right), 0.5);
        nnz = solve(alpha);
        if ((double)nnz > P.rbound) left = alpha;
        else if ((double)nnz < P.lbound) right = alpha;
        else break;
    }
    for (int e = tid; e < c; e += LT) {
        P.out_idxs[e] = w[e] != 0.0 ? 1 : 0;
        P.out_coef[e] = w[e];
    }
    if (tid == 0) {
        P.out_scalars[0] = alpha;
        P.out_scalars[1] = (double)probe;
        P.out_scalars[2] = (double)status;
        P.out_scalars[3] = (double)nnz;
    }
}

}  // namespace

extern "C" int cp_lasso_build(cp_handle_t h, const double *Gs, const double *Bs, const double *sxs,
                              const double *sys, const double *yys, const double *WW, const double *sw,
                              const float *W2, int c, int k2, int n, int S, double *Q, double *qv, double *yn2,
                              cp_stream_t stream_) {
    CP_REQUIRE(h && Gs && Bs && sxs && sys && yys && WW && sw && W2 && Q && qv && yn2, "cp_lasso_build: NULL argument");
    CP_REQUIRE(c > 0 && k2 > 0 && n > 0 && S > 0, "cp_lasso_build: bad shape");
    cudaStream_t stream = (cudaStream_t)stream_;
    const double m = (double)S * (double)n;
    dim3 grid(cp_cdiv(c, 16), cp_cdiv(c, 16));
    lasso_build_Q<<<grid, 256, 0, stream>>>(Gs, WW, sxs, sw, c, k2, m, Q);
    CP_CHECK_LAUNCH();
    lasso_build_q<<<c, 128, 0, stream>>>(W2, Bs, sxs, sw, sys, yys, c, k2, n, m, qv, yn2);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_lasso_select(cp_handle_t h, const double *Q, const double *qv, const double *yn2, int c, double m,
                               int rank, double lbound, double rbound, double right0, double tol, int max_iter,
                               const uint32_t *seeds, int max_probes, uint8_t *out_idxs, double *out_coef,
                               double *out_scalars, double *out_probe_log, cp_stream_t stream_) {
    CP_REQUIRE(h && Q && qv && yn2 && seeds && out_idxs && out_coef && out_scalars && out_probe_log,
               "cp_lasso_select: NULL argument");
    CP_REQUIRE(c > 0 && c <= MAXC, "cp_lasso_select: c=%d outside 1..%d", c, MAXC);
    CP_REQUIRE(max_probes > 0 && max_iter > 0 && right0 > 0 && m > 0, "cp_lasso_select: bad parameters");
    SelectParams P{Q, qv, yn2, c, m, rank, lbound, rbound, right0, tol, max_iter, seeds, max_probes,
                   out_idxs, out_coef, out_scalars, out_probe_log};
    const size_t smem = (size_t)c * (5 + RING) * sizeof(double) + (size_t)c * (2 * sizeof(uint32_t) + 1) + 16;
    static bool configured = false;
    if (!configured) {
        CP_CUDA(cudaFuncSetAttribute(lasso_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
        configured = true;
    }
    CP_REQUIRE(smem <= 227 * 1024 - 256, "cp_lasso_select: c=%d needs %zu bytes of shared memory", c, smem);
    lasso_select_kernel<<<1, LT, smem, (cudaStream_t)stream_>>>(P);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
