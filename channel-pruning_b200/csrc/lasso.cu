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
// |yc|^2).  The kernel reproduces that model BIT FOR BIT: every floating-point operation
// the model rounds separately is issued with an explicit *_rn intrinsic (no FMA
// contraction), the gap reductions run in the model's "warp order", the division is
// correctly rounded, and the random coordinate order comes from the same 32-bit xorshift.
//
// Coordinate descent is one long serial dependency chain (each coordinate update needs the
// previous one), so the whole search is latency bound; it runs in ONE WARP per problem:
//   * Qw (= Q w) lives in shared memory, each lane owning interleaved pairs of elements;
//     a step applies  Qw += delta * Q[j,:]  with LDS.128 / DMUL / DADD / STS.128 per pair;
//   * the pair that holds the NEXT coordinate is updated first and its value is broadcast
//     with a shuffle, so the serial chain (update -> soft threshold -> divide -> delta) of
//     step f+1 is issued in the shadow of step f's remaining pair updates;
//   * rows of Q stream from L2 through an 8-deep cp.async ring that runs ahead along the
//     (data-independent) random coordinate sequence; every lane copies exactly the pairs it
//     will read back, so no barrier is needed for the ring;
//   * the division  num / Q_jj  uses a pre-computed correctly rounded reciprocal and one
//     Markstein correction step (DMUL + 2 DFMA) -- IEEE-exact, 3 dependent ops instead of ~10.
// No CTA barrier anywhere; __syncwarp() only orders the lane-0 scalar stores.
#include "common.cuh"

namespace {

// prefetch depth (rows of Q in flight): 16 while the ring fits next to the state, 8 for c > 1024
template <int NP> struct RingDepth { static constexpr int value = NP <= 16 ? 16 : 8; };
constexpr int MAXC = 2048;  // largest channel count (shared-memory bound)

// ------------------------------------------------------------------ build
__global__ void __launch_bounds__(256)
lasso_build_Q(const double *__restrict__ Gs, const double *__restrict__ WW, const double *__restrict__ sxs,
              const double *__restrict__ sw, int c, int k2, double m, double *__restrict__ Q, int ldq) {
    const int b = blockIdx.x * 16 + (threadIdx.x & 15);
    const int a = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (a >= c || b >= ldq) return;
    if (b >= c) {  // padding column (keeps rows 16-byte aligned for the selection kernel)
        Q[(int64_t)a * ldq + b] = 0.0;
        return;
    }
    const int64_t K = (int64_t)c * k2;
    double s = 0.0, za = 0.0, zb = 0.0;
    for (int p = 0; p < k2; ++p) {
        const double *g = Gs + ((int64_t)a * k2 + p) * K + (int64_t)b * k2;
        const double *w = WW + ((int64_t)a * k2 + p) * K + (int64_t)b * k2;
        for (int q = 0; q < k2; ++q) s = fma(g[q], w[q], s);
        za = fma(sxs[a * k2 + p], sw[a * k2 + p], za);
        zb = fma(sxs[b * k2 + p], sw[b * k2 + p], zb);
    }
    Q[(int64_t)a * ldq + b] = s - za * zb / m;  // - m * zbar_a * zbar_b
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
    const double *Q;
    int ldq;
    const double *qv, *yn2;
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

__device__ __forceinline__ uint32_t xorshift(uint32_t &s) {  // sklearn/utils/_random.pxd:20-34
    if (s == 0) s = 1;
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s & 0x7fffffffu;  // % (RAND_R_MAX + 1)
}
// a % d for 32-bit a through a pre-computed M = floor(2^64 / d) + 1 (Lemire's fastmod; exact)
__device__ __forceinline__ uint32_t fastmod(uint32_t a, uint64_t M, uint32_t d) {
    return (uint32_t)__umul64hi(M * (uint64_t)a, (uint64_t)d);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ double warp_sum_butterfly(double v) {  // model: p[l] + p[l ^ off], off = 16..1
#pragma unroll
    for (int off = 16; off; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// correctly rounded num / d given rc = RN(1/d):  q0 = RN(num*rc), r = num - q0*d (exact), q = RN(q0 + r*rc)
__device__ __forceinline__ double div_markstein(double num, double d, double rc) {
    const double q0 = __dmul_rn(num, rc);
    const double r = __fma_rn(-q0, d, num);
    return __fma_rn(r, rc, q0);
}

template <int NP>  // pairs per lane; padded channel count CP = 64 * NP
__global__ void __launch_bounds__(32, 1) lasso_select_kernel(const SelectParams P) {
    constexpr int CP = 64 * NP;
    constexpr int RING = RingDepth<NP>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = P.c, lane = threadIdx.x;
    double *w = reinterpret_cast<double *>(smem_raw);  // [CP]
    double *Qw = w + CP;
    double *qv = Qw + CP;
    double *dg = qv + CP;
    double *rc = dg + CP;
    double *ring = rc + CP;  // [RING][CP]
    uint32_t *active = reinterpret_cast<uint32_t *>(ring + (size_t)RING * CP);
    uint32_t *zlist = active + CP;
    uint32_t *jq = zlist + CP;  // [RING] upcoming coordinates
    uint8_t *excluded = reinterpret_cast<uint8_t *>(jq + RING);

    const double *__restrict__ Q = P.Q;
    const int ldq = P.ldq;
    const double yn2 = *P.yn2;
    for (int e = lane; e < CP; e += 32) {
        w[e] = 0.0;
        Qw[e] = 0.0;
        const bool in = e < c;
        const double d = in ? Q[(int64_t)e * ldq + e] : 0.0;
        qv[e] = in ? P.qv[e] : 0.0;
        dg[e] = d;
        rc[e] = d != 0.0 ? __drcp_rn(d) : 0.0;
        active[e] = e;
        excluded[e] = 0;
    }
    for (int e = lane; e < RING * CP; e += 32) ring[e] = 0.0;  // padding pairs stay zero
    __syncwarp();

    const double tolS = __dmul_rn(P.tol, yn2);
    int probe = 0, status = 0;

    // this lane's pair `s` lives at element 64*s + 2*lane
    auto prefetch_row = [&](uint32_t j, int slot) {
        const double *src = Q + (int64_t)j * ldq;
        double *dst = ring + (size_t)slot * CP;
#pragma unroll
        for (int s = 0; s < NP; ++s) {
            const int e = 64 * s + 2 * lane;
            if (e < c) cp_async16(dst + e, src + e);
        }
    };

    // Qw += a * Q[j,:], row fetched directly (rare paths: screening evictions)
    auto axpy_row_direct = [&](uint32_t j, double a) {
        const double *src = Q + (int64_t)j * ldq;
#pragma unroll
        for (int s = 0; s < NP; ++s) {
            const int e = 64 * s + 2 * lane;
            if (e < c) {
                const double2 r = *reinterpret_cast<const double2 *>(src + e);
                double2 v = *reinterpret_cast<double2 *>(Qw + e);
                v.x = __dadd_rn(v.x, __dmul_rn(a, r.x));
                v.y = __dadd_rn(v.y, __dmul_rn(a, r.y));
                *reinterpret_cast<double2 *>(Qw + e) = v;
            }
        }
    };

    // ---- one Lasso.fit (warm start) at l1 = alpha*m; returns nnz (uniform across lanes)
    auto solve = [&](double alpha_user) -> int {
        const double l1 = __dmul_rn(alpha_user, P.m);
        uint32_t la = P.seeds[probe];  // coordinate-order RNG, runs RING draws ahead of the update
        int n_active = c;
        int n_iter_ret = 0;
        double gap = 0.0, dual_norm = 0.0;

        // gap_enet_gram + dual_gap_formulation_A (beta = 0) in warp order
        auto gap_check = [&]() {
            double a1 = 0.0, a2 = 0.0, a3 = 0.0, dn = 0.0;
            for (int i = lane; i < c; i += 32) {
                const double wi = w[i];
                a1 = __dadd_rn(a1, __dmul_rn(wi, qv[i]));
                a2 = __dadd_rn(a2, __dmul_rn(wi, Qw[i]));
                a3 = __dadd_rn(a3, fabs(wi));
                dn = fmax(dn, fabs(__dadd_rn(qv[i], -Qw[i])));
            }
            const double q_dot_w = warp_sum_butterfly(a1);
            const double wQw = warp_sum_butterfly(a2);
            const double l1n = warp_sum_butterfly(a3);
            dn = warp_max(dn);
            const double R_norm2 = __dadd_rn(__dadd_rn(yn2, wQw), -__dmul_rn(2.0, q_dot_w));
            const double Ry = __dadd_rn(yn2, -q_dot_w);
            const double primal = __dadd_rn(__dmul_rn(0.5, R_norm2), __dmul_rn(l1, l1n));
            const double scale = dn > l1 ? __ddiv_rn(l1, dn) : 1.0;
            const double dual = __dadd_rn(__dmul_rn(__dmul_rn(-0.5, __dmul_rn(scale, scale)), R_norm2),
                                          __dmul_rn(scale, Ry));
            gap = __dadd_rn(primal, -dual);
            dual_norm = dn;
        };
        // gap-safe screening: keep j iff (1 - |XtA_j / max(l1, dn)|) / sqrt(Q_jj) <= sqrt(2|gap|) / l1
        auto screen = [&](bool first) {
            const double radius = __ddiv_rn(sqrt(__dmul_rn(2.0, fabs(gap))), l1);
            const double den = l1 > dual_norm ? l1 : dual_norm;
            int na = 0, nz = 0;
            for (int base = 0; base < c; base += 32) {
                const int j = base + lane;
                bool keep = false, evict_nonzero = false;
                if (j < c) {
                    if (first && dg[j] == 0.0) {
                        w[j] = 0.0;  // zero column
                        excluded[j] = 1;
                    } else if (!first && excluded[j]) {
                        // stays excluded
                    } else {
                        const double th = __ddiv_rn(__dadd_rn(qv[j], -Qw[j]), den);
                        const double d_j = __ddiv_rn(__dadd_rn(1.0, -fabs(th)), sqrt(dg[j]));
                        keep = d_j <= radius;
                        if (!keep) {
                            evict_nonzero = w[j] != 0.0;
                            excluded[j] = 1;
                        } else excluded[j] = 0;
                    }
                }
                const uint32_t mk = __ballot_sync(0xffffffffu, keep);
                const uint32_t mz = __ballot_sync(0xffffffffu, evict_nonzero);
                const uint32_t lt = (1u << lane) - 1u;
                if (keep) active[na + __popc(mk & lt)] = j;
                if (evict_nonzero) zlist[nz + __popc(mz & lt)] = j;
                na += __popc(mk);
                nz += __popc(mz);
            }
            __syncwarp();
            for (int z = 0; z < nz; ++z) {  // Qw -= w[j] * Q[j,:], ascending j like the model
                const uint32_t j = zlist[z];
                axpy_row_direct(j, -w[j]);
            }
            __syncwarp();
            for (int z = lane; z < nz; z += 32) w[zlist[z]] = 0.0;
            __syncwarp();
            n_active = na;
        };

        gap_check();
        if (!(gap <= tolS)) {
            screen(true);
            bool broke = false;
            int n_iter = 0;
            for (n_iter = 0; n_iter < P.max_iter; ++n_iter) {
                double w_max = 0.0, d_w_max = 0.0;
                if (n_active > 0) {
                    const uint64_t M = ~0ull / (uint32_t)n_active + 1ull;
                    // prime the ring and the coordinate queue
#pragma unroll
                    for (int d = 0; d < RING; ++d) {
                        if (d < n_active) {
                            const uint32_t jd = active[fastmod(xorshift(la), M, (uint32_t)n_active)];
                            if (lane == 0) jq[d] = jd;
                            prefetch_row(jd, d);
                        }
                        cp_async_commit();
                    }
                    __syncwarp();
                    // soft-threshold update of one coordinate from the current x = Qw[j]; all lanes redundantly
                    auto cd_update = [&](double qj, double Qjj, double rj, double x, double w_j, double &delta,
                                         double &aw, double &w_new) {
                        const double tmp = __dadd_rn(__dadd_rn(qj, -x), __dmul_rn(w_j, Qjj));
                        const double mag = __dadd_rn(fabs(tmp), -l1);
                        // fsign(tmp) * fmax(|tmp| - l1, 0) / Qjj  (Qjj > 0): signed zero when thresholded away
                        const double wn = mag > 0.0 ? div_markstein(copysign(mag, tmp), Qjj, rj) : (tmp < 0.0 ? -0.0 : 0.0);
                        const bool live = Qjj != 0.0;  // model: "if Qjj == 0: continue"
                        w_new = live ? wn : w_j;
                        delta = live ? __dadd_rn(wn, -w_j) : 0.0;
                        aw = live ? fabs(wn) : -1.0;
                    };
                    // value of this lane's pair in slot(jx) that corresponds to column jx (garbage unless owner)
                    auto own_elem = [&](const double *base, uint32_t jx) -> double {
                        const double2 v = *reinterpret_cast<const double2 *>(base + 64 * (jx >> 6) + 2 * lane);
                        return (jx & 1) ? v.y : v.x;
                    };
                    // ---- software pipeline: the serial chain of step f+1 runs in the shadow of step f's pair updates.
                    // w[] is read and written by lane 0 only during a sweep (broadcast by shuffle): no races.
                    cp_async_wait<RING - 1>();  // row 0 (own pairs)
                    uint32_t j = jq[0];
                    double delta, aw, w_new;
                    cd_update(qv[j], dg[j], rc[j], Qw[j], __shfl_sync(0xffffffffu, lane == 0 ? w[j] : 0.0, 0), delta, aw,
                              w_new);
                    if (lane == 0) w[j] = w_new;
                    // package of step 1
                    bool valid_n = n_active > 1;
                    uint32_t jn = valid_n ? jq[1] : j;
                    double qn = qv[jn], dn_ = dg[jn], rn = rc[jn];
                    double wn_ = __shfl_sync(0xffffffffu, lane == 0 ? w[jn] : 0.0, 0);
                    double xo = __shfl_sync(0xffffffffu, own_elem(Qw, jn), (jn >> 1) & 31);
                    double rr = __shfl_sync(0xffffffffu, own_elem(ring, jn), (jn >> 1) & 31);  // row 0 sits in slot 0
                    for (int f = 0; f < n_active; ++f) {
                        const int slot = f & (RING - 1);
                        cp_async_wait<RING - 2>();  // rows <= f+1 have landed (own pairs)
                        // (1) serial chain of step f+1
                        const double xnew = __dadd_rn(xo, __dmul_rn(delta, rr));
                        double delta_n, aw_n, w_new_n;
                        cd_update(qn, dn_, rn, xnew, wn_, delta_n, aw_n, w_new_n);
                        if (valid_n && lane == 0) w[jn] = w_new_n;
                        // (2) pair updates of step f:  Qw += delta * Q[j_f, :]
                        const double *row = ring + (size_t)slot * CP;
#pragma unroll
                        for (int s = 0; s < NP; ++s) {
                            const int e = 64 * s + 2 * lane;
                            const double2 r = *reinterpret_cast<const double2 *>(row + e);
                            double2 v = *reinterpret_cast<double2 *>(Qw + e);
                            v.x = __dadd_rn(v.x, __dmul_rn(delta, r.x));
                            v.y = __dadd_rn(v.y, __dmul_rn(delta, r.y));
                            *reinterpret_cast<double2 *>(Qw + e) = v;
                        }
                        // (3) bookkeeping of step f
                        {
                            const double d = fabs(delta);
                            d_w_max = (aw >= 0.0 && d > d_w_max) ? d : d_w_max;
                            w_max = (aw > w_max) ? aw : w_max;
                        }
                        // (4) package of step f+2 (needs the pair updates above and row f+1)
                        const bool valid_nn = f + 2 < n_active;
                        const uint32_t jnn = valid_nn ? jq[(f + 2) & (RING - 1)] : jn;
                        const double qnn = qv[jnn], dnn = dg[jnn], rnn = rc[jnn];
                        const double wnn = __shfl_sync(0xffffffffu, lane == 0 ? w[jnn] : 0.0, 0);
                        const double xo2 = __shfl_sync(0xffffffffu, own_elem(Qw, jnn), (jnn >> 1) & 31);
                        const double rr2 = __shfl_sync(0xffffffffu, own_elem(ring + (size_t)((f + 1) & (RING - 1)) * CP, jnn),
                                                       (jnn >> 1) & 31);
                        // (5) refill the slot just consumed with the row of step f + RING
                        if (f + RING < n_active) {
                            const uint32_t jf = active[fastmod(xorshift(la), M, (uint32_t)n_active)];
                            if (lane == 0) jq[slot] = jf;
                            prefetch_row(jf, slot);
                        }
                        cp_async_commit();
                        __syncwarp();
                        j = jn; delta = delta_n; aw = aw_n;
                        jn = jnn; valid_n = valid_nn; qn = qnn; dn_ = dnn; rn = rnn; wn_ = wnn; xo = xo2; rr = rr2;
                    }
                    cp_async_wait<0>();
                }
                if (w_max == 0.0 || __ddiv_rn(d_w_max, w_max) <= P.tol || n_iter == P.max_iter - 1) {
                    gap_check();
                    if (gap <= tolS) { broke = true; break; }
                    screen(false);
                }
            }
            n_iter_ret = broke ? n_iter + 1 : P.max_iter;
        }
        int cnt = 0;
        for (int i = lane; i < c; i += 32) cnt += (w[i] != 0.0);
#pragma unroll
        for (int off = 16; off; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) {
            double *lg = P.out_probe_log + (size_t)probe * 4;
            lg[0] = alpha_user; lg[1] = (double)cnt; lg[2] = (double)n_iter_ret; lg[3] = gap;
        }
        ++probe;
        return cnt;
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
    for (int e = lane; e < c; e += 32) {
        P.out_idxs[e] = w[e] != 0.0 ? 1 : 0;
        P.out_coef[e] = w[e];
    }
    if (lane == 0) {
        P.out_scalars[0] = alpha;
        P.out_scalars[1] = (double)probe;
        P.out_scalars[2] = (double)status;
        P.out_scalars[3] = (double)nnz;
    }
}

template <int NP>
int launch_select(const SelectParams &P, cudaStream_t stream) {
    constexpr int CP = 64 * NP;
    constexpr int RING = RingDepth<NP>::value;
    const size_t smem = (size_t)CP * (5 + RING) * sizeof(double) + (size_t)CP * (2 * sizeof(uint32_t) + 1) +
                        RING * sizeof(uint32_t) + 16;
    static bool configured = false;
    if (!configured) {
        CP_CUDA(cudaFuncSetAttribute(lasso_select_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    lasso_select_kernel<NP><<<1, 32, smem, stream>>>(P);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

}  // namespace

extern "C" int cp_lasso_build(cp_handle_t h, const double *Gs, const double *Bs, const double *sxs,
                              const double *sys, const double *yys, const double *WW, const double *sw,
                              const float *W2, int c, int k2, int n, int S, double *Q, int ldq, double *qv,
                              double *yn2, cp_stream_t stream_) {
    CP_REQUIRE(h && Gs && Bs && sxs && sys && yys && WW && sw && W2 && Q && qv && yn2, "cp_lasso_build: NULL argument");
    CP_REQUIRE(c > 0 && k2 > 0 && n > 0 && S > 0, "cp_lasso_build: bad shape");
    CP_REQUIRE(ldq >= c && (ldq % 2) == 0, "cp_lasso_build: ldq must be even and >= c (got %d)", ldq);
    cudaStream_t stream = (cudaStream_t)stream_;
    const double m = (double)S * (double)n;
    dim3 grid(cp_cdiv(ldq, 16), cp_cdiv(c, 16));
    lasso_build_Q<<<grid, 256, 0, stream>>>(Gs, WW, sxs, sw, c, k2, m, Q, ldq);
    CP_CHECK_LAUNCH();
    lasso_build_q<<<c, 128, 0, stream>>>(W2, Bs, sxs, sw, sys, yys, c, k2, n, m, qv, yn2);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_lasso_select(cp_handle_t h, const double *Q, int ldq, const double *qv, const double *yn2, int c,
                               double m, int rank, double lbound, double rbound, double right0, double tol,
                               int max_iter, const uint32_t *seeds, int max_probes, uint8_t *out_idxs,
                               double *out_coef, double *out_scalars, double *out_probe_log, cp_stream_t stream_) {
    CP_REQUIRE(h && Q && qv && yn2 && seeds && out_idxs && out_coef && out_scalars && out_probe_log,
               "cp_lasso_select: NULL argument");
    CP_REQUIRE(c > 0 && c <= MAXC, "cp_lasso_select: c=%d outside 1..%d", c, MAXC);
    CP_REQUIRE(ldq >= c && (ldq % 2) == 0 && (((uintptr_t)Q) & 15) == 0,
               "cp_lasso_select: Q rows must be 16-byte aligned (even ldq >= c, aligned base)");
    CP_REQUIRE(max_probes > 0 && max_iter > 0 && right0 > 0 && m > 0, "cp_lasso_select: bad parameters");
    SelectParams P{Q, ldq, qv, yn2, c, m, rank, lbound, rbound, right0, tol, max_iter, seeds, max_probes,
                   out_idxs, out_coef, out_scalars, out_probe_log};
    cudaStream_t stream = (cudaStream_t)stream_;
    if (c <= 64) return launch_select<1>(P, stream);
    if (c <= 128) return launch_select<2>(P, stream);
    if (c <= 256) return launch_select<4>(P, stream);
    if (c <= 512) return launch_select<8>(P, stream);
    if (c <= 1024) return launch_select<16>(P, stream);
    return launch_select<32>(P, stream);
}
