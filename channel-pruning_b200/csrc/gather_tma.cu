// Sparse-point im2col on NHWC feature maps, TMA in / TMA out -- the HBM-roofline path of cp_patch_gather
// (replaces Net.extract_XY, reference lib/net.py:534-684, + the relu of lib/net.py:1720).
//
// One sampled output point needs a k x k x c window of the bottom blob.  In NHWC that window is k runs of k*c
// contiguous floats (6 KB at c = 512): a 4-D tensor map over (c, W, H, image) with box (c_box, k, k, 1) lets ONE
// cp.async.bulk.tensor request fetch it, zero-filling the taps that fall into the padding (out-of-range coordinates,
// net.py:631-632), and the finished patch row (K = c k k contiguous floats of X) leaves through a bulk shared->global
// copy.  Persistent CTAs (as many per SM as shared memory allows: the latency of one row -- TMA flight, two CTA-wide
// hand-offs, the transposition -- is hidden by the other CTAs of the SM) each keep a ring of windows in flight:
//     producer warp     the 32 lanes prefetch the sampled coordinates of the next 32 rows (one global-load latency per
//                       32 rows instead of per row); one lane arms the mbarrier and issues the TMA loads (ring of NS stages)
//     128 consumers     [tap][channel] -> [channel][tap] (the reference's column order a*k*k + p) with the ReLU folded in,
//                       conflict-free both ways (lanes walk channels; the tap stride k*k is odd)
//     one consumer      bulk store of the row, two rows in flight
// Bytes: the window is read once and the row written once -- 8 N K bytes, the algorithmic figure of SURVEY.md 8(d).
// Bound: HBM.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int GT_CONS = 128;             // consumer threads
constexpr int GT_THREADS = GT_CONS + 32; // + producer warp
constexpr int GT_OUT = 2;                // output rows in flight

struct GtParams {
    const int32_t *randx, *randy;
    float *X;
    int64_t ldx, rows;
    int B, P, c, k, pad, stride, relu, cbox, nbox, nstage;
    int box_f, stage_f, out_f;  // strides in floats, each a multiple of 32 (TMA destinations are 128-byte aligned)
};

__device__ __forceinline__ uint32_t g_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug traps
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void g_tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void g_bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_cons_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GT_CONS) : "memory"); }

template <int K2>  // k*k known at compile time (1, 9, 25) or 0
__global__ void __launch_bounds__(GT_THREADS)
patch_gather_nhwc_tma(const __grid_constant__ CUtensorMap map, const GtParams P) {
    extern __shared__ __align__(128) unsigned char gsm_raw[];
    const int k2 = K2 > 0 ? K2 : P.k * P.k, K = P.c * k2;
    const uint32_t stage_bytes = (uint32_t)K * 4u;
    // layout: [nstage][K] input windows ([box][tap][c_box]), [GT_OUT][K] output rows, mbarriers
    unsigned char *base = (unsigned char *)(((uintptr_t)gsm_raw + 127) & ~(uintptr_t)127);
    float *in = reinterpret_cast<float *>(base);
    float *out = in + (size_t)P.nstage * P.stage_f;
    uint64_t *bars = reinterpret_cast<uint64_t *>(out + (size_t)GT_OUT * P.out_f);  // full[nstage], empty[nstage]
    const int tid = threadIdx.x;
    const uint32_t bar0 = g_smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bar0 + 8u * (uint32_t)(P.nstage + s); };
    if (tid == 0) {
        for (int s = 0; s < P.nstage; ++s) {
            g_mbar_init(full(s), 1);
            g_mbar_init(empty(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    if (tid >= GT_CONS) {
        // ---------------- producer warp
        const int lane = tid & 31;
        int it = 0;
        for (int64_t rb = first; rb < P.rows; rb += 32 * step) {
            // lane l looks up the window of row rb + l * step
            const int64_t r = rb + (int64_t)lane * step;
            int x0 = 0, y0 = 0, img = 0;
            if (r < P.rows) {
                const int img_in_batch = (int)(r % P.B);
                const int64_t bp = r / P.B;  // batch * P + point
                const int batch = (int)(bp / P.P);
                y0 = P.stride * P.randx[bp] - P.pad;  // window origin, rows  (feat[:,:,x,y]: x indexes H)
                x0 = P.stride * P.randy[bp] - P.pad;
                img = batch * P.B + img_in_batch;
            }
            const int64_t left = (P.rows - rb + step - 1) / step;
            const int nb = left < 32 ? (int)left : 32;
            for (int j = 0; j < nb; ++j, ++it) {
                const int xs = __shfl_sync(0xffffffffu, x0, j), ys = __shfl_sync(0xffffffffu, y0, j);
                const int is = __shfl_sync(0xffffffffu, img, j);
                if (lane == 0) {
                    const int s = it % P.nstage;
                    const uint32_t ph = (uint32_t)((it / P.nstage) & 1);
                    if (it >= P.nstage) g_mbar_wait(empty(s), ph ^ 1);  // the consumers have released the stage
                    g_mbar_expect_tx(full(s), stage_bytes);
                    const uint32_t dst = g_smem_u32(in + (size_t)s * P.stage_f);
                    for (int b = 0; b < P.nbox; ++b)
                        g_tma_load_4d(dst + (uint32_t)b * (uint32_t)P.box_f * 4u, &map, full(s), b * P.cbox, xs, ys, is);
                }
                __syncwarp();
            }
        }
        return;
    }
    // ---------------- consumers
    int it = 0;
    for (int64_t r = first; r < P.rows; r += step, ++it) {
        const int s = it % P.nstage, o = it % GT_OUT;
        const uint32_t ph = (uint32_t)((it / P.nstage) & 1);
        g_mbar_wait(full(s), ph);
        if (tid == 0 && it >= GT_OUT) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(GT_OUT - 1) : "memory");
        g_cons_barrier();  // out[o] is no longer being read by the store of row it - GT_OUT
        const float *src = in + (size_t)s * P.stage_f;
        float *dst = out + (size_t)o * P.out_f;
        for (int a = tid; a < P.c; a += GT_CONS) {
            const int b = a / P.cbox, al = a - b * P.cbox;
            const float *sp = src + (size_t)b * P.box_f + al;
            float *dp = dst + (size_t)a * k2;
            if (K2 > 0) {
                float v[K2 > 0 ? K2 : 1];
#pragma unroll
                for (int p = 0; p < K2; ++p) v[p] = sp[p * P.cbox];
#pragma unroll
                for (int p = 0; p < K2; ++p) dp[p] = P.relu ? fmaxf(v[p], 0.f) : v[p];
            } else {
                for (int p = 0; p < k2; ++p) {
                    float v = sp[(size_t)p * P.cbox];
                    if (P.relu) v = fmaxf(v, 0.f);
                    dp[p] = v;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
        g_cons_barrier();
        if (tid == 0) {
            g_mbar_arrive(empty(s));  // every consumer has finished reading in[s]
            g_bulk_store(P.X + r * P.ldx, g_smem_u32(dst), stage_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
}

inline size_t gt_round128(size_t b) { return (b + 127) & ~(size_t)127; }

typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// true when the TMA path applies (device memory, 16-byte rules); the caller falls back to the SIMT kernel otherwise
bool cp_gather_tma_eligible(const float *fmap, int c, int k, float *X_out, int64_t ldx) {
    if (c % 4 || c < 16 || k > 16) return false;
    if (((uintptr_t)fmap & 15) || ((uintptr_t)X_out & 15) || (ldx % 4)) return false;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, fmap) != cudaSuccess || pa.type != cudaMemoryTypeDevice) {
        (void)cudaGetLastError();
        return false;
    }
    int cbox = 0;
    for (int d = 256; d >= 16; d -= 4)
        if (c % d == 0) { cbox = d; break; }
    if (!cbox) return false;
    const size_t row = gt_round128((size_t)cbox * k * k * 4) * (c / cbox);
    return (2 + GT_OUT) * row + 1024 <= 200 * 1024;  // at least two input stages in one CTA
}

int cp_patch_gather_tma(cp_handle_t h, const float *fmap, int nbatch, int B, int c, int H, int W, const int32_t *randx,
                        const int32_t *randy, int P, int k, int pad, int stride, int relu, float *X_out, int64_t ldx,
                        cudaStream_t stream) {
    if (!h->tmap_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) CP_FAIL(CP_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        h->tmap_encode = fn;
    }
    int cbox = 0;
    for (int d = 256; d >= 16; d -= 4)
        if (c % d == 0) { cbox = d; break; }
    const int64_t nimg = (int64_t)nbatch * B;
    CUtensorMap map;
    const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)W * c * 4, (cuuint64_t)H * W * c * 4};
    const cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)k, (cuuint32_t)k, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = ((encode_fn_t)h->tmap_encode)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)fmap, dims, strides, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) CP_FAIL(CP_ERR_CUDA, "cuTensorMapEncodeTiled (4-D feature map) failed (%d)", (int)cr);
    GtParams Pm{};
    Pm.randx = randx; Pm.randy = randy; Pm.X = X_out; Pm.ldx = ldx;
    Pm.rows = (int64_t)nbatch * P * B;
    Pm.B = B; Pm.P = P; Pm.c = c; Pm.k = k; Pm.pad = pad; Pm.stride = stride; Pm.relu = relu;
    Pm.cbox = cbox; Pm.nbox = c / cbox;
    const size_t box_b = gt_round128((size_t)cbox * k * k * 4), row = box_b * Pm.nbox;
    const size_t out_b = gt_round128((size_t)c * k * k * 4);
    // CTAs per SM: as many as fit with >= 2 input stages each (up to 4), then the stages fill what is left
    const size_t budget = 216 * 1024;
    int per_sm = (int)(budget / (2 * row + GT_OUT * out_b + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    static const int env_per_sm = [] { const char *e = getenv("CPB200_GATHER_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    static const int env_stages = [] { const char *e = getenv("CPB200_GATHER_STAGES"); return e ? atoi(e) : 0; }();
    if (env_per_sm > 0 && env_per_sm < per_sm) per_sm = env_per_sm;  // tuning knobs (profiles/r2_run9.sh)
    int nstage = (int)((budget / per_sm - 1024 - GT_OUT * out_b) / row);
    if (nstage > 6) nstage = 6;
    if (env_stages >= 2 && env_stages < nstage) nstage = env_stages;
    Pm.nstage = nstage;
    Pm.box_f = (int)(box_b / 4); Pm.stage_f = (int)(row / 4); Pm.out_f = (int)(out_b / 4);
    const size_t smem = (size_t)nstage * row + GT_OUT * out_b + 2 * nstage * 8 + 256;
    auto kern = k == 3 ? patch_gather_nhwc_tma<9> : k == 1 ? patch_gather_nhwc_tma<1> : k == 5 ? patch_gather_nhwc_tma<25>
                                                                                              : patch_gather_nhwc_tma<0>;
    static cp_per_device_flag configured[4];
    const int which = k == 3 ? 0 : k == 1 ? 1 : k == 5 ? 2 : 3;
    if (bool *done = configured[which].slot(); !*done) {
        CP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        *done = true;
    }
    int64_t grid = (int64_t)h->num_sms * per_sm;
    if (grid > Pm.rows) grid = Pm.rows;
    kern<<<(unsigned)grid, GT_THREADS, smem, stream>>>(map, Pm);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
