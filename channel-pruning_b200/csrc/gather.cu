// Sparse-point gathers: the "im2col" of the pruning path.
//
//   cp_patch_gather  <- Net.extract_XY          (reference lib/net.py:534-684) + relu (:1720)
//   cp_point_gather  <- Net.extract_features    (reference lib/net.py:509-519)
//
// Both are pure data movement (HBM bound).  Row r of the output is
// (batch, point, image) = ((r / B) / P, (r / B) % P, r % B), the reference's order.
//
// NCHW path: one CTA per output row; consecutive threads write consecutive columns
// (a*k*k + py*k + px), so stores are fully coalesced; the loads are the sparse part
// (k floats per (channel,row) segment) and are served through L1/L2 at sector
// granularity -- that over-fetch is inherent to reading k-wide windows out of NCHW.
// NHWC path: a window row is k*c contiguous floats; the CTA stages the k*k x c tile
// in shared memory with coalesced loads (channel fastest) and writes it back
// transposed to (c, k*k) column order, again coalesced.
#include <cstdlib>

#include "common.cuh"

namespace {

template <int KS>
__global__ void __launch_bounds__(256)
patch_gather_nchw(const float *__restrict__ fmap, const int32_t *__restrict__ randx,
                  const int32_t *__restrict__ randy, float *__restrict__ X, int64_t ldx, int64_t rows, int B, int c,
                  int H, int W, int P, int k_rt, int pad, int stride, int relu) {
    const int k = KS > 0 ? KS : k_rt;
    const int k2 = k * k;
    const int K = c * k2;
    // one CTA per output row when the map is in HBM; a small persistent grid strides over the rows when the
    // map is read in place from pinned host memory (PCIe-bound: more CTAs only block SMs other layers need)
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const int img_in_batch = (int)(r % B);
        const int64_t bp = r / B;  // batch*P + point
        const int batch = (int)(bp / P);
        const int y0 = stride * randx[bp] - pad;  // window origin, rows   (net.py: feat[:,:,x,y], x indexes H)
        const int x0 = stride * randy[bp] - pad;  // window origin, cols
        const float *src = fmap + ((int64_t)batch * B + img_in_batch) * c * H * W;
        float *dst = X + r * ldx;
#pragma unroll 4
        for (int col = threadIdx.x; col < K; col += blockDim.x) {
            const int a = col / k2;
            const int p = col - a * k2;
            const int py = p / k;
            const int px = p - py * k;
            const int yy = y0 + py, xx = x0 + px;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(src + ((int64_t)a * H + yy) * W + xx);
            if (relu) v = fmaxf(v, 0.f);
            dst[col] = v;
        }
    }
}

constexpr int64_t CP_HOST_GATHER_CTAS = 64;  // grid of the in-place (zero-copy) reader

// NHWC: tile = k2 spatial taps x CT channels staged through shared memory.
constexpr int NHWC_CT = 128;  // channels per tile

__global__ void __launch_bounds__(256)
patch_gather_nhwc(const float *__restrict__ fmap, const int32_t *__restrict__ randx,
                  const int32_t *__restrict__ randy, float *__restrict__ X, int64_t ldx, int B, int c, int H,
                  int W, int P, int k, int pad, int stride, int relu) {
    extern __shared__ float tile[];  // [k2][NHWC_CT + 1]
    const int k2 = k * k;
    const int64_t r = blockIdx.x;
    const int a0 = blockIdx.y * NHWC_CT;
    const int ct = min(NHWC_CT, c - a0);
    const int img_in_batch = (int)(r % B);
    const int64_t bp = r / B;
    const int batch = (int)(bp / P);
    const int y0 = stride * randx[bp] - pad;
    const int x0 = stride * randy[bp] - pad;
    const float *src = fmap + ((int64_t)batch * B + img_in_batch) * H * W * c;
    for (int e = threadIdx.x; e < k2 * ct; e += blockDim.x) {
        const int p = e / ct;
        const int a = e - p * ct;
        const int py = p / k, px = p - py * k;
        const int yy = y0 + py, xx = x0 + px;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(src + ((int64_t)yy * W + xx) * c + a0 + a);
        if (relu) v = fmaxf(v, 0.f);
        tile[p * (NHWC_CT + 1) + a] = v;
    }
    __syncthreads();
    float *dst = X + r * ldx + (int64_t)a0 * k2;
    for (int e = threadIdx.x; e < k2 * ct; e += blockDim.x) {
        const int a = e / k2;
        const int p = e - a * k2;
        dst[e] = tile[p * (NHWC_CT + 1) + a];
    }
}

__global__ void __launch_bounds__(256)
point_gather(const float *__restrict__ fmap, const int32_t *__restrict__ randx,
             const int32_t *__restrict__ randy, float *__restrict__ Y, int64_t ldy, int B, int n, int H, int W,
             int P, int nhwc) {
    const int64_t r = blockIdx.x;
    const int img_in_batch = (int)(r % B);
    const int64_t bp = r / B;
    const int batch = (int)(bp / P);
    const int yy = randx[bp], xx = randy[bp];
    const float *src = fmap + ((int64_t)batch * B + img_in_batch) * n * H * W;
    float *dst = Y + r * ldy;
    if (nhwc) {
        const float *s = src + ((int64_t)yy * W + xx) * n;
        for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = __ldg(s + j);
    } else {
        const float *s = src + (int64_t)yy * W + xx;
        for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = __ldg(s + (int64_t)j * H * W);
    }
}

}  // namespace

// gather_tma.cu
bool cp_gather_tma_eligible(const float *fmap, int c, int k, float *X_out, int64_t ldx);
int cp_patch_gather_tma(cp_handle_t h, const float *fmap, int nbatch, int B, int c, int H, int W, const int32_t *randx,
                        const int32_t *randy, int P, int k, int pad, int stride, int relu, float *X_out, int64_t ldx,
                        cudaStream_t stream);
static bool tma_enabled() {  // CPB200_GATHER_TMA=0 keeps the SIMT kernel (A/B measurements)
    static const bool on = [] {
        const char *e = getenv("CPB200_GATHER_TMA");
        return !(e && e[0] == '0');
    }();
    return on;
}

extern "C" int cp_patch_gather(cp_handle_t h, const float *fmap, int nbatch, int B, int c, int H, int W,
                               int layout, const int32_t *randx, const int32_t *randy, int P, int k, int pad,
                               int stride, int relu, float *X_out, int64_t ldx, cp_stream_t stream_) {
    CP_REQUIRE(h && fmap && randx && randy && X_out, "cp_patch_gather: NULL argument");
    CP_REQUIRE(nbatch >= 0 && B > 0 && c > 0 && H > 0 && W > 0 && P > 0, "cp_patch_gather: bad shape");
    CP_REQUIRE(k >= 1 && (k & 1) == 1, "cp_patch_gather: kernel_size must be odd (reference net.py:604-605), got %d", k);
    CP_REQUIRE(pad >= 0 && stride >= 1, "cp_patch_gather: bad pad/stride");
    CP_REQUIRE(ldx >= (int64_t)c * k * k, "cp_patch_gather: ldx %lld < c*k*k", (long long)ldx);
    CP_REQUIRE(layout == CP_LAYOUT_NCHW || layout == CP_LAYOUT_NHWC, "cp_patch_gather: unknown layout %d", layout);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t rows = (int64_t)nbatch * P * B;
    if (rows == 0) return CP_OK;
    CP_REQUIRE(rows < (1ll << 31), "cp_patch_gather: too many rows");
    if (layout == CP_LAYOUT_NCHW) {
        // map in (pinned, UVA-mapped) host memory?  then the kernel is a PCIe reader: keep its footprint small
        cudaPointerAttributes pa;
        const bool host_src = cudaPointerGetAttributes(&pa, fmap) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        (void)cudaGetLastError();
        static const int64_t host_ctas = [] {
            const char *e = getenv("CPB200_HOST_GATHER_CTAS");  // tuning knob (profiles/e2e_breakdown.py)
            const long v = e ? atol(e) : 0;
            return (int64_t)(v > 0 ? v : CP_HOST_GATHER_CTAS);
        }();
        const int64_t ncta = host_src ? (rows < host_ctas ? rows : host_ctas) : rows;
        dim3 grid((unsigned)ncta);
        if (k == 3)
            patch_gather_nchw<3><<<grid, 256, 0, stream>>>(fmap, randx, randy, X_out, ldx, rows, B, c, H, W, P, k, pad, stride, relu);
        else if (k == 1)
            patch_gather_nchw<1><<<grid, 256, 0, stream>>>(fmap, randx, randy, X_out, ldx, rows, B, c, H, W, P, k, pad, stride, relu);
        else
            patch_gather_nchw<0><<<grid, 256, 0, stream>>>(fmap, randx, randy, X_out, ldx, rows, B, c, H, W, P, k, pad, stride, relu);
    } else if (tma_enabled() && cp_gather_tma_eligible(fmap, c, k, X_out, ldx)) {
        // NHWC map in HBM: whole windows by TMA, rows out by bulk store (gather_tma.cu)
        return cp_patch_gather_tma(h, fmap, nbatch, B, c, H, W, randx, randy, P, k, pad, stride, relu, X_out, ldx, stream);
    } else {
        const size_t smem = (size_t)k * k * (NHWC_CT + 1) * sizeof(float);
        CP_REQUIRE(smem <= 48 * 1024, "cp_patch_gather: kernel_size %d too large for the NHWC tile", k);
        dim3 grid((unsigned)rows, (unsigned)cp_cdiv(c, NHWC_CT));
        patch_gather_nhwc<<<grid, 256, smem, stream>>>(fmap, randx, randy, X_out, ldx, B, c, H, W, P, k, pad, stride, relu);
    }
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_point_gather(cp_handle_t h, const float *fmap, int nbatch, int B, int n, int H, int W,
                               int layout, const int32_t *randx, const int32_t *randy, int P, float *Y_out,
                               int64_t ldy, cp_stream_t stream_) {
    CP_REQUIRE(h && fmap && randx && randy && Y_out, "cp_point_gather: NULL argument");
    CP_REQUIRE(nbatch >= 0 && B > 0 && n > 0 && H > 0 && W > 0 && P > 0, "cp_point_gather: bad shape");
    CP_REQUIRE(ldy >= n, "cp_point_gather: ldy < n");
    CP_REQUIRE(layout == CP_LAYOUT_NCHW || layout == CP_LAYOUT_NHWC, "cp_point_gather: unknown layout %d", layout);
    const int64_t rows = (int64_t)nbatch * P * B;
    if (rows == 0) return CP_OK;
    CP_REQUIRE(rows < (1ll << 31), "cp_point_gather: too many rows");
    point_gather<<<(unsigned)rows, 256, 0, (cudaStream_t)stream_>>>(fmap, randx, randy, Y_out, ldy, B, n, H, W, P,
                                                                   layout == CP_LAYOUT_NHWC);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
