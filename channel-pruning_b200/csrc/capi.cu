// libcpb200: handle lifetime, version, error string.
#include "common.cuh"

thread_local char cp_err_buf[512] = "";
std::atomic<unsigned long long> cp_launch_counter{0};

extern "C" int64_t cp_launch_count(void) { return (int64_t)cp_launch_counter.load(std::memory_order_relaxed); }

extern "C" int cp_version(void) { return 200; }  // 0.2.0

extern "C" const char *cp_last_error(void) { return cp_err_buf; }

extern "C" int cp_create(cp_handle_t *out, int device) {
    CP_REQUIRE(out != nullptr, "cp_create: out is NULL");
    int ndev = 0;
    CP_CUDA(cudaGetDeviceCount(&ndev));
    CP_REQUIRE(device >= 0 && device < ndev, "cp_create: device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    CP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        CP_FAIL(CP_ERR_CUDA, "cp_create: device %d is sm_%d%d; libcpb200 is built for sm_100a only", device,
                prop.major, prop.minor);
    cp_handle_s *h = new cp_handle_s();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->ws = nullptr;
    h->ws_bytes = 0;
    h->tmap_encode = nullptr;
    h->side = nullptr;
    h->bulk = nullptr;
    h->ev_panel = nullptr;
    h->ev_side = nullptr;
    h->ev_bulk = nullptr;
    h->potrf_configured = false;
    h->fac = nullptr;
    h->fac_bytes = 0;
    h->fac_K = h->fac_Kfull = 0;
    h->fac_N = 0;
    h->fac_rows = 0;
    h->aux = nullptr;
    h->aux_bytes = 0;
    h->gram_profile = false;
    h->ev_gram0 = h->ev_gram1 = nullptr;
    h->ls_tc = false;
    for (int i = 0; i < 3; ++i) {
        h->tcbuf[i] = nullptr;
        h->tcbuf_bytes[i] = 0;
    }
    *out = h;
    return CP_OK;
}

extern "C" int cp_destroy(cp_handle_t h) {
    if (!h) return CP_OK;
    if (h->ws || h->side || h->fac || h->aux || h->ev_gram0 || h->tcbuf[0] || h->tcbuf[1] || h->tcbuf[2]) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(h->device);
        if (h->ws) cudaFree(h->ws);
        if (h->fac) cudaFree(h->fac);
        if (h->aux) cudaFree(h->aux);
        for (int i = 0; i < 3; ++i)
            if (h->tcbuf[i]) cudaFree(h->tcbuf[i]);
        if (h->ev_gram0) {
            cudaEventDestroy(h->ev_gram0);
            cudaEventDestroy(h->ev_gram1);
        }
        if (h->side) {
            cudaStreamDestroy(h->side);
            cudaStreamDestroy(h->bulk);
            cudaEventDestroy(h->ev_panel);
            cudaEventDestroy(h->ev_side);
            cudaEventDestroy(h->ev_bulk);
        }
        cudaSetDevice(cur);
    }
    delete h;
    return CP_OK;
}

extern "C" int cp_gram_profile(cp_handle_t h, int enable) {
    CP_REQUIRE(h != nullptr, "cp_gram_profile: NULL handle");
    CP_DEVICE_GUARD(h);
    if (enable && !h->ev_gram0) {
        CP_CUDA(cudaEventCreate(&h->ev_gram0));
        CP_CUDA(cudaEventCreate(&h->ev_gram1));
    }
    h->gram_profile = enable != 0;
    return CP_OK;
}

extern "C" int cp_ls_tensor_cores(cp_handle_t h, int enable) {
    CP_REQUIRE(h != nullptr, "cp_ls_tensor_cores: NULL handle");
    h->ls_tc = enable != 0;
    return CP_OK;
}

extern "C" int cp_gram_kernel_ms(cp_handle_t h, float *ms) {
    CP_REQUIRE(h != nullptr && ms != nullptr, "cp_gram_kernel_ms: NULL argument");
    CP_REQUIRE(h->gram_profile && h->ev_gram0, "cp_gram_kernel_ms: profiling is off (cp_gram_profile)");
    CP_DEVICE_GUARD(h);
    CP_CUDA(cudaEventSynchronize(h->ev_gram1));
    CP_CUDA(cudaEventElapsedTime(ms, h->ev_gram0, h->ev_gram1));
    return CP_OK;
}

extern "C" int64_t cp_workspace_bytes(cp_handle_t h) { return h ? (int64_t)h->ws_bytes : 0; }

int cp_ws_reserve(cp_handle_t h, size_t bytes, void **out) {
    if (bytes > h->ws_bytes) {
        // cudaFree synchronises the device, so no in-flight kernel can still be using the old block
        if (h->ws) CP_CUDA(cudaFree(h->ws));
        h->ws = nullptr;
        h->ws_bytes = 0;
        size_t want = cp_align_up(bytes + bytes / 8, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&h->ws, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            CP_FAIL(CP_ERR_WORKSPACE, "workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        h->ws_bytes = want;
    }
    *out = h->ws;
    return CP_OK;
}

int cp_aux_reserve(cp_handle_t h, size_t bytes, void **out) {
    if (bytes > h->aux_bytes) {
        if (h->aux) CP_CUDA(cudaFree(h->aux));
        h->aux = nullptr;
        h->aux_bytes = 0;
        size_t want = cp_align_up(bytes + bytes / 8, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&h->aux, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            CP_FAIL(CP_ERR_WORKSPACE, "auxiliary workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        h->aux_bytes = want;
    }
    *out = h->aux;
    return CP_OK;
}
