// Shared host-side plumbing of libcpb200: handle, scratch workspace, error reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cpb200.h"

struct cp_handle_s {
    int device;
    int num_sms;
    void *ws;          // scratch, grown on demand
    size_t ws_bytes;
    void *tmap_encode; // cuTensorMapEncodeTiled entry point (resolved lazily)
    // look-ahead of the blocked Cholesky (ls.cu): low-priority side stream + fork/join events, created lazily
    cudaStream_t side;
    cudaEvent_t ev_panel, ev_side;
};

extern thread_local char cp_err_buf[512];

#define CP_FAIL(code, ...)                                  \
    do {                                                    \
        snprintf(cp_err_buf, sizeof(cp_err_buf), __VA_ARGS__); \
        return (code);                                      \
    } while (0)

#define CP_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            CP_FAIL(CP_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

// every kernel launch of the library goes through one of these two, so the counter is exact
extern unsigned long long cp_launch_counter;
#define CP_CHECK_LAUNCH()              \
    do {                               \
        ++cp_launch_counter;           \
        CP_CUDA(cudaGetLastError());   \
    } while (0)
#define CP_GEMM_LAUNCH(call)           \
    do {                               \
        ++cp_launch_counter;           \
        CP_CUDA((call));               \
    } while (0)

#define CP_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) CP_FAIL(CP_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// Returns scratch of at least `bytes` (256-byte aligned); grows (synchronising) if needed.
int cp_ws_reserve(cp_handle_t h, size_t bytes, void **out);

static inline size_t cp_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cp_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Carves aligned sub-buffers out of one reservation.
struct cp_carver {
    char *base;
    size_t off;
    explicit cp_carver(void *b) : base((char *)b), off(0) {}
    template <typename T>
    T *take(size_t count) {
        T *p = (T *)(base + off);
        off += cp_align_up(count * sizeof(T), 256);
        return p;
    }
    static size_t need(size_t count, size_t elt) { return cp_align_up(count * elt, 256); }
};
