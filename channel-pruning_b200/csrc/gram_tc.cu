// cp_gram, mode CP_GRAM_3XTF32: tcgen05 tensor-core path (under construction in this file).
#include "common.cuh"

int cp_gram_tc(cp_handle_t h, const float *X, int64_t N, int K, int64_t ldx, const void *Yraw, int y_dtype, int n, int64_t ldy,
               const float *y_bias, const int32_t *rows, int nrows, double *G, double *Bxy, double *sx,
               double *sy, double *yy, cudaStream_t stream) {
    CP_FAIL(CP_ERR_INVALID, "cp_gram: CP_GRAM_3XTF32 is not available in this build");
}
