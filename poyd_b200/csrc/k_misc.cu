// Translation unit of the generic (any-shape) fills, the tracebacks, the stand-alone medians and the INT32 probe.
#include "launch.h"
#include "generic_kernels.cuh"
#include "peak.cuh"
#include "trace_kernels.cuh"

namespace poyb200 {

cudaError_t aff_generic_launch(bool bt, int blocks, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, int4 *state,
                               int state_stride, uint8_t *dir, int *cost, cudaStream_t stream) {
    if (bt) aff_generic_kernel<true><<<blocks, 128, 0, stream>>>(d_tasks, n, cm, pool, state, state_stride, dir, cost);
    else aff_generic_kernel<false><<<blocks, 128, 0, stream>>>(d_tasks, n, cm, pool, state, state_stride, dir, cost);
    return cudaGetLastError();
}

cudaError_t lin_generic_launch(bool bt, int blocks, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, int *state,
                               int state_stride, uint8_t *dir, int *cost, cudaStream_t stream) {
    if (bt) lin_generic_kernel<true><<<blocks, 128, 0, stream>>>(d_tasks, n, cm, pool, state, state_stride, dir, cost);
    else lin_generic_kernel<false><<<blocks, 128, 0, stream>>>(d_tasks, n, cm, pool, state, state_stride, dir, cost);
    return cudaGetLastError();
}

cudaError_t traceback_launch(bool affine, int blocks, int threads, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool,
                             const uint8_t *dir, OutPtrs out, int *work_counter, int wpw, cudaStream_t stream) {
    if (affine) aff_traceback_kernel<<<blocks, threads, 0, stream>>>(d_tasks, n, cm, pool, dir, out, work_counter, wpw);
    else lin_traceback_kernel<<<blocks, threads, 0, stream>>>(d_tasks, n, cm, pool, dir, out, work_counter, wpw);
    return cudaGetLastError();
}

cudaError_t median_2_launch(int which, DevCM cm, const uint8_t *a, const uint8_t *b, long long in_stride, const int *len, int n,
                            uint8_t *out, long long out_stride, int *out_len, cudaStream_t stream) {
    median_2_kernel<<<(n + 127) / 128, 128, 0, stream>>>(which, cm, a, b, in_stride, len, n, out, out_stride, out_len);
    return cudaGetLastError();
}

cudaError_t calc_aligned_2_launch(const int *matrix, DevCM cm, const uint8_t *a, const uint8_t *b, long long in_stride, const int *len,
                                  int n, int *out, cudaStream_t stream) {
    calc_aligned_2_kernel<<<(n + 127) / 128, 128, 0, stream>>>(matrix, cm, a, b, in_stride, len, n, out);
    return cudaGetLastError();
}

cudaError_t median_3_launch(const uint8_t *median3, int lcm, const uint8_t *a, const uint8_t *b, const uint8_t *c, long long in_stride,
                            const int *len, int n, uint8_t *out, long long out_stride, int *out_len, cudaStream_t stream) {
    median_3_kernel<<<(n + 127) / 128, 128, 0, stream>>>(median3, lcm, a, b, c, in_stride, len, n, out, out_stride, out_len);
    return cudaGetLastError();
}

cudaError_t store_row_stats_launch(const uint8_t *rows, long long stride, const int *outlen4, int n, int gap, int *stats, cudaStream_t s) {
    store_row_stats_kernel<<<(n + 3) / 4, 128, 0, s>>>(rows, stride, outlen4, n, gap, stats);
    return cudaGetLastError();
}
cudaError_t store_append_launch(const uint8_t *rows, long long stride, const int *outlen4, const long long *newoff, int n, uint8_t *pool,
                                cudaStream_t s) {
    store_append_kernel<<<(n + 3) / 4, 128, 0, s>>>(rows, stride, outlen4, newoff, n, pool);
    return cudaGetLastError();
}
cudaError_t store_equal_launch(const uint8_t *pool, const uint4 *jobs, int n, uint8_t *eq, cudaStream_t s) {
    store_equal_kernel<<<(n + 3) / 4, 128, 0, s>>>(pool, jobs, n, eq);
    return cudaGetLastError();
}
cudaError_t store_closest_same_launch(DevCM cm, const uint8_t *pool, const uint2 *jobs, int n, uint8_t *rows, long long stride,
                                      int *outlen4, cudaStream_t s) {
    store_closest_same_kernel<<<(n + 127) / 128, 128, 0, s>>>(cm, pool, jobs, n, rows, stride, outlen4);
    return cudaGetLastError();
}

cudaError_t int32_peak_launch(int kind, int blocks, int threads, int *out, int seed, cudaStream_t stream) {
    if (kind == 0) int32_peak_kernel<0><<<blocks, threads, 0, stream>>>(out, seed);
    else if (kind == 1) int32_peak_kernel<1><<<blocks, threads, 0, stream>>>(out, seed);
    else int32_peak_kernel<2><<<blocks, threads, 0, stream>>>(out, seed);
    return cudaGetLastError();
}

}  // namespace poyb200
