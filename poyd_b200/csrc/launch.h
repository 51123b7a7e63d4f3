// Host-callable launchers of every kernel family.  Each family lives in its own translation unit (k_*.cu) so that the
// library builds in parallel; api.cu (planning, memory, streams) sees only these declarations.
#pragma once
#include <cuda_runtime.h>

#include "classes.h"

namespace poyb200 {

// ---- affine_3 fills (k_aff_stripe.cu, k_aff_fast.cu) -----------------------------------------------------------------
cudaError_t stripe_launch(uint32_t klass, bool affine, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                          int *cost, int sm_count, int seq_bytes, int allow_noeb, int *work_counter, const int *batch_list,
                          const int *batch_count, cudaStream_t stream);
cudaError_t fast_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                        int sm_count, int seq_bytes, int *work_counter, const int *batch_list, const int *batch_count, int *slow_list,
                        int *slow_count, bool dir6, cudaStream_t stream);
// two pairs per lane group on 16-bit halves (aff_x2_kernels.cuh): shape (5, 8), 6-bit band; declines whole batches into slow_list
bool x2_usable(int seq_bytes, int max_unit4, int gap_open);
cudaError_t x2_launch(const Task *d_tasks, int n, DevCM cm, int max_unit4, const uint8_t *pool, uint8_t *dir, int *cost, int sm_count,
                      int seq_bytes, int *work_counter, int *slow_list, int *slow_count, cudaStream_t stream);
// ring kernels (k_aff_ring.cu): fill + traceback of the pairs whose stripe has no spare diagonals.  ebf = false takes the
// batches without gap bits and lists the others in slow_list; ebf = true takes everything it is given.
size_t ring_scratch_bytes(int sm_count, size_t slot_bytes);
cudaError_t ring_launch(uint32_t klass, bool bt, bool ebf, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *scratch,
                        size_t scratch_bytes, size_t slot_bytes, OutPtrs out, int sm_count, int seq_bytes, int *work_counter,
                        const int *batch_list, const int *batch_count, int *slow_list, int *slow_count, cudaStream_t stream);
// ---- linear fills (k_lin_stripe.cu) -----------------------------------------------------------------------------------
cudaError_t lin_stripe_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                              int *cost, int sm_count, int seq_bytes, int custom_tail, int *work_counter, cudaStream_t stream);
// full matrices, column-striped (k_lin_rows.cu)
cudaError_t lin_rows_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                            int sm_count, int seq_bytes, int custom_tail, int *work_counter, cudaStream_t stream);
// ---- generic fills, tracebacks, medians, INT32 probe (k_misc.cu) -------------------------------------------------------
cudaError_t aff_generic_launch(bool bt, int blocks, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, int4 *state,
                               int state_stride, uint8_t *dir, int *cost, cudaStream_t stream);
cudaError_t lin_generic_launch(bool bt, int blocks, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, int *state,
                               int state_stride, uint8_t *dir, int *cost, cudaStream_t stream);
cudaError_t traceback_launch(bool affine, int blocks, int threads, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool,
                             const uint8_t *dir, OutPtrs out, int *work_counter, int wpw, cudaStream_t stream);
cudaError_t median_2_launch(int which, DevCM cm, const uint8_t *a, const uint8_t *b, long long in_stride, const int *len, int n,
                            uint8_t *out, long long out_stride, int *out_len, cudaStream_t stream);
cudaError_t calc_aligned_2_launch(const int *matrix, DevCM cm, const uint8_t *a, const uint8_t *b, long long in_stride, const int *len,
                                  int n, int *out, cudaStream_t stream);
cudaError_t median_3_launch(const uint8_t *median3, int lcm, const uint8_t *a, const uint8_t *b, const uint8_t *c, long long in_stride,
                            const int *len, int n, uint8_t *out, long long out_stride, int *out_len, cudaStream_t stream);
// device-resident sequence store (store.cu)
cudaError_t store_row_stats_launch(const uint8_t *rows, long long stride, const int *outlen4, int n, int gap, int *stats, cudaStream_t s);
cudaError_t store_append_launch(const uint8_t *rows, long long stride, const int *outlen4, const long long *newoff, int n, uint8_t *pool,
                                cudaStream_t s);
cudaError_t store_equal_launch(const uint8_t *pool, const uint4 *jobs, int n, uint8_t *eq, cudaStream_t s);
cudaError_t store_closest_same_launch(DevCM cm, const uint8_t *pool, const uint2 *jobs, int n, uint8_t *rows, long long stride,
                                      int *outlen4, cudaStream_t s);
constexpr int PEAK_ITERS = 4096;
constexpr int PEAK_CHAINS = 8;
cudaError_t int32_peak_launch(int kind, int blocks, int threads, int *out, int seed, cudaStream_t stream);

// ---- three sequences (k_cube.cu) ------------------------------------------------------------------------------------------
struct Task3 {
    uint32_t off1, off2, off3;
    int32_t l1, l2, l3;      // stored lengths (leading gap included)
    uint32_t triple;         // index in the caller's list
    uint32_t pad;
    uint64_t dir_off;        // byte offset of this triple's direction cube
};

struct DevCM3 {
    int lcm, gap;
    const int *cost;         // (1 << lcm)^3
    const uint8_t *median;
};

struct Out3 {
    int *cost, *out_len, *status;
    uint8_t *r1, *r2, *r3, *median;
    long long stride;
    uint32_t want;
};

constexpr int CUBE_THREADS = 512;
cudaError_t cube_fill_launch(int grid, int threads, const Task3 *tasks, int n, DevCM3 cm, const uint8_t *pool, int *ring, size_t ring_ints,
                             uint8_t *dir, int *cost, int bt, cudaStream_t stream);
cudaError_t cube_traceback_launch(const Task3 *tasks, int n, DevCM3 cm, const uint8_t *pool, const uint8_t *dir, Out3 out,
                                  cudaStream_t stream);

}  // namespace poyb200
