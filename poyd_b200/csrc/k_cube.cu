// Translation unit of the three-sequence kernels (cube_kernels.cuh).
#include "../../include/poyb200.h"
#include "launch.h"
#include "cube_kernels.cuh"

namespace poyb200 {

cudaError_t cube_fill_launch(int grid, int threads, const Task3 *tasks, int n, DevCM3 cm, const uint8_t *pool, int *ring, size_t ring_ints,
                             uint8_t *dir, int *cost, int bt, cudaStream_t stream) {
    cube_fill_kernel<<<grid, threads, 0, stream>>>(tasks, n, cm, pool, ring, ring_ints, dir, cost, bt);
    return cudaGetLastError();
}

cudaError_t cube_traceback_launch(const Task3 *tasks, int n, DevCM3 cm, const uint8_t *pool, const uint8_t *dir, Out3 out,
                                  cudaStream_t stream) {
    cube_traceback_kernel<<<(n + 127) / 128, 128, 0, stream>>>(tasks, n, cm, pool, dir, out);
    return cudaGetLastError();
}

}  // namespace poyb200
