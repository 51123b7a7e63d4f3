// Control experiment for `compute-sanitizer --tool racecheck` on the staging ring (poyd_b200/csrc/staging.cuh).
//
// The product kernels stage operands with cp.async.bulk completing on an mbarrier.  Round 1's racecheck runs reported
// hazards between the asynchronous-proxy write and the later generic reads.  This program runs the SAME StageRing code
// in a minimal kernel, two ways:
//
//   mode 0  correct protocol          (wait FULL, read, arrive EMPTY; producer waits EMPTY)
//   mode 1  read BEFORE the FULL wait (a genuine read-after-write race)
//
// If racecheck reports the same hazards for mode 0 as for mode 1, it does not model the mbarrier completion of bulk
// copies (tool limitation); if mode 0 is clean and mode 1 is flagged, the tool sees the protocol and a clean report on
// the product kernels means what it says.  Mode 0 also checks every byte it read against global memory (must be exact).
// Second argument: ring depth (2 default, 1 = the one-slot protocol).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O2 -o build/racecheck_control tools/racecheck_control.cu
//   compute-sanitizer --tool racecheck build/racecheck_control 0     (then 1; then `0 1`)
// Result on B200 (profiles/r02_sanitizer_racecheck.txt): identical reports for both modes -> tool limitation.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../poyd_b200/csrc/staging.cuh"

using namespace poyb200;

constexpr int G = 8, GPW = 4, WARPS = 2, SEQ = 256, NB = 64;

template <int MODE>
__global__ void control_kernel(const Task *tasks, int ntasks, const uint8_t *pool, int nslots, int *work_counter, unsigned *sums,
                               int *bad) {
    extern __shared__ __align__(16) uint8_t smem[];
    StageBars *s_bar = reinterpret_cast<StageBars *>(smem);
    uint8_t *s_seq = smem + WARPS * GPW * sizeof(StageBars);
    if (threadIdx.x < WARPS * GPW) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31, grp = lane32 / G, lane = lane32 % G;
    StageRing<G> ring;
    ring.attach(&s_bar[warp * GPW + grp], s_seq + (size_t) ((warp * GPW + grp) * 2 * nslots) * SEQ, SEQ, nslots, lane);
    const int nbatches = (ntasks + GPW - 1) / GPW;
    int slot = 0;
    int batch = fetch_batch(work_counter, nbatches, nullptr, nullptr);
    if (batch >= 0) ring.produce_task(0, tasks, ntasks, batch * GPW + grp, pool, 16);
    while (batch >= 0) {
        int next = -1;
        if (nslots == 2) {
            next = fetch_batch(work_counter, nbatches, nullptr, nullptr);
            if (next >= 0) ring.produce_task(slot ^ 1, tasks, ntasks, next * GPW + grp, pool, 16);
        }
        const int ti = batch * GPW + grp;
        const bool valid = ti < ntasks;
        const int lr = valid ? tasks[ti].lr : 1, lc = valid ? tasks[ti].lc : 1;
        unsigned acc = 0;
        if (MODE == 1) {  // deliberately wrong: touch the slot before the copy is known to have landed
            for (int k = lane; k < lr; k += G) acc += ring.rows(slot)[k];
        }
        ring.wait_full(slot);
        const uint8_t *r = ring.rows(slot), *c = ring.cols(slot);
        for (int k = lane; k < lr; k += G) acc += r[k] * 3u;
        for (int k = lane; k < lc; k += G) acc += c[k] * 5u;
        if (valid) {
            int wrong = 0;
            for (int k = lane; k < lr; k += G) wrong |= (r[k] != pool[tasks[ti].off_r + k]);
            for (int k = lane; k < lc; k += G) wrong |= (c[k] != pool[tasks[ti].off_c + k]);
            if (wrong) atomicAdd(bad, 1);
            atomicAdd(&sums[ti], acc);
        }
        ring.release(slot);
        if (nslots == 1) {
            next = fetch_batch(work_counter, nbatches, nullptr, nullptr);
            if (next >= 0) {
                ring.produce_task(0, tasks, ntasks, next * GPW + grp, pool, 16);
            }
        } else {
            slot ^= 1;
        }
        batch = next;
    }
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int nslots = argc > 2 ? atoi(argv[2]) : 2;
    const int ntasks = NB * GPW;
    std::vector<uint8_t> pool((size_t) ntasks * 2 * SEQ);
    std::vector<Task> tasks(ntasks);
    unsigned x = 12345;
    for (auto &b : pool) { x = x * 1664525u + 1013904223u; b = (uint8_t) (1u << ((x >> 24) & 3)); }
    for (int t = 0; t < ntasks; t++) {
        tasks[t] = Task{};
        tasks[t].off_r = (uint32_t) (t * 2 * SEQ);
        tasks[t].off_c = (uint32_t) (t * 2 * SEQ + SEQ);
        tasks[t].lr = 100 + (t * 7) % 150;
        tasks[t].lc = 100 + (t * 13) % 150;
    }
    uint8_t *d_pool; Task *d_tasks; int *d_counter, *d_bad; unsigned *d_sums;
    cudaMalloc(&d_pool, pool.size() + 64); cudaMalloc(&d_tasks, sizeof(Task) * ntasks);
    cudaMalloc(&d_counter, 4); cudaMalloc(&d_bad, 4); cudaMalloc(&d_sums, 4 * ntasks);
    cudaMemcpy(d_pool, pool.data(), pool.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_tasks, tasks.data(), sizeof(Task) * ntasks, cudaMemcpyHostToDevice);
    cudaMemset(d_counter, 0, 4); cudaMemset(d_bad, 0, 4); cudaMemset(d_sums, 0, 4 * ntasks);
    const size_t smem = WARPS * GPW * sizeof(StageBars) + (size_t) WARPS * GPW * 2 * nslots * SEQ;
    if (mode == 0) control_kernel<0><<<4, WARPS * 32, smem>>>(d_tasks, ntasks, d_pool, nslots, d_counter, d_sums, d_bad);
    else control_kernel<1><<<4, WARPS * 32, smem>>>(d_tasks, ntasks, d_pool, nslots, d_counter, d_sums, d_bad);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned> sums(ntasks);
    cudaMemcpy(sums.data(), d_sums, 4 * ntasks, cudaMemcpyDeviceToHost);
    int wrong_sums = 0;
    for (int t = 0; t < ntasks; t++) {
        unsigned want = 0;
        for (int k = 0; k < tasks[t].lr; k++) want += pool[tasks[t].off_r + k] * 3u;
        for (int k = 0; k < tasks[t].lc; k++) want += pool[tasks[t].off_c + k] * 5u;
        if (mode != 1 && sums[t] != want) wrong_sums++;
    }
    printf("racecheck_control mode %d nslots %d: %s, groups with a stale byte %d, wrong checksums %d\n", mode, nslots,
           cudaGetErrorString(e), bad, wrong_sums);
    return (e != cudaSuccess) || (mode == 0 && (bad != 0 || wrong_sums != 0));
}
