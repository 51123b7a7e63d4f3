// The context behind the C ABI (include/poyb200.h), shared by api.cu (planning, launches, one-shot calls), store.cu (the
// device-resident sequence store) and multi.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/poyb200.h"
#include "launch.h"

using namespace poyb200;

enum Mode { MODE_COST_2 = 0, MODE_ALIGN_2 = 1, MODE_COST_AFF = 2, MODE_ALIGN_AFF = 3 };

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Grow-only array in pinned host memory (so that its upload is a true asynchronous DMA).
template <typename T>
struct PinnedVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool resize(size_t m) {
        if (m > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            const size_t want = m + m / 8 + 64;
            if (cudaHostAlloc((void **) &p, want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) return false;
            cap = want;
        }
        n = m;
        return true;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = cap = 0;
    }
    size_t size() const { return n; }
    T *data() { return p; }
    T *begin() { return p; }
    T *end() { return p + n; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

struct Chunk {
    size_t begin, end;  // task range
    size_t dir_bytes;
};

struct poyb200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    // cost matrix
    bool has_cm = false;
    poyb200_cm hcm{};  // scalars only; pointers below are device pointers
    DevCM dcm{};
    DevBuf<int> d_cost, d_prepend, d_tail, d_worst;
    DevBuf<uint8_t> d_median;
    // staged batch
    bool staged = false;
    int mode = 0;
    poyb200_batch hb{};
    PinnedVec<Task> tasks, tasks_tmp;
    std::vector<Chunk> chunks;
    std::vector<size_t> class_begin;  // per chunk x class boundaries are recomputed at launch time
    DevBuf<uint8_t> d_pool, d_dir, d_out[4], d_bits[3];
    const uint8_t *cur_pool = nullptr;  // what the kernels read operands from: d_pool, or a device-resident store's pool
    bool device_store = false;          // operands live in a poyb200_store, results stay on the device (store_batch)
    DevBuf<Task> d_tasks;
    DevBuf<int> d_costs, d_outlen, d_lin_state, d_counters, d_slow_list, d_slow_list2;
    size_t counter_next = 0;  // work counters handed to launches of the current call (zeroed once per call)
    DevBuf<int4> d_aff_state;
    long long dstride = 0, bstride = 0;
    size_t dir_budget = 0;
    int state_stride = 0;
    int stripe_seq_bytes = 16;
    poyb200_config cfg{};  // every tunable of the context (include/poyb200.h); fixed at creation
    int ring_mode = 2;     // cfg.use_ring as it applies to the call being staged (small calls may take the ring kernels)
    // Shard view (poyb200_multi_*, multi.cu): the batch's `pool` pointer is the caller's pool + view_lo and holds only the
    // bytes this shard's pairs reference; seq_off[] stays the caller's array, so view_lo is subtracted per pair and the
    // sequences are validated per pair instead of per pool entry.
    bool view = false;
    int64_t view_lo = 0;
    int x2_unit4 = 1 << 30;  // 4 * the largest table entry aff_x2_kernel can read (its 16-bit range guard)
    int custom_tail = 0;   // tail_cost[a] != cost[a][gap] for some a: the last-column rule is not a no-op
    int lin_natural = 0;   // default prepend and tail costs: the first row and column follow from the ordinary recurrence
    int host_threads = 8;
    size_t chunk_pairs = 1u << 16;  // pairs per chunk (pipelining granularity of the one-shot calls); with three direction
                                    // buffers 65 536 and 131 072 give the same device time, and the smaller chunk lets the
                                    // download of the four sequences keep up (581 against 543 GCUPS end to end)
    bool in_order = true;           // tasks[k].pair == k
    cudaStream_t s_in = nullptr, s_out = nullptr, s_tb = nullptr, s_len = nullptr;
    DevBuf<uint8_t> d_scratch;       // ring kernels: per-warp band slots (aff_ring_kernels.cuh)
    DevBuf<uint8_t> d_walked;        // use_ring = 2: pairs the full ring instance walked
    size_t ring_slot_bytes = 0;      // largest band of a ring-class pair of the staged batch
    DevBuf<uint8_t> d_dir2, d_dir3;  // further direction buffers: the traceback of chunk k runs under the fills of chunks k+1, k+2
    uint8_t *cur_dir = nullptr;
    std::vector<cudaEvent_t> ev_fill, ev_tb;
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_done, ev_pool;
    // 3-D
    bool has_cm3 = false;
    DevCM3 dcm3{};
    DevBuf<int> d_cost3, d_ring, d_status;
    DevBuf<uint8_t> d_median3;
    DevBuf<Task3> d_tasks3;
    DevBuf<uint8_t> d_pw_arena, d_pw_seq;  // Powell kernel: per-CTA workspaces, kept between calls (powell.cu)
    // stats
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> chunk_ev;  // 3 per chunk: before fill, after fill, after traceback
    size_t timed_chunks = 0;
};

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            return (e__ == cudaErrorMemoryAllocation) ? POYB200_ENOMEM : POYB200_ECUDA;    \
        }                                                                                  \
    } while (0)

static int fail(poyb200_ctx *ctx, int code, const char *msg) {
    ctx->err = msg;
    return code;
}


// internal entry points shared between the translation units
int poyb200_stage_internal(poyb200_ctx *ctx, int mode, const poyb200_batch *b, bool upload);
int poyb200_run_staged(poyb200_ctx *ctx);  // = poyb200_run
