// HBM -> shared-memory staging of a pair's two operands with the bulk-copy engine (TMA, cp.async.bulk), as a ring of
// one or two slots per lane group with a FULL / EMPTY mbarrier pair per slot:
//
//   producer (lane 0 of the group)             consumers (all G lanes of the group)
//   ---------------------------------------    ------------------------------------------------
//   wait EMPTY[s]   (the slot's last reader    wait FULL[s]    (complete_tx of both bulk copies)
//                    has arrived)              ... read the operands from slot s ...
//   expect_tx FULL[s]; cp.async.bulk x 2       arrive EMPTY[s] (after the LAST read of slot s)
//
// so the asynchronous-proxy write of the next pair is ordered after every generic-proxy read of the previous pair by
// an mbarrier phase (release at the readers' arrive, acquire at the producer's wait), not by a warp barrier, and with
// two slots the operands of pair k+1 travel while pair k is being computed.  A kernel whose shared memory would lose
// a resident CTA to the second slot runs the same protocol with one slot (no prefetch, same ordering).
//
// All waits are WARP-UNIFORM: every lane polls its own group's barrier and the loop ends on an __all_sync vote.
// (Per-lane polling loops let the lanes of a warp leave at different polls; the warp then kept running as separate
// fragments and issued every later instruction once per fragment -- a measured 2x slowdown, profiles/README.md.)
#pragma once
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace poyb200 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// One poll of the barrier: true once the phase with the given parity has completed.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_uniform(uint64_t *bar, uint32_t parity) {
    bool done;
    do {
        done = mbar_try_wait(bar, parity);
    } while (!__all_sync(0xffffffffu, done));
}

struct StageBars {
    uint64_t full[2], empty[2];
};
constexpr int STAGE_BAR_BYTES = (int) sizeof(StageBars);  // per lane group

// Per-thread view of its group's ring.  Every member function must be called by all 32 lanes of the warp.
template <int G>
struct StageRing {
    StageBars *bars;
    uint8_t *base;                 // slot 0, row operand; column operand at + op_stride; slot 1 at + 2 * op_stride
    int op_stride, nslots, lane;   // lane inside the group
    uint32_t fpar, epar;           // parity to wait for next, one bit per slot

    // One thread per group initialises the barriers; the caller runs __syncthreads() before first use.
    static __device__ __forceinline__ void init_bars(StageBars *b) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&b->full[s], 1);
            mbar_init(&b->empty[s], G);
        }
    }
    __device__ __forceinline__ void attach(StageBars *b, uint8_t *group_base, int op_stride_, int nslots_, int lane_) {
        bars = b; base = group_base; op_stride = op_stride_; nslots = nslots_; lane = lane_;
        fpar = 0u;   // FULL: wait for the completion of phase 0, 1, 0, ...
        epar = 3u;   // EMPTY: a fresh barrier reports its "previous" phase complete, i.e. the slot starts free
    }
    __device__ __forceinline__ uint8_t *rows(int slot) const { return base + (size_t) slot * 2 * op_stride; }
    __device__ __forceinline__ uint8_t *cols(int slot) const { return rows(slot) + op_stride; }

    // Fills `slot` with the operands of one pair.  Bulk copies need 16-byte aligned sources (SeqPool guarantees that);
    // otherwise, and for the padding groups of a last batch (valid = false: one element of pad_code each), plain stores.
    __device__ __forceinline__ void produce(int slot, const uint8_t *gr, const uint8_t *gc, int lr, int lc, bool valid, int pad_code) {
        mbar_wait_uniform(&bars->empty[slot], (epar >> slot) & 1u);
        epar ^= 1u << slot;
        uint8_t *dst_r = rows(slot), *dst_c = cols(slot);
        const bool bulk = valid && ((((uintptr_t) gr | (uintptr_t) gc) & 15) == 0);
        if (bulk) {
            if (lane == 0) {
                const uint32_t br = (uint32_t) (lr + 15) & ~15u, bc = (uint32_t) (lc + 15) & ~15u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars->full[slot], br + bc);
                bulk_g2s(dst_r, gr, br, &bars->full[slot]);
                bulk_g2s(dst_c, gc, bc, &bars->full[slot]);
            }
        } else if (valid) {
            for (int k = lane; k < lr; k += G) dst_r[k] = __ldg(gr + k);
            for (int k = lane; k < lc; k += G) dst_c[k] = __ldg(gc + k);
        } else if (lane == 0) {
            dst_r[0] = (uint8_t) pad_code;
            dst_c[0] = (uint8_t) pad_code;
        }
        __syncwarp();  // plain path: the group's stores are ordered before lane 0's (releasing) arrive
        if (!bulk && lane == 0) mbar_arrive(&bars->full[slot]);
    }
    // produce() for task `ti` of the launch (ti >= ntasks: a padding group)
    __device__ __forceinline__ void produce_task(int slot, const Task *__restrict__ tasks, int ntasks, int ti,
                                                 const uint8_t *__restrict__ pool, int pad_code) {
        const bool valid = ti < ntasks;
        uint32_t off_r = 0, off_c = 0;
        int lr = 1, lc = 1;
        if (valid) {
            off_r = tasks[ti].off_r; off_c = tasks[ti].off_c;
            lr = tasks[ti].lr; lc = tasks[ti].lc;
        }
        produce(slot, pool + off_r, pool + off_c, lr, lc, valid, pad_code);
    }
    __device__ __forceinline__ void wait_full(int slot) {
        mbar_wait_uniform(&bars->full[slot], (fpar >> slot) & 1u);
        fpar ^= 1u << slot;
    }
    // After the calling lane's LAST read of the slot.
    __device__ __forceinline__ void release(int slot) { mbar_arrive(&bars->empty[slot]); }
};

// Hands out batches of a persistent launch: one atomic per warp and batch (no wave-quantisation tail).  With a batch
// list (the batches another kernel declined) the counter indexes the list.  Returns -1 when the work is exhausted.
__device__ __forceinline__ int fetch_batch(int *work_counter, int nbatches, const int *__restrict__ batch_list,
                                           const int *__restrict__ batch_count) {
    int b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (batch_list != nullptr) {
        if (b >= *batch_count) return -1;
        b = batch_list[b];
    }
    return (b < nbatches) ? b : -1;
}

// Shared-memory size and ring depth of a launch: two slots unless the second one costs a resident CTA (long operands) or
// does not fit at all.  fixed = tables + barriers, ring1 = bytes of ONE slot for all groups of the CTA.  The answer depends
// only on (kernel, device, sizes) and costs three driver calls, so it is remembered (small batches are launch-latency bound).
struct StageCfg { size_t smem; int nslots, per_sm; };
template <typename KernelT>
static inline cudaError_t stage_ring_config(KernelT kern, size_t fixed, size_t ring1, int threads, size_t &smem, int &nslots,
                                            int &per_sm) {
    constexpr size_t SMEM_MAX = 227 * 1024;
    static std::mutex mu;
    static std::map<std::tuple<const void *, int, size_t, size_t>, StageCfg> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple((const void *) kern, dev, fixed, ring1);
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            smem = it->second.smem; nslots = it->second.nslots; per_sm = it->second.per_sm;
            return cudaSuccess;
        }
    }
    const size_t s1 = fixed + ring1, s2 = fixed + 2 * ring1;
    // always the architectural maximum: the attribute belongs to the (function, device), which several contexts and host
    // threads share -- a per-launch value would race between them
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM_MAX);
    if (e != cudaSuccess) return e;
    int occ1 = 0, occ2 = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, kern, threads, s1);
    if (e != cudaSuccess) return e;
    if (s2 <= SMEM_MAX) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, threads, s2);
        if (e != cudaSuccess) return e;
    }
    nslots = (occ2 >= occ1 && occ2 >= 1) ? 2 : 1;
    smem = (nslots == 2) ? s2 : s1;
    per_sm = std::max(1, nslots == 2 ? occ2 : occ1);
    std::lock_guard<std::mutex> g(mu);
    cache[key] = StageCfg{smem, nslots, per_sm};
    return cudaSuccess;
}

}  // namespace poyb200
